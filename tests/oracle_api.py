"""ctypes front end shared by the two CPU checkers (TEST INFRASTRUCTURE):

* ``Oracle("dabo")``   -> oracle/libdab_oracle.so, the C restatement (travels to the GPU box)
* ``Oracle("dabref")`` -> oracle/_ref/libdabref.so, the UNMODIFIED reference sources (built here only)

Both export the same entry points (oracle/dab_oracle.h, oracle/ref_build/dabref.h).
"""
from __future__ import annotations

import ctypes
import tempfile
import os

import numpy as np

from dabstar_b200 import build

c_p = ctypes.c_void_p


class ChainCfg(ctypes.Structure):
    _fields_ = [("soft_bit_type", ctypes.c_int), ("threshold", ctypes.c_float), ("strongest_peak", ctypes.c_int),
                ("scan_mode", ctypes.c_int), ("tap_soft_bits", ctypes.c_int), ("tap_fft", ctypes.c_int),
                ("n_subch", ctypes.c_int), ("subch", c_p), ("eti_path", ctypes.c_char_p),
                ("eti_cif_hi", ctypes.c_int), ("eti_cif_lo", ctypes.c_int),  # these two: restatement only (the dabref stub returns 0, 0)
                ("track_cif", ctypes.c_int)]


class FrameInfo(ctypes.Structure):
    _fields_ = [("sym0_pos", ctypes.c_int64), ("start_index", ctypes.c_int32), ("fbb_sym0", ctypes.c_float),
                ("fbb_data", ctypes.c_float), ("fbb_null", ctypes.c_float), ("fsync", ctypes.c_float),
                ("phase_cp", ctypes.c_float), ("clock_err", ctypes.c_float), ("fic_ratio_before", ctypes.c_int32),
                ("fic_ratio_after", ctypes.c_int32), ("fic_valid", ctypes.c_uint8 * 4)]


class SuperFrame(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("first_frame", "ok", "rs_errors", "rs_corrections", "fc_corrected", "dac_rate", "sbr_flag",
                                              "aac_channel_mode", "ps_flag", "mpeg_surround", "num_aus")] + [("au_start", ctypes.c_int32 * 7), ("au_state", ctypes.c_int32 * 6)]

    def key(self):
        return (self.first_frame, self.ok, self.rs_errors, self.rs_corrections, self.fc_corrected, self.dac_rate, self.sbr_flag, self.aac_channel_mode,
                self.ps_flag, self.mpeg_surround, self.num_aus, tuple(self.au_start), tuple(self.au_state))


def ref_available() -> bool:
    return os.path.exists(build.LIB_REF)


def _ptr(a: np.ndarray) -> c_p:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_p)


class Oracle:
    def __init__(self, prefix: str = "dabo"):
        self.prefix = prefix
        if prefix == "dabo":
            path = build.build_oracle()
        elif prefix in ("dabref", "dabref_fast"):
            # dabref_fast: the TIMING build of the same sources (AVX2 Viterbi, x86-64-v3, float FFT shim): bench.py only
            path = build.LIB_REF if prefix == "dabref" else build.LIB_REF_FAST
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} is not built (needs /root/reference)")
            self.prefix = prefix = "dabref"
        else:
            raise ValueError(prefix)
        self.lib = ctypes.CDLL(path)
        for name, res in (("fic_new", c_p), ("backend_new", c_p), ("ofdm_new", c_p), ("phaseref_new", c_p), ("chain_run", c_p),
                          ("calc_crc", ctypes.c_uint16), ("chain_msc_size", ctypes.c_int64), ("chain_msc_copy", ctypes.c_int64),
                          ("chain_seconds", ctypes.c_double)):
            self.f(name).restype = res

    def f(self, name: str):
        return getattr(self.lib, f"{self.prefix}_{name}")

    # ---- tables
    def freq_interleaver(self) -> np.ndarray:
        t = np.zeros(1536, np.int16)
        self.f("freq_interleaver")(_ptr(t))
        return t

    def phase_table(self) -> np.ndarray:
        t = np.zeros(2048, np.complex64)
        self.f("phase_table")(_ptr(t))
        return t

    def fft2048(self, x: np.ndarray, sign: int = -1) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64)
        y = np.zeros(2048, np.complex64)
        self.f("fft2048")(_ptr(x), _ptr(y), ctypes.c_int(sign))
        return y

    def prbs(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.uint8)
        self.f("prbs")(_ptr(out), ctypes.c_int(n))
        return out

    def fic_addresses(self) -> np.ndarray:
        a = np.zeros(2304, np.int32)
        n = self.f("fic_addresses")(_ptr(a), ctypes.c_int(2304))
        assert n == 2304
        return a

    def to_cf32(self, iq: np.ndarray) -> np.ndarray:
        """File sample formats -> complex float exactly as the reference's readers do: u8 (v - 127.38f) / 128
        (raw_reader.cpp:66-70), i16 v / 32768 (xml_reader.cpp:254-372). IEEE float32 arithmetic, same as dabo_convert_*."""
        if iq.dtype == np.complex64:
            return np.ascontiguousarray(iq)
        if iq.dtype == np.uint8:
            f = (iq.astype(np.float32) - np.float32(127.38)) / np.float32(128.0)
        else:
            f = iq.astype(np.float32) / np.float32(32768.0)
        return np.ascontiguousarray(f).view(np.complex64).reshape(-1)

    def dc_iq_correct(self, x: np.ndarray, do_iq: bool, state: np.ndarray | None = None) -> np.ndarray:
        """SampleReader's DC / IQ correction. dabo: serial restatement (state = meanI, meanQ, meanII, meanQQ, meanIQ, updated in
        place); dabref: the reference's own SampleReader (fresh state; self.last_dc = its get_dc_offset())."""
        x = np.ascontiguousarray(x, np.complex64).reshape(-1)
        out = np.zeros_like(x)
        if self.prefix == "dabo":
            st = state if state is not None else np.array([0, 0, 1, 1, 0], np.float32)
            self.f("dc_iq_correct")(_ptr(x), ctypes.c_int64(x.size), int(do_iq), _ptr(st), _ptr(out))
            self.last_dc = (float(st[0]), float(st[1]))
        else:
            dc = np.zeros(2, np.float32)
            self.f("dc_iq_correct")(_ptr(x), ctypes.c_int64(x.size), int(do_iq), _ptr(out), _ptr(dc))
            self.last_dc = (float(dc[0]), float(dc[1]))
        return out

    def resample(self, x: np.ndarray, sample_rate: int, reader: str) -> np.ndarray:
        """XmlReader / WavReader linear-interpolation resampling (C restatement only: the readers need Qt)."""
        x = np.ascontiguousarray(x, np.complex64).reshape(-1)
        out = np.zeros((x.size // (sample_rate // 1000) + 1) * 2048, np.complex64)
        fn = self.f("resample_" + reader)
        fn.restype = ctypes.c_int64
        n = fn(_ptr(x), ctypes.c_int64(x.size), int(sample_rate), _ptr(out))
        return out[:n]

    def convert_samples(self, raw: np.ndarray, container: int, bits: int, msb_first: int, order: int, n_samples: int) -> np.ndarray:
        """XmlReader::readElements_* (xml_reader.cpp:254-800); only in the C restatement (the reader needs Qt)."""
        raw = np.ascontiguousarray(raw, np.uint8)
        out = np.zeros(n_samples, np.complex64)
        fn = self.f("convert_samples")
        fn.restype = ctypes.c_int
        r = fn(raw.ctypes.data_as(ctypes.c_void_p), int(container), int(bits), int(msb_first), int(order), ctypes.c_int64(n_samples),
               out.ctypes.data_as(ctypes.c_void_p))
        assert r == 0
        return out

    def xml_reader_run(self, raw: np.ndarray, sample_rate: int, container: str, bits: int, byte_order: str, iq_order: str, n_samples: int) -> np.ndarray:
        """oracle/_ref only: the reference's own XmlReader thread over a file image `raw` (sample data only, as after the
        descriptor): format conversion + (sample_rate != 2 048 000) resampling, what it put into its ring buffer. Paces itself in
        real time: 1 ms per millisecond of recording."""
        assert self.prefix == "dabref"
        raw = np.ascontiguousarray(raw, np.uint8)
        cap = (n_samples // max(1, sample_rate // 1000) + 4) * 2048
        out = np.zeros(cap, np.complex64)
        fn = self.f("xml_reader_run")
        fn.restype = ctypes.c_int64
        n = fn(_ptr(raw), ctypes.c_int64(raw.size), int(sample_rate), int(bits), container.encode(), byte_order.encode(), iq_order.encode(),
               ctypes.c_int64(n_samples), _ptr(out), ctypes.c_int64(cap))
        assert n >= 0
        return out[:n]

    def wav_reader_run(self, frames: np.ndarray, sample_rate: int) -> np.ndarray:
        """oracle/_ref only: the reference's own WavReader thread over the float frames libsndfile would deliver."""
        assert self.prefix == "dabref"
        x = np.ascontiguousarray(frames, np.complex64).reshape(-1)
        cap = (x.size // max(1, sample_rate // 1000) + 4) * 2048 + 4 * 32768
        out = np.zeros(cap, np.complex64)
        fn = self.f("wav_reader_run")
        fn.restype = ctypes.c_int64
        n = fn(_ptr(x), ctypes.c_int64(x.size), int(sample_rate), _ptr(out), ctypes.c_int64(cap))
        assert n >= 0
        return out[:n]

    # ---- DAB+ outer code
    def rs_decode(self, cw: np.ndarray):
        cw = np.ascontiguousarray(cw, np.uint8)
        out = np.zeros(110, np.uint8)
        return int(self.f("rs_decode")(_ptr(cw), _ptr(out))), out

    def rs_encode(self, data: np.ndarray) -> np.ndarray:
        data = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(120, np.uint8)
        self.f("rs_encode")(_ptr(data), _ptr(out))
        return out

    def firecode_check(self, x: np.ndarray) -> bool:
        return bool(self.f("firecode_check")(_ptr(np.ascontiguousarray(x, np.uint8))))

    def firecode_check_and_correct(self, x: np.ndarray):
        y = np.ascontiguousarray(x, np.uint8).copy()
        return bool(self.f("firecode_check_and_correct")(_ptr(y))), y

    def dabplus_run(self, frame_bits: np.ndarray, bit_rate: int):
        """Mp4Processor over a run of logical frames (restatement only): list of SuperFrame records and the decoded super-frames."""
        b = np.ascontiguousarray(frame_bits, np.uint8).reshape(-1, 24 * bit_rate)
        cap = max(b.shape[0], 1)
        rec = (SuperFrame * cap)()
        pay = np.zeros((cap, 110 * (bit_rate // 8)), np.uint8)
        n = int(self.f("dabplus_run")(_ptr(b), int(bit_rate), int(b.shape[0]), rec, cap, _ptr(pay)))
        return list(rec[:n]), pay[:n]

    def mp4_events(self, frame_bits: np.ndarray, bit_rate: int) -> list[tuple]:
        """oracle/_ref only: the reference's own Mp4Processor (mp4processor.cpp) over a run of logical frames, with a recording
        stand-in for its AAC decoder. Returns what the processor handed to the decoder, in order:
        ("au", stream parameter bits, access-unit bytes) or ("conceal", samples)."""
        assert self.prefix == "dabref"
        b = np.ascontiguousarray(frame_bits, np.uint8).reshape(-1, 24 * bit_rate)
        self.f("mp4_new").restype = c_p
        h = c_p(self.f("mp4_new")(int(bit_rate)))
        try:
            n = int(self.f("mp4_add_frames")(h, _ptr(b), int(b.shape[0])))
            out = []
            buf = np.zeros(1024, np.uint8)
            for i in range(n):
                kind, value = ctypes.c_int(0), ctypes.c_int(0)
                ln = int(self.f("mp4_event")(h, i, ctypes.byref(kind), ctypes.byref(value), _ptr(buf), buf.size))
                out.append(("au", value.value, buf[:ln].tobytes()) if kind.value == 1 else ("conceal", value.value))
            return out
        finally:
            self.f("mp4_free")(h)

    def fib_decoder(self) -> "RefFibDecoder":
        """oracle/_ref only: the reference's own FibDecoder object (decoder/fib_decoder*.cpp)."""
        assert self.prefix == "dabref"
        return RefFibDecoder(self)

    # ---- channel decoding
    def viterbi(self, soft: np.ndarray, frame_bits: int) -> np.ndarray:
        soft = np.ascontiguousarray(soft, np.int16)
        assert soft.size == 4 * (frame_bits + 6)
        out = np.zeros(frame_bits, np.uint8)
        self.f("viterbi")(_ptr(soft), ctypes.c_int(frame_bits), _ptr(out))
        return out

    def viterbi_ber(self, soft: np.ndarray, punct: np.ndarray, bits: np.ndarray) -> tuple[int, int]:
        b, e = ctypes.c_int(0), ctypes.c_int(0)
        self.f("viterbi_ber")(_ptr(np.ascontiguousarray(soft, np.int16)), _ptr(np.ascontiguousarray(punct, np.uint8)),
                              _ptr(np.ascontiguousarray(bits, np.uint8)), ctypes.c_int(bits.size), ctypes.byref(b), ctypes.byref(e))
        return b.value, e.value

    def protection(self, short_form: int, bit_rate: int, prot_level: int, soft: np.ndarray) -> np.ndarray:
        soft = np.ascontiguousarray(soft, np.int16)
        out = np.zeros(24 * bit_rate, np.uint8)
        self.f("protection")(short_form, bit_rate, prot_level, _ptr(soft), ctypes.c_int(soft.size), _ptr(out))
        return out

    def protection_addresses(self, short_form: int, bit_rate: int, prot_level: int) -> np.ndarray:
        cap = 4 * 24 * bit_rate + 24
        a = np.zeros(cap, np.int32)
        n = self.f("protection_addresses")(short_form, bit_rate, prot_level, _ptr(a), cap)
        return a[:n].copy()

    def check_crc_bits(self, bits: np.ndarray) -> bool:
        bits = np.ascontiguousarray(bits, np.uint8)
        return bool(self.f("check_crc_bits")(_ptr(bits), ctypes.c_int(bits.size)))

    def calc_crc(self, data: np.ndarray) -> int:
        data = np.ascontiguousarray(data, np.uint8)
        return int(self.f("calc_crc")(_ptr(data), ctypes.c_int(data.size)))

    def fic_decode_frames(self, soft: np.ndarray):
        """soft: int16[n_frames, 3, 3072] (symbols 1..3). Returns bits[n,3072], valid[n,4], ratio[n]."""
        soft = np.ascontiguousarray(soft, np.int16)
        n = soft.shape[0]
        h = c_p(self.f("fic_new")())
        bits = np.zeros((n, 3072), np.uint8)
        valid = np.zeros((n, 4), np.uint8)
        ratio = np.zeros(n, np.int32)
        ber = np.zeros((n, 2), np.int32)
        r, bb, be = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        for i in range(n):
            for s in range(3):
                self.f("fic_process_block")(h, _ptr(soft[i, s]), ctypes.c_int(s + 1))
            self.f("fic_get")(h, _ptr(bits[i]), _ptr(valid[i]), ctypes.byref(r), ctypes.byref(bb), ctypes.byref(be))
            ratio[i] = r.value
            ber[i] = (bb.value, be.value)
        self.f("fic_free")(h)
        return bits, valid, ratio, ber

    def backend_run(self, start_cu: int, size_cu: int, short_form: int, prot_level: int, bit_rate: int, cifs: np.ndarray):
        """cifs: int16[n, 55296]. Returns (uint8[n_emitted, 24*bitRate], index of first emitting CIF)."""
        cifs = np.ascontiguousarray(cifs, np.int16)
        h = c_p(self.f("backend_new")(1, start_cu, size_cu, short_form, prot_level, bit_rate))
        outs, first = [], None
        for i in range(cifs.shape[0]):
            out = np.zeros(24 * bit_rate, np.uint8)
            frag = np.ascontiguousarray(cifs[i, start_cu * 64:(start_cu + size_cu) * 64])
            if self.f("backend_process")(h, _ptr(frag), _ptr(out)):
                outs.append(out)
                first = i if first is None else first
        self.f("backend_free")(h)
        return (np.stack(outs) if outs else np.zeros((0, 24 * bit_rate), np.uint8)), first

    # ---- OFDM
    def ofdm_new(self, soft_bit_type: int = 0):
        return c_p(self.f("ofdm_new")(soft_bit_type))

    def ofdm_free(self, h):
        self.f("ofdm_free")(h)

    def ofdm_reset(self, h):
        self.f("ofdm_reset")(h)

    def ofdm_store_reference(self, h, fft: np.ndarray):
        self.f("ofdm_store_reference_symbol_0")(h, _ptr(np.ascontiguousarray(fft, np.complex64)))

    def ofdm_store_null(self, h, fft: np.ndarray):
        self.f("ofdm_store_null_symbol_without_tii")(h, _ptr(np.ascontiguousarray(fft, np.complex64)))

    def ofdm_decode_symbol(self, h, fft: np.ndarray, sym_idx: int, phase_corr: float = 0.0, clock_err: float = 0.0) -> np.ndarray:
        out = np.zeros(3072, np.int16)
        self.f("ofdm_decode_symbol")(h, _ptr(np.ascontiguousarray(fft, np.complex64)), ctypes.c_int(sym_idx),
                                     ctypes.c_float(phase_corr), ctypes.c_float(clock_err), _ptr(out))
        return out

    def ofdm_state(self, h, which: int) -> np.ndarray:
        n = {0: 1536, 1: 1536, 2: 1536, 3: 1536, 4: 2048, 5: 2}[which]
        out = np.zeros(n, np.float32)
        self.f("ofdm_get_state")(h, ctypes.c_int(which), _ptr(out))
        return out

    QUALITY = ("mer_db", "snr_db", "mean_value", "mean_power_overall", "noise_power", "sigma_freq_corr")

    def ofdm_quality(self, h) -> dict:
        """SLcdData figures from the current state (restatement only: the reference keeps them private)."""
        out = np.zeros(6, np.float32)
        self.f("ofdm_quality")(h, _ptr(out))
        return dict(zip(self.QUALITY, (float(v) for v in out)))

    def phaseref_correlate(self, samples: np.ndarray, threshold: float, strongest: int = 0) -> int:
        h = c_p(self.f("phaseref_new")())
        r = self.f("phaseref_correlate")(h, _ptr(np.ascontiguousarray(samples[:2048], np.complex64)), ctypes.c_float(threshold), ctypes.c_int(strongest))
        self.f("phaseref_free")(h)
        return int(r)

    def phaseref_estimate_offset(self, fft: np.ndarray) -> int:
        h = c_p(self.f("phaseref_new")())
        r = self.f("phaseref_estimate_offset")(h, _ptr(np.ascontiguousarray(fft, np.complex64)))
        self.f("phaseref_free")(h)
        return int(r)

    # ---- whole chain
    def chain_run(self, iq: np.ndarray, subch_table: np.ndarray | None = None, n_subch: int = 0, soft_bit_type: int = 0,
                  threshold: float = 3.0, strongest_peak: int = 0, scan_mode: int = 0, tap_soft: bool = False, tap_fft: bool = False,
                  eti: bool = False, eti_cif: tuple[int, int] = (0, 0), track_cif: bool = False):
        """eti: also run the reference's EtiGenerator (start_eti_generator before run); the ETI-NI stream is in result.eti.
        track_cif: get_cif_count() follows the recording's FIG 0/0, so DabProcessor treats every second null symbol as a TII symbol."""
        iq = np.ascontiguousarray(iq, np.complex64)
        tab = np.ascontiguousarray(subch_table if subch_table is not None else np.zeros((1, 7)), np.int32)
        eti_path = None
        if eti:
            fd, eti_path = tempfile.mkstemp(suffix=".eti")
            os.close(fd)
        cfg = ChainCfg(soft_bit_type, threshold, strongest_peak, scan_mode, int(tap_soft), int(tap_fft), n_subch, tab.ctypes.data,
                       eti_path.encode() if eti_path else None, int(eti_cif[0]), int(eti_cif[1]), int(track_cif))
        h = c_p(self.f("chain_run")(_ptr(iq), ctypes.c_int64(iq.size), ctypes.byref(cfg)))
        res = ChainResult(self, h, tab[:n_subch].copy())
        if eti_path:
            res.eti = np.fromfile(eti_path, np.uint8)
            os.unlink(eti_path)
        return res


class TiiResultC(ctypes.Structure):
    _fields_ = [("main_id", ctypes.c_int32), ("sub_id", ctypes.c_int32), ("strength", ctypes.c_float), ("phase_deg", ctypes.c_float), ("non_etsi", ctypes.c_int32)]


class TiiDetector:
    """ofdm/tii_detector.h:30-45 through the oracle restatement (dabo) or the reference's own object (dabref)."""

    def __init__(self, o: "Oracle"):
        self.o = o
        o.f("tii_new").restype = c_p
        self.h = c_p(o.f("tii_new")())

    def __del__(self):
        try:
            self.o.f("tii_free")(self.h)
        except Exception:
            pass

    def reset(self):
        self.o.f("tii_reset")(self.h)

    def set_collisions(self, on: bool, sub_id: int = 0):
        self.o.f("tii_set_collisions")(self.h, int(on), int(sub_id))

    def add(self, fft: np.ndarray):
        fft = np.ascontiguousarray(fft, np.complex64).reshape(-1, 2048)
        for row in fft:
            self.o.f("tii_add")(self.h, _ptr(row))

    def process(self, threshold_db: int) -> list[tuple]:
        out = (TiiResultC * 2048)()
        n = int(self.o.f("tii_process")(self.h, int(threshold_db), out, 2048))
        return [(r.main_id, r.sub_id, r.strength, r.phase_deg, r.non_etsi) for r in out[:n]]

    def decoded(self) -> np.ndarray:
        out = np.zeros(768, np.complex64)
        self.o.f("tii_decoded")(self.h, _ptr(out))
        return out


class RefFibDecoder:
    """FibDecoder of the reference behind the harness: process_FIB, then the multiplex configuration it filed."""

    def __init__(self, o: "Oracle"):
        self.o = o
        o.f("fibdec_new").restype = c_p
        self.h = c_p(o.f("fibdec_new")())

    def __del__(self):
        try:
            self.o.f("fibdec_free")(self.h)
        except Exception:
            pass

    def process_FIB(self, fib_bits: np.ndarray):
        b = np.ascontiguousarray(fib_bits, np.uint8).reshape(-1, 256)
        for i, row in enumerate(b):
            self.o.f("fibdec_process")(self.h, _ptr(row), ctypes.c_int(i % 12))

    def get_cif_count(self) -> tuple[int, int]:
        hi, lo = ctypes.c_int(0), ctypes.c_int(0)
        self.o.f("fibdec_cif_count")(self.h, ctypes.byref(hi), ctypes.byref(lo))
        return hi.value, lo.value

    def eid(self) -> int:
        return int(self.o.f("fibdec_eid")(self.h))

    def sub_channels(self) -> list[tuple]:
        """[(sub_ch_id, start_cu, size_cu, short_form, prot_level, bit_rate)] in the order FIG 0/1 filed them."""
        ids = np.zeros(64, np.int8)
        n = int(self.o.f("fibdec_subch_list")(self.h, _ptr(ids), 64))
        out = []
        for i in ids[:n]:
            v = (ctypes.c_int * 7)()
            self.o.f("fibdec_subch")(self.h, int(i), v)
            assert v[0] == 1 and v[1] == int(i)
            out.append((v[1], v[2], v[5], v[3], v[4], v[6]))
        return out

    def components(self) -> list[tuple]:
        """[(SId, TMId, SubChId or SCId, ASCTy / DSCTy, primary, index within the service)]"""
        buf = (ctypes.c_int * (6 * 256))()
        n = int(self.o.f("fibdec_components")(self.h, buf, 256))
        return [(buf[6 * i] & 0xffffffff,) + tuple(buf[6 * i + k] for k in range(1, 6)) for i in range(min(n, 256))]


class ChainResult:
    def __init__(self, o: Oracle, h, tab: np.ndarray):
        self.o, self.h = o, h
        f = o.f
        self.n_frames = int(f("chain_n_frames")(h))
        self.info = []
        self.fib_bits = np.zeros((self.n_frames, 3072), np.uint8)
        for i in range(self.n_frames):
            fi = FrameInfo()
            f("chain_frame_info")(h, ctypes.c_int(i), ctypes.byref(fi))
            self.info.append(fi)
            f("chain_fib_bits")(h, ctypes.c_int(i), _ptr(self.fib_bits[i]))
        self.fic_valid = np.array([list(fi.fic_valid) for fi in self.info], np.uint8).reshape(self.n_frames, 4)
        self.n_good_fibs = int(f("chain_n_good_fibs")(h))
        self.seconds = float(f("chain_seconds")(h))
        cnt = np.zeros(8, np.int64)
        f("chain_counters")(h, _ptr(cnt))
        self.counters = cnt
        self.quality, self.lcd, self.n_lcd = None, None, 0
        if o.prefix == "dabo":
            q = np.zeros(6, np.float32)
            f("chain_quality")(h, _ptr(q))
            self.quality = dict(zip(Oracle.QUALITY, (float(v) for v in q)))
            lcd = np.zeros(6, np.float32)
            self.n_lcd = int(f("chain_lcd")(h, _ptr(lcd)))
            self.lcd = dict(zip(("symbol_no", "sigma_freq_corr", "snr_db", "mer_db", "test1", "test2"), (float(v) for v in lcd)))
        self.msc = {}
        for row in tab:
            sid, br = int(row[0]), int(row[5])
            n = int(f("chain_msc_size")(h, ctypes.c_int(sid)))
            buf = np.zeros(n, np.uint8)
            if n:
                f("chain_msc_copy")(h, ctypes.c_int(sid), _ptr(buf), ctypes.c_int64(n))
            self.msc[sid] = buf.reshape(-1, 24 * br)

    def soft_bits(self, frame: int) -> np.ndarray | None:
        out = np.zeros((75, 3072), np.int16)
        return out if self.o.f("chain_soft_bits")(self.h, ctypes.c_int(frame), _ptr(out)) else None

    def fft(self, frame: int) -> np.ndarray | None:
        out = np.zeros((77, 2048), np.complex64)
        return out if self.o.f("chain_fft")(self.h, ctypes.c_int(frame), _ptr(out)) else None

    def close(self):
        if self.h is not None:
            self.o.f("chain_free")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
