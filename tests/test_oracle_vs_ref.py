"""Direct comparison of the C restatement with the reference's own objects (oracle/_ref), where /root/reference was
available to build them. Skipped on the GPU box."""
import ctypes

import numpy as np
import pytest

import helpers
from dabstar_b200 import synth


def test_tables_and_crc(oracle, refo):
    assert np.array_equal(oracle.freq_interleaver(), refo.freq_interleaver())
    assert np.array_equal(oracle.phase_table(), refo.phase_table())
    rng = np.random.default_rng(0)
    for n in (1, 30, 32, 255):
        d = rng.integers(0, 256, n).astype(np.uint8)
        assert oracle.calc_crc(d) == refo.calc_crc(d)
    for _ in range(50):
        bits = rng.integers(0, 2, 256).astype(np.uint8)
        assert oracle.check_crc_bits(bits) == refo.check_crc_bits(bits)
    good = np.zeros(256, np.uint8)
    crc = refo.calc_crc(np.zeros(30, np.uint8))
    good[240:] = [(crc >> (15 - b)) & 1 for b in range(16)]
    assert oracle.check_crc_bits(good) and refo.check_crc_bits(good)


def test_all_protection_profiles(oracle, refo):
    uep = [(32, l) for l in range(1, 6)] + [(56, l) for l in range(2, 6)] + [(128, l) for l in range(1, 6)] + [(320, 5), (320, 4), (320, 2)]
    # 384 kbit/s is left out: the reference's i16 viterbiCounter overflows there (undefined behaviour)
    for br, lvl in uep:
        assert np.array_equal(oracle.protection_addresses(1, br, lvl), refo.protection_addresses(1, br, lvl)), (br, lvl)
    for lvl in range(8):
        for br in ((8, 16, 64, 72, 128, 192) if lvl < 4 else (32, 64, 128, 192)):
            assert np.array_equal(oracle.protection_addresses(0, br, lvl), refo.protection_addresses(0, br, lvl)), (br, lvl)


@pytest.mark.parametrize("frame_bits,sigma", [(768, 0.0), (768, 60.0), (768, 120.0), (768, 200.0), (1728, 180.0), (3072, 200.0)])
def test_viterbi(oracle, refo, frame_bits, sigma):
    _, soft = helpers.random_codewords(10, frame_bits, sigma, seed=int(sigma) + frame_bits)
    for s in soft:
        assert np.array_equal(oracle.viterbi(s, frame_bits), refo.viterbi(s, frame_bits))


def test_fic_and_backend(oracle, refo):
    rng = np.random.default_rng(3)
    soft = rng.integers(-300, 300, (6, 3, 3072)).astype(np.int16)
    a, b = oracle.fic_decode_frames(soft), refo.fic_decode_frames(soft)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    cifs = rng.integers(-100, 100, (20, 55296)).astype(np.int16)
    for sf, lvl, br, cu in (helpers.EEP_A_72[2], helpers.UEP_128[4]):
        x, fx = oracle.backend_run(5, cu, sf, lvl, br, cifs)
        y, fy = refo.backend_run(5, cu, sf, lvl, br, cifs)
        assert fx == fy and np.array_equal(x, y)


@pytest.mark.parametrize("cfo,snr", [(0.0, 20.0), (12345.0, 12.0), (-800.0, 8.0)])
def test_whole_chain(oracle, refo, cfo, snr):
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(5, 300, 84, 1, 4, 128, start_frame=1)]
    rec = synth.generate(8, seed=int(snr), snr_db=snr, cfo_hz=cfo, subch=sc, fmt=synth.FMT_CF32)
    a = oracle.chain_run(rec.iq, synth.subch_table(sc), 2, tap_soft=True)
    b = refo.chain_run(rec.iq, synth.subch_table(sc), 2, tap_soft=True)
    assert a.n_frames == b.n_frames and a.n_good_fibs == b.n_good_fibs
    assert [i.sym0_pos for i in a.info] == [i.sym0_pos for i in b.info]
    assert [round(i.fbb_null) for i in a.info] == [round(i.fbb_null) for i in b.info]
    assert np.array_equal(a.fic_valid, b.fic_valid)
    for s in sc:
        assert np.array_equal(a.msc[s.sub_ch_id], b.msc[s.sub_ch_id])
    for f in range(a.n_frames):
        d = np.abs(a.soft_bits(f).astype(np.int32) - b.soft_bits(f).astype(np.int32))
        assert (d > 1).mean() <= 1e-4
    assert np.array_equal(a.counters[:4], b.counters[:4])
    # SLcdData as the reference emitted it last (signal_show_lcd_data; the harness keeps SNR and MER x 1000, truncated)
    assert a.n_lcd >= 1
    assert abs(a.lcd["snr_db"] - b.counters[4] / 1000.0) < 3e-3 and abs(a.lcd["mer_db"] - b.counters[5] / 1000.0) < 3e-3


def test_eti_generator(oracle, refo):
    """EtiGenerator restatement against the reference's own generator (its output file), incl. a recording that ends inside a frame."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(5, 300, 84, 1, 4, 128), synth.SubChannel(9, 400, 30, 0, 7, 64)]
    rec = synth.generate(8, seed=31, snr_db=13.0, subch=sc, fmt=synth.FMT_CF32)
    iq = rec.iq[:rec.iq.size - 120000]  # cut the last frame after its second CIF
    a = oracle.chain_run(iq, synth.subch_table(sc), 3, eti=True)
    b = refo.chain_run(iq, synth.subch_table(sc), 3, eti=True)
    assert a.n_frames == b.n_frames
    assert b.eti.size > 0 and b.eti.size % 6144 == 0
    assert np.array_equal(a.eti, b.eti)


def _tii_close(a, b):
    assert len(a) == len(b), (a, b)
    for x, y in zip(a, b):
        assert x[:2] == y[:2] and x[4] == y[4], (x, y)
        assert abs(x[2] - y[2]) <= 1e-5 * max(1.0, abs(y[2])) and abs(((x[3] - y[3] + 180.0) % 360.0) - 180.0) < 1e-2, (x, y)


@pytest.mark.parametrize("case", ["one", "three", "non_etsi", "collision", "noise_only", "lone_carriers"])
def test_tii_detector(oracle, refo, case):
    """TiiDetector restatement against the reference's own object: same identifications, strengths and phases over several
    processing rounds (the pair-product IIR carries over), incl. non-ETSI phases, collisions and rejected lone carriers."""
    from oracle_api import TiiDetector
    rng = np.random.default_rng(len(case))
    prs = oracle.phase_table()
    ids = {"one": [(12, 5, 0.3)], "three": [(3, 0, 0.0), (44, 17, 1.0), (69, 23, -2.0)], "non_etsi": [(20, 9, 0.5)],
           "collision": [(10, 4, 0.2), (33, 4, 0.9)], "noise_only": [], "lone_carriers": [(7, 2, 0.1)]}[case]
    a, b = TiiDetector(oracle), TiiDetector(refo)
    if case == "collision":
        a.set_collisions(True, 4)
        b.set_collisions(True, 4)
    for rnd in range(4):
        amps = [40.0, 25.0, 60.0, 5.0][rnd]
        for _ in range(3):
            x = helpers.tii_spectrum(ids, rng, amp=amps, non_etsi=case == "non_etsi", prs=prs,
                                     single_carriers=[(100, 300.0), (555, 200.0j)] if case == "lone_carriers" else ())
            a.add(x)
            b.add(x)
        ra, rb = a.process(6 + rnd), b.process(6 + rnd)
        _tii_close(ra, rb)
        if rnd == 0 and ids:
            assert {(r[0], r[1]) for r in ra if r[0] != 99} >= {(m, s) for m, s, _ in ids} or case == "collision"
    a.reset()
    b.reset()
    _tii_close(a.process(8), b.process(8))


def test_tii_tables_are_derived_correctly(oracle):
    """The two constant tables of tii_detector.cpp (:19-125) are derived in the restatement, not copied: check them against the
    values extracted from the reference source (tests/golden/tii_tables.npz, tools/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tii_tables.npz"))
    from oracle_api import TiiDetector
    t = TiiDetector(oracle)
    pat, pc = np.zeros(70, np.uint8), np.zeros(768, np.uint8)
    oracle.f("tii_tables")(t.h, pat.ctypes.data_as(ctypes.c_void_p), pc.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(pat, g["main_id_pattern"]) and np.array_equal(pc, g["phase_corr"])
    assert list(pat) == helpers.TII_PATTERNS


def test_whole_chain_with_tii_null_symbols(oracle, refo):
    """get_cif_count() following the recording's own FIG 0/0 (track_cif): the reference's DabProcessor then treats the null symbol
    after every second frame as a TII symbol and leaves the null power alone (dab_processor.cpp:273-285). Restatement = reference,
    and the result differs from the run whose FIB decoder holds no counter."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(10, seed=21, snr_db=14.0, subch=sc, fmt=synth.FMT_CF32, fig_mode=1, tii=(12, 5))
    a = oracle.chain_run(rec.iq, synth.subch_table(sc), 1, tap_soft=True, track_cif=True)
    b = refo.chain_run(rec.iq, synth.subch_table(sc), 1, tap_soft=True, track_cif=True)
    plain = oracle.chain_run(rec.iq, synth.subch_table(sc), 1, tap_soft=True)
    assert a.n_frames == b.n_frames == 10 and np.array_equal(a.fic_valid, b.fic_valid) and np.array_equal(a.msc[3], b.msc[3])
    differs = 0
    for f in range(a.n_frames):
        d = np.abs(a.soft_bits(f).astype(np.int32) - b.soft_bits(f).astype(np.int32))
        assert (d > 1).mean() <= 1e-4
        differs += int((a.soft_bits(f) != plain.soft_bits(f)).sum())
    assert differs > 1000  # the TII energy in every second null symbol does change the soft-bit weights


@pytest.mark.parametrize("soft_type", [0, 1, 2])
def test_ofdm_decoder_three_soft_bit_types(oracle, refo, soft_type):
    """OfdmDecoder::decode_symbol (ofdm_decoder.cpp:147-355) for SOFTDEC1 / 2 / 3: the restatement against the reference's own object,
    same FFT rows (those of the reference chain), four frames incl. the start-up transient, with a non-zero clock error."""
    rec = synth.generate(5, seed=40 + soft_type, snr_db=14.0, fmt=synth.FMT_CF32)
    r = refo.chain_run(rec.iq, scan_mode=1, tap_fft=True)
    assert r.n_frames >= 4
    ho, hr = oracle.ofdm_new(soft_type), refo.ofdm_new(soft_type)
    worst, differ, total = 0.0, 0, 0
    for f in range(4):
        fft = r.fft(f)
        ce = float(r.info[f].clock_err) + 3.5 * f
        oracle.ofdm_store_reference(ho, fft[0]); refo.ofdm_store_reference(hr, fft[0])
        for s in range(1, 76):
            a = oracle.ofdm_decode_symbol(ho, fft[s], s, 0.0, ce).astype(np.int32)
            b = refo.ofdm_decode_symbol(hr, fft[s], s, 0.0, ce).astype(np.int32)
            d = np.abs(a - b)
            assert d.max() <= 1, (soft_type, f, s, d.max())
            worst = max(worst, float((d > 0).mean()))
            differ += int((d > 0).sum()); total += d.size
        oracle.ofdm_store_null(ho, fft[76]); refo.ofdm_store_null(hr, fft[76])
    # a last-bit difference of the float pipeline moves a truncation now and then, nothing more (measured: 3e-6 / 1.6e-5 / 4e-6 of the
    # soft bits differ, by one, for SOFTDEC1 / 2 / 3)
    assert worst <= 5e-3 and differ <= 1e-4 * total, (worst, differ, total)  # (seven of a symbol's 3072 in the SOFTDEC2 start-up)
    for which in range(5):
        sa, sb = oracle.ofdm_state(ho, which), refo.ofdm_state(hr, which)
        assert np.allclose(sa, sb, rtol=2e-3, atol=1e-6), which
    oracle.ofdm_free(ho); refo.ofdm_free(hr)


@pytest.mark.parametrize("soft_type", [1, 2])
def test_whole_chain_other_soft_bit_types(oracle, refo, soft_type):
    """The chain with SOFTDEC2 / SOFTDEC3 selected (slot_soft_bit_gen_type): decoded bits identical, soft bits within 1 LSB."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(7, seed=50 + soft_type, snr_db=13.0, cfo_hz=400.0, subch=sc, fmt=synth.FMT_CF32)
    a = oracle.chain_run(rec.iq, synth.subch_table(sc), 1, soft_bit_type=soft_type, tap_soft=True)
    b = refo.chain_run(rec.iq, synth.subch_table(sc), 1, soft_bit_type=soft_type, tap_soft=True)
    assert a.n_frames == b.n_frames >= 6 and a.n_good_fibs == b.n_good_fibs
    assert [i.sym0_pos for i in a.info] == [i.sym0_pos for i in b.info]
    assert np.array_equal(a.fic_valid, b.fic_valid) and np.array_equal(a.fib_bits, b.fib_bits)
    assert np.array_equal(a.msc[3], b.msc[3])
    for f in range(a.n_frames):
        d = np.abs(a.soft_bits(f).astype(np.int32) - b.soft_bits(f).astype(np.int32))
        assert (d > 1).mean() <= 1e-4, (f, (d > 1).mean())


def test_phase_reference_stage_functions(oracle, refo):
    """PhaseReference::correlate_with_phase_ref_and_find_max_peak (phasereference.cpp:87-213) and
    estimate_carrier_offset_from_sync_symbol_0 (:223-280) called directly on the reference's object and on the restatement:
    timing offsets across the search range, both thresholds, first / strongest peak, two echoes, noise only, all zeros;
    carrier offsets up to +-34 kHz incl. fractional ones and a symbol without a PRS."""
    rng = np.random.default_rng(7)
    base = 2656  # first sample of the PRS cyclic prefix in frame 0
    for snr in (20.0, 6.0, 0.0):
        rec = synth.generate(2, seed=int(60 + snr), snr_db=snr, fmt=synth.FMT_CF32, lead_samples=0, tail_samples=0)
        echo = rec.iq.copy()
        echo[37:] += 0.8 * rec.iq[:-37]  # a second path 37 samples later: two candidates
        for src in (rec.iq, echo):
            for off in (0, 1, 100, 250, 254, 400, 503):
                w = src[base + off: base + off + 2048]
                for thr in (3.0, 6.0):
                    for strongest in (0, 1):
                        assert oracle.phaseref_correlate(w, thr, strongest) == refo.phaseref_correlate(w, thr, strongest), (snr, off, thr, strongest)
        n = np.arange(2048)
        for cfo in (0.0, 1000.0, -3000.0, 2400.0, 2499.0, 17000.0, -34000.0, 34400.0, 123.0):
            sym0 = (rec.iq[base + 504: base + 504 + 2048] * np.exp(2j * np.pi * cfo * n / 2048000.0)).astype(np.complex64)
            X = refo.fft2048(sym0, -1)
            a, b = oracle.phaseref_estimate_offset(X), refo.phaseref_estimate_offset(X)
            assert abs(a - b) <= 1, (snr, cfo, a, b)  # (int)(offset * 1000): a last-bit difference may move the truncation by one
    noise = (rng.normal(size=2048) + 1j * rng.normal(size=2048)).astype(np.complex64)
    zeros = np.zeros(2048, np.complex64)
    for thr in (3.0, 6.0):
        for strongest in (0, 1):
            assert oracle.phaseref_correlate(noise, thr, strongest) == refo.phaseref_correlate(noise, thr, strongest)  # (3 x the mean is within reach of noise)
        assert oracle.phaseref_correlate(zeros, thr) == refo.phaseref_correlate(zeros, thr) == -1  # phasereference.cpp:126-129
    assert oracle.phaseref_estimate_offset(noise) == refo.phaseref_estimate_offset(noise)


T_F = 196608


def _damaged_recordings():
    """A 12-frame recording damaged in the middle in ways that make DabProcessor lose or shift its synchronisation
    (dab_processor.cpp:144-182: PRS peak below threshold -> back to the time syncer; :205-265: AFC and clock loops)."""
    rng = np.random.default_rng(5)
    rec = synth.generate(12, seed=71, snr_db=15.0, cfo_hz=300.0, fmt=synth.FMT_CF32)
    iq = rec.iq
    s = 60000 + 4 * T_F + 5000
    noise = ((rng.normal(size=2 * T_F) + 1j * rng.normal(size=2 * T_F)) * 0.25 / np.sqrt(2)).astype(np.complex64)
    x = iq.copy(); x[s:s + 2 * T_F] = noise
    yield "noise burst of two frames", x
    yield "1000 samples dropped", np.concatenate([iq[:s], iq[s + 1000:]])
    yield "7 samples dropped", np.concatenate([iq[:s], iq[s + 7:]])
    x = iq.copy(); n = np.arange(x.size - s); x[s:] *= np.exp(2j * np.pi * 2700.0 * n / 2048000.0).astype(np.complex64)
    yield "carrier jumps by 2.7 kHz", x
    x = iq.copy(); x[s:s + 3 * T_F] *= 0.02
    yield "deep fade of three frames", x
    x = iq.copy(); x[s:s + T_F // 2] = 0
    yield "half a frame of zeros", x
    x = iq.copy(); x[s:] = 0
    yield "signal ends, zeros follow", x


def test_synchronisation_loss_and_recovery(oracle, refo):
    """The control loops of DabProcessor::run under damage: the restatement follows the reference's own DabProcessor frame for frame
    (positions, integer-Hz derotation, FIC CRCs, soft bits) and signal for signal (dip found / no dip found, reads, stream position)."""
    seen_resync = False
    for name, iq in _damaged_recordings():
        a = oracle.chain_run(iq, scan_mode=1, tap_soft=True)
        b = refo.chain_run(iq, scan_mode=1, tap_soft=True)
        assert a.n_frames == b.n_frames, name
        assert [i.sym0_pos for i in a.info] == [i.sym0_pos for i in b.info], name
        assert [round(i.fbb_null) for i in a.info] == [round(i.fbb_null) for i in b.info], name
        assert np.array_equal(a.fic_valid, b.fic_valid) and np.array_equal(a.fib_bits, b.fib_bits), name
        assert np.array_equal(a.counters[:4], b.counters[:4]), (name, a.counters[:4], b.counters[:4])
        for f in range(a.n_frames):
            d = np.abs(a.soft_bits(f).astype(np.int32) - b.soft_bits(f).astype(np.int32))
            assert (d > 1).mean() <= 1e-4, (name, f)
        seen_resync = seen_resync or a.counters[0] > 1
    assert seen_resync  # at least one of the cases went back through the time syncer


@pytest.mark.parametrize("cfo,snr,seed", [(0.0, 20.0, 20), (12345.0, 12.0, 12), (-800.0, 8.0, 8), (300.0, 15.0, 71)])
def test_control_loop_values_per_frame(oracle, refo, cfo, snr, seed):
    """What DabProcessor carries from frame to frame (dab_processor.cpp:191-265), frame by frame: PRS start index, the clock-error
    IIR and the FIC success ratio identical; the AFC frequencies (fSync, the fBB of each frame segment) within 1e-3 Hz and the
    cyclic-prefix phase within 1e-5 rad of the reference's -ffast-math floats (measured: 6e-5 Hz, 3.4e-7 rad)."""
    rec = synth.generate(10, seed=seed, snr_db=snr, cfo_hz=cfo, fmt=synth.FMT_CF32)
    a, b = oracle.chain_run(rec.iq, scan_mode=1), refo.chain_run(rec.iq, scan_mode=1)
    assert a.n_frames == b.n_frames >= 8
    get = lambda res, f: np.array([getattr(i, f) for i in res.info], dtype=np.float64)
    for f in ("sym0_pos", "start_index", "clock_err", "fic_ratio_before", "fic_ratio_after"):
        assert np.array_equal(get(a, f), get(b, f)), f
    for f in ("fbb_sym0", "fbb_data", "fbb_null", "fsync"):
        assert np.abs(get(a, f) - get(b, f)).max() <= 1e-3, f
        assert np.array_equal(np.round(get(a, f)), np.round(get(b, f))), f  # what the derotation uses: integer Hz
    assert np.abs(get(a, "phase_cp") - get(b, "phase_cp")).max() <= 1e-5


@pytest.mark.parametrize("strongest,threshold", [(1, 3.0), (0, 4.5), (1, 6.0)])
def test_whole_chain_with_an_echo_and_sync_options(oracle, refo, strongest, threshold):
    """set_sync_on_strongest_peak and the correlation threshold through the whole chain, on a two-path channel (an echo 45 samples
    late and 2 dB stronger than the first path: first-peak and strongest-peak synchronisation settle on different positions)."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(8, seed=91, snr_db=16.0, cfo_hz=-150.0, subch=sc, fmt=synth.FMT_CF32)
    iq = rec.iq.copy()
    iq[45:] += 1.26 * rec.iq[:-45]
    a = oracle.chain_run(iq, synth.subch_table(sc), 1, threshold=threshold, strongest_peak=strongest)
    b = refo.chain_run(iq, synth.subch_table(sc), 1, threshold=threshold, strongest_peak=strongest)
    assert a.n_frames == b.n_frames >= 6
    assert [(i.sym0_pos, i.start_index, round(i.fbb_null)) for i in a.info] == [(i.sym0_pos, i.start_index, round(i.fbb_null)) for i in b.info]
    assert np.array_equal(a.fic_valid, b.fic_valid) and np.array_equal(a.fib_bits, b.fib_bits)
    assert np.array_equal(a.msc[3], b.msc[3])
    assert np.array_equal(a.counters[:4], b.counters[:4])


def test_strongest_peak_changes_the_position(oracle):
    rec = synth.generate(6, seed=91, snr_db=16.0, fmt=synth.FMT_CF32)
    iq = rec.iq.copy()
    iq[45:] += 1.26 * rec.iq[:-45]
    first = oracle.chain_run(iq, scan_mode=1, strongest_peak=0)
    strong = oracle.chain_run(iq, scan_mode=1, strongest_peak=1)
    assert first.n_frames == strong.n_frames >= 4
    d = [s.sym0_pos - f.sym0_pos for f, s in zip(first.info, strong.info)]
    assert all(x == 45 for x in d[1:]), d  # the option is live: the strongest path is the late one
