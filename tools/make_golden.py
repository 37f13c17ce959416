#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libdabref.so, built from /root/reference by
oracle/ref_build/Makefile). Run in the build container only; the fixtures are committed so that the GPU box (where
/root/reference does not exist) can pin the C restatement and the CUDA path against the reference's own results.

    python tools/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from dabstar_b200 import synth  # noqa: E402
from oracle_api import Oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference"


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = Oracle("dabref")
    os.makedirs(OUT, exist_ok=True)

    # ---- constant tables (SURVEY.md appendix C lists the same values)
    prs = ref.phase_table()
    quad = np.where(np.abs(prs) < 0.5, 7, np.round(np.angle(prs) / (np.pi / 2)).astype(np.int64) % 4).astype(np.uint8)
    prof = {}
    for sf, lvl, br, _ in helpers.ALL_PROFILES + [(0, 1, 8, 6), (0, 0, 128, 192), (1, 5, 32, 16), (1, 1, 320, 416), (1, 4, 64, 42)]:
        prof[f"addr_{sf}_{lvl}_{br}"] = ref.protection_addresses(sf, br, lvl)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), freq_interleaver=ref.freq_interleaver(), prs_quadrant=quad, **prof)

    # ---- Viterbi known answers: noisy code words, scalar reference kernel
    vit = {}
    for n_bits, sigma, seed in ((768, 0.0, 1), (768, 120.0, 2), (768, 200.0, 3), (1728, 160.0, 4), (192, 220.0, 5)):
        _, soft = helpers.random_codewords(3, n_bits, sigma, seed)
        vit[f"soft_{n_bits}_{seed}"] = soft
        vit[f"bits_{n_bits}_{seed}"] = np.packbits(np.stack([ref.viterbi(s, n_bits) for s in soft]), axis=1)
    rng = np.random.default_rng(6)
    wild = rng.integers(-32768, 32767, (3, 4 * 774)).astype(np.int16)
    wild[0] = 0
    vit["soft_768_wild"] = wild
    vit["bits_768_wild"] = np.packbits(np.stack([ref.viterbi(s, 768) for s in wild]), axis=1)
    np.savez_compressed(os.path.join(OUT, "viterbi.npz"), **vit)

    # ---- EEP/UEP deconvolve and Backend (time de-interleave + dispersal) on random soft bits
    msc = {}
    rng = np.random.default_rng(7)
    for sf, lvl, br, cu in helpers.ALL_PROFILES:
        soft = rng.integers(-180, 180, cu * 64).astype(np.int16)
        msc[f"prot_in_{sf}_{lvl}_{br}"] = soft
        msc[f"prot_out_{sf}_{lvl}_{br}"] = np.packbits(ref.protection(sf, br, lvl, soft))
    frag = rng.integers(-120, 120, (19, 54 * 64)).astype(np.int8)  # only the sub-channel's capacity units; int8 keeps it small
    cifs = np.zeros((19, 55296), np.int16)
    cifs[:, 11 * 64:(11 + 54) * 64] = frag
    out, first = ref.backend_run(11, 54, 0, 2, 72, cifs)
    msc["backend_frag_i8"] = frag
    msc["backend_out"] = np.packbits(out, axis=1)
    msc["backend_first"] = np.int32(first)
    np.savez_compressed(os.path.join(OUT, "msc.npz"), **msc)

    # ---- whole chain on synthetic recordings regenerated from a seed (sha256 of the IQ pins the input)
    chain = {}
    cases = {
        "c1": dict(n_frames=20, seed=1, snr_db=20.0, cfo_hz=0.0, subch=[synth.SubChannel(3, 100, 54, 0, 2, 72)]),
        "cfo": dict(n_frames=8, seed=5, snr_db=15.0, cfo_hz=-4300.0, subch=[synth.SubChannel(3, 100, 54, 0, 2, 72)], lead_samples=77777),
        "mix": dict(n_frames=7, seed=7, snr_db=14.0, cfo_hz=0.0,
                    subch=[synth.SubChannel(1, 0, 108, 0, 0, 72), synth.SubChannel(2, 108, 42, 0, 5, 64), synth.SubChannel(4, 150, 96, 1, 3, 128)]),
    }
    for name, kw in cases.items():
        rec = synth.generate(fmt=synth.FMT_U8, **kw)
        iq_f = ((rec.iq.astype(np.float32) - np.float32(127.38)) / np.float32(128.0)).view(np.complex64).reshape(-1)
        r = ref.chain_run(iq_f, synth.subch_table(kw["subch"]), len(kw["subch"]), tap_soft=True, tap_fft=True)
        chain[f"{name}_iq_sha256"] = np.array(sha(rec.iq))
        chain[f"{name}_n_frames"] = np.int32(r.n_frames)
        chain[f"{name}_sym0_pos"] = np.array([i.sym0_pos for i in r.info], np.int64)
        chain[f"{name}_start_index"] = np.array([i.start_index for i in r.info], np.int32)
        chain[f"{name}_fbb"] = np.array([[i.fbb_sym0, i.fbb_data, i.fbb_null, i.fsync, i.phase_cp, i.clock_err] for i in r.info], np.float32)
        chain[f"{name}_ratio"] = np.array([[i.fic_ratio_before, i.fic_ratio_after] for i in r.info], np.int32)
        chain[f"{name}_fic_valid"] = r.fic_valid
        chain[f"{name}_fib"] = np.packbits(r.fib_bits, axis=1)
        chain[f"{name}_good_fibs"] = np.int32(r.n_good_fibs)
        for s in kw["subch"]:
            chain[f"{name}_msc_{s.sub_ch_id}"] = np.packbits(r.msc[s.sub_ch_id], axis=1)
        # soft bits: symbols 1..3 (FIC) of frames 0 and 2, and the FFT bins of frame 1 symbols 0,1,76 for the FFT tolerance check
        chain[f"{name}_soft_f0"] = r.soft_bits(0)[:3]
        chain[f"{name}_soft_f2"] = r.soft_bits(2)[:3]
        chain[f"{name}_soft_f2_hist"] = np.bincount(np.clip(r.soft_bits(2).astype(np.int32).ravel() + 512, 0, 1023), minlength=1024).astype(np.int32)
        chain[f"{name}_fft_f1"] = r.fft(1)[[0, 1, 76]]
        r.close()
    np.savez_compressed(os.path.join(OUT, "chain.npz"), **chain)

    # ---- ETI-NI stream of the reference's own EtiGenerator (eti_generator.cpp) for the mixed ensemble, 9 frames
    kw = dict(n_frames=9, seed=7, snr_db=14.0, cfo_hz=0.0,
              subch=[synth.SubChannel(1, 0, 108, 0, 0, 72), synth.SubChannel(2, 108, 42, 0, 5, 64), synth.SubChannel(4, 150, 96, 1, 3, 128)])
    rec = synth.generate(fmt=synth.FMT_U8, **kw)
    iq_f = ((rec.iq.astype(np.float32) - np.float32(127.38)) / np.float32(128.0)).view(np.complex64).reshape(-1)
    r = ref.chain_run(iq_f, synth.subch_table(kw["subch"]), len(kw["subch"]), eti=True)
    np.savez_compressed(os.path.join(OUT, "eti.npz"), iq_sha256=np.array(sha(rec.iq)), n_frames=np.int32(r.n_frames), eti=r.eti.reshape(-1, 6144))
    r.close()
    # ---- TII detector: its two constant tables, extracted from the reference source text (tii_detector.cpp:19-125), and the
    #      identifications of the reference's own TiiDetector on seeded spectra
    import re
    from oracle_api import TiiDetector
    src = open(os.path.join(REF_ROOT, "src", "base", "ofdm", "tii_detector.cpp")).read()
    pc = np.array([int(x) for x in re.findall(r"\d+", re.search(r"cPhaseCorrTable = \{(.*?)\};", src, re.S).group(1))], np.uint8)
    pat = np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]{2})", re.search(r"cMainIdPatternTable = \{(.*?)\};", src, re.S).group(1))], np.uint8)
    np.savez_compressed(os.path.join(OUT, "tii_tables.npz"), main_id_pattern=pat, phase_corr=pc)
    rng = np.random.default_rng(99)
    ids = [(3, 0, 0.0), (44, 17, 1.0), (69, 23, -2.0)]
    det = TiiDetector(ref)
    spectra, results = [], []
    for rnd in range(3):
        for _ in range(2):
            x = helpers.tii_spectrum(ids, rng, amp=[40.0, 25.0, 60.0][rnd])
            spectra.append(x)
            det.add(x)
        results.append(np.array(det.process(8), np.float64).reshape(-1, 5))
    np.savez_compressed(os.path.join(OUT, "tii_known_answers.npz"), spectra=np.stack(spectra), r0=results[0], r1=results[1], r2=results[2])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
