"""Next row f4: TII detection. The CUDA detector (one CTA per recording) against the oracle restatement of
ofdm/tii_detector.cpp, which is pinned against the reference's own TiiDetector object (tests/test_oracle_vs_ref.py,
tests/golden/tii_known_answers.npz): identifications and flags identical, strengths / phases within float tolerance."""
import os

import numpy as np
import pytest

import helpers
from dabstar_b200 import api

pytestmark = pytest.mark.gpu

CASES = {"one": [(12, 5, 0.3)], "three": [(3, 0, 0.0), (44, 17, 1.0), (69, 23, -2.0)], "non_etsi": [(20, 9, 0.5)],
         "collision": [(10, 4, 0.2), (33, 4, 0.9)], "noise_only": [], "lone_carriers": [(7, 2, 0.1)]}


def _close(a, b):
    assert len(a) == len(b), (a, b)
    for x, y in zip(a, b):
        assert (x[0], x[1], x[4]) == (y[0], y[1], y[4]), (x, y)
        assert abs(x[2] - y[2]) <= 1e-5 * max(1.0, abs(y[2])) and abs(((x[3] - y[3] + 180.0) % 360.0) - 180.0) < 1e-2, (x, y)


@pytest.mark.parametrize("collisions", [False, True])
def test_tii_batch_against_oracle(ctx, oracle, collisions):
    """Six recordings with different transmitter constellations in one batch, four processing rounds (the filtered pair products
    carry over), then reset."""
    from oracle_api import TiiDetector
    rng = np.random.default_rng(11)
    prs = oracle.phase_table()
    names = list(CASES)
    dets = [TiiDetector(oracle) for _ in names]
    gpu = api.TiiDetector(len(names), ctx)
    if collisions:
        gpu.set_detect_collisions(True, 4)
        for d in dets:
            d.set_collisions(True, 4)
    for rnd in range(4):
        amp = [40.0, 25.0, 60.0, 5.0][rnd]
        n_sym = 1 + rnd % 3
        x = np.stack([np.stack([helpers.tii_spectrum(CASES[n], rng, amp=amp, non_etsi=n == "non_etsi", prs=prs,
                                                     single_carriers=[(100, 300.0), (555, 200.0j)] if n == "lone_carriers" else ())
                                for _ in range(n_sym)]) for n in names])
        gpu.add_to_tii_buffer(x)
        for d, xs in zip(dets, x):
            d.add(xs)
        got = gpu.process_tii_data(6 + rnd)
        for i, (d, n) in enumerate(zip(dets, names)):
            want = d.process(6 + rnd)
            _close(got[i], want)
            assert np.allclose(gpu.decoded(i), d.decoded(), rtol=1e-6, atol=1e-6 * np.abs(d.decoded()).max())
            if rnd == 0 and n not in ("collision", "noise_only"):
                assert {(r[0], r[1]) for r in got[i]} == {(m, s) for m, s, _ in CASES[n]}, n
        if rnd == 0:
            assert got[names.index("noise_only")] == []
            assert got[names.index("non_etsi")][0][4] == 1
    gpu.reset()
    for d in dets:
        d.reset()
    for i, d in enumerate(dets):
        _close(gpu.process_tii_data(8)[i], d.process(8))


def test_tii_against_reference_known_answers(ctx):
    """What the reference's own TiiDetector reported for these spectra (tools/make_golden.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tii_known_answers.npz"))
    det = api.TiiDetector(1, ctx)
    for rnd in range(3):
        det.add_to_tii_buffer(g["spectra"][2 * rnd:2 * rnd + 2])
        got, want = det.process_tii_data(8)[0], g[f"r{rnd}"]
        _close(got, [(int(r[0]), int(r[1]), r[2], r[3], int(r[4])) for r in want])


def test_tii_result_cap_and_arguments(ctx):
    det = api.TiiDetector(2, ctx)
    rng = np.random.default_rng(2)
    ids = [(m, s, 0.0) for m, s in ((1, 1), (2, 2), (3, 3), (4, 4), (5, 5))]
    det.add_to_tii_buffer(np.stack([np.stack([helpers.tii_spectrum(ids, rng)]), np.stack([helpers.tii_spectrum([], rng)])]))
    out = (api._lib.TiiResultC * (2 * 3))()
    cnt = np.zeros(2, np.int32)
    ctx.check(ctx.lib.dabstar_tii_process(det.h, 8, out, 3, cnt.ctypes.data_as(api.c_p)), "dabstar_tii_process")
    assert list(cnt) == [5, 0]  # five found, room for three: the count tells
    with pytest.raises(api.DabstarError):
        det.set_detect_collisions(True, 24)


def test_decoder_tii_from_own_cif_counter(ctx, oracle):
    """Self-configured recordings whose transmitter sends TII in the null symbols the CIF counter selects: the decoder finds the
    identification, at the frames the reference's counting predicts, and its result equals the oracle detector fed with the
    null-symbol spectra of the CPU chain for the same frames."""
    from dabstar_b200 import synth
    from oracle_api import TiiDetector
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    recs = [synth.generate(12, seed=50 + i, snr_db=20.0, subch=sc, fmt=synth.FMT_U8, fig_mode=1, tii=t) for i, t in enumerate([(12, 5), (40, 23)])]
    recs.append(synth.generate(8, seed=60, snr_db=20.0, subch=sc, fmt=synth.FMT_U8, fig_mode=1))  # shorter, no TII sent
    dp = api.DabProcessor(3, input_format=api.FMT_U8, ctx=ctx)
    for r in range(3):
        dp.set_auto_config(r)
        dp.set_tii_processing(r, True, frames_to_count=2, threshold_db=8)
    dp.run([r.iq for r in recs])
    for r, want_id in ((0, (12, 5)), (1, (40, 23))):
        ev = dp.tii_events(r)
        nf = dp.result(r).n_frames
        # CIF counter of frame f is 4 f: TII null symbols follow the odd frames; two per search
        assert [e[0] for e in ev] == [f for f in range(3, nf, 4)], (ev, nf)
        for fr, res in ev:
            assert [(x[0], x[1], x[4]) for x in res] == [(want_id[0], want_id[1], 0)], (fr, res)
        # the same searches on the CPU: null-symbol FFTs of the oracle chain (row 76), oracle detector
        chain = oracle.chain_run(oracle.to_cf32(recs[r].iq), synth.subch_table(sc), 1, tap_fft=True)
        det = TiiDetector(oracle)
        for (fr, res) in ev:
            det.add(np.stack([chain.fft(fr - 2)[76], chain.fft(fr)[76]]))
            want = det.process(8)
            assert [(x[0], x[1], x[4]) for x in res] == [(y[0], y[1], y[4]) for y in want]
            assert abs(res[0][2] - want[0][2]) < 1e-3 and abs(((res[0][3] - want[0][3] + 180.0) % 360.0) - 180.0) < 0.5
        chain.close()
    ev2 = dp.tii_events(2)
    assert [e[0] for e in ev2] == [3, 7] and all(res == [] for _, res in ev2)  # searched, nothing above the noise
    # switched off, or without self-configuration: no searches
    dp.set_tii_processing(0, False)
    dp.set_tii_processing(1, False)
    dp.set_tii_processing(2, False)
    dp.run([r.iq for r in recs])
    assert dp.tii_events(0) == []


def test_tii_null_symbols_leave_the_null_power_alone(ctx, oracle):
    """Self-configuration level 2: the demapper skips the null-power update on the null symbols the recording's own CIF counter
    marks as TII symbols, like DabProcessor with a real FIB decoder (dab_processor.cpp:273-285). Equal to the CPU chain whose
    get_cif_count() follows FIG 0/0 (itself pinned against the reference's DabProcessor), different from level 1."""
    from dabstar_b200 import synth
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    recs = [synth.generate(10, seed=21 + i, snr_db=14.0, subch=sc, fmt=synth.FMT_U8, fig_mode=1, tii=(12, 5)) for i in range(2)]
    dp = api.DabProcessor(2, input_format=api.FMT_U8, ctx=ctx)
    dp.set_auto_config(0, True, tii_null_symbols=True)
    dp.set_auto_config(1, True)                       # level 1 beside it: plain null symbols
    dp.run([r.iq for r in recs])
    for r, track in ((0, True), (1, False)):
        want = oracle.chain_run(oracle.to_cf32(recs[r].iq), synth.subch_table(sc), 1, tap_soft=True, track_cif=track)
        got = dp.result(r)
        assert got.n_frames == want.n_frames == 10 and np.array_equal(got.fic_valid, want.fic_valid)
        for f in range(got.n_frames):
            d = np.abs(dp.soft_bits(r, f).astype(np.int32) - want.soft_bits(f).astype(np.int32))
            assert (d > 1).mean() <= 1e-4, (r, f, d.max(), (d > 1).mean())
        n = min(got.msc[3].shape[0], want.msc[3].shape[0])  # the self-configured Backend starts a frame later than the configured one
        assert n >= 16 and np.array_equal(got.msc[3][-n:], want.msc[3][-n:])
        want.close()
    # and the two levels do differ on the same recording
    plain = oracle.chain_run(oracle.to_cf32(recs[0].iq), synth.subch_table(sc), 1, tap_soft=True)
    assert sum(int((dp.soft_bits(0, f) != plain.soft_bits(f)).sum()) for f in range(3, 10)) > 1000
    plain.close()


def test_cpp_harness_reports_tii(ctx, tmp_path):
    """The C++ facade (DabProcessor::set_tii_processing / tii_events) through the headless harness: -a -t on a recording whose
    transmitter sends TII prints the identification found in the null symbols its CIF counter selects."""
    import shutil, subprocess
    from dabstar_b200 import synth
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dab_file_decode")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(root, "dabstar_b200", "host", "dab_file_decode.cpp"), "-I", os.path.join(root, "include"),
                    "-L", os.path.join(root, "dabstar_b200"), "-ldabstar_b200", "-Wl,-rpath," + os.path.join(root, "dabstar_b200"), "-o", exe], check=True)
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(22, seed=71, snr_db=18.0, subch=sc, fmt=synth.FMT_U8, fig_mode=1, tii=(12, 5))
    path = str(tmp_path / "rec.iq")
    rec.iq.tofile(path)
    r = subprocess.run([exe, "-f", "u8", "-a", "-t", "-o", str(tmp_path / "o_"), path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if "TII main id" in l]
    assert len(lines) == 2 and all("TII main id 12 sub id 5" in l for l in lines), r.stdout  # five TII null symbols per search: frames 9 and 19
    assert "frame 9:" in lines[0] and "frame 19:" in lines[1]
