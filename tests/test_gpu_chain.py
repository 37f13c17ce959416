"""-m gpu: whole-path parity (IQ -> FIB / MSC bits) of the CUDA decoder against the CPU oracle chain."""
import numpy as np
import pytest

from dabstar_b200 import api, synth

pytestmark = pytest.mark.gpu

SC_3A = synth.SubChannel(3, 100, 54, 0, 2, 72)


def _run_both(oracle, ctx, rec, subch, fmt, **kw):
    iq_f = oracle.to_cf32(rec.iq)
    want = oracle.chain_run(iq_f, synth.subch_table(subch), len(subch), tap_soft=True, scan_mode=kw.get("scan_mode", 0))
    dp = api.DabProcessor(1, input_format=fmt, ctx=ctx, **kw)
    dp.set_audio_channel(0, subch)
    dp.run([rec.iq])
    return want, dp, dp.result(0)


def _compare(want, dp, got, subch, soft_frames=3):
    assert got.n_frames == want.n_frames
    for a, b in zip(got.info, want.info):
        assert (a.sym0_pos, a.start_index) == (b.sym0_pos, b.start_index)
        assert round(a.fbb_data) == round(b.fbb_data) and round(a.fbb_null) == round(b.fbb_null)
        assert abs(a.fsync - b.fsync) < 0.05 and abs(a.clock_err - b.clock_err) < 1e-3
        assert (a.fic_ratio_before, a.fic_ratio_after) == (b.fic_ratio_before, b.fic_ratio_after)
    assert np.array_equal(got.fic_valid, want.fic_valid)
    ok = want.fic_valid.astype(bool).repeat(768, axis=1)
    assert np.array_equal(got.fib_bits[ok], want.fib_bits[ok])  # CRC-good FIC blocks are what the reference hands on
    assert got.n_good_fibs == want.n_good_fibs
    for s in subch:
        assert np.array_equal(got.msc[s.sub_ch_id], want.msc[s.sub_ch_id]), s
    for f in range(min(soft_frames, got.n_frames)):
        d = np.abs(dp.soft_bits(0, f).astype(np.int32) - want.soft_bits(f).astype(np.int32))
        assert (d > 1).mean() <= 1e-4, (f, d.max(), (d > 1).mean())
    # signal-quality figures at the end of the run (MER, SNR, mMeanValue, mMeanPowerOvrAll, noise, frequency-correction sigma)
    if want.quality is not None and got.n_frames > 0:
        q = dp.quality(0)
        for k, tol in (("mer_db", 0.01), ("snr_db", 0.01), ("sigma_freq_corr", 0.05)):
            assert abs(q[k] - want.quality[k]) < tol, (k, q, want.quality)
        for k in ("mean_value", "mean_power_overall", "noise_power"):
            assert np.isclose(q[k], want.quality[k], rtol=2e-4), (k, q, want.quality)


@pytest.mark.parametrize("fmt", [synth.FMT_U8, synth.FMT_I16, synth.FMT_CF32])
def test_config1_one_dabplus_subchannel(ctx, oracle, fmt, viterbi_path):
    rec = synth.generate(22, seed=1, snr_db=20.0, subch=[SC_3A], fmt=fmt)
    want, dp, got = _run_both(oracle, ctx, rec, [SC_3A], fmt)
    assert want.n_frames == 22 and want.fic_valid.all()
    _compare(want, dp, got, [SC_3A])
    # payload is what was transmitted
    assert np.array_equal(got.fib_bits, rec.fib_truth[:got.n_frames])
    assert np.array_equal(dp.fib_packed(0).reshape(got.n_frames, 384), np.packbits(got.fib_bits, axis=1))  # the read-back format
    assert np.array_equal(got.msc[3], rec.msc_truth[0][:got.msc[3].shape[0]])


@pytest.mark.parametrize("cfo", [700.0, -4300.0, 31000.0])
def test_carrier_offset_acquisition(ctx, oracle, cfo):
    rec = synth.generate(12, seed=5, snr_db=15.0, cfo_hz=cfo, subch=[SC_3A], fmt=synth.FMT_U8, lead_samples=77777)
    want, dp, got = _run_both(oracle, ctx, rec, [SC_3A], synth.FMT_U8)
    assert want.n_frames >= 11
    _compare(want, dp, got, [SC_3A], soft_frames=2)


def test_mixed_eep_uep_ensemble(ctx, oracle, viterbi_path):
    subch = [synth.SubChannel(1, 0, 108, 0, 0, 72), synth.SubChannel(2, 108, 42, 0, 5, 64), synth.SubChannel(4, 150, 96, 1, 3, 128),
             synth.SubChannel(9, 246, 64, 1, 5, 128, start_frame=2), synth.SubChannel(11, 310, 36, 0, 3, 72), synth.SubChannel(12, 346, 140, 1, 1, 128)]
    rec = synth.generate(9, seed=7, snr_db=14.0, subch=subch, fmt=synth.FMT_U8)
    want, dp, got = _run_both(oracle, ctx, rec, subch, synth.FMT_U8)
    _compare(want, dp, got, subch, soft_frames=1)
    assert got.msc[9].shape[0] == got.msc[1].shape[0] - 8  # Backend created two frames later emits 8 logical frames fewer
    # the payload's read-back format: packed 8 bits per byte
    import ctypes
    for s_ in subch:
        buf = np.zeros(got.msc[s_.sub_ch_id].size // 8, np.uint8)
        n = ctx.lib.dabstar_decoder_msc_packed(dp.h, 0, s_.sub_ch_id, buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size))
        assert n == buf.size and np.array_equal(buf, np.packbits(got.msc[s_.sub_ch_id].reshape(-1)))


def test_low_snr_same_crc_counts(ctx, oracle):
    rec = synth.generate(10, seed=11, snr_db=6.0, subch=[SC_3A], fmt=synth.FMT_U8)
    want, dp, got = _run_both(oracle, ctx, rec, [SC_3A], synth.FMT_U8)
    assert got.n_frames == want.n_frames
    assert [i.sym0_pos for i in got.info] == [i.sym0_pos for i in want.info]
    # below the 10 dB parity bar: report, require the CRC pass counts to be close
    assert abs(got.n_good_fibs - want.n_good_fibs) <= max(2, want.n_good_fibs // 50)


def test_scan_mode_and_small_windows(ctx, oracle):
    rec = synth.generate(14, seed=13, snr_db=18.0, fmt=synth.FMT_U8)
    want, dp, got = _run_both(oracle, ctx, rec, [], synth.FMT_U8, scan_mode=True, max_window=3)
    _compare(want, dp, got, [])
    assert got.counters[4] >= 5  # several speculation windows were needed


def test_batch_of_recordings_matches_single(ctx, oracle):
    recs = [synth.generate(6 + i, seed=20 + i, snr_db=12.0 + 3 * i, cfo_hz=(-1) ** i * 900.0 * i, subch=[SC_3A], fmt=synth.FMT_U8, lead_samples=50000 + 7919 * i)
            for i in range(5)]
    dp = api.DabProcessor(len(recs), input_format=synth.FMT_U8, ctx=ctx)
    for i in range(len(recs)):
        dp.set_audio_channel(i, [SC_3A])
    dp.run([r.iq for r in recs])
    for i, rec in enumerate(recs):
        want = oracle.chain_run(oracle.to_cf32(rec.iq), synth.subch_table([SC_3A]), 1)
        got = dp.result(i)
        assert got.n_frames == want.n_frames, i
        assert np.array_equal(got.fic_valid, want.fic_valid)
        assert np.array_equal(got.msc[3], want.msc[3]), i


def test_empty_and_short_inputs(ctx):
    dp = api.DabProcessor(2, input_format=synth.FMT_U8, ctx=ctx)
    dp.run([np.zeros((0, 2), np.uint8), np.full((50000, 2), 127, np.uint8)])
    assert dp.result(0).n_frames == 0 and dp.result(1).n_frames == 0


def test_no_replay_on_a_clean_recording(ctx):
    """A window ends before the first frame whose PRS peak is not where the layout assumed (here: the noise after the last
    frame), instead of being decoded and replayed: every complete frame goes through the FFT/demap/FIC pass exactly once."""
    rec = synth.generate(14, seed=13, snr_db=18.0, fmt=synth.FMT_U8)
    dp = api.DabProcessor(1, input_format=synth.FMT_U8, scan_mode=True, ctx=ctx)
    dp.run([rec.iq])
    got = dp.result(0)
    assert got.n_frames == 14 and got.fic_valid.all()
    assert got.counters[7] <= got.n_frames + 1, got.counters  # + at most one trailing partial frame


def _same_result(a, b):
    assert a.n_frames == b.n_frames
    for x, y in zip(a.info, b.info):
        for f in ("sym0_pos", "start_index", "fbb_sym0", "fbb_data", "fbb_null", "fsync", "phase_cp", "clock_err", "fic_ratio_before", "fic_ratio_after"):
            assert getattr(x, f) == getattr(y, f), f
    assert np.array_equal(a.fic_valid, b.fic_valid) and np.array_equal(a.fib_bits, b.fib_bits)
    assert a.counters[0] == b.counters[0]
    for k in a.msc:
        assert np.array_equal(a.msc[k], b.msc[k])


def test_chunked_host_upload_matches_device_resident(ctx):
    """MEM_HOST input is uploaded in chunks on a second stream and decoded as it arrives (windows sized to what is
    resident); the result must not depend on the chunking or on where the input lives."""
    import torch
    recs = [synth.generate(9 + 2 * i, seed=40 + i, snr_db=14.0 + i, cfo_hz=350.0 * i, subch=[SC_3A], fmt=synth.FMT_U8, lead_samples=30000 + 4111 * i)
            for i in range(4)]
    results = []
    for chunk in (0, 2, 5):
        dp = api.DabProcessor(len(recs), input_format=synth.FMT_U8, upload_chunk_frames=chunk, ctx=ctx)
        for i in range(len(recs)):
            dp.set_audio_channel(i, [SC_3A])
        pinned = [torch.from_numpy(np.ascontiguousarray(r.iq)).pin_memory() for r in recs]
        dp.run_ptrs([t.data_ptr() for t in pinned], [t.shape[0] for t in pinned], api.MEM_HOST)
        results.append([dp.result(i) for i in range(len(recs))])
    dp = api.DabProcessor(len(recs), input_format=synth.FMT_U8, ctx=ctx)
    for i in range(len(recs)):
        dp.set_audio_channel(i, [SC_3A])
    dev = [torch.from_numpy(np.ascontiguousarray(r.iq)).cuda() for r in recs]
    torch.cuda.synchronize()
    dp.run_ptrs([t.data_ptr() for t in dev], [t.shape[0] for t in dev], api.MEM_DEVICE)
    want = [dp.result(i) for i in range(len(recs))]
    for res in results:
        for a, b in zip(res, want):
            _same_result(a, b)
    assert all(w.n_frames >= 8 for w in want)


def test_eti_generator_stream(ctx, oracle, viterbi_path):
    """Row E: the ETI(NI) stream (every sub-channel of every CIF through EtiGenerator's de-interleaver, ETS 300 799 framing)
    is byte for byte what the reference's generator writes, incl. its start-up and a recording that ends inside a frame."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(5, 300, 84, 1, 4, 128), synth.SubChannel(9, 400, 30, 0, 7, 64)]
    rec = synth.generate(10, seed=31, snr_db=13.0, subch=sc, fmt=synth.FMT_U8)
    iq = rec.iq[:rec.iq.shape[0] - 120000]  # the last frame is cut after its second CIF
    want = oracle.chain_run(oracle.to_cf32(iq), synth.subch_table(sc), 3, eti=True)
    dp = api.DabProcessor(1, input_format=api.FMT_U8, ctx=ctx)
    dp.set_audio_channel(0, sc)
    dp.start_eti_generator(0)
    dp.run([iq])
    got = dp.eti(0)
    assert want.eti.size > 0 and got.shape == (want.eti.size // 6144, 6144)
    assert np.array_equal(got, want.eti.reshape(-1, 6144))
    # the Backend outputs of the same run are unaffected
    for s in sc:
        assert np.array_equal(dp.result(0).msc[s.sub_ch_id], want.msc[s.sub_ch_id])
    # scan mode: no Backend output, the ETI stream is the same (dab_processor.cpp:352-360)
    dp2 = api.DabProcessor(1, input_format=api.FMT_U8, scan_mode=True, ctx=ctx)
    dp2.set_audio_channel(0, sc)
    dp2.start_eti_generator(0)
    dp2.run([iq])
    assert np.array_equal(dp2.eti(0), got)
    dp2.stop_eti_generator(0)
    dp2.run([iq])
    assert dp2.eti(0).shape[0] == 0


def test_cpp_harness_end_to_end(ctx, tmp_path):
    """The C++ facades + headless harness (dabstar_b200/host/): self-configured decode of a big-endian QI int16 file, FIC dump,
    sub-channel files and ETI stream equal what the Python mirror of the same ABI produces."""
    import os, shutil, subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dab_file_decode")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(root, "dabstar_b200", "host", "dab_file_decode.cpp"), "-I", os.path.join(root, "include"),
                    "-L", os.path.join(root, "dabstar_b200"), "-ldabstar_b200", "-Wl,-rpath," + os.path.join(root, "dabstar_b200"), "-o", exe], check=True)
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72), synth.SubChannel(5, 300, 84, 1, 4, 128)]
    rec = synth.generate(8, seed=44, snr_db=17.0, subch=sc, fmt=synth.FMT_I16, fig_mode=1)
    raw = np.ascontiguousarray(rec.iq[:, ::-1]).astype(">i2")  # Ordering MSB, iqOrder QI
    path = str(tmp_path / "rec.iq")
    raw.tofile(path)
    prefix = str(tmp_path / "out_")
    r = subprocess.run([exe, "-x", "int16,16,MSB,QI", "-a", "-e", "-o", prefix, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "sub-channel 3: CU 100+54, EEP level 2, 72 kbit/s" in r.stdout and "sub-channel 5: CU 300+84, UEP level 4, 128 kbit/s" in r.stdout
    dp = api.DabProcessor(1, input_format=api.FMT_CF32, ctx=ctx)
    dp.set_auto_config(0)
    dp.start_eti_generator(0)
    dp.run_files([raw.view(np.uint8)], api.SampleFormat("int16", 16, "MSB", "QI"))
    got = dp.result(0)
    assert got.n_frames == 8 and got.fic_valid.all()
    assert np.array_equal(np.fromfile(prefix + "0.eti", np.uint8).reshape(-1, 6144), dp.eti(0))
    assert np.fromfile(prefix + "0.fic", np.uint8).size == 32 * got.n_good_fibs
    for s in sc:
        assert np.array_equal(np.fromfile(prefix + f"0.sub{s.sub_ch_id}", np.uint8), np.packbits(got.msc[s.sub_ch_id].reshape(-1)))
    # the same recording as a RIFF/WAVE file (16-bit PCM, 2.048 MS/s) through the container path (-c): header probe on the host,
    # conversion on the GPU; prints the container and the reference's LCD figures
    from test_file_containers import wav_bytes
    wav = str(tmp_path / "rec.wav")
    with open(wav, "wb") as f:
        f.write(wav_bytes(np.ascontiguousarray(rec.iq), 2048000, 1, 16))
    prefix2 = str(tmp_path / "wav_")
    r = subprocess.run([exe, "-c", "-a", "-o", prefix2, wav], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "wav container, 2048000 S/s" in r.stdout and "MER" in r.stdout and "SNR" in r.stdout
    assert np.array_equal(np.fromfile(prefix2 + "0.fic", np.uint8), np.fromfile(prefix + "0.fic", np.uint8))
    for s in sc:
        assert np.array_equal(np.fromfile(prefix2 + f"0.sub{s.sub_ch_id}", np.uint8), np.fromfile(prefix + f"0.sub{s.sub_ch_id}", np.uint8))


def test_config3_full_ensemble_18_subchannels(ctx, oracle):
    """BASELINE configs[2] in small: 18 sub-channels filling 864 CU (EEP-A, EEP-B and UEP, DAB+ and MP2 bit rates), all decoded
    from one recording in one MSC batch: every sub-channel's logical frames equal the CPU chain's."""
    spec = [(0, 0, 72, 108), (0, 1, 72, 72), (0, 2, 72, 54), (0, 3, 72, 36), (0, 4, 64, 54), (0, 5, 64, 42), (0, 6, 64, 36), (0, 7, 64, 30),
            (1, 3, 128, 96), (1, 4, 128, 84), (1, 5, 128, 64), (0, 2, 48, 36), (0, 2, 32, 24), (0, 6, 32, 18), (1, 5, 32, 16), (0, 3, 32, 16),
            (0, 2, 8, 6), (1, 4, 64, 42)]
    subch, cu = [], 0
    for i, (sf, lvl, br, size) in enumerate(spec):
        subch.append(synth.SubChannel(i + 1, cu, size, sf, lvl, br))
        cu += size
    assert len(subch) == 18 and cu <= 864
    rec = synth.generate(20, seed=3, snr_db=16.0, subch=subch, fmt=synth.FMT_U8)
    want, dp, got = _run_both(oracle, ctx, rec, subch, synth.FMT_U8)
    assert want.n_frames == 20
    _compare(want, dp, got, subch, soft_frames=1)
    for j, s in enumerate(subch):
        assert got.msc[s.sub_ch_id].shape == (4 * 20 - 16, 24 * s.bit_rate)
        assert np.array_equal(got.msc[s.sub_ch_id], rec.msc_truth[j][:got.msc[s.sub_ch_id].shape[0]]), s  # and they are what was sent


def test_config5_batch_snr_cfo_timing_sweep(ctx, oracle):
    """BASELINE configs[4] in small: a batch of independent recordings over SNR 3..30 dB, carrier offsets within +-5 kHz plus the
    +-30 kHz corner, and arbitrary start offsets: positions and frequencies identical, payload identical from 10 dB up, CRC pass
    counts close below."""
    rng = np.random.default_rng(5)
    snrs = [3.0, 6.0, 9.0, 12.0, 15.0, 18.0, 21.0, 24.0, 27.0, 30.0, 14.0, 20.0]
    cfos = list(rng.uniform(-5000.0, 5000.0, 10)) + [30000.0, -30000.0]
    recs = [synth.generate(9, seed=500 + i, snr_db=snrs[i], cfo_hz=float(cfos[i]), subch=[SC_3A], fmt=synth.FMT_U8,
                           lead_samples=int(rng.integers(45000, 45000 + 196608))) for i in range(12)]
    dp = api.DabProcessor(len(recs), input_format=synth.FMT_U8, ctx=ctx)
    for i in range(len(recs)):
        dp.set_audio_channel(i, [SC_3A])
    dp.run([r.iq for r in recs])
    for i, rec in enumerate(recs):
        want = oracle.chain_run(oracle.to_cf32(rec.iq), synth.subch_table([SC_3A]), 1)
        got = dp.result(i)
        if snrs[i] >= 10.0:
            assert got.n_frames == want.n_frames, (i, snrs[i], cfos[i])
            assert [(a.sym0_pos, round(a.fbb_null)) for a in got.info] == [(b.sym0_pos, round(b.fbb_null)) for b in want.info], i
            assert np.array_equal(got.fic_valid, want.fic_valid), i
            assert np.array_equal(got.msc[3], want.msc[3]), i
            assert got.n_good_fibs == want.n_good_fibs
        else:
            assert abs(got.n_frames - want.n_frames) <= 1, (i, snrs[i])
            assert abs(got.n_good_fibs - want.n_good_fibs) <= max(3, want.n_good_fibs // 20), (i, got.n_good_fibs, want.n_good_fibs)
        want.close()


# ---- parity in the shape the bench runs: 104-frame recordings (10 s), a batch in lock step, max_window = 128, scan mode
@pytest.mark.parametrize("fmt,n_rec", [(synth.FMT_U8, 16), (synth.FMT_I16, 6)])
def test_bench_shape_batch_of_10s_recordings(ctx, oracle, fmt, n_rec):
    recs = [synth.generate(104, seed=300 + i, snr_db=15.0, fmt=fmt) for i in range(n_rec)]
    dp = api.DabProcessor(n_rec, input_format=fmt, scan_mode=True, max_window=128, ctx=ctx)
    dp.run([r.iq for r in recs])
    for i, rec in enumerate(recs):
        want = oracle.chain_run(oracle.to_cf32(rec.iq), scan_mode=1, tap_soft=(i % 4 == 0))
        got = dp.result(i)
        assert got.n_frames == want.n_frames == 104, i
        assert [(a.sym0_pos, a.start_index, round(a.fbb_data)) for a in got.info] == [(b.sym0_pos, b.start_index, round(b.fbb_data)) for b in want.info]
        assert np.array_equal(got.fic_valid, want.fic_valid) and want.fic_valid.all()
        assert np.array_equal(got.fib_bits, want.fib_bits) and np.array_equal(got.fib_bits, rec.fib_truth)
        assert got.n_good_fibs == want.n_good_fibs == 12 * 104
        if i % 4 == 0:  # soft bits of the LAST frames: 103 frames of IIR history behind them
            for f in (102, 103):
                d = np.abs(dp.soft_bits(i, f).astype(np.int32) - want.soft_bits(f).astype(np.int32))
                assert (d > 1).mean() <= 1e-4, (i, f, d.max())
        want.close()


def test_config0_ten_seconds_one_dabplus_subchannel(ctx, oracle):
    """BASELINE configs[0] at its full size: 10 s = 104 frames, one DAB+ EEP 3-A 72 kbit/s sub-channel + FIC."""
    rec = synth.generate(104, seed=1, snr_db=20.0, subch=[SC_3A], fmt=synth.FMT_U8)
    want, dp, got = _run_both(oracle, ctx, rec, [SC_3A], synth.FMT_U8, max_window=128)
    assert want.n_frames == 104
    _compare(want, dp, got, [SC_3A], soft_frames=1)
    for f in (52, 103):
        d = np.abs(dp.soft_bits(0, f).astype(np.int32) - want.soft_bits(f).astype(np.int32))
        assert (d > 1).mean() <= 1e-4, (f, d.max())
    assert got.msc[3].shape[0] == 4 * 104 - 16 and np.array_equal(got.msc[3], rec.msc_truth[0][:got.msc[3].shape[0]])


def test_snr_cfo_batch_config4_small(ctx, oracle):
    """BASELINE configs[4] in small: recordings across the SNR sweep with kHz carrier offsets and random timing, in one batch.
    At >= 10 dB everything is identical to the oracle; below, positions and CRC counts are compared."""
    snrs = [3.0, 6.0, 9.0, 12.0, 15.0, 18.0, 21.0, 24.0, 27.0, 30.0]
    rng = np.random.default_rng(9)
    recs = [synth.generate(10, seed=500 + i, snr_db=s, cfo_hz=float(rng.uniform(-5000, 5000)), fmt=synth.FMT_U8, lead_samples=int(rng.integers(45000, 45000 + 196608)))
            for i, s in enumerate(snrs)]
    dp = api.DabProcessor(len(recs), input_format=synth.FMT_U8, scan_mode=True, ctx=ctx)
    dp.run([r.iq for r in recs])
    for i, (rec, snr) in enumerate(zip(recs, snrs)):
        want = oracle.chain_run(oracle.to_cf32(rec.iq), scan_mode=1)
        got = dp.result(i)
        assert got.n_frames == want.n_frames, (snr, got.n_frames, want.n_frames)
        assert [a.sym0_pos for a in got.info] == [b.sym0_pos for b in want.info], snr
        if snr >= 10.0:
            assert np.array_equal(got.fic_valid, want.fic_valid) and [round(a.fbb_data) for a in got.info] == [round(b.fbb_data) for b in want.info]
            ok = want.fic_valid.astype(bool).repeat(768, axis=1)
            assert np.array_equal(got.fib_bits[ok], want.fib_bits[ok])
        else:
            assert abs(got.n_good_fibs - want.n_good_fibs) <= max(3, want.n_good_fibs // 20), (snr, got.n_good_fibs, want.n_good_fibs)
        want.close()
