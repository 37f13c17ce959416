"""-m gpu: stage-level parity of the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): integer stages bit exact; FFT bins within 1e-4 relative; soft bits within
+-1 LSB for >= 99.99 %.
"""
import numpy as np
import pytest

import helpers
from dabstar_b200 import api, synth

pytestmark = pytest.mark.gpu


def test_tables(ctx, oracle):
    assert (api.FreqInterleaver(ctx).table == oracle.freq_interleaver()).all()
    assert np.array_equal(api.PhaseReference(ctx).mRefTable, oracle.phase_table())
    for sf, lvl, br, _ in helpers.ALL_PROFILES:
        assert np.array_equal(api.Protection(sf, br, lvl, ctx).addresses(), oracle.protection_addresses(sf, br, lvl)), (sf, lvl, br)


def test_fft_bins(ctx, oracle):
    rng = np.random.default_rng(1)
    x = (rng.normal(size=(8, 2048)) + 1j * rng.normal(size=(8, 2048))).astype(np.complex64)
    for sign in (-1, +1):
        y = ctx.fft2048(x, sign)
        ref = np.stack([oracle.fft2048(r, sign) for r in x])
        rel = np.abs(y - ref).max() / np.abs(ref).max()
        assert rel < 1e-4, rel          # north_star tolerance
        assert rel < 2e-6, rel          # what a float32 FFT should reach


@pytest.mark.parametrize("frame_bits,sigma", [(768, 0.0), (768, 60.0), (768, 120.0), (768, 200.0), (1728, 150.0), (192, 100.0), (3072, 180.0), (9216, 150.0)])
def test_viterbi_bit_exact(ctx, oracle, frame_bits, sigma, viterbi_path):
    n = 24 if frame_bits < 4000 else 3
    _, soft = helpers.random_codewords(n, frame_bits, sigma, seed=frame_bits + int(sigma))
    got = api.ViterbiSpiral(frame_bits, ctx).deconvolve(soft)
    want = np.stack([oracle.viterbi(s, frame_bits) for s in soft])
    assert np.array_equal(got, want), f"{(got != want).sum()} differing bits"


def test_viterbi_saturating_inputs_and_ties(ctx, oracle, viterbi_path):
    # clamp to [0,255] after +127 (viterbi_scalar.h:34-40); all-zero input = all erasures = every comparison ties
    rng = np.random.default_rng(5)
    soft = rng.integers(-32768, 32767, (6, 4 * (768 + 6))).astype(np.int16)
    soft[0] = 0
    soft[1] = 32767
    soft[2] = -32768
    got = api.ViterbiSpiral(768, ctx).deconvolve(soft)
    want = np.stack([oracle.viterbi(s, 768) for s in soft])
    assert np.array_equal(got, want)


def test_viterbi_ragged_batch(ctx, oracle, viterbi_path):
    sizes = [192, 768, 1536, 768, 2304, 192, 190, 331, 5]  # incl. lengths that are not a multiple of 4 (byte-wise output path)
    softs = [helpers.random_codewords(1, n, 140.0, seed=100 + i)[1][0] for i, n in enumerate(sizes)]
    got = api.viterbi_ragged(ctx, softs, sizes)
    for g, s, n in zip(got, softs, sizes):
        assert np.array_equal(g, oracle.viterbi(s, n))
    assert api.viterbi_ragged(ctx, [], []) == []


@pytest.mark.parametrize("sf,lvl,br,size_cu", helpers.ALL_PROFILES)
def test_protection_deconvolve(ctx, oracle, sf, lvl, br, size_cu, viterbi_path):
    rng = np.random.default_rng(br + lvl)
    n = 6
    soft = rng.integers(-200, 200, (n, size_cu * 64)).astype(np.int16)
    got = api.Protection(sf, br, lvl, ctx).deconvolve(soft, size_cu)
    want = np.stack([oracle.protection(sf, br, lvl, s) for s in soft])
    assert np.array_equal(got, want)


def test_fic_decode(ctx, oracle, viterbi_path):
    rng = np.random.default_rng(9)
    n = 12
    fic_addr = oracle.fic_addresses()
    prbs = oracle.prbs(768)
    soft = np.zeros((n, 3, 3072), np.int16)
    for f in range(n):
        flat = soft[f].reshape(-1)
        for blk in range(4):
            bits = rng.integers(0, 2, 768, dtype=np.uint8)
            for fib in range(3):
                crc = oracle.calc_crc(np.packbits(bits[fib * 256:fib * 256 + 240]))
                bits[fib * 256 + 240:fib * 256 + 256] = [(crc >> (15 - b)) & 1 for b in range(16)]
            sigma = [0.0, 90.0, 200.0, 320.0][(f + blk) % 4]
            coded = helpers.soft_from_coded(helpers.conv_encode(bits ^ prbs), sigma, rng)
            flat[blk * 2304:(blk + 1) * 2304] = coded[fic_addr]
    bits, crc, ber, ratio = api.FicDecoder(ctx).process_frames(soft.reshape(n, -1))
    obits, ovalid, oratio, ober = oracle.fic_decode_frames(soft)
    assert np.array_equal(bits, obits)
    assert np.array_equal(crc.reshape(n, 4, 3).all(axis=2).astype(np.uint8), ovalid)
    assert np.array_equal(ratio, oratio)
    assert crc.any() and not crc.all()  # the noise levels above must exercise both outcomes
    # per-block BER counters against ViterbiSpiral::calculate_BER on the same block
    punct = np.zeros(3096, np.uint8)
    punct[fic_addr] = 1
    for f in range(0, n, 5):
        for blk in range(4):
            block = np.zeros(3096, np.int16)
            block[fic_addr] = soft[f].reshape(-1)[blk * 2304:(blk + 1) * 2304]
            raw = bits[f, blk * 768:(blk + 1) * 768] ^ prbs
            assert tuple(ber[f, blk]) == oracle.viterbi_ber(block, punct, raw)


@pytest.mark.parametrize("sf,lvl,br,size_cu", [helpers.EEP_A_72[2], helpers.EEP_B_64[1], helpers.UEP_128[2]])
def test_backend_time_deinterleave(ctx, oracle, sf, lvl, br, size_cu, viterbi_path):
    rng = np.random.default_rng(br)
    n_cifs = 22
    cifs = rng.integers(-150, 150, (n_cifs, 55296)).astype(np.int16)
    start_cu = 37
    got = api.Backend(synth.SubChannel(1, start_cu, size_cu, sf, lvl, br), ctx).process(cifs)
    want, first = oracle.backend_run(start_cu, size_cu, sf, lvl, br, cifs)
    assert first == 16 and got.shape == want.shape
    assert np.array_equal(got, want)
    assert api.Backend(synth.SubChannel(1, start_cu, size_cu, sf, lvl, br), ctx).process(cifs[:10]).shape[0] == 0  # fewer than 17 CIFs: nothing


def test_viterbi_large_batch(ctx, oracle):
    """Default dispatch: >= 2048 code words per launch take the thread-per-code-word kernels (gather, decode)."""
    n = 2304
    sig = np.array([0.0, 80.0, 150.0, 220.0, 400.0])
    rng = np.random.default_rng(77)
    bits = rng.integers(0, 2, (n, 768), dtype=np.uint8)
    coded = np.stack([helpers.conv_encode(b) for b in bits[:64]])
    coded = coded[rng.integers(0, 64, n)]  # 64 distinct code words, independent noise per row
    soft = np.clip(np.trunc((2.0 * coded - 1.0) * 127.0 + rng.normal(size=coded.shape) * sig[np.arange(n) % 5, None]), -32768, 32767).astype(np.int16)
    soft[7] = 0
    soft[8] = 32767
    soft[9] = -32768
    launches = ctx.kernel_launches
    got = api.ViterbiSpiral(768, ctx).deconvolve(soft)
    assert ctx.kernel_launches - launches >= 2  # gather + decode (one warp-per-code-word launch would count 1)
    want = np.stack([oracle.viterbi(s, 768) for s in soft])
    assert np.array_equal(got, want), f"{(got != want).sum()} differing bits"


def _frames_fft(oracle, n_frames, snr_db, seed, cfo=0.0):
    """FFT rows of synthetic frames as the reference's own chain produced them (tap of the oracle chain)."""
    rec = synth.generate(n_frames + 1, seed=seed, snr_db=snr_db, cfo_hz=cfo, fmt=synth.FMT_CF32)
    r = oracle.chain_run(rec.iq, scan_mode=1, tap_fft=True, tap_soft=True)
    assert r.n_frames >= n_frames
    fft = np.stack([r.fft(i) for i in range(n_frames)])
    soft = np.stack([r.soft_bits(i) for i in range(n_frames)])
    ce = np.array([r.info[i].clock_err for i in range(n_frames)], np.float32)
    return fft, soft, ce


@pytest.mark.parametrize("soft_type", [0, 1, 2])
def test_ofdm_decoder_soft_bits(ctx, oracle, soft_type):
    fft, _, ce = _frames_fft(oracle, 6, 18.0, seed=30 + soft_type)
    dec = api.OfdmDecoder(soft_type, ctx)
    got = dec.decode_frames(fft, ce)
    h = oracle.ofdm_new(soft_type)
    want = np.zeros_like(got)
    for f in range(fft.shape[0]):
        oracle.ofdm_store_reference(h, fft[f, 0])
        for s in range(1, 76):
            want[f, s - 1] = oracle.ofdm_decode_symbol(h, fft[f, s], s, 0.0, float(ce[f]))
        oracle.ofdm_store_null(h, fft[f, 76])
    diff = np.abs(got.astype(np.int32) - want.astype(np.int32))
    frac_gt1 = (diff > 1).mean()
    assert frac_gt1 <= 1e-4, f"{frac_gt1:.2e} of soft bits differ by more than 1 LSB (max {diff.max()})"
    # state vectors after the run (nominal carrier order; null power is stored per carrier on the device)
    bins = oracle.freq_interleaver().astype(np.int64)
    idx = np.where(bins < 0, bins + 2048, bins)
    for which, tol in ((0, 1e-4), (1, 1e-3), (2, 1e-3), (3, 1e-3)):
        a, b = dec.state(which), oracle.ofdm_state(h, which)
        assert np.allclose(a, b, rtol=tol, atol=1e-6), which
    assert np.allclose(dec.state(4), oracle.ofdm_state(h, 4)[idx], rtol=1e-4, atol=1e-12)
    assert np.isclose(dec.state(5)[0], oracle.ofdm_state(h, 5)[0], rtol=1e-4)
    # mMeanPowerOvrAll (a serial IIR over carriers in the reference, a per-carrier form on the device) and the LCD figures
    assert np.isclose(dec.state(5)[1], oracle.ofdm_state(h, 5)[1], rtol=1e-4)
    q, qo = dec.quality(), oracle.ofdm_quality(h)
    assert abs(q["mer_db"] - qo["mer_db"]) < 0.01 and abs(q["snr_db"] - qo["snr_db"]) < 0.01, (q, qo)
    assert np.isclose(q["noise_power"], qo["noise_power"], rtol=1e-4) and np.isclose(q["mean_power_overall"], qo["mean_power_overall"], rtol=1e-4)
    oracle.ofdm_free(h)


def test_prs_correlate_and_coarse_afc(ctx, oracle):
    rec = synth.generate(3, seed=41, snr_db=12.0, fmt=synth.FMT_CF32, lead_samples=0, tail_samples=0)
    base = 2656  # first sample of the PRS cyclic prefix in frame 0
    windows, want = [], []
    for off in (0, 100, 250, 504 - 254, 1, 503):
        w = rec.iq[base + off: base + off + 2048]
        windows.append(w)
        want.append(oracle.phaseref_correlate(w, 3.0))
    windows.append(np.zeros(2048, np.complex64))  # all-zero input: the reference returns -1 before dividing (phasereference.cpp:126-129)
    want.append(oracle.phaseref_correlate(windows[-1], 3.0))
    pr = api.PhaseReference(ctx)
    got = pr.correlate_with_phase_ref_and_find_max_peak(np.stack(windows), 3.0)
    assert list(got) == want and want[-1] == -1 and want[0] == 504 and want[5] == -1
    pr.set_sync_on_strongest_peak(True)
    got_s = pr.correlate_with_phase_ref_and_find_max_peak(np.stack(windows), 3.0)
    assert list(got_s) == [oracle.phaseref_correlate(w, 3.0, 1) for w in windows]
    # coarse AFC on symbol 0 with integer and fractional carrier offsets
    ffts, want_o = [], []
    n = np.arange(2048)
    for cfo in (0.0, 1000.0, -3000.0, 2400.0, 17000.0, -34000.0):
        sym0 = rec.iq[base + 504: base + 504 + 2048] * np.exp(2j * np.pi * cfo * n / 2048000.0).astype(np.complex64)
        X = oracle.fft2048(sym0.astype(np.complex64), -1)
        ffts.append(X)
        want_o.append(oracle.phaseref_estimate_offset(X))
    got_o = pr.estimate_carrier_offset_from_sync_symbol_0(np.stack(ffts))
    assert np.abs(got_o - np.array(want_o)).max() <= 1, (got_o, want_o)  # (int)(offset*1000) may flip by one on float rounding
    noise = (np.random.default_rng(3).normal(size=2048) + 1j * np.random.default_rng(4).normal(size=2048)).astype(np.complex64)
    assert pr.estimate_carrier_offset_from_sync_symbol_0(noise[None])[0] == oracle.phaseref_estimate_offset(noise)
