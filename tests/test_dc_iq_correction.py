"""Row I2 (optional part): SampleReader's DC removal and IQ-imbalance correction (sample_reader.cpp:216-243, off by default).
CPU: the oracle's serial restatement is bit exact against the reference's own SampleReader (oracle/_ref). GPU: the scan
kernels evaluate the same recurrences in exact arithmetic (double); the reference's float rounding is the difference, so
the bar is an absolute tolerance on the samples (written below) and identical decoded bits for a recording."""
import numpy as np
import pytest

from dabstar_b200 import api, synth


def impaired(n, seed, dc=(0.031, -0.022), gain_q=1.12, phi=0.07):
    """A noise-like complex stream with a DC offset and an IQ imbalance (Q = gain (q + phi i))."""
    rng = np.random.default_rng(seed)
    i = rng.normal(size=n).astype(np.float32) * 0.2
    q = rng.normal(size=n).astype(np.float32) * 0.2
    return ((i + dc[0]) + 1j * (gain_q * (q + phi * i) + dc[1])).astype(np.complex64)


@pytest.mark.parametrize("do_iq", [False, True])
def test_oracle_equals_reference_sample_reader(oracle, refo, do_iq):
    x = impaired(300000, 3)
    a, b = oracle.dc_iq_correct(x, do_iq), refo.dc_iq_correct(x, do_iq)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert oracle.last_dc == refo.last_dc


def test_oracle_state_carries_over(oracle):
    x = impaired(100000, 4)
    st = np.array([0, 0, 1, 1, 0], np.float32)
    a = np.concatenate([oracle.dc_iq_correct(x[:33333], True, st), oracle.dc_iq_correct(x[33333:], True, st)])
    assert np.array_equal(a, oracle.dc_iq_correct(x, True))


@pytest.mark.gpu
@pytest.mark.parametrize("do_iq", [False, True])
@pytest.mark.parametrize("n", [1, 1023, 1024, 5000, 3_000_001])
def test_dc_iq_correct_against_oracle(ctx, oracle, do_iq, n):
    x = impaired(n, 5 + n % 7)
    got, want = ctx.dc_iq_correct(x, do_iq), oracle.dc_iq_correct(x, do_iq)
    # DC removal: samples are O(0.2..1); the serial float filters carry a rounding random walk of about 1e-7 of that.
    # IQ correction: meanII and meanQQ start at 1 and settle near the signal power (0.04 here); while they are large, one step
    # of the reference's float recurrence moves them by only ~8 ulps, so its round-off is a systematic part (up to ~1 %) of
    # the decay rate and gainQ = sqrt(meanII / meanQQ) differs from the exact recurrence by a few 1e-3 during the first seconds.
    err = np.abs(got - want).max()
    assert err < (2e-3 if do_iq else 3e-6), err
    assert np.abs(got.real - want.real).max() < 3e-6  # the I branch only depends on the DC filters


@pytest.mark.gpu
def test_dc_iq_state_and_convergence(ctx, oracle):
    from dabstar_b200 import _lib
    x = impaired(4_500_000, 9)
    st = _lib.DcIqStateC(0.0, 0.0, 1.0, 1.0, 0.0)
    a = np.concatenate([ctx.dc_iq_correct(x[:1_700_001], True, st), ctx.dc_iq_correct(x[1_700_001:], True, st)])
    b = ctx.dc_iq_correct(x, True)
    assert np.abs(a - b).max() < 1e-6            # float state between the calls
    so = np.array([0, 0, 1, 1, 0], np.float32)
    oracle.dc_iq_correct(x, True, so)
    assert np.allclose([st.mean_i, st.mean_q], so[:2], rtol=2e-4, atol=1e-6)
    assert np.allclose([st.mean_ii, st.mean_qq, st.mean_iq], so[2:], rtol=2e-2, atol=1e-4)  # float start-up transient of the reference, see above
    # after two time constants the DC estimate is within 15 % of the offset put in
    assert abs(st.mean_i - 0.031) < 0.006 and abs(st.mean_q + 0.022) < 0.006


@pytest.mark.gpu
def test_corrected_recording_decodes_like_the_cpu_chain(ctx, oracle):
    """DC offset added to a recording, removed by the pre-pass, decoded: same FIBs and MSC bits as the CPU chain fed with the
    oracle-corrected samples (the correction's float noise is far below one soft-bit step)."""
    sc = [synth.SubChannel(3, 100, 54, 0, 2, 72)]
    rec = synth.generate(8, seed=12, snr_db=18.0, subch=sc, fmt=synth.FMT_CF32)
    x = (rec.iq + np.complex64(0.05 - 0.03j)).astype(np.complex64)
    got_x, want_x = ctx.dc_iq_correct(x, False), oracle.dc_iq_correct(x, False)
    want = oracle.chain_run(want_x, synth.subch_table(sc), 1)
    dp = api.DabProcessor(1, input_format=api.FMT_CF32, ctx=ctx)
    dp.set_audio_channel(0, sc)
    dp.run([got_x])
    got = dp.result(0)
    assert got.n_frames == want.n_frames == 8
    assert np.array_equal(got.fic_valid, want.fic_valid) and got.fic_valid[1:].all()
    ok = want.fic_valid.astype(bool).repeat(768, axis=1)
    assert np.array_equal(got.fib_bits[ok], want.fib_bits[ok])
    assert np.array_equal(got.msc[3], want.msc[3])
