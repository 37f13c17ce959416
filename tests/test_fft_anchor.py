"""The FFT boundary has no known answers in the reference (FFTW3f is third party and un-vendored): the oracle's
transform (oracle/dab_oracle.c) and the stand-in `fftwf_*` shim the reference's own sources are linked against (oracle/ref_build)
are anchored here against an independent library transform, numpy's pocketfft in double precision, with the conventions of the
call sites (dab_processor.cpp:63,199-201: forward, unnormalised; phasereference.cpp:92-105: backward, unnormalised)."""
import numpy as np
import pytest


def _cases():
    rng = np.random.default_rng(11)
    x = (rng.normal(size=(6, 2048)) + 1j * rng.normal(size=(6, 2048))).astype(np.complex64)
    x[4] = 0
    x[4, 1] = 1.0           # one sample: a pure phasor across the bins, sign visible
    x[5] = np.exp(2j * np.pi * 37 * np.arange(2048) / 2048).astype(np.complex64)  # one carrier: all energy in one bin
    return x


def _numpy_fft(x, sign):
    x = x.astype(np.complex128)
    return np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * x.shape[-1]


@pytest.mark.parametrize("sign", [-1, +1])
def test_oracle_fft_against_numpy(oracle, sign):
    for x in _cases():
        want = _numpy_fft(x, sign)
        got = oracle.fft2048(x, sign)
        assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max() + 1e-30


@pytest.mark.parametrize("sign", [-1, +1])
def test_reference_fft_shim_against_numpy(refo, sign):
    for x in _cases():
        want = _numpy_fft(x, sign)
        got = refo.fft2048(x, sign)
        # (the shim is built with the reference's own flags: -fsingle-precision-constant makes its 2 pi a float, measured 3.3e-7)
        assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max() + 1e-30


def test_single_carrier_lands_in_its_bin(oracle):
    x = _cases()[5]
    y = oracle.fft2048(x, -1)
    assert np.argmax(np.abs(y)) == 37 and abs(abs(y[37]) - 2048.0) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("sign", [-1, +1])
def test_cuda_fft_against_numpy(ctx, sign):
    x = _cases()
    got = ctx.fft2048(x, sign)
    want = _numpy_fft(x, sign)
    rel = np.abs(got - want).max(axis=1) / np.abs(want).max(axis=1)
    assert rel.max() < 1e-4, rel   # north_star tolerance
    assert rel.max() < 1e-5, rel   # a float32 transform of this length (2e-6 measured against the oracle, which is within 4e-8 of numpy)
