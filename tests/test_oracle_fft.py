"""The oracle's own FFT (cuFFT is closed, FFTW absent) against numpy's float64 rfft."""
import numpy as np
import pytest


@pytest.mark.parametrize("n", [12500, 2500, 500, 20, 10, 1000, 96])
def test_rfft_matches_numpy(orc, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((3, n)).astype(np.float32)
    got = orc.rfft(x)
    ref = np.fft.rfft(x.astype(np.float64), axis=-1)
    scale = np.sqrt((np.abs(ref) ** 2).mean())
    assert np.abs(got - ref).max() / scale < 2e-6


def test_rfft_unnormalised_dc_and_nyquist(orc):
    x = np.ones(12500, np.float32)
    X = orc.rfft(x)
    assert abs(X[0].real - 12500) < 1e-2 and np.abs(X[1:]).max() < 1e-2
    x[1::2] = -1
    X = orc.rfft(x)
    assert abs(X[6250].real - 12500) < 1e-2 and np.abs(X[:6250]).max() < 1e-2


def test_rfft_linearity_and_parseval(orc):
    rng = np.random.default_rng(5)
    a = rng.standard_normal(12500).astype(np.float32)
    b = rng.standard_normal(12500).astype(np.float32)
    A, B, S = orc.rfft(a), orc.rfft(b), orc.rfft(a + b)
    assert np.abs(S - (A + B)).max() < 2e-3
    e_t = float((a.astype(np.float64) ** 2).sum())
    e_f = (np.abs(A[0]) ** 2 + np.abs(A[-1]) ** 2 + 2 * (np.abs(A[1:-1]) ** 2).sum()) / 12500
    assert abs(e_t - e_f) / e_t < 1e-5
