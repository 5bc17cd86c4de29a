import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return ge.load_package()


@pytest.fixture(scope="session")
def orc():
    return ge.load_oracle()


def make_input(pkg, T, seed=1, antenna=0, sample0=0, **kw):
    """pol0, pol1 of one segment of T FFTs from the deterministic generator."""
    g = pkg.GenParams.default(seed=seed, **kw)
    n = T * pkg.NFFT
    return (pkg.gen_samples(g, antenna, 0, sample0, n), pkg.gen_samples(g, antenna, 1, sample0, n))


RFI = dict(rfi_amp=60, rfi_burst_every=3)


def byte_diff(a, b, nbit):
    """(max code difference, fraction of samples that differ) of packed filterbank bytes."""
    a = np.asarray(a, np.uint8); b = np.asarray(b, np.uint8)
    per = 8 // nbit
    m = (1 << nbit) - 1
    worst, ndiff = 0, 0
    for j in range(per):
        ca = (a >> (nbit * j)) & m
        cb = (b >> (nbit * j)) & m
        d = np.abs(ca.astype(int) - cb.astype(int))
        worst = max(worst, int(d.max()) if d.size else 0)
        ndiff += int(np.count_nonzero(d))
    return worst, ndiff / max(1, a.size * per)
