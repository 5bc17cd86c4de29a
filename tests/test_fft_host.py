"""Index arithmetic and butterflies of the kernel's 12500-point two-for-one FFT
(csrc/vf_fft12500.cuh), executed on the CPU by build/vf_fft_hosttest: the same
__host__ __device__ functions the kernel calls, one loop per barrier interval."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, make_input

EXE = os.path.join(ROOT, "build", "vf_fft_hosttest")


def run(b0, b1, mask=0):
    r = subprocess.run([EXE, "%x" % mask, "0", "12499"], input=b0.tobytes() + b1.tobytes(), stdout=subprocess.PIPE, check=True)
    out = np.frombuffer(r.stdout, np.float32)
    P = out[: 2 * 4096].reshape(4096, 2)
    Z = out[2 * 4096:].reshape(12500, 2)
    return P, Z[:, 0] + 1j * Z[:, 1]


def volts(u):
    x = u.astype(np.float64) / 128 - 1
    x[u == 0] = 0
    return x


@pytest.mark.skipif(not os.path.exists(EXE), reason="build/vf_fft_hosttest not built")
@pytest.mark.parametrize("mask", [0, 0x1, 0x1555555, 0x1FFFFFE])
def test_two_for_one_fft(pkg, mask):
    b0, b1 = make_input(pkg, 1, seed=21)
    b0 = b0.copy(); b0[100:105] = 0            # dropped samples unpack to 0.0
    P, Z = run(b0, b1, mask)
    x0, x1 = volts(b0), volts(b1)
    for j in range(25):
        if (mask >> j) & 1:
            x0[500 * j:500 * (j + 1)] = 0
            x1[500 * j:500 * (j + 1)] = 0
    Zr = np.fft.fft(x0 + 1j * x1)
    scale = np.sqrt((np.abs(Zr) ** 2).mean()) + 1e-30
    assert np.abs(Z - Zr).max() / scale < 3e-6
    for pol, x in enumerate((x0, x1)):
        Pr = np.abs(np.fft.rfft(x)[2155:2155 + 4096]) ** 2
        assert np.abs(P[:, pol] - Pr).max() / (Pr.mean() + 1e-30) < 1e-5


EXE6 = os.path.join(ROOT, "build", "vf_fft6250_hosttest")


@pytest.mark.skipif(not os.path.exists(EXE6), reason="build/vf_fft6250_hosttest not built")
@pytest.mark.parametrize("fused", ["plain", "fused"])
@pytest.mark.parametrize("mask", [0, 0x1, 0x1555555, 0x1FFFFFE])
def test_real_input_fft(pkg, mask, fused):
    """the product kernel's transform (csrc/vf_fft6250.cuh): one polarisation as a 6250-point complex FFT plus the
    split pass, separately ("plain") and fused with pass 3 as the kernel runs it ("fused": every one of the 4096
    channels must be written exactly by the 313 units)"""
    b0, _ = make_input(pkg, 1, seed=21)
    b0 = b0.copy(); b0[100:105] = 0
    r = subprocess.run([EXE6, "%x" % mask, fused], input=b0.tobytes(), stdout=subprocess.PIPE, check=True)
    out = np.frombuffer(r.stdout, np.float32)
    P, Z = out[:4096], out[4096:].reshape(6250, 2)
    Z = 2 * (Z[:, 0] + 1j * Z[:, 1])            # the transform runs at half scale (vf6_unpack_pair)
    x = volts(b0)
    for j in range(25):
        if (mask >> j) & 1:
            x[500 * j:500 * (j + 1)] = 0
    Zr = np.fft.fft(x[0::2] + 1j * x[1::2])
    assert np.abs(Z - Zr).max() / (np.sqrt((np.abs(Zr) ** 2).mean()) + 1e-30) < 3e-6
    Pr = np.abs(np.fft.rfft(x)[2155:2155 + 4096]) ** 2
    assert (P >= 0).all()
    assert np.abs(P - Pr).max() / (Pr.mean() + 1e-30) < 1e-5


@pytest.mark.skipif(not os.path.exists(EXE6), reason="build/vf_fft6250_hosttest not built")
def test_fused_pass_writes_every_channel_exactly_once(pkg):
    """the 313 units of the fused pass 3 partition the 4096 kept channels of a polarisation: every channel is written
    by exactly one unit (two would race in the kernel), and nothing lands in the other polarisation's slots"""
    b0, _ = make_input(pkg, 1, seed=22)
    r = subprocess.run([EXE6, "0", "count"], input=b0.tobytes(), stdout=subprocess.PIPE, check=True)
    cnt = np.frombuffer(r.stdout, np.int32)
    assert cnt.shape == (4096,) and (cnt == 1).all(), np.unique(cnt)
