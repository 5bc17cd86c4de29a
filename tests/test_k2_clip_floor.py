"""The normaliser's clip pre-test (csrc/vf_kernels.cu: vf_k2_clip_floor and phase A of vf_k2_stream) skips the per-step
test p > 11 bp (src/pb_kernels.cu:493-494) for a 16-step chunk whose largest power does not exceed
clip_floor * (bandpass at the start of the chunk).  That is only bit-safe if the bound really is a lower bound of
float32 (11 * bp_j) over the chunk for every sequence of non-negative powers: checked here in float32 arithmetic, with
random chunks and with the worst case (the bandpass decays for 15 steps, the largest power comes last, sitting
exactly on the bound)."""
import numpy as np

C = 16                                          # VF_K2_C
S = np.float32((12500.0 / 128000000 * 8) / 1.0)  # bp_scale, vf_api.cu (src/process_baseband.cu:737-741)


def clip_floor(s):
    """restatement of vf_k2_clip_floor (host side, double)"""
    oms = np.float32(np.float32(1.0) - s)
    if not (oms > 0 and oms <= 1):
        return np.float32(0)
    f = 11.0 * (float(oms) * (1.0 - 2.0 ** -24)) ** C * (1.0 - 2.0 ** -20)
    ff = np.float32(f)
    if float(ff) > f:
        ff = np.nextafter(ff, np.float32(0))
    return ff


def run_chunk(bp0, p, s):
    """the kernel's recursion x <- fma (x, 1 - s, s p) in float32 (fma emulated in float64: the product and sum of
    float32 values are exact in float64 up to one rounding, which is what fma does); returns whether any step clips"""
    oms = np.float32(np.float32(1.0) - s)
    x = np.float32(bp0)
    clipped = False
    for pj in p:
        lim = np.float32(x * np.float32(11.0))
        clipped |= bool(pj > lim)
        sp = np.float32(s * pj)
        x = np.float32(float(x) * float(oms) + float(sp))
    return clipped


def test_floor_value():
    f = clip_floor(S)
    assert 10.8 < f < 11.0 * (1 - float(S)) ** C


def test_no_clip_below_the_floor_random():
    rng = np.random.default_rng(5)
    f = clip_floor(S)
    for trial in range(3000):
        bp0 = np.float32(10.0 ** rng.uniform(-20, 20))
        lim = np.float32(bp0 * f)                            # the kernel's vf_mul2 (bp, clip_floor)
        p = (rng.random(C) ** rng.integers(1, 6)).astype(np.float32) * lim
        p[rng.integers(0, C)] = lim                          # the largest power sits exactly on the bound
        assert p.max() <= lim
        assert not run_chunk(bp0, p, S), (trial, bp0)


def test_worst_case_decay_then_peak():
    f = clip_floor(S)
    for bp0 in (np.float32(1e-25), np.float32(0.37), np.float32(1.0), np.float32(4097.123), np.float32(3e30)):
        lim = np.float32(bp0 * f)
        p = np.zeros(C, np.float32)
        p[-1] = lim                                          # 15 steps of pure decay, then the peak
        assert not run_chunk(bp0, p, S)
        p[-1] = np.nextafter(np.float32(bp0 * np.float32(11.0)), np.float32(np.inf))
        assert run_chunk(bp0, p, S)                          # and a power above 11 bp0 does clip (sanity of the model)


def test_other_scales():
    for s in (np.float32(1e-5), np.float32(0.01), np.float32(0.2), np.float32(0.9)):
        f = clip_floor(s)
        rng = np.random.default_rng(int(float(s) * 1e6))
        for trial in range(300):
            bp0 = np.float32(10.0 ** rng.uniform(-10, 10))
            lim = np.float32(bp0 * f)
            p = np.zeros(C, np.float32)
            p[rng.integers(0, C):] = lim
            assert not run_chunk(bp0, p, s), (s, trial)
