"""TEST INFRASTRUCTURE -- numpy restatement of the GPU baseband generator (libvlitegen,
vlite-fast_b200/csrc/vf_genbase_gpu.cu), which follows src/genbase.cu of the reference: noise + pulse
profile (:375-384, :554-585), overlap-save convolution with the dispersion kernel (:391-398, :525-552,
:587-598), side-band swap (:651-661), RFI (:671-687), digitisation (:690-708).  Only tests/ may use it.

The reference draws its noise from cuRAND (not reproducible off the GPU), so there is no reference output
to pin this to: parity of the generator is "unpinned" by the reference and pinned between the two
implementations here; the Philox-4x32-10 core is checked against the published known-answer vectors."""
import numpy as np

RATE = 128000000
M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """arrays (or scalars) of uint32 counters, scalar keys -> four uint32 arrays"""
    c0, c1, c2, c3 = (np.asarray(c, np.uint32).astype(np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)) & mask
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def sweep_samples(dm):
    """(n_lo, n_hi) after the swap of src/genbase.cu:188-197, with the reference's truncations"""
    tsamp = 1.0 / RATE
    t_lo = dm / 2.41e-10 * (1. / (320. * 320.) - 1. / (352. * 352.))
    t_hi = dm / 2.41e-10 * (1. / (352. * 352.) - 1. / (384. * 384.))
    n_lo = int(int(t_lo) * 1e-6 / tsamp)
    n_hi = int(int(t_hi) * 1e-6 / tsamp)
    n_lo += n_lo & 1
    n_hi += n_hi & 1
    return n_hi, n_lo


def noise(base, n, pol, seed, period, skip_period, ampl):
    """input samples base .. base+n-1 of pol: N(0,1) by Box-Muller on Philox, pulse profile on top"""
    g0, g1 = base >> 2, (base + n - 1) >> 2
    g = np.arange(g0, g1 + 1, dtype=np.int64)
    r = philox4x32_10((g & 0xFFFFFFFF).astype(np.uint32), (g >> 32).astype(np.uint32), np.uint32(pol), np.uint32(0),
                      seed & 0xFFFFFFFF, seed >> 32)
    z = np.empty((g.size, 4), np.float32)
    for k in range(2):
        rad = np.sqrt(np.float32(-2.0) * np.log(u01(r[2 * k])))
        th = np.float32(6.283185307179586) * u01(r[2 * k + 1])
        z[:, 2 * k] = rad * np.cos(th)
        z[:, 2 * k + 1] = rad * np.sin(th)
    z = z.reshape(-1)[base - 4 * g0: base - 4 * g0 + n]
    sample = base + np.arange(n, dtype=np.int64)
    phasei = sample // period
    phasef = (sample - phasei * period).astype(np.float32) / np.float32(period)
    on = (phasef < np.float32(0.03)) & (phasei % skip_period == 0)
    return np.where(on, z * np.float32(ampl), z).astype(np.float32)


def dm_kernel(dm, n):
    i = np.arange(n, dtype=np.float64)
    freq = 64. * i / n
    arg = (2 * np.pi * dm / 2.41e-10) * freq * freq / (320. * 320. * (320. + freq))
    k = (np.cos(arg) + 1j * np.sin(arg)) / (2 * (n - 1))
    f = freq / 64.
    scale = 1 - np.exp(-(f * f) / (0.05 * 0.05))
    scale -= np.exp(-((1 - f) * (1 - f)) / (0.10 * 0.10))
    scale *= (1 + 0.20 * f)
    return (k * scale).astype(np.complex64)


def block(blk, pol, dm=30.0, pulse_period=0.5, ampl=0.05, skip_period=1, add_rfi=0, seed=42, buflen=RATE // 4):
    """(uint8 samples, float32 voltages) of output block blk of pol"""
    n_lo, n_hi = sweep_samples(dm)
    n_dm = n_lo + n_hi
    new = buflen - n_dm
    period = int(pulse_period * RATE)
    base = blk * new
    x = noise(base, buflen, pol, seed, period, skip_period, 1.0 + ampl)
    y = np.fft.irfft(np.fft.rfft(x.astype(np.float64)) * dm_kernel(dm, buflen // 2 + 1).astype(np.complex128), buflen) * buflen
    i = n_lo + np.arange(new, dtype=np.int64)
    v = y[n_lo:n_lo + new].astype(np.float32)
    v = np.where(i & 1, -v, v)
    if add_rfi:
        o = base + np.arange(new, dtype=np.int64)
        phase = np.fmod((o.astype(np.float64) * ((1e6 / RATE) / 11.3)).astype(np.float32), np.float32(1.0))
        g = o >> 2
        r = philox4x32_10((g & 0xFFFFFFFF).astype(np.uint32), (g >> 32).astype(np.uint32), np.uint32(pol), np.uint32(1),
                          seed & 0xFFFFFFFF, seed >> 32)
        pick = np.choose(o & 3, r)
        v = np.where(phase < np.float32(0.1), v + np.float32(5.0) * (u01(pick) - np.float32(0.5)), v).astype(np.float32)
    tmp = v / np.float32(0.02957) / np.float32(2) + np.float32(128.5)
    u = np.where(tmp <= 0, 0, np.where(tmp >= 255, 255, np.minimum(tmp, 255).astype(np.uint8))).astype(np.uint8)
    return u, v
