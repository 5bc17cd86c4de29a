"""CPU oracle stage by stage against independent numpy restatements of the
reference kernels (src/pb_kernels.cu), plus the facts SURVEY.md section 8a pins."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_input, RFI

f32 = np.float32


def fma32(a, b, c):
    ld = np.longdouble
    return (a.astype(ld) * b.astype(ld) + c.astype(ld)).astype(f32)


def np_convert(u):
    # convertarray, src/pb_kernels.cu:23-33
    x = u.astype(f32) / f32(128) - f32(1)
    x[u == 0] = 0
    return x


def np_kurtosis(x):
    # kurtosis, src/pb_kernels.cu:35-107: 256-slot tree, strides 128..1
    v = x.reshape(-1, 500)
    a2 = v[:, :250] * v[:, :250]
    b2 = v[:, 250:] * v[:, 250:]
    d4 = np.zeros((v.shape[0], 256), f32)
    d2 = np.zeros((v.shape[0], 256), f32)
    d4[:, :250] = fma32(b2, b2, a2 * a2)
    d2[:, :250] = a2 + b2
    s = 128
    while s >= 1:
        d2[:, :s] = d2[:, :s] + d2[:, s:2 * s]
        d4[:, :s] = d4[:, :s] + d4[:, s:2 * s]
        s //= 2
    with np.errstate(all="ignore"):
        pw = d2[:, 0] / f32(500)
        kur = d4[:, 0] / f32(500) / (pw * pw)
    return pw, kur


def test_convert(orc):
    u = np.arange(256, dtype=np.uint8)
    x = np.empty(256, f32)
    orc.liba().orc_stage_convert(u.ctypes.data, x.ctypes.data, 256)
    assert np.array_equal(x, np_convert(u))
    assert x[0] == 0 and x[1] == f32(-0.9921875) and x[128] == 0 and x[255] == f32(0.9921875)


def test_kurtosis_tree_order(pkg, orc):
    p0, _ = make_input(pkg, 8, seed=3, **RFI)
    x = np_convert(p0)
    n = x.size // 500
    pw = np.empty(n, f32); kur = np.empty(n, f32)
    orc.liba().orc_stage_kurtosis(x.ctypes.data, pw.ctypes.data, kur.ctypes.data, n)
    rp, rk = np_kurtosis(x)
    assert np.array_equal(pw, rp)
    assert np.array_equal(kur, rk)
    assert 2.5 < np.median(kur) < 3.2


def test_kurtosis_all_zero_block_is_nan_then_excised(orc):
    x = np.zeros(500, f32)
    pw = np.empty(1, f32); kur = np.empty(1, f32)
    orc.liba().orc_stage_kurtosis(x.ctypes.data, pw.ctypes.data, kur.ctypes.data, 1)
    assert pw[0] == 0 and np.isnan(kur[0])
    k2 = np.array([np.nan, 3.0], f32)       # [pol0 | pol1], n = 1
    dag = np.empty(2, f32)
    orc.liba().orc_stage_dagostino(k2.ctypes.data, dag.ctypes.data, 1, 500)
    assert dag[0] == 9.0 and dag[1] == 9.0   # DAG_INF, max over pols, duplicated


def test_dagostino_constants(orc):
    c = (C.c_double * 5)()
    orc.liba().orc_dagostino_constants(500, c)
    # SURVEY.md section 8a row A7
    for got, want in zip(c, (-0.0119760479, 86.4183965, 19.7201112, 0.997428531, 0.721750072)):
        assert abs(got - want) < 1e-7 * max(1, abs(want))
    orc.liba().orc_dagostino_constants(12500, c)
    for got, want in zip(c, (-0.000479961603, 1864.14982, 91.5897056, 0.999880792, 0.748774391)):
        assert abs(got - want) < 1e-7 * max(1, abs(want))


def test_dagostino_pass_band(orc):
    k = np.linspace(1.5, 6.0, 4001).astype(f32)
    kk = np.concatenate([k, np.full_like(k, 3.0)])
    dag = np.empty(kk.size, f32)
    orc.liba().orc_stage_dagostino(kk.ctypes.data, dag.ctypes.data, k.size, 500)
    ok = k[dag[:k.size] <= 3.0]
    assert abs(ok.min() - 2.494) < 5e-3 and abs(ok.max() - 3.843) < 5e-3
    assert np.array_equal(dag[:k.size], dag[k.size:])


def test_weight_values():
    # k-fold float sums of float(500)/12500, SURVEY.md 8a row A10
    inc = f32(500) / f32(12500)
    acc = f32(0); vals = [acc]
    for _ in range(25):
        acc = f32(acc + inc); vals.append(acc)
    assert float(vals[5]) == 0.19999998807907104 and float(vals[5]) < 0.2
    assert float(vals[6]) == 0.23999997973442078
    assert float(vals[25]) == 1.0000001192092896


@pytest.mark.parametrize("nbit", [2, 4, 8])
def test_digitise(orc, nbit):
    rng = np.random.default_rng(nbit)
    ntime, npol = 4, 2
    ave = rng.standard_normal((npol, ntime, 6251)).astype(f32) * f32(2)
    ave[0, 0, 2155:2165] = [-100, 100, -0.6109, 0.3970, 1.4050, 0, -3.77, 3.77, 1e-9, -1e-9]
    out = np.empty(ntime * npol * 4096 * nbit // 8, np.uint8)
    orc.liba().orc_digitise(ave.ctypes.data, out.ctypes.data, ntime, npol, nbit)
    x = ave[:, :, 2155:2155 + 4096].transpose(1, 0, 2).reshape(-1)     # [t][pol][chan]
    if nbit == 8:
        tmp = (x.astype(np.float64) / 0.02957 + 127.5).astype(f32)
        code = np.where(tmp <= 0, 0, np.where(tmp >= 255, 255, np.trunc(tmp))).astype(np.uint8)
        assert np.array_equal(out, code)
    elif nbit == 4:
        tmp = (x.astype(np.float64) / 0.3188 + 7.5).astype(f32)
        code = np.where(tmp <= 0, 0, np.where(tmp >= 15, 15, np.trunc(tmp))).astype(np.uint8)
        assert np.array_equal(out, code[0::2] | (code[1::2] << 4))
    else:
        xd = x.astype(np.float64)
        code = ((xd >= -0.6109).astype(np.uint8) + (xd >= 0.3970) + (xd >= 1.4050)).astype(np.uint8)
        packed = code[0::4] | (code[1::4] << 2) | (code[2::4] << 4) | (code[3::4] << 6)
        assert np.array_equal(out, packed)


def np_chain_raw(pkg, pol0, pol1, T, bp=None):
    """float64-FFT restatement of the raw stream (rfi_mode 0, npol 1): returns fft_ave [T/8][6251]."""
    s = f32((12500 / 128000000 * 8) / 1.0)
    oms = f32(1) - s
    P = []
    for p in (pol0, pol1):
        X = np.fft.rfft(np_convert(p).reshape(T, 12500).astype(np.float64), axis=1)
        P.append((X.real ** 2 + X.imag ** 2).astype(f32))
    out = []
    for Pp in P:
        b = Pp.sum(axis=0, dtype=np.float64).astype(f32) / f32(T) if bp is None else bp
        o = np.empty_like(Pp)
        for t in range(T):
            b = fma32(b, np.full_like(b, oms), s * Pp[t])
            o[t] = Pp[t] / b - f32(1)
        out.append(o)
    ps = (np.float64(0.7071067811865476) * (out[0] + out[1]).astype(np.float64)).astype(f32)
    ave = ps.reshape(T // 8, 8, -1).sum(axis=1, dtype=f32) * f32(np.sqrt(1 / 8))
    return ave


def test_chain_raw_against_numpy(pkg, orc):
    T = 16
    p0, p1 = make_input(pkg, T, seed=11)
    o = orc.OracleChain(T, 8, 1, 0)
    o.process_segment(p0, p1)
    ave = o.get("ave_main").reshape(T // 8, 6251)
    ref = np_chain_raw(pkg, p0, p1, T)
    # the bandpass initial mean is summed in a different order and the FFT is float32: tolerance, not identity
    assert np.abs(ave - ref)[:, 1:6250].max() < 2e-3
    assert np.abs(ave - ref)[:, 1:6250].mean() < 5e-5


def test_chain_modes_consistent(pkg, orc):
    """rfi_mode 1 equals the excised stream of rfi_mode 2; rfi_mode 0 its raw stream."""
    T = 16
    p0, p1 = make_input(pkg, T, seed=5, **RFI)
    o2 = orc.OracleChain(T, 8, 1, 2); m2, r2 = o2.process_segment(p0, p1)
    o1 = orc.OracleChain(T, 8, 1, 1); m1, _ = o1.process_segment(p0, p1)
    o0 = orc.OracleChain(T, 8, 1, 0); m0, _ = o0.process_segment(p0, p1)
    assert np.array_equal(m1, m2) and np.array_equal(m0, r2)
    assert np.count_nonzero(o2.mask()) > 0
    assert not np.array_equal(m2, r2)


def test_chain_clean_input_streams_agree_where_unmasked(pkg, orc):
    T = 8
    p0, p1 = make_input(pkg, T, seed=9)
    o = orc.OracleChain(T, 8, 2, 2)
    o.process_segment(p0, p1)
    mask = o.mask()
    w = o.get("weights")[:T]
    inc = np.cumsum(np.full(25, f32(500) / f32(12500), f32), dtype=f32)
    for t in range(T):
        kept = 25 - bin(int(mask[t])).count("1")
        assert w[t] == (inc[kept - 1] if kept else 0)
    pk = o.power_trimmed("main"); pr = o.power_trimmed("raw")
    for t in range(T):
        if mask[t] == 0:
            assert np.array_equal(pk[t], pr[t])


def test_chain_all_dropped_input(orc):
    """A segment of dropped frames (byte 0): everything excised, weights 0, output = code of 0."""
    T = 8
    z = np.zeros(T * 12500, np.uint8)
    o = orc.OracleChain(T, 2, 1, 1)
    main, _ = o.process_segment(z, z)
    assert np.all(o.mask() == (1 << 25) - 1)
    assert np.all(o.get("weights") == 0)
    assert np.all(o.get("bp_main") == 1)           # src/pb_kernels.cu:454-458
    assert np.all(main == 0b01010101)              # 0.0 -> level 1 in every 2-bit field


def test_bandpass_state_carries(pkg, orc):
    T = 8
    a0, a1 = make_input(pkg, T, seed=2, sample0=0)
    b0, b1 = make_input(pkg, T, seed=2, sample0=T * 12500)
    o = orc.OracleChain(T, 8, 1, 1)
    o.process_segment(a0, a1)
    bp1 = o.get("bp_main")
    second, _ = o.process_segment(b0, b1)
    o.reset_bandpass()
    fresh, _ = o.process_segment(b0, b1)
    assert not np.array_equal(second, fresh)
    o.set_bandpass(0, bp1)
    again, _ = o.process_segment(b0, b1)
    assert np.array_equal(again, second)
