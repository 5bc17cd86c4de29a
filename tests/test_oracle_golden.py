"""Pins the CPU oracle (oracle A) against outputs of the reference's own kernels
(tests/golden/*.npz, made on a B200 by scripts/make_golden.py from
/root/reference/src/pb_kernels.cu compiled unmodified + cuFFT).  The reference
ships no golden vectors of its own."""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT, byte_diff

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))


def load(path):
    z = np.load(path, allow_pickle=False)
    gen = {str(k): int(v) for k, v in zip(z["gen_keys"], z["gen_vals"])}
    return z, gen, int(z["nbit"]), int(z["npol"]), int(z["mode"]), int(z["T"]), int(z["nseg"])


def gen_segment(pkg, gen, T, s):
    g = pkg.GenParams.default(**gen)
    return (pkg.gen_samples(g, 0, 0, s * T * 12500, T * 12500), pkg.gen_samples(g, 0, 1, s * T * 12500, T * 12500))


def canon(a):
    a = a.copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a


def check_rows(got, want, nbit, what):
    worst, frac = byte_diff(got, want, nbit)
    assert worst <= 1, "%s: code differs by %d" % (what, worst)
    assert frac < 1e-4 or frac * got.size * (8 // nbit) <= 3, "%s: %.3g of samples differ" % (what, frac)


def test_fixtures_present():
    assert len(FILES) >= 3


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_matches_reference_kernels(pkg, orc, path):
    z, gen, nbit, npol, mode, T, nseg = load(path)
    rows, arows, steps = z["rows"], z["arows"], z["steps"]
    o = orc.OracleChain(T, nbit, npol, mode)
    for s in range(nseg):
        p0, p1 = gen_segment(pkg, gen, T, s)
        pre = "s%d_" % s
        # the generator is part of the fixture: same bytes as on the box that made it
        assert hashlib.sha256(p0.tobytes() + p1.tobytes()).hexdigest() == str(z[pre + "in_sha256"])
        main, raw = o.process_segment(p0, p1)
        rb = o.out_bytes // 128 // npol
        check_rows(main.reshape(128, npol, rb)[rows], z[pre + "fb_main_rows"], nbit, "main seg %d" % s)
        if mode == 2:
            check_rows(raw.reshape(128, npol, rb)[rows], z[pre + "fb_raw_rows"], nbit, "raw seg %d" % s)
        if mode:
            assert np.array_equal(o.mask(), z[pre + "mask"])
            assert np.array_equal(o.get("weights")[:T], z[pre + "weights"])
        bp = o.get("bp_main").reshape(2, 6251)[:, 2155:2155 + 4096:8]
        np.testing.assert_allclose(bp, z[pre + "bp_main"], rtol=2e-5)
        if s > 0:
            continue
        if mode:
            for k in ("pow", "kur"):        # bit-exact: summation order and FMA placement
                full = canon(o.get(k))
                assert hashlib.sha256(full.tobytes()).hexdigest() == str(z[pre + k + "_sha256"]), k
            for k in ("pow_fb", "kur_fb"):
                assert np.array_equal(o.get(k), z[pre + k], equal_nan=True), k
            # powf: glibc here, libdevice in the reference
            np.testing.assert_allclose(o.get("dag").reshape(2, -1)[:, :500], z[pre + "dag_head"], rtol=1e-5, atol=5e-6)
            np.testing.assert_allclose(o.get("dag_fb"), z[pre + "dag_fb"], rtol=1e-5, atol=2e-5)
        assert np.array_equal(o.get("histo"), z[pre + "histo"])
        for which in (("main", "raw") if mode == 2 else ("main",)):
            det, want = o.power_trimmed(which)[steps], z[pre + "det_%s_steps" % which]
            assert np.abs(det - want).max() / want.mean() < 1e-5, which       # north_star: 1e-5 relative
            ave, want = o.ave_trimmed(which)[:, arows], z[pre + "ave_%s_rows" % which]
            assert np.abs(ave - want).max() < 2e-4, which


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_library_matches_golden(pkg, path):
    z, gen, nbit, npol, mode, T, nseg = load(path)
    rows, arows, steps = z["rows"], z["arows"], z["steps"]
    with pkg.Pipeline(ffts_per_seg=T, nbit=nbit, npol=npol, rfi_mode=mode, keep_stats=1, keep_power=1, do_histo=1) as p:
        for s in range(nseg):
            p0, p1 = gen_segment(pkg, gen, T, s)
            pre = "s%d_" % s
            main, raw = p.process_segment(p0, p1)
            rb = p.out_bytes // 128 // npol
            check_rows(main.reshape(128, npol, rb)[rows], z[pre + "fb_main_rows"], nbit, "main seg %d" % s)
            if mode == 2:
                check_rows(raw.reshape(128, npol, rb)[rows], z[pre + "fb_raw_rows"], nbit, "raw seg %d" % s)
            if mode:
                assert np.array_equal(p.get_mask(), z[pre + "mask"])
                assert np.array_equal(p.get_stats()["weights"][:T], z[pre + "weights"])
            np.testing.assert_allclose(p.get_bandpass()[:, ::8], z[pre + "bp_main"], rtol=2e-5)
            if s > 0:
                continue
            st = p.get_stats()
            if mode:
                for k in ("pow", "kur", "dag"):
                    assert hashlib.sha256(canon(st[k]).tobytes()).hexdigest() == str(z[pre + k + "_sha256"]), k
                for k in ("pow_fb", "kur_fb", "dag_fb"):
                    assert np.array_equal(st[k], z[pre + k], equal_nan=True), k
            assert np.array_equal(st["histo"], z[pre + "histo"])
            for wi, which in enumerate(("main", "raw") if mode == 2 else ("main",)):
                det, want = p.get_detected_power(0, wi)[steps], z[pre + "det_%s_steps" % which]
                assert np.abs(det - want).max() / want.mean() < 1e-5, which
                ave, want = p.get_power_f32(0, wi)[:, arows], z[pre + "ave_%s_rows" % which]
                assert np.abs(ave - want).max() < 2e-4, which
