"""-m gpu: libvlitefast against the reference's OWN kernels (oracle B:
/root/reference/src/pb_kernels.cu compiled unmodified + cuFFT, launch sequence
of src/process_baseband.cu:1108-1375) on the same GPU, at the reference's
compile-time geometry (1024 FFTs per segment), and the CPU oracle against both."""
import os

import numpy as np
import pytest

from conftest import ROOT, make_input, RFI
from test_gpu_parity import check_bytes, REL

pytestmark = pytest.mark.gpu

HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvlite_ref.so"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("nbit,npol,mode", [(2, 1, 2), (8, 1, 2), (8, 2, 1), (4, 1, 0)])
def test_full_segment_vs_reference_kernels(pkg, orc, nbit, npol, mode):
    T = 1024
    r = orc.RefChain(nbit, npol, mode, keep_det=True, do_histo=True)
    p = pkg.Pipeline(ffts_per_seg=T, nbit=nbit, npol=npol, rfi_mode=mode, keep_stats=1, keep_power=1, do_histo=1)
    for s in range(2):       # second segment exercises the carried bandpass
        p0, p1 = make_input(pkg, T, seed=100 + nbit, sample0=s * T * 12500, **RFI)
        main, raw = p.process_segment(p0, p1)
        rmain, rraw = r.process_segment(p0, p1)
        if mode:
            assert np.array_equal(p.get_mask(), r.mask())
            st = p.get_stats()
            for k in ("pow", "kur", "dag", "pow_fb", "kur_fb", "dag_fb", "weights"):
                assert np.array_equal(st[k], r.get(k), equal_nan=True), (k, s)
        assert np.array_equal(p.get_stats()["histo"], r.get("histo"))
        det, rdet = p.get_detected_power(0, 0), r.power_trimmed("main")
        assert np.abs(det - rdet).max() / rdet.mean() < REL
        l2 = np.sqrt(((det - rdet).astype(np.float64) ** 2).sum() / (rdet.astype(np.float64) ** 2).sum())
        assert l2 < 1e-6
        assert np.abs(p.get_power_f32(0, 0) - r.ave_trimmed("main")).max() < 2e-4
        check_bytes(main, rmain, nbit, "main seg %d" % s)
        if mode == 2:
            check_bytes(raw, rraw, nbit, "raw seg %d" % s)
    p.close(); r.close()


@needs_ref
def test_cpu_oracle_vs_reference_kernels(pkg, orc):
    """pins oracle A against the reference itself (same check as tests/golden, live)"""
    T = 1024
    p0, p1 = make_input(pkg, T, seed=5, **RFI)
    r = orc.RefChain(8, 1, 2, keep_det=True, do_histo=True)
    o = orc.OracleChain(T, 8, 1, 2)
    rmain, rraw = r.process_segment(p0, p1)
    omain, oraw = o.process_segment(p0, p1)
    for k in ("pow", "kur", "pow_fb", "kur_fb", "weights", "histo"):
        assert np.array_equal(o.get(k), r.get(k), equal_nan=True), k
    # powf differs by a few ulp between glibc and libdevice
    np.testing.assert_allclose(o.get("dag"), r.get("dag"), rtol=1e-5, atol=5e-6)
    assert np.array_equal(o.mask(), r.mask())
    odet, rdet = o.power_trimmed("main"), r.power_trimmed("main")
    assert np.abs(odet - rdet).max() / rdet.mean() < REL
    check_bytes(omain, rmain, 8, "main")
    check_bytes(oraw, rraw, 8, "raw")
    r.close()


@needs_ref
def test_frb_injection_vs_reference(pkg, orc):
    T = 1024
    p0, p1 = make_input(pkg, T, seed=9)
    r = orc.RefChain(8, 1, 0, keep_det=True, inject_frb=True)
    with pkg.Pipeline(ffts_per_seg=T, nbit=8, npol=1, rfi_mode=0, inject_frb=1) as p:
        plain, _ = p.process_segment(p0, p1)
        p.reset_bandpass()
        # the DM 80 sweep reaches the kept channels (bins >= 2155) ~2900 FFT steps after
        # the FRB second starts: third segment, nfft_since_frb = 2 * 1024
        p.set_frb_injection(2048, 80.0, 20.48, 1.05)
        inj, _ = p.process_segment(p0, p1)
        det = p.get_detected_power(0, 0)
    rinj, _ = r.process_segment(p0, p1, inject_frb_now=3)
    rdet = r.power_trimmed("main")
    assert not np.array_equal(plain, inj)
    assert np.abs(det - rdet).max() / rdet.mean() < REL
    check_bytes(inj, rinj, 8, "injected")
    r.close()
