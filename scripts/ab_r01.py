"""A/B on the GPU box: the current libvlitefast against a saved round-1 build (build/r01/libvlitefast_r01.so, same
C ABI up to the tail of vf_config) on one antenna-second, segment by segment: masks, weights and f32 tiles must be
identical where the arithmetic is meant to be the same (normaliser), detected power within the FFT's tolerance."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
T, NSEG = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 4
nbit, npol, mode = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (2, 1, 2)))

old = C.CDLL(os.path.join(ROOT, "build", "r01", "libvlitefast_r01.so"), mode=C.RTLD_LOCAL)
vp = C.c_void_p
old.vf_create.argtypes = [vp, C.POINTER(vp)]
old.vf_process_segment.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, vp, vp, vp]
old.vf_get_power_f32.argtypes = [vp, C.c_int, C.c_int, vp]
old.vf_get_detected_power.argtypes = [vp, C.c_int, C.c_int, vp]
old.vf_get_mask.argtypes = [vp, C.c_int, vp]
old.vf_segment_out_bytes.argtypes = [vp]; old.vf_segment_out_bytes.restype = C.c_size_t
cfg = (C.c_int * 24)()
old.vf_config_default(cfg)
cfg[3], cfg[7], cfg[8], cfg[9], cfg[12] = T, nbit, npol, mode, 1      # ffts_per_seg, nbit, npol, rfi_mode, keep_power
ho = vp()
assert old.vf_create(cfg, C.byref(ho)) == 0
ob = old.vf_segment_out_bytes(ho)

g = pkg.GenParams.default(seed=102, rfi_amp=60, rfi_burst_every=16)
worst = 0.0
with pkg.Pipeline(ffts_per_seg=T, nbit=nbit, npol=npol, rfi_mode=mode, keep_power=1) as p:
    for s in range(NSEG):
        p0 = pkg.gen_samples(g, 0, 0, s * T * 12500, T * 12500); p1 = pkg.gen_samples(g, 0, 1, s * T * 12500, T * 12500)
        m, r = p.process_segment(p0, p1)
        om, orr = np.empty(ob, np.uint8), np.empty(ob, np.uint8)
        assert old.vf_process_segment(ho, 0, p0.ctypes.data, p1.ctypes.data, p0.size, om.ctypes.data, orr.ctypes.data, None) == 0
        msk = np.empty(T, np.uint32); old.vf_get_mask(ho, 0, msk.ctypes.data)
        assert np.array_equal(p.get_mask(), msk)
        for which in range(2 if mode == 2 else 1):
            a = p.get_power_f32(0, which)
            b = np.empty_like(a); old.vf_get_power_f32(ho, 0, which, b.ctypes.data)
            d = p.get_detected_power(0, which)
            e = np.empty_like(d); old.vf_get_detected_power(ho, 0, which, e.ctypes.data)
            print("seg %d stream %d: det max rel %.2e  ave max abs %.2e  ave identical %s  bytes differ %d" % (
                s, which, np.abs(d - e).max() / e.mean(), np.abs(a - b).max(), np.array_equal(a, b),
                int(((m if which == 0 else r) != (om if which == 0 else orr)).sum())))
            worst = max(worst, float(np.abs(a - b).max()))
print("worst ave abs diff", worst)
