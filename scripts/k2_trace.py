"""TESTING BUILD: timeline of the normaliser's chunk pipeline from clock64 stamps (vf_debug_k2_trace)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import torch
pkg = ge.load_package()
T, NSEG = 1024, 10
g = pkg.GenParams.default(seed=102, rfi_amp=60, rfi_burst_every=16)
host = np.empty((NSEG, 1, 2, T * 12500), np.uint8)
for s in range(NSEG):
    for pol in range(2):
        pkg.gen_samples(g, 0, pol, s * T * 12500, T * 12500, host[s, 0, pol])
d_in = torch.from_numpy(host).cuda()
p = pkg.Pipeline(testing=True, ffts_per_seg=T, nbit=2, npol=1, rfi_mode=2, max_batch_segments=NSEG)
d_main = torch.zeros((NSEG, 1, p.out_bytes), dtype=torch.uint8, device="cuda"); d_raw = torch.zeros_like(d_main)
L = pkg.lib(testing=True)
L.vf_debug_k2_trace.argtypes = [C.c_void_p, C.c_void_p]
for _ in range(3):
    p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
p.sync()
assert L.vf_debug_k2_trace(p.h, None) == 0
p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
p.sync()
tr = np.zeros((2, 4096, 6), np.int64)
assert L.vf_debug_k2_trace(p.h, tr.ctypes.data) == 0
n = NSEG * T // 16
for sid, name in ((0, "excised"), (1, "raw")):
    t = tr[sid, :n].astype(np.float64)
    t -= t[0, 0]
    step = np.diff(t[:, 4])                      # hand-over to hand-over
    print("%s: chunks %d  total %.0f cycles  chain step median %.0f  mean %.0f" % (name, n, t[-1, 5], np.median(step), step.mean()))
    print("   token wait (1->2) med %.0f | A (2->3) med %.0f | hand-over (3->4) med %.0f | B etc (4->5) med %.0f | top+prep (0->1) med %.0f" % (
        np.median(t[:, 2] - t[:, 1]), np.median(t[:, 3] - t[:, 2]), np.median(t[:, 4] - t[:, 3]), np.median(t[:, 5] - t[:, 4]), np.median(t[:, 1] - t[:, 0])))
    lat = t[1:, 2] - t[:-1, 4]                   # published by g-1 -> received by g
    print("   publish(g-1) -> received(g) med %.0f  p90 %.0f ; ready(g) before publish(g-1): %.0f%% of chunks" % (
        np.median(lat), np.percentile(lat, 90), 100 * np.mean(t[1:, 1] < t[:-1, 4])))
    for q in range(40, 56):
        print("   g=%d top %.0f ready %.0f recv %.0f Aend %.0f pub %.0f end %.0f" % ((q,) + tuple(t[q])))
