"""time the two kernels of a given libvlitefast build (path in argv[1]) on one antenna-second, launches serialised"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
import torch
pkg = ge.load_package()
from vlite_fast_b200 import binding
if len(sys.argv) > 1:
    L = C.CDLL(sys.argv[1], mode=C.RTLD_LOCAL)
    ref = binding.lib()
    for n in dir(ref):
        pass
    # reuse argtypes by declaring the few calls used here
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.vf_config_default.argtypes = [C.POINTER(binding.VfConfig)]
    L.vf_create.argtypes = [C.POINTER(binding.VfConfig), C.POINTER(vp)]
    L.vf_destroy.argtypes = [vp]; L.vf_last_error.argtypes = [vp]; L.vf_last_error.restype = C.c_char_p
    L.vf_strerror.argtypes = [i]; L.vf_strerror.restype = C.c_char_p
    L.vf_segment_out_bytes.argtypes = [vp]; L.vf_segment_out_bytes.restype = sz
    L.vf_segment_in_samples.argtypes = [vp]; L.vf_segment_in_samples.restype = sz
    L.vf_process_device.argtypes = [vp, i, i, vp, vp, vp]; L.vf_sync.argtypes = [vp]; L.vf_set_serial.argtypes = [vp, i]
    fp = C.POINTER(C.c_float); L.vf_last_elapsed_ms.argtypes = [vp, fp, fp, fp]
    binding._lib[False] = L
T, NSEG = 1024, 10
g = pkg.GenParams.default(seed=102, rfi_amp=60, rfi_burst_every=16)
host = np.empty((NSEG, 1, 2, T * 12500), np.uint8)
for s in range(NSEG):
    for pol in range(2):
        pkg.gen_samples(g, 0, pol, s * T * 12500, T * 12500, host[s, 0, pol])
d_in = torch.from_numpy(host).cuda()
p = pkg.Pipeline(ffts_per_seg=T, nbit=2, npol=1, rfi_mode=2, max_batch_segments=NSEG)
d_main = torch.zeros((NSEG, 1, p.out_bytes), dtype=torch.uint8, device="cuda"); d_raw = torch.zeros_like(d_main)
p.set_serial(1)
for _ in range(3):
    p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr()); p.sync()
k1 = k2 = 0
for _ in range(5):
    p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr()); p.sync()
    _, a, b = p.last_elapsed_ms(); k1 += a; k2 += b
print("%s: K1 %.1f us/segment  K2 %.1f us/segment" % (sys.argv[1] if len(sys.argv) > 1 else "product", k1 / 5 / NSEG * 1e3, k2 / 5 / NSEG * 1e3))
