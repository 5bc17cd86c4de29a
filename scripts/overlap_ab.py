import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import __graft_entry__ as ge
pkg = ge.load_package()
T, NSEG = 1024, 10
g = pkg.GenParams.default(seed=102, rfi_amp=60, rfi_burst_every=16)
host = np.empty((NSEG, 1, 2, T * 12500), np.uint8)
for s in range(NSEG):
    for pol in range(2):
        pkg.gen_samples(g, 0, pol, s * T * 12500, T * 12500, host[s, 0, pol])
d_in = torch.from_numpy(host).cuda()
p = pkg.Pipeline(ffts_per_seg=T, nbit=2, npol=1, rfi_mode=2, max_batch_segments=NSEG)
d_main = torch.zeros((NSEG, 1, p.out_bytes), dtype=torch.uint8, device="cuda"); d_raw = torch.zeros_like(d_main)
for serial in (1, 0, 1, 0):
    p.set_serial(serial)
    for _ in range(5):
        p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
    p.sync()
    p.timer_begin()
    for _ in range(64):
        p.process_device(1, NSEG, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr())
    ms = p.timer_end()
    print("serial", serial, "ms per second %.4f" % (ms / 64))
