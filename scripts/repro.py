"""small driver for compute-sanitizer / ncu runs: one pipeline, a few segments"""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
import numpy as np
ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=16); ap.add_argument("--mode", type=int, default=2)
ap.add_argument("--nbit", type=int, default=2); ap.add_argument("--npol", type=int, default=1)
ap.add_argument("--histo", type=int, default=0); ap.add_argument("--stats", type=int, default=0)
ap.add_argument("--nseg", type=int, default=1); ap.add_argument("--ant", type=int, default=1)
ap.add_argument("--threads", type=int, default=0); ap.add_argument("--power", type=int, default=1)
a = ap.parse_args()
pkg = ge.load_package()
g = pkg.GenParams.default(seed=3, rfi_amp=60, rfi_burst_every=3)
with pkg.Pipeline(ffts_per_seg=a.T, nbit=a.nbit, npol=a.npol, rfi_mode=a.mode, do_histo=a.histo, keep_stats=a.stats,
                  keep_power=a.power, n_antennas=a.ant, k1_threads=a.threads) as p:
    for s in range(a.nseg):
        p0 = [pkg.gen_samples(g, i, 0, s * a.T * 12500, a.T * 12500) for i in range(a.ant)]
        p1 = [pkg.gen_samples(g, i, 1, s * a.T * 12500, a.T * 12500) for i in range(a.ant)]
        m, r = p.process_batch(p0, p1)
    print("ok", m[0][:8], p.last_elapsed_ms())
