"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the reference's
OWN kernels (oracle/_ref: /root/reference/src/pb_kernels.cu compiled unmodified
+ cuFFT, launch sequence of src/process_baseband.cu:1108-1375) on a B200:

    gpurun -- python scripts/make_golden.py gpurun_out/golden

and then copying gpurun_out/golden/*.npz into tests/golden/.  The input is the
deterministic generator (vlite-fast_b200/host/vf_genbase.c), so the fixtures
hold only generator parameters and (subsampled) reference outputs."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

CASES = {
    # name: (nbit, npol, rfi_mode, generator keyword arguments)
    "m2_b8_p1_rfi": (8, 1, 2, dict(seed=102, rfi_amp=60, rfi_burst_every=16)),
    "m1_b2_p2_rfi": (2, 2, 1, dict(seed=7, rfi_amp=40, rfi_burst_every=4, drop_period=997, drop_len=3, drop_pol_skew=1)),
    "m0_b4_p1_clean": (4, 1, 0, dict(seed=33)),
}
ROWS = list(range(0, 128, 16)) + [127]        # scrunched time rows kept (packed bytes)
AROWS = [0, 127]                              # scrunched time rows kept (f32 tile)
STEPS = [511]                                 # FFT time steps kept (detected power)
NSEG = 2


def canon(a):
    """one NaN bit pattern (0/0 is 0x7fffffff on the GPU, 0xffc00000 on x86)"""
    a = a.copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    pkg = ge.load_package()
    orc = ge.load_oracle()
    T = 1024
    for name, (nbit, npol, mode, gen) in CASES.items():
        r = orc.RefChain(nbit, npol, mode, keep_det=True, do_histo=True)
        assert r.T == T
        g = pkg.GenParams.default(**gen)
        out = {"nbit": nbit, "npol": npol, "mode": mode, "T": T, "nseg": NSEG,
               "gen_keys": np.array(sorted(gen)), "gen_vals": np.array([gen[k] for k in sorted(gen)], np.int64),
               "rows": np.array(ROWS), "arows": np.array(AROWS), "steps": np.array(STEPS)}
        for s in range(NSEG):
            p0 = pkg.gen_samples(g, 0, 0, s * T * 12500, T * 12500)
            p1 = pkg.gen_samples(g, 0, 1, s * T * 12500, T * 12500)
            main_b, raw_b = r.process_segment(p0, p1)
            pre = "s%d_" % s
            out[pre + "in_sha256"] = hashlib.sha256(p0.tobytes() + p1.tobytes()).hexdigest()
            rb = r.out_bytes // 128 // npol            # bytes per (time, pol) row
            mb = main_b.reshape(128, npol, rb)
            out[pre + "fb_main_rows"] = mb[ROWS].copy()
            out[pre + "fb_main_sha256"] = hashlib.sha256(main_b.tobytes()).hexdigest()
            if mode == 2:
                out[pre + "fb_raw_rows"] = raw_b.reshape(128, npol, rb)[ROWS].copy()
            if mode:
                out[pre + "mask"] = r.mask()
                out[pre + "weights"] = r.get("weights")[:T].copy()
            out[pre + "bp_main"] = r.get("bp_main").reshape(2, 6251)[:, 2155:2155 + 4096:8].copy()
            if s > 0:
                continue            # later segments: bytes, mask and bandpass pin the carried state
            if mode:
                for k in ("pow_fb", "kur_fb", "dag_fb"):
                    out[pre + k] = r.get(k)
                for k in ("pow", "kur", "dag"):
                    full = canon(r.get(k))
                    out[pre + k + "_head"] = full.reshape(2, -1)[:, :500].copy()      # first 20 FFT blocks per pol
                    out[pre + k + "_sha256"] = hashlib.sha256(full.tobytes()).hexdigest()
            out[pre + "histo"] = r.get("histo")
            out[pre + "ave_main_rows"] = r.ave_trimmed("main")[:, AROWS].copy()
            out[pre + "det_main_steps"] = r.power_trimmed("main")[STEPS].copy()
            if mode == 2:
                out[pre + "ave_raw_rows"] = r.ave_trimmed("raw")[:, AROWS].copy()
                out[pre + "det_raw_steps"] = r.power_trimmed("raw")[STEPS].copy()
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        r.close()
        print("wrote", name, os.path.getsize(os.path.join(outdir, name + ".npz")), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
