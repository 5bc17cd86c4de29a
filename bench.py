#!/usr/bin/env python
"""bench.py -- baseband antenna-seconds per second (x real time) per B200.

Workload (BASELINE.json configs[1], the same string in both arms): per GPU, A antennas (default 1) of synthetic
8-bit two-polarisation baseband at 128 MS/s; one antenna-second = 10 segments of 1024 FFTs of 12500 samples,
kurtosis excision on, reference defaults nbit 2 / npol 1 / rfi_mode 2.  A "step" is S (default 16) consecutive
antenna-seconds per antenna through the chain; every rank runs the SAME code at every N: per antenna-second one
vf_process_device call (one launch pair over the 10 segments, f32 tiles kept) and one vf_coadd_batch (local sum +
count, NCCL reduce over the ranks when N > 1, scale and digitise on the root).

  value     device-resident: inputs already in HBM; CUDA events on the library's streams (vf_timer_begin/end,
            every stream of the handle including the co-add) around the K timed steps, max over ranks.
  e2e       same metric through the host-buffer C ABI (vf_submit_block_async / vf_wait, pinned host memory allocated
            on the GPU's NUMA node, double-buffered H2D and D2H inside the timed region, co-add included), wall
            clock between barriers, max over ranks.
  roofline  dominant kernel (the channeliser): algorithmic bytes per launch / its mean launch duration (CUDA
            events around each launch, launches serialised) against MEASURED_PEAKS.json.  The kernel is bound by
            fp32 issue, not HBM: the same duration is also set against the fp32 peak.
  cpu_baseline  the CPU oracle (oracle/liboracle.so, OpenMP, all host cores, thread count stated).
  coadd_check   before and after the timed region, outside it: the root's co-added bytes and f32 sum of two
            segments against the sum of the CPU oracle's tiles of ALL ranks' antennas; a mismatch exits non-zero.

--impl reference times the CPU oracle as the whole arm (the reference has no CPU path).
--antennas-total M shards M antennas over the ranks (BASELINE.json configs[4]: 16 over 2/4/8 GPUs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

# the co-add's reduce must fit into the SMs the normaliser leaves free (vf_coadd_init, vf_api.cu): NCCL reads this when the
# process creates its first communicator, which is torch.distributed's below
os.environ.setdefault("NCCL_MAX_NCHANNELS", "8")

SEG_PER_SEC = 10
T = 1024
NSAMP = T * 12500
METRIC = "baseband antenna-seconds/s per B200 (x real-time)"
UNIT = "antenna-seconds/s"
GEN = dict(seed=102, rfi_amp=60, rfi_burst_every=16)      # scripts/baseband_test:21 uses -r 102
WORKLOAD = ("configs[1]: 1 antenna-second per antenna = 10 segments x 1024 FFTs x 12500 samples x 2 pols of "
            "genbase-style synthetic baseband, default channelisation, kurtosis excision on")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, t_lo=None, t_hi=None):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for (t, r) in self.rows if t_lo is None or (t_lo <= t <= t_hi)] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_second(pkg, antenna, out):
    """one antenna-second, layout [seg][pol][NSAMP]"""
    g = pkg.GenParams.default(**GEN)
    for s in range(SEG_PER_SEC):
        for pol in range(2):
            pkg.gen_samples(g, antenna, pol, s * NSAMP, NSAMP, out[s, pol])


def cpu_oracle_run(nseg_sample, pkg, nthreads):
    """oracle A over nseg_sample segments of antenna 0; returns (antenna-seconds/s, threads, seconds)"""
    orc = ge.load_oracle()
    o = orc.OracleChain(T, 2, 1, 2, nthreads)
    g = pkg.GenParams.default(**GEN)
    segs = []
    for s in range(min(nseg_sample, SEG_PER_SEC)):
        segs.append((pkg.gen_samples(g, 0, 0, s * NSAMP, NSAMP), pkg.gen_samples(g, 0, 1, s * NSAMP, NSAMP)))
    o.process_segment(*segs[0])            # warm-up (page faults, bandpass init)
    t0 = time.perf_counter()
    for s in range(nseg_sample):
        o.process_segment(*segs[s % len(segs)])
    dt = time.perf_counter() - t0
    return (nseg_sample / SEG_PER_SEC) / dt, nthreads, dt


def config_block(args, n_ant, world):
    """identical in both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "antennas_per_gpu": n_ant, "antennas_total": n_ant * world, "nbit": args.nbit,
            "npol": args.npol, "rfi_mode": args.rfi_mode, "antenna_seconds_per_step_per_antenna": args.seconds_per_step,
            "generator": GEN,
            "l2": "inputs (%.0f MB per antenna-second per GPU) exceed the 126 MB L2; no flush" % (n_ant * 2 * NSAMP * SEG_PER_SEC / 1e6)}


def run_reference(args, rank, world):
    """the reference arm: the CPU restatement of the chain (the reference is CUDA-only), every host thread this
    process may use (torchrun exports OMP_NUM_THREADS=1: the count is passed explicitly), rank 0 only"""
    if rank != 0:
        return
    pkg = ge.load_package()
    nthreads = host_threads()
    per_step = 5                            # segments per step: a bounded sample (0.5 antenna-second)
    for _ in range(args.warmup):
        cpu_oracle_run(1, pkg, nthreads)
    vals = []
    for _ in range(args.steps):
        v, _, dt = cpu_oracle_run(per_step, pkg, nthreads)
        vals.append((per_step / SEG_PER_SEC, dt))
    tot_as = sum(v[0] for v in vals); tot_t = sum(v[1] for v in vals)
    value = tot_as / tot_t
    n_ant = args.antennas
    sample = "%d x %d segments of 100 ms of antenna 0 (rfi_mode %d, nbit %d), CPU oracle with %d OpenMP threads" % (
        args.steps, per_step, args.rfi_mode, args.nbit, nthreads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args, n_ant, world),
        "notes": {"arm": "the reference has no CPU implementation; this is the C restatement (oracle/) of its chain; "
                         "each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--antennas", type=int, default=1, help="antennas batched per GPU (configs[3] uses 8)")
    ap.add_argument("--antennas-total", type=int, default=0, help="antennas sharded over all ranks (configs[4] uses 16)")
    ap.add_argument("--seconds-per-step", type=int, default=16, help="antenna-seconds per antenna per step")
    ap.add_argument("--nbit", type=int, default=2)
    ap.add_argument("--npol", type=int, default=1)
    ap.add_argument("--rfi-mode", type=int, default=2)
    ap.add_argument("--max-batch", type=int, default=SEG_PER_SEC,
                    help="segments per launch pair (default: the 10 of an antenna-second; 1 = per segment)")
    ap.add_argument("--clean", action="store_true", help="no impulsive RFI in the synthetic input (no time step needs the second FFT)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legacy", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle check of the co-added output")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.clean:
        GEN["rfi_amp"] = 0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.antennas_total:
        if args.antennas_total % world:
            raise SystemExit("--antennas-total must be a multiple of the number of ranks")
        args.antennas = args.antennas_total // world
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libvlitefast has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = ge.load_package()
    n_ant, S, K = args.antennas, args.seconds_per_step, args.steps
    nthreads_all = host_threads()
    # pinned buffers are allocated from here on: sit on the GPU's NUMA node first
    cpulist = pkg.bind_thread_to_gpu(local)
    p = pkg.Pipeline(ffts_per_seg=T, nbit=args.nbit, npol=args.npol, rfi_mode=args.rfi_mode, gpu_id=local,
                     n_antennas=n_ant, keep_power=1, max_batch_segments=args.max_batch, power_segments=4 * SEG_PER_SEC)
    out_bytes = p.out_bytes
    nstream = 2 if args.rfi_mode == 2 else 1
    coadd_bytes = args.npol * (T // 8) * 4096 * args.nbit // 8       # per segment

    # ---- inputs: one antenna-second per antenna, pinned on the host and resident on the device
    host = torch.empty((SEG_PER_SEC, n_ant, 2, NSAMP), dtype=torch.uint8, pin_memory=True)
    hn = host.numpy()
    tmp = np.empty((SEG_PER_SEC, 2, NSAMP), np.uint8)
    ants = [rank + world * a for a in range(n_ant)]                   # antenna a lives on rank a % world
    for a in range(n_ant):
        make_second(pkg, ants[a], tmp)
        hn[:, a] = tmp
    d_in = host.cuda(non_blocking=False)
    d_main = torch.zeros((SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, device="cuda")
    d_raw = torch.zeros_like(d_main) if args.rfi_mode == 2 else None
    h_main = torch.empty((SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, pin_memory=True)
    h_raw = torch.empty((SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, pin_memory=True) if args.rfi_mode == 2 else None
    h_coadd = torch.empty((SEG_PER_SEC, coadd_bytes), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    uid = [p.coadd_unique_id() if (rank == 0 and world > 1) else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    p.coadd_init(world, rank, uid[0])
    total_ant = world * n_ant

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def second_device():
        """one antenna-second per antenna of this rank, device-resident, and its co-add"""
        p.process_device(n_ant, SEG_PER_SEC, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr() if d_raw is not None else None)
        p.coadd_batch(0, total_ant, SEG_PER_SEC, want=False, wait=False)

    # ---- correctness of what is timed: co-added output of two segments against the oracle's tiles ----------
    def coadd_check():
        orc = ge.load_oracle()
        nchk = 2
        p.sync()
        p.reset_bandpass()
        p.process_device(n_ant, nchk, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr() if d_raw is not None else None)
        fb, sm = p.coadd_batch(0, total_ant, nchk, want=(rank == 0))
        nthr = max(1, nthreads_all // world)
        osum = np.zeros((nchk, args.npol, T // 8, 4096), np.float32)
        ocnt = np.zeros((nchk, T // 8), np.float32)
        for a in range(n_ant):
            o = orc.OracleChain(T, args.nbit, args.npol, args.rfi_mode, nthr)
            for s in range(nchk):
                o.process_segment(hn[s, a, 0], hn[s, a, 1])
                osum[s] += o.ave_trimmed("main")
                if args.rfi_mode:
                    w = o.get("weights")[:T].reshape(T // 8, 8)
                    keep = (w != 0) & (w.astype(np.float64) >= 0.2)
                    ws = np.zeros(T // 8, np.float32)
                    for j in range(8):                      # sequential float adds, src/pb_kernels.cu:616-619
                        ws = np.where(keep[:, j], (ws + w[:, j]).astype(np.float32), ws)
                    ocnt[s] += ((ws / np.float32(8)).astype(np.float64) >= 0.2)
                else:
                    ocnt[s] += 1
            o.close()
        if world > 1:
            ts, tc = torch.from_numpy(osum).cuda(), torch.from_numpy(ocnt).cuda()
            dist.reduce(ts, dst=0, op=dist.ReduceOp.SUM); dist.reduce(tc, dst=0, op=dist.ReduceOp.SUM)
            osum, ocnt = ts.cpu().numpy(), tc.cpu().numpy()
        res = None
        if rank == 0:
            max_abs = float(np.abs(sm - osum).max())
            with np.errstate(divide="ignore", invalid="ignore"):
                x = np.where(ocnt[:, None, :, None] > 0, osum / np.sqrt(ocnt)[:, None, :, None], 0).astype(np.float32)
            ndiff = worst = nsamp = 0
            for s in range(nchk):
                full = np.zeros((args.npol, T // 8, 6251), np.float32)
                full[:, :, 2155:2155 + 4096] = x[s]
                want = np.empty(coadd_bytes, np.uint8)
                orc.liba().orc_digitise(full.ctypes.data, want.ctypes.data, T // 8, args.npol, args.nbit)
                per, m = 8 // args.nbit, (1 << args.nbit) - 1
                for j in range(per):
                    d = np.abs(((fb[s] >> (args.nbit * j)) & m).astype(int) - ((want >> (args.nbit * j)) & m).astype(int))
                    worst = max(worst, int(d.max())); ndiff += int(np.count_nonzero(d)); nsamp += d.size
            res = {"segments": nchk, "antennas": total_ant, "max_abs": max_abs, "byte_frac": ndiff / nsamp, "byte_max": worst,
                   "ok": bool(max_abs < 5e-4 * max(1.0, np.sqrt(total_ant)) and worst <= 1 and ndiff / nsamp < 1e-4)}
        p.reset_bandpass()
        return res

    check = None
    if not args.no_check:
        check = coadd_check()
        if rank == 0 and not check["ok"]:
            print(json.dumps({"error": "coadd_check failed", "coadd_check": check}), flush=True)
        if world > 1:
            flag = torch.tensor([0 if (rank != 0 or check["ok"]) else 1], device="cuda")
            dist.broadcast(flag, src=0)
            if int(flag.item()):
                sys.exit(3)
        elif not check["ok"]:
            sys.exit(3)

    # ---- device-resident timing -------------------------------------------------
    for _ in range(args.warmup):
        for _ in range(S):
            second_device()
    p.sync()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.4)
    t_lo = time.perf_counter()
    p.timer_begin()
    for _ in range(K * S):
        second_device()
    dev_ms = p.timer_end()
    wall = time.perf_counter() - t_lo
    t_hi = time.perf_counter()
    barrier()
    time.sleep(0.1)
    clocks = sampler.finish(t_lo, t_hi)
    if world > 1:
        tt = torch.tensor([dev_ms, wall * 1e3], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, wall = float(tt[0].item()), float(tt[1].item()) / 1e3
    ant_seconds = K * S * n_ant * world
    value = ant_seconds / (dev_ms / 1e3)

    # ---- pure kernel durations for the roofline: same calls, launches not overlapped -------------
    k1_ms = k2_ms = 0.0
    p.set_serial(1)
    second_device(); p.sync()
    nser = 5
    for _ in range(nser):
        second_device(); p.sync()
        _, b, c = p.last_elapsed_ms()
        k1_ms += b; k2_ms += c
    p.set_serial(0)
    seg_per_launch = min(SEG_PER_SEC, args.max_batch if args.max_batch > 0 else 16)
    launches_per_second = -(-SEG_PER_SEC // seg_per_launch)
    k1_ms /= nser * launches_per_second; k2_ms /= nser * launches_per_second       # per launch
    second_device(); p.sync()

    # ---- end to end through host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        nblk = [0]

        def second_e2e():
            """one antenna-second per antenna from pinned host memory: vf_submit_block_async takes the whole second
            of all the rank's antennas in one copy and one launch pair; two seconds in flight (slots)"""
            slot = nblk[0] & 1
            if nblk[0] >= 2:
                p.wait(slot)
                # the co-add of the second that has just completed on this slot ... is issued right after its submit
            p.submit_block_async(slot, n_ant, SEG_PER_SEC, hn, h_main.numpy(), h_raw.numpy() if h_raw is not None else None)
            p.coadd_batch_into(0, total_ant, SEG_PER_SEC, h_coadd.numpy() if rank == 0 else None, wait=False)
            nblk[0] += 1

        def drain_e2e():
            for sl in (0, 1):
                if nblk[0] > sl:
                    p.wait((nblk[0] - 1 - sl) & 1)
            nblk[0] = 0
        for _ in range(max(2, args.warmup * S // 8)):
            second_e2e()
        drain_e2e()
        p.sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K * S):
            second_e2e()
        drain_e2e()
        p.sync()
        barrier()
        e_t = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([e_t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_t = float(tt.item())
        e2e = {"value": ant_seconds / e_t, "unit": UNIT,
               "h2d_bytes_per_step": int(S * n_ant * 2 * NSAMP * SEG_PER_SEC),
               "d2h_bytes_per_step": int(S * (n_ant * out_bytes * nstream + (coadd_bytes if rank == 0 else 0)) * SEG_PER_SEC),
               "api": "vf_submit_block_async (one antenna-second of every antenna of the rank per call) / vf_wait + vf_coadd_batch, pinned host buffers, 2 slots",
               "h2d_gb_per_s_per_gpu": S * K * n_ant * 2 * NSAMP * SEG_PER_SEC / e_t / 1e9,
               "numa_cpulist": cpulist}

    # ---- the same check after the timed regions: the state the timing left behind computes the same thing ------
    check_after = None
    if not args.no_check:
        check_after = coadd_check()
        if rank == 0 and not (check_after["ok"] and check_after["max_abs"] == check["max_abs"] and check_after["byte_frac"] == check["byte_frac"]):
            print(json.dumps({"error": "coadd_check after timing differs or failed", "before": check, "after": check_after}), flush=True)
            check_after["ok"] = False

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------
    peak, peak_src = peaks()
    alg_bytes_per_seg = n_ant * (2 * NSAMP + out_bytes * nstream)
    alg_bytes_per_launch = alg_bytes_per_seg * seg_per_launch
    roofline = None
    if k1_ms > 0:
        k1_avg_s = k1_ms / 1e3                      # mean duration of one launch
        achieved = alg_bytes_per_launch / k1_avg_s / 1e9
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch_1ant")
                if traffic is not None:     # the captured launch covered tj["segments_per_launch"] segments of 1 antenna
                    traffic = traffic * n_ant * seg_per_launch / tj.get("segments_per_launch", 1)
                    traffic_src = "static: ncu --set full capture of one launch (%s), scaled to this launch's segments; not measured in this run" % tj.get("source", "profiles/k1_traffic.json")
        except Exception:
            pass
        # secondary view: the reference's arithmetic for this launch (real FFTs of both streams,
        # 2.5 N log2 N each) against the fp32 peak of the SMs at the clock measured above
        alg_flops = seg_per_launch * n_ant * nstream * 2 * T * 2.5 * 12500 * np.log2(12500)
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        fp32 = {"alg_flops_per_launch": alg_flops, "achieved_tflops": alg_flops / k1_avg_s / 1e12, "peak_tflops": fp32_peak,
                "frac": alg_flops / k1_avg_s / 1e12 / fp32_peak,
                "peak_source": "148 SMs x 128 fp32 lanes x 2 (FMA) x SM clock; the packed FFMA2 the kernel uses reaches "
                               "116-128 lanes/clk/SM in scripts/ubench/fp32_rate.cu (profiles/r01_fp32_rate_ubench.txt)"}
        roofline = {"bound": "hbm", "kernel": "vf_k1_pipelined", "fp32": fp32, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg_bytes_per_launch, "segments_per_launch": seg_per_launch,
                    "k1_ms_per_launch": k1_ms, "k2_ms_per_launch": k2_ms,
                    "k1_ms_per_segment": k1_ms / seg_per_launch, "k2_ms_per_segment": k2_ms / seg_per_launch,
                    "launch_timing": "CUDA events around each launch on the library's stream, launches serialised (vf_set_serial) so that no queueing is included",
                    "whole_chain_achieved": alg_bytes_per_seg * SEG_PER_SEC * K * S / (dev_ms / 1e3) / 1e9}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, dt = cpu_oracle_run(20, pkg, nthreads_all)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "20 segments of 100 ms (2 antenna-seconds) of the same workload, %.1f s, %d OpenMP threads" % (dt, cores)}

    legacy = None
    if not args.no_legacy and world == 1:
        try:
            orc = ge.load_oracle()
            r = orc.RefChain(args.nbit, args.npol, args.rfi_mode)
            one = d_in[:, 0].contiguous()
            r.time_device(one.data_ptr(), 2)
            ms = min(r.time_device(one.data_ptr(), SEG_PER_SEC) for _ in range(3))
            legacy = {"value": 1.0 / (ms / 1e3), "unit": UNIT, "ms_per_antenna_second": ms,
                      "what": "reference pb_kernels.cu compiled unmodified for sm_100a + cuFFT, device-resident, same GPU"}
            r.close()
        except Exception as e:      # oracle/_ref absent on this box
            legacy = {"unavailable": str(e)[:200]}

    # launches of this library's kernels inside the timed region, per rank 0: per antenna-second one channeliser and one
    # normaliser launch per batch of segments, one local co-add sum, and (root) one scale + digitise
    gpu_launches = int(K * S * (2 * launches_per_second + 2))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": args.warmup,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_block(args, n_ant, world),
        "notes": {"coadd": "every rank: local sum + count of its antennas' f32 tiles per antenna-second; %s; scale + digitise on the root"
                           % ("one NCCL reduce of the 10 segments" if world > 1 else "single rank, no collective"),
                  "timing": "CUDA events on the library's streams (vf_timer_begin/end: slot streams and co-add stream) around the K steps, max over ranks"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "legacy_cuda": legacy,
        "coadd_check": check, "coadd_check_after": check_after,
        "wall_ms_per_step": 1e3 * wall / K, "timed_region_ms": dev_ms,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if check_after is not None and not check_after["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
