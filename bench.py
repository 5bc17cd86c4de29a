#!/usr/bin/env python
"""bench.py -- baseband antenna-seconds per second (x real time) per B200.

A "step" is one pass of the baseband -> filterbank chain over one antenna-second
of synthetic baseband per antenna of the rank (BASELINE.json configs[1]: 1
antenna, 2 pols, 128 MS/s, 10 segments of 1024 FFTs of 12500 samples,
kurtosis excision on, reference defaults nbit 2 / npol 1 / rfi_mode 2).

  value     device-resident: inputs already in HBM, CUDA events on the
            library's streams around the K timed steps.
  e2e       same metric through the host-buffer C ABI (vf_submit_async /
            vf_wait, pinned host memory, double-buffered H2D and D2H inside the
            timed region, wall clock around a synchronised region).
  roofline  dominant kernel (vf_k1_pipelined, the channeliser): algorithmic
            bytes per launch divided by its mean launch duration, against
            MEASURED_PEAKS.json.  The kernel is bound by fp32 issue, not by HBM
            (DESIGN.md section 4), so the same duration is also set against the
            fp32 peak (roofline.fp32).
  cpu_baseline  the CPU oracle (oracle/liboracle.so, OpenMP) on the host cores.

--impl reference times that CPU oracle as the whole arm (the reference has no
CPU path).  With N > 1 ranks every rank channelises
its own antennas (weak scaling, no data-path collective) and the co-added
filterbank of all antennas is produced by an NCCL reduce of the f32 tiles
(vf_coadd_segment) once per segment inside the timed region.
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

SEG_PER_SEC = 10
T = 1024
NSAMP = T * 12500
METRIC = "baseband antenna-seconds/s per B200 (x real-time)"
UNIT = "antenna-seconds/s"
GEN = dict(seed=102, rfi_amp=60, rfi_burst_every=16)      # scripts/baseband_test:21 uses -r 102


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
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


def cpu_oracle_run(nseg_sample, pkg, nthreads=0):
    """oracle A over nseg_sample segments of antenna 0; returns (antenna-seconds/s, cores, seconds)"""
    orc = ge.load_oracle()
    cores = nthreads or os.cpu_count()
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
    return (nseg_sample / SEG_PER_SEC) / dt, cores, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    pkg = ge.load_package()
    per_step = 5                            # segments per step: a bounded sample (0.5 antenna-second)
    vals = []
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_oracle_run(1, pkg)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, dt = cpu_oracle_run(per_step, pkg)
        vals.append((per_step / SEG_PER_SEC, dt))
    tot_as = sum(v[0] for v in vals); tot_t = sum(v[1] for v in vals)
    value = tot_as / tot_t
    sample = "%d x %d segments of 100 ms (1 antenna, rfi_mode 2, nbit 2), CPU oracle with OpenMP" % (args.steps, per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1 antenna, 1 s genbase-style VDIF samples, default channelisation, kurtosis excision on",
                   "nbit": 2, "npol": 1, "rfi_mode": 2, "note": "the reference has no CPU implementation; this is the C restatement (oracle/) of its chain"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    ap.add_argument("--nbit", type=int, default=2)
    ap.add_argument("--npol", type=int, default=1)
    ap.add_argument("--rfi-mode", type=int, default=2)
    ap.add_argument("--k1-threads", type=int, default=0)
    ap.add_argument("--max-batch", type=int, default=SEG_PER_SEC,
                    help="segments per launch pair (default: the 10 of an antenna-second; 1 = per segment; 0 = library default, 16)")
    ap.add_argument("--clean", action="store_true", help="no impulsive RFI in the synthetic input (no time step needs the second FFT)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legacy", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.clean:
        GEN["rfi_amp"] = 0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    n_ant = args.antennas
    p = pkg.Pipeline(ffts_per_seg=T, nbit=args.nbit, npol=args.npol, rfi_mode=args.rfi_mode, gpu_id=local,
                     n_antennas=n_ant, keep_power=1 if world > 1 else 0, k1_threads=args.k1_threads,
                     max_batch_segments=args.max_batch,
                     power_segments=2 * SEG_PER_SEC if world > 1 else 0)
    out_bytes = p.out_bytes

    # ---- inputs: one antenna-second per antenna, pinned on the host and resident on the device
    host = torch.empty((SEG_PER_SEC, n_ant, 2, NSAMP), dtype=torch.uint8, pin_memory=True)
    hn = host.numpy()
    tmp = np.empty((SEG_PER_SEC, 2, NSAMP), np.uint8)
    for a in range(n_ant):
        make_second(pkg, rank * n_ant + a, tmp)
        hn[:, a] = tmp
    d_in = host.cuda(non_blocking=False)
    d_main = torch.zeros((SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, device="cuda")
    d_raw = torch.zeros_like(d_main) if args.rfi_mode == 2 else None
    h_main = torch.empty((SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, pin_memory=True)
    h_raw = torch.empty_like(h_main).pin_memory() if args.rfi_mode == 2 else None
    torch.cuda.synchronize()

    if world > 1:
        uid = [p.coadd_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        p.coadd_init(world, rank, uid[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # world == 1: the K timed steps go to the library as ONE call over K consecutive antenna-seconds (the second of
    # synthetic baseband repeated K times in HBM), so that its launches overlap across steps as they do in a stream
    if world == 1:
        d_in_k = d_in.repeat(args.steps, 1, 1, 1).contiguous()
        d_main_k = torch.zeros((args.steps * SEG_PER_SEC, n_ant, out_bytes), dtype=torch.uint8, device="cuda")
        d_raw_k = torch.zeros_like(d_main_k) if args.rfi_mode == 2 else None
        torch.cuda.synchronize()

    def steps_device_all():
        p.process_device(n_ant, SEG_PER_SEC * args.steps, d_in_k.data_ptr(), d_main_k.data_ptr(),
                         d_raw_k.data_ptr() if d_raw_k is not None else None)

    def step_device():
        if world == 1:
            p.process_device(n_ant, SEG_PER_SEC, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr() if d_raw is not None else None)
        else:
            p.process_device(n_ant, SEG_PER_SEC, d_in.data_ptr(), d_main.data_ptr(), d_raw.data_ptr() if d_raw is not None else None)
            p.coadd_batch(0, world * n_ant, SEG_PER_SEC, want=False, wait=False)      # one reduce per second

    # ---- device-resident timing -------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    p.sync()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    k1_ms = k2_ms = dev_ms = 0.0
    t0 = time.perf_counter()
    if world == 1:
        steps_device_all()
        p.sync()
        dev_ms, k1_ms, k2_ms = p.last_elapsed_ms()
    else:
        for _ in range(args.steps):
            step_device()
    p.sync()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.finish()
    if world == 1:
        t_ms = dev_ms                        # CUDA events on the library's control stream, per step
    else:
        t_ms = wall * 1e3
        tt = torch.tensor([t_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    ant_seconds = args.steps * n_ant * world
    value = ant_seconds / (t_ms / 1e3)

    # ---- pure kernel durations for the roofline: same steps, segments not overlapped -------------
    if world == 1:
        p.set_serial(1)
        step_device(); p.sync()
        k1_ms = k2_ms = 0.0
        nser = max(2, min(args.steps, 5))
        for _ in range(nser):
            step_device(); p.sync()
            _, b, c = p.last_elapsed_ms()
            k1_ms += b; k2_ms += c
        p.set_serial(0)
        k1_ms /= nser * SEG_PER_SEC; k2_ms /= nser * SEG_PER_SEC     # per segment
        step_device(); p.sync()

    # ---- end to end through host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            for s in range(SEG_PER_SEC):
                slot = s & 1
                if s >= 2:
                    p.wait(slot)
                p.submit_async(slot, [hn[s, a, 0] for a in range(n_ant)], [hn[s, a, 1] for a in range(n_ant)],
                               [h_main[s, a].numpy() for a in range(n_ant)],
                               [h_raw[s, a].numpy() for a in range(n_ant)] if h_raw is not None else None)
            p.wait(0); p.wait(1)
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e_t = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([e_t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_t = float(tt.item())
        nstream = 2 if args.rfi_mode == 2 else 1
        e2e = {"value": ant_seconds / e_t, "unit": UNIT,
               "h2d_bytes_per_step": int(n_ant * 2 * NSAMP * SEG_PER_SEC),
               "d2h_bytes_per_step": int(n_ant * out_bytes * SEG_PER_SEC * nstream),
               "api": "vf_submit_async/vf_wait, pinned host buffers, 2 slots"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------
    peak, peak_src = peaks()
    nstream = 2 if args.rfi_mode == 2 else 1
    seg_per_launch = min(SEG_PER_SEC, args.max_batch if args.max_batch > 0 else 16)     # vf_process_device batches segments
    alg_bytes_per_seg = n_ant * (2 * NSAMP + out_bytes * nstream)
    alg_bytes_per_launch = alg_bytes_per_seg * seg_per_launch
    roofline = None
    if world == 1 and k1_ms > 0:
        k1_avg_s = k1_ms * seg_per_launch / 1e3                      # mean duration of one launch
        achieved = alg_bytes_per_launch / k1_avg_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
                tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch_1ant")
                if traffic is not None:     # the captured launch covered tj["segments_per_launch"] segments of 1 antenna
                    traffic = traffic * n_ant * seg_per_launch / tj.get("segments_per_launch", 1)
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
                               "116-128 lanes/clk/SM in scripts/ubench/fp32_rate.cu (gpurun_out/fp32_rate.log)"}
        roofline = {"bound": "hbm", "kernel": "vf_k1_pipelined" if args.k1_threads == 0 else "vf_k1_channelise<%d>" % args.k1_threads,
                    "fp32": fp32, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": alg_bytes_per_launch, "segments_per_launch": seg_per_launch,
                    "k1_ms_per_launch": k1_ms * seg_per_launch, "k2_ms_per_launch": k2_ms * seg_per_launch,
                    "k1_ms_per_segment": k1_ms, "k2_ms_per_segment": k2_ms,
                    "launch_timing": "CUDA events around each launch on the library's stream, segments serialised (vf_set_serial) so that no queueing is included",
                    "whole_chain_achieved": alg_bytes_per_seg * SEG_PER_SEC * args.steps / (t_ms / 1e3) / 1e9}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        v, cores, dt = cpu_oracle_run(20, pkg)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "20 segments of 100 ms (2 antenna-seconds) of the same workload, %.1f s" % dt}

    legacy = None
    if not args.no_legacy and world == 1 and n_ant >= 1:
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1 antenna-second per antenna (10 segments x 1024 FFTs x 12500 samples x 2 pols), "
                               "default channelisation, kurtosis excision on",
                   "antennas_per_gpu": n_ant, "nbit": args.nbit, "npol": args.npol, "rfi_mode": args.rfi_mode,
                   "l2": "inputs (%.0f MB per step per GPU) exceed the 126 MB L2; no flush" % (n_ant * 2 * NSAMP * SEG_PER_SEC / 1e6),
                   "generator": GEN, "coadd": "one NCCL reduce of the f32 tiles of the 10 segments per step" if world > 1 else "none",
                   "timing": "CUDA events on the library's stream (fork/join over its 2 slot streams) around one call that covers the K steps" if world == 1
                             else "wall clock between barrier+synchronize, max over ranks"},
        "clocks": clocks, "e2e": e2e,
        "gpu_launches": int(args.steps * -(-SEG_PER_SEC // seg_per_launch) * 2 + (args.steps * SEG_PER_SEC * n_ant if world > 1 else 0)),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "legacy_cuda": legacy,
        "wall_ms_per_step": 1e3 * wall / args.steps,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
