"""Concurrent pinned host-to-device copy rate of the box: the ceiling of the bench's e2e line at N ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      scripts/ubench/h2d_concurrent.py [--numa 0|1]

Every rank copies a pinned 256 MB buffer (one antenna-second of baseband) to its GPU 40 times, all ranks started
together; rank 0 prints per-rank and aggregate GB/s, plus the GPU <-> CPU/NUMA affinity sysfs reports.  --numa 1
binds the rank to the GPU's local CPUs (vf_bind_thread_to_gpu) BEFORE the pinned allocation, as bench.py does."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--numa", type=int, default=1)
ap.add_argument("--mb", type=int, default=256)
ap.add_argument("--reps", type=int, default=40)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cpul = ge.load_package().bind_thread_to_gpu(local) if a.numa else ""
n = a.mb * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(a.reps):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
gbs = n * a.reps / (e0.elapsed_time(e1) / 1e3) / 1e9
res = torch.tensor([gbs, wall], device="cuda")
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    rates = [float(r[0]) for r in allr]
    walls = [float(r[1]) for r in allr]
    agg = world * n * a.reps / max(walls) / 1e9
    print("ranks %d  numa_bind %d  cpulist[rank0] %r  per-rank GB/s %s  aggregate %.1f GB/s (%.1f per GPU)" % (
        world, a.numa, cpul, " ".join("%.1f" % r for r in rates), agg, agg / world), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
