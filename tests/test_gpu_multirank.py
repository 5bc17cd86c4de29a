"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the multi-rank co-add -- one process
per GPU, each channelising its own antennas, ONE ncclReduce of the f32 tiles and the row counts of a batch of
segments (vf_coadd_batch), digitised on the root -- against the sum of the CPU oracle's tiles (SURVEY.md 8d
config 5).  The NCCL unique id travels through a file: no torch.distributed in this test."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

T, NSEG, PER_RANK, NBIT = 64, 3, 2, 8

RANK_SCRIPT = r'''
import os, sys, time
import numpy as np
sys.path.insert(0, %(root)r)
import __graft_entry__ as ge
pkg = ge.load_package()
rank, world, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
T, NSEG, PER_RANK, NBIT = %(T)d, %(NSEG)d, %(PER_RANK)d, %(NBIT)d
p = pkg.Pipeline(ffts_per_seg=T, nbit=NBIT, rfi_mode=2, gpu_id=rank, n_antennas=PER_RANK, keep_power=1, power_segments=NSEG)
idf = os.path.join(out, "nccl_id")
if rank == 0:
    uid = p.coadd_unique_id()
    open(idf + ".tmp", "wb").write(uid); os.rename(idf + ".tmp", idf)
else:
    for _ in range(600):
        if os.path.exists(idf): break
        time.sleep(0.1)
    uid = open(idf, "rb").read()
p.coadd_init(world, rank, uid)
g = pkg.GenParams.default(seed=95, rfi_amp=60, rfi_burst_every=3)
ants = [a for a in range(world * PER_RANK) if a %% world == rank]
for s in range(NSEG):
    ins = []
    for a in ants:
        p0 = pkg.gen_samples(g, a, 0, s * T * 12500, T * 12500); p1 = pkg.gen_samples(g, a, 1, s * T * 12500, T * 12500)
        if a == 1 and s == 1:
            p0[:8 * 12500] = 0; p1[:8 * 12500] = 0         # one zeroed row on one antenna of rank 1: the count path
        ins.append((p0, p1))
    p.process_batch([i[0] for i in ins], [i[1] for i in ins])
fb, sm = p.coadd_batch(0, world * PER_RANK, NSEG, want=(rank == 0))
if rank == 0:
    np.save(os.path.join(out, "fb.npy"), fb); np.save(os.path.join(out, "sum.npy"), sm)
p.sync(); p.close()
print("rank", rank, "done")
'''


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_coadd_matches_the_sum_of_oracle_tiles(pkg, orc, tmp_path):
    from test_gpu_parity import _row_kept, check_bytes
    world = 2
    script = tmp_path / "rank.py"
    script.write_text(RANK_SCRIPT % dict(root=ROOT, T=T, NSEG=NSEG, PER_RANK=PER_RANK, NBIT=NBIT))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    for pr in procs:
        out, _ = pr.communicate(timeout=600)
        assert pr.returncode == 0, out[-2000:]
    fb, sm = np.load(tmp_path / "fb.npy"), np.load(tmp_path / "sum.npy")
    g = pkg.GenParams.default(seed=95, rfi_amp=60, rfi_burst_every=3)
    n = world * PER_RANK
    oracles = [orc.OracleChain(T, NBIT, 1, 2) for _ in range(n)]
    for s in range(NSEG):
        osum = np.zeros((1, T // 8, 4096), np.float32)
        cnt = np.zeros(T // 8, np.float32)
        for a in range(n):
            p0 = pkg.gen_samples(g, a, 0, s * T * 12500, T * 12500); p1 = pkg.gen_samples(g, a, 1, s * T * 12500, T * 12500)
            if a == 1 and s == 1:
                p0[:8 * 12500] = 0; p1[:8 * 12500] = 0
            oracles[a].process_segment(p0, p1)
            osum += oracles[a].ave_trimmed("main")
            cnt += _row_kept(oracles[a].get("weights")[:T])
        assert np.abs(sm[s] - osum).max() < 5e-4, s
        if s == 1:
            assert cnt[0] == n - 1
        full = np.zeros((1, T // 8, 6251), np.float32)
        full[:, :, 2155:2155 + 4096] = osum / np.sqrt(cnt)[None, :, None]
        want = np.empty(fb[s].size, np.uint8)
        orc.liba().orc_digitise(full.ctypes.data, want.ctypes.data, T // 8, 1, NBIT)
        check_bytes(fb[s], want, NBIT, "coadd seg %d" % s)
