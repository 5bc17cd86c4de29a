"""World-size-2 run of the multi-rank flow on CPU (gloo): each rank processes its
antenna shard, the f32 tiles are summed across ranks, the root scales and
digitises.  The per-antenna chain here is the CPU oracle (this is a test of the
host-side sharding / reduce / digitise logic, not of the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, make_input, RFI

T, NANT, NBIT = 8, 5, 8


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def local_sum(pkg, orc, ants):
    acc = np.zeros((1, T // 8, 4096), np.float32)
    for a in ants:
        o = orc.OracleChain(T, NBIT, 1, 1, 1)
        o.process_segment(*make_input(pkg, T, seed=90, antenna=a, **RFI))
        acc += o.ave_trimmed("main")
    return acc


def digitise(orc, tile):
    full = np.zeros((1, T // 8, 6251), np.float32)
    full[:, :, 2155:2155 + 4096] = tile
    out = np.empty(tile.size * NBIT // 8, np.uint8)
    orc.liba().orc_digitise(full.ctypes.data, out.ctypes.data, T // 8, 1, NBIT)
    return out


def worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg, orc = ge.load_package(), ge.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ants = pkg.antennas_of_rank(NANT, world, rank)
    t = torch.from_numpy(local_sum(pkg, orc, ants))
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        tile = t.numpy() * np.float32(pkg.coadd_scale(NANT))
        np.save(os.path.join(outdir, "coadd.npy"), digitise(orc, tile))
        np.save(os.path.join(outdir, "sum.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_antenna_sharding(pkg):
    for world in (1, 2, 4, 8):
        shards = [pkg.antennas_of_rank(16, world, r) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(16))
        assert all(len(s) == 16 // world for s in shards)
    assert pkg.antennas_of_rank(5, 2, 0) == [0, 2, 4] and pkg.antennas_of_rank(5, 2, 1) == [1, 3]
    with pytest.raises(ValueError):
        pkg.antennas_of_rank(4, 2, 2)


def test_two_rank_coadd_matches_single_process(pkg, orc, tmp_path):
    port = free_port()
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "coadd.npy")
    gsum = np.load(tmp_path / "sum.npy")
    ref_sum = local_sum(pkg, orc, list(range(NANT)))
    assert np.abs(gsum - ref_sum).max() < 1e-5          # float addition order differs across ranks
    want = digitise(orc, ref_sum * np.float32(pkg.coadd_scale(NANT)))
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and np.count_nonzero(d) <= 3
