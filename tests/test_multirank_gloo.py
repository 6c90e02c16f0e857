"""CPU test of the N>1 host logic with world_size-2 gloo: particle sharding + reduce of the private
accumulators + finalisation on the root.  (The per-rank accumulation itself is done by the CPU oracle
here; on the GPU box the same flow runs through rfb200_reduce_nccl, see tests/test_gpu_multigpu.py.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xmipp3_b200 import sharding, synth


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            r = sharding.all_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    d = synth.make_dataset(n, N, seed=3)
    b, e = sharding.shard_range(n, world, rank)
    p = O.make_particles(e - b, rot=d["rot"][b:e], tilt=d["tilt"][b:e], psi=d["psi"][b:e])
    o = O.Oracle(N)
    o.insert(d["images"][b:e], p, threads=1)
    V, W = o.accumulators()
    tV = torch.from_numpy(np.ascontiguousarray(V.view(np.float64)))
    tW = torch.from_numpy(W)
    dist.reduce(tV, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(tW, dst=0, op=dist.ReduceOp.SUM)
    if rank == 0:
        root = O.Oracle(N)
        root.add_accumulators(tV.numpy().view(np.complex128), tW.numpy())
        np.save(out_path, root.finalize())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_rank(tmp_path, oracle_mod):
    N, n = 16, 24
    out = str(tmp_path / "vol.npy")
    mp.spawn(_worker, args=(2, _free_port(), N, n, out), nprocs=2, join=True)
    vol2 = np.load(out)
    d = synth.make_dataset(n, N, seed=3)
    o = oracle_mod.Oracle(N)
    o.insert(d["images"], oracle_mod.make_particles(n, rot=d["rot"], tilt=d["tilt"], psi=d["psi"]), threads=1)
    vol1 = o.finalize()
    assert synth.rel_l2(vol2, vol1) < 1e-12
