"""Multi-GPU path (needs >= 2 CUDA devices; skipped otherwise): one process per GPU, particle sharding, private
accumulators, a single ncclReduce of V and W onto rank 0 over NVLink (rfb200_reduce_nccl), finalisation on rank 0.
The result must equal the single-GPU reconstruction of the whole set."""
import os
import tempfile

import numpy as np
import pytest

from xmipp3_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _exchange(tmp, name, rank, world, blob):
    """every rank publishes `blob` as a file and reads the others' (the host program's job: any transport will do)"""
    import time
    path = lambda k: os.path.join(tmp, "%s_%d" % (name, k))
    with open(path(rank) + ".tmp", "wb") as f:
        f.write(blob)
    os.rename(path(rank) + ".tmp", path(rank))
    out = []
    for k in range(world):
        while not os.path.exists(path(k)):
            time.sleep(0.05)
        out.append(open(path(k), "rb").read())
    return out


def _worker(rank, world, N, n, tmp, p2p=False):
    from xmipp3_b200._lib import Reconstructor, make_particles
    d = synth.make_dataset(n, N, seed=7, ctf=True)
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], **d["ctf"])
    p = make_particles(n, **cols)
    b, e = sharding.shard_range(n, world, rank)
    r = Reconstructor(N, use_ctf=True, sampling=1.5, device=rank)
    idfile = os.path.join(tmp, "nccl_id")
    if rank == 0:
        uid = Reconstructor.nccl_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        import time
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    if p2p == "host":
        r.set_ranks(world, rank)          # no NCCL at all: the ranks order themselves (here: through files)
    else:
        r.nccl_init(uid, world, rank)
    if p2p:
        for k, blob in enumerate(_exchange(tmp, "ipc", rank, world, r.ipc_export())):
            if k != rank:
                r.ipc_import(k, blob)
    r.insert(d["images"][b:e], p[b:e])
    if p2p == "host":
        r.reduce_p2p_prepare()
        _exchange(tmp, "b0", rank, world, b"1")
        r.reduce_p2p_run(0)
        _exchange(tmp, "b1", rank, world, b"1")
    elif p2p:
        r.reduce_p2p(0)
    else:
        r.reduce(0)
    r.sync()
    if rank == 0:
        np.save(os.path.join(tmp, "vol.npy"), r.finalize())
    if p2p:
        r.ipc_release()
    _exchange(tmp, "done", rank, world, b"1")     # nobody frees memory a peer has mapped before everybody has unmapped it
    r.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_reduce_matches_single_gpu():
    import torch.multiprocessing as mp
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 400
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, N, n, tmp), nprocs=2, join=True)
        vol2 = np.load(os.path.join(tmp, "vol.npy"))
    d = synth.make_dataset(n, N, seed=7, ctf=True)
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], **d["ctf"])
    r = Reconstructor(N, use_ctf=True, sampling=1.5, device=0)
    r.insert(d["images"], make_particles(n, **cols))
    vol1 = r.finalize()
    r.close()
    assert synth.rel_l2(vol2, vol1) <= 2e-5     # FP32 sums in a different order (two partial volumes)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_peer_memory_reduce_matches_single_gpu():
    """rfb200_reduce_p2p: every rank sums its half of V and W out of both GPUs' memory (CUDA IPC over NVLink) into rank 0."""
    import torch.multiprocessing as mp
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 400
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, N, n, tmp, True), nprocs=2, join=True)
        vol2 = np.load(os.path.join(tmp, "vol.npy"))
    d = synth.make_dataset(n, N, seed=7, ctf=True)
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], **d["ctf"])
    r = Reconstructor(N, use_ctf=True, sampling=1.5, device=0)
    r.insert(d["images"], make_particles(n, **cols))
    vol1 = r.finalize()
    r.close()
    assert synth.rel_l2(vol2, vol1) <= 2e-5


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_gpu_peer_memory_reduce_without_nccl():
    """rfb200_set_ranks + rfb200_reduce_p2p_prepare / _run with the host program's own barriers: no communicator."""
    import torch.multiprocessing as mp
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 400
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, N, n, tmp, "host"), nprocs=2, join=True)
        vol2 = np.load(os.path.join(tmp, "vol.npy"))
    d = synth.make_dataset(n, N, seed=7, ctf=True)
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], **d["ctf"])
    r = Reconstructor(N, use_ctf=True, sampling=1.5, device=0)
    r.insert(d["images"], make_particles(n, **cols))
    vol1 = r.finalize()
    r.close()
    assert synth.rel_l2(vol2, vol1) <= 2e-5
