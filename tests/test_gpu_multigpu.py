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


def _worker(rank, world, N, n, tmp):
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
    r.nccl_init(uid, world, rank)
    r.insert(d["images"][b:e], p[b:e])
    r.reduce(0)
    r.sync()
    if rank == 0:
        np.save(os.path.join(tmp, "vol.npy"), r.finalize())
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
