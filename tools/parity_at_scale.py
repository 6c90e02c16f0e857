#!/usr/bin/env python
"""North-star check at scale: N x 256^2 particles with CTF (BASELINE config 3 as stated: 100k) reconstructed by the GPU
library (FP32 accumulators, one handle) and by the double-precision CPU oracle (race-free parallel mode, bit-identical
to its single-thread result) on the SAME particles; reports rel-L2 and the FSC curve of the two maps.

    python tools/parity_at_scale.py --n 100000 --out gpurun_out/r2_parity_100k.json

Test infrastructure (uses oracle/); the particles are generated chunk by chunk on the GPU with torch."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--box", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=2000)
    ap.add_argument("--sym", default="c1")
    ap.add_argument("--no-ctf", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_at_scale.json"))
    args = ap.parse_args()
    import torch
    import bench
    from oracle import oracle as O
    from xmipp3_b200 import geometry, synth
    from xmipp3_b200._lib import Reconstructor, make_particles
    O.build()
    dev = torch.device("cuda", 0)
    cores = os.cpu_count() or 1
    ctf = not args.no_ctf
    mats = geometry.point_group_matrices(args.sym) if args.sym.lower() != "c1" else None
    kw = dict(use_ctf=ctf, sampling=bench.SAMPLING, sym_matrices=mats)
    r = Reconstructor(args.box, **kw)
    o = O.Oracle(args.box, **kw)
    t_gpu = t_cpu = 0.0
    done = 0
    t00 = time.time()
    while done < args.n:
        n = min(args.chunk, args.n - done)
        img, cols = bench.synth_batch_torch(n, args.box, 90000 + done, dev, ctf=ctf)
        p = make_particles(n, **cols)
        torch.cuda.synchronize()
        t = time.time()
        r.insert_device_ptr(img.data_ptr(), p)
        r.sync()
        t_gpu += time.time() - t
        h = img.cpu().numpy()
        del img
        t = time.time()
        o.insert(h, O.make_particles(n, **cols), threads=cores, scheme="slabs")
        t_cpu += time.time() - t
        done += n
        print("[parity_at_scale] %d / %d particles, gpu %.1f s, oracle %.1f s, wall %.0f s" % (done, args.n, t_gpu, t_cpu, time.time() - t00),
              file=sys.stderr, flush=True)
    v = r.finalize()
    V, W = r.accumulators()
    Vo, Wo = o.accumulators()
    acc_v = float(np.linalg.norm(V[:, :, 1:] - Vo[:, :, 1:]) / np.linalg.norm(Vo[:, :, 1:]))
    acc_w = float(np.linalg.norm(W[:, :, 1:] - Wo[:, :, 1:]) / np.linalg.norm(Wo[:, :, 1:]))
    w_max = float(Wo.max())
    del V, W, Vo, Wo
    vo = o.finalize()
    f = synth.fsc(v, vo)
    res = {
        "what": "GPU library (FP32 accumulation, one handle) vs CPU oracle (FP64, scheme=slabs) on the same particles",
        "particles": args.n, "box": args.box, "padding": 2, "ctf": ctf, "sym": args.sym.lower(),
        "rel_l2_map": float(synth.rel_l2(v, vo)), "min_fsc": float(np.nanmin(f[1:])),
        "fsc_first_last": [float(f[1]), float(f[len(f) // 2]), float(f[-1])],
        "acc_rel_l2_V": acc_v, "acc_rel_l2_W": acc_w, "largest_weight": w_max,
        "gate": {"rel_l2": 1e-4, "fsc": 0.999},
        "passed": bool(synth.rel_l2(v, vo) <= 1e-4 and np.nanmin(f[1:]) >= 0.999),
        "gpu_insert_s": t_gpu, "oracle_insert_s": t_cpu, "oracle_threads": cores,
        "gpu_particles_per_s": args.n / t_gpu, "oracle_particles_per_s": args.n / t_cpu,
    }
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res))
    return 0 if res["passed"] else 1


if __name__ == "__main__":
    sys.exit(main())
