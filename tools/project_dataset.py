#!/usr/bin/env python
"""Make a particle set on the GPU the way xmipp_phantom_project does (reconstruction/project.cpp:992: FourierProjector over
random orientations), ready for xmipp_reconstruct_fourier_b200:

    python tools/project_dataset.py -i volume.vol -o particles --n 100000 [--ctf --sampling 1.5] [--degree 3]
    python tools/project_dataset.py --phantom 256 -o particles --n 100000 --ctf

writes particles.stk (Spider stack) and particles.xmd (image, angles, CTF columns).  The projector is the library's
rfb200_projector_* (no CPU fallback); the CTF is passed as the projector's per-image Fourier multiplier."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xmipp3_b200 import io, synth                      # noqa: E402
from xmipp3_b200._lib import FourierProjector          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-i", help="input volume (Spider / MRC), cubic")
    ap.add_argument("--phantom", type=int, default=0, help="instead of -i: Gaussian phantom of this box size")
    ap.add_argument("-o", required=True, help="output root: <root>.stk and <root>.xmd")
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--padding", type=float, default=2.0)
    ap.add_argument("--max_freq", type=float, default=0.5)
    ap.add_argument("--degree", type=int, default=3, choices=(0, 1, 3))
    ap.add_argument("--ctf", action="store_true")
    ap.add_argument("--sampling", type=float, default=1.5)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--batch", type=int, default=2048)
    a = ap.parse_args()
    if a.phantom:
        vol = synth.phantom_volume(synth.make_phantom(n_gauss=30, box=a.phantom, seed=a.seed), a.phantom).astype(np.float32)
    elif a.i:
        vol = io.read_volume(a.i).astype(np.float32)
    else:
        ap.error("give -i <volume> or --phantom <box>")
    N = vol.shape[0]
    rot, tilt, psi = synth.random_orientations(a.n, a.seed + 1)
    cp = synth.random_ctf_params(a.n, a.seed + 3) if a.ctf else None
    t0 = time.perf_counter()
    pr = FourierProjector(vol, a.padding, a.max_freq, a.degree, device=a.device)
    t1 = time.perf_counter()
    stack = a.o + ".stk"
    imgs = np.empty((a.n, N, N), dtype=np.float32)
    for b0 in range(0, a.n, a.batch):
        b1 = min(a.n, b0 + a.batch)
        ctf = None
        if cp is not None:
            ctf = np.stack([synth.ctf_2d(N, a.sampling, cp["kV"][k], cp["defocusU"][k], cp["defocusV"][k], cp["defocus_angle"][k],
                                         cp["Cs"][k], cp["Q0"][k])[:, :N // 2 + 1] for k in range(b0, b1)]).astype(np.float32)
        imgs[b0:b1] = pr.project(rot[b0:b1], tilt[b0:b1], psi[b0:b1], ctf)
    t2 = time.perf_counter()
    pr.close()
    io.write_spider_stack(stack, imgs)
    cols = {"image": ["%06d@%s" % (k + 1, os.path.basename(stack)) for k in range(a.n)], "enabled": [1] * a.n,
            "angleRot": rot, "angleTilt": tilt, "anglePsi": psi, "shiftX": np.zeros(a.n), "shiftY": np.zeros(a.n)}
    if cp is not None:
        cols.update({"ctfVoltage": cp["kV"], "ctfDefocusU": cp["defocusU"], "ctfDefocusV": cp["defocusV"],
                     "ctfDefocusAngle": cp["defocus_angle"], "ctfSphericalAberration": cp["Cs"], "ctfQ0": cp["Q0"]})
    io.write_xmd(a.o + ".xmd", cols)
    print("projector build %.2f s, %d projections of %d^2 in %.2f s (%.0f /s incl. CTF images and D2H), wrote %s + .xmd"
          % (t1 - t0, a.n, N, t2 - t1, a.n / max(t2 - t1, 1e-9), stack))


if __name__ == "__main__":
    main()
