"""Throughput of the drop-in CLI from files (metadata + MRC stack on local disk -> volume file), i.e. including
file I/O and metadata parsing: generates n synthetic box x box particles with CTF, writes them as one .mrcs stack and
an .xmd, runs xmipp_reconstruct_fourier_b200 with the given loader thread counts and prints its own timing line."""
import argparse
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=32768)
    ap.add_argument("--box", type=int, default=256)
    ap.add_argument("--thr", type=int, nargs="+", default=[4, 16])
    ap.add_argument("--buffer", type=int, default=1024)
    ap.add_argument("--dir", default=None)
    ap.add_argument("--gpus", nargs="+", default=["1"], help="values passed to the CLI as --gpus (numbers or 'all'), one run each")
    a = ap.parse_args()
    import torch
    from bench import synth_batch_torch
    from xmipp3_b200 import _build, io
    _build.build_host()
    d = a.dir or tempfile.mkdtemp(prefix="rfb200_cli_")
    os.makedirs(d, exist_ok=True)
    stack = os.path.join(d, "particles.mrcs")
    md = os.path.join(d, "input.xmd")
    reuse = os.path.exists(md) and os.path.exists(stack) and os.path.getsize(stack) >= a.n * a.box * a.box * 4
    dev = torch.device("cuda", 0)
    imgs = np.empty((a.n if not reuse else 1, a.box, a.box), np.float32)
    cols = None
    for b0 in range(0, a.n if not reuse else 0, 4096):
        b1 = min(a.n, b0 + 4096)
        t, c = synth_batch_torch(b1 - b0, a.box, 7 + b0, dev, ctf=True)
        imgs[b0:b1] = t.cpu().numpy()
        cols = c if cols is None else {k: np.concatenate([cols[k], c[k]]) for k in c}
    if not reuse:
        io.write_mrc(stack, imgs)
    if not reuse:
        io.write_xmd(md, {"image": ["%06d@particles.mrcs" % (k + 1) for k in range(a.n)], "enabled": [1] * a.n,
                      "angleRot": cols["rot"], "angleTilt": cols["tilt"], "anglePsi": cols["psi"],
                      "shiftX": cols.get("shift_x", np.zeros(a.n)), "shiftY": cols.get("shift_y", np.zeros(a.n)),
                      "ctfVoltage": cols["kV"], "ctfDefocusU": cols["defocusU"], "ctfDefocusV": cols["defocusV"],
                      "ctfDefocusAngle": cols["defocus_angle"], "ctfSphericalAberration": cols["Cs"], "ctfQ0": cols["Q0"]})
    del imgs
    exe = _build.CLI_BIN
    for thr, gpus in [(t, g) for t in a.thr for g in a.gpus]:
        t0 = time.perf_counter()
        out = subprocess.run([exe, "-i", md, "-o", os.path.join(d, "rec.vol"), "--useCTF", "--sampling", "1.5", "--thr", str(thr),
                              "--bufferSize", str(a.buffer), "-v", "1"] + (["--gpus", gpus] if gpus != "1" else []),
                             capture_output=True, text=True)
        dt = time.perf_counter() - t0
        tail = [l for l in out.stdout.replace("\r", "\n").splitlines() if ("images in" in l and "inserted" not in l) or "GPU time" in l or "wall (s)" in l]
        print("gpus=%s thr=%d wall %.2f s (%.0f images/s incl. process start) rc=%d | %s" % (gpus, thr, dt, a.n / dt, out.returncode, " | ".join(t.strip() for t in tail)), flush=True)
        if out.returncode:
            print(out.stderr[-500:])


if __name__ == "__main__":
    main()
