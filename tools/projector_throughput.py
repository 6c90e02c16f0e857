import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xmipp3_b200 import synth
from xmipp3_b200._lib import FourierProjector
import torch
N, n = 256, 2048
vol = synth.phantom_volume(synth.make_phantom(n_gauss=30, box=N, seed=0), N).astype(np.float32)
rot, tilt, psi = synth.random_orientations(n, 1)
w = FourierProjector(vol[:32, :32, :32].copy(), 2.0, 0.5, 3)      # context, cuFFT load
w.close()
for deg in (3, 1, 3):
    t = time.perf_counter()
    g = FourierProjector(vol, 2.0, 0.5, deg)
    tb = time.perf_counter() - t
    out = torch.empty((n, N, N), device="cuda", dtype=torch.float32)
    g.project_device_ptr(rot, tilt, psi, out.data_ptr())
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        g.project_device_ptr(rot, tilt, psi, out.data_ptr())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print("degree %d: build %.2f s, %d projections of %d^2 in %.1f ms (%.0f /s), finite %s" % (deg, tb, n, N, dt * 1e3, n / dt, bool(torch.isfinite(out).all())))
    g.close()
