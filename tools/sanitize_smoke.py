import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xmipp3_b200 import synth, geometry
from xmipp3_b200._lib import Reconstructor, FourierProjector, make_particles
N, n = 24, 24
d = synth.make_dataset(n, N, seed=3, ctf=True, shifts=True)
cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"] + 0.3, shift_y=d["shift_y"] - 0.7, **d["ctf"])
p = make_particles(n, **cols)
for fast in (False, True):
    for sym in (None, "d7"):
        mats = geometry.point_group_matrices(sym) if sym else None
        r = Reconstructor(N, use_ctf=True, sampling=d["sampling"], fast=fast, sym_matrices=mats, max_batch=16)
        r.insert(d["images"], p)
        v = r.finalize()
        r.close()
        print("fast", fast, "sym", sym, "finite", bool(np.isfinite(v).all()))
r = Reconstructor(25, padding=(1.0, 1.5), max_resolution=0.3, n_iter_weight=2)
d2 = synth.make_dataset(10, 25, seed=4)
r.insert(d2["images"], make_particles(10, rot=d2["rot"], tilt=d2["tilt"], psi=d2["psi"]))
r.halfset_push(); r.insert(d2["images"], make_particles(10, rot=d2["rot"], tilt=d2["tilt"], psi=d2["psi"])); r.halfset_merge()
print("odd box finite", bool(np.isfinite(r.finalize()).all())); r.close()
vol = synth.phantom_volume(synth.make_phantom(n_gauss=5, box=N, seed=1), N).astype(np.float32)
for deg in (0, 1, 3):
    g = FourierProjector(vol, 2.0, 0.5, deg)
    out = g.project(d["rot"][:6], d["tilt"][:6], d["psi"][:6])
    g.close()
    print("projector degree", deg, "finite", bool(np.isfinite(out).all()))
