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
# power-of-two padded size with integer shifts: the fused K1r -> K1c chain (per-sequence named barriers, cut-off table in
# shared memory, CTF variants) instead of cuFFT; with and without CTF, and with envelope terms (general CTF variant)
d3 = synth.make_dataset(20, 32, seed=5, ctf=True, shifts=True)
c3 = dict(rot=d3["rot"], tilt=d3["tilt"], psi=d3["psi"], shift_x=np.round(d3["shift_x"]), shift_y=np.round(d3["shift_y"]))
for mode in ("none", "fast", "general"):
    extra = dict(d3["ctf"]) if mode != "none" else {}
    if mode == "general":
        extra["espr"] = np.full(20, 1.0)
        extra["DeltaF"] = np.full(20, 50.0)
    r = Reconstructor(32, use_ctf=mode != "none", sampling=d3["sampling"], max_batch=16)
    r.insert(d3["images"], make_particles(20, **c3, **extra))
    print("fused chain, ctf", mode, "finite", bool(np.isfinite(r.finalize()).all()))
    r.close()
# P = 512 (the headline size): sequences of 64 threads with their own named barriers, 4 columns per CTA
d4 = synth.make_dataset(2, 256, seed=6, ctf=True, shifts=False)
r = Reconstructor(256, use_ctf=True, sampling=d4["sampling"], max_batch=16)
r.insert(d4["images"], make_particles(2, rot=d4["rot"], tilt=d4["tilt"], psi=d4["psi"], **d4["ctf"]))
print("box 256 fused chain finite", bool(np.isfinite(r.finalize()).all()))
r.close()
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
