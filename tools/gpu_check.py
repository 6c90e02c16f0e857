"""Developer check: CUDA path vs CPU oracle on small seeded problems (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from xmipp3_b200 import synth, geometry
from xmipp3_b200._lib import Reconstructor, make_particles


def run(N, n, ctf=False, shifts=False, sym=None, max_res=0.5, pad=(2.0, 2.0), blob=(1.9, 0, 15.0), seed=0, tag=""):
    d = synth.make_dataset(n, N, seed=seed, ctf=ctf, shifts=shifts, sym=sym)
    cols = dict(rot=d['rot'], tilt=d['tilt'], psi=d['psi'], shift_x=d['shift_x'], shift_y=d['shift_y'])
    if ctf:
        cols.update(d['ctf'])
    mats = geometry.point_group_matrices(sym) if sym else None
    kw = dict(padding=pad, max_resolution=max_res, blob=blob, sym_matrices=mats, use_ctf=ctf, sampling=d['sampling'])
    o = O.Oracle(N, **kw)
    t = time.time(); o.insert(d['images'], O.make_particles(n, **cols), threads=1); tc = time.time() - t
    Vo, Wo = o.accumulators()
    volo = o.finalize()
    r = Reconstructor(N, **kw)
    t = time.time(); r.insert(d['images'], make_particles(n, **cols)); r.sync(); tg = time.time() - t
    V, W = r.accumulators()
    vol = r.finalize()
    # compare accumulators away from the x=0 plane (stored post-symmetrisation there)
    eV = np.linalg.norm(V[:, :, 1:] - Vo[:, :, 1:]) / np.linalg.norm(Vo[:, :, 1:])
    eW = np.linalg.norm(W[:, :, 1:] - Wo[:, :, 1:]) / np.linalg.norm(Wo[:, :, 1:])
    eVmax = np.abs(V[:, :, 1:] - Vo[:, :, 1:]).max() / np.abs(Vo).max()
    eWmax = np.abs(W[:, :, 1:] - Wo[:, :, 1:]).max() / np.abs(Wo).max()
    # x = Z/2 plane and others separately
    Z = o.Z
    eWlast = np.abs(W[:, :, Z // 2] - Wo[:, :, Z // 2]).max() / np.abs(Wo).max()
    ev = synth.rel_l2(vol, volo)
    f = synth.fsc(vol, volo)
    print("%-28s N=%d n=%d relV=%.2e relW=%.2e maxV=%.2e maxW=%.2e Wlast=%.2e | vol rel-L2=%.3e minFSC=%.6f | cpu %.2fs gpu %.3fs  %s" % (
        tag, N, n, eV, eW, eVmax, eWmax, eWlast, ev, np.nanmin(f[1:]), tc, tg, r.timings()), flush=True)
    r.close()
    return ev


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(16, 20, tag="c1")
        run(32, 100, tag="c1")
        run(32, 100, ctf=True, shifts=True, tag="ctf+shifts")
        run(32, 30, sym='d7', tag="d7")
        run(32, 100, max_res=0.3, tag="maxres0.3")
        run(24, 50, pad=(2.0, 2.0), tag="N=24")
        run(25, 50, tag="odd N=25 (Z=50)")
        run(32, 50, pad=(1.0, 2.0), tag="pad 1 2")
        run(32, 50, blob=(1.5, 0, 10.0), tag="blob 1.5/10")
    elif which == "mid":
        run(64, 1000, tag="config1")
        run(128, 1000, ctf=True, shifts=True, tag="config2-subset")
