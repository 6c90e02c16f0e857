"""Writes tests/golden/fast_oracle_box16.npz: a regression pin of the --fast CPU restatement on a seeded box-16 case
(the reference holds no golden vector for this mode).  Run from the repository root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from xmipp3_b200 import synth           # noqa: E402

N, n = 16, 40
d = synth.make_dataset(n, N, seed=21, ctf=True, shifts=True)
cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"], **d["ctf"])
f = O.FastOracle(N, use_ctf=True, sampling=d["sampling"], min_ctf=0.1)
f.insert(d["images"], O.make_particles(n, **cols))
V, W = f.temp_spaces()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fast_oracle_box16.npz"),
                    n=n, nnz_w=np.count_nonzero(W), sum_w=W.sum(dtype=np.float64), sum_abs_v=np.abs(V).sum(dtype=np.float64),
                    vol=f.finalize().astype(np.float64))
print("written")
