import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from xmipp3_b200 import synth
from xmipp3_b200._lib import Reconstructor, make_particles
N, n = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 600
d = synth.make_dataset(n, N, seed=0, ctf=True, shifts=True)
cols = dict(rot=d['rot'], tilt=d['tilt'], psi=d['psi'], shift_x=d['shift_x'], shift_y=d['shift_y'], **d['ctf'])
o = O.Oracle(N, use_ctf=True, sampling=1.5)
op = O.make_particles(n, **cols)
o.insert(d['images'], op, threads=1)
Vo, Wo = o.accumulators()
r = Reconstructor(N, use_ctf=True, sampling=1.5)
r.insert(d['images'], make_particles(n, **cols)); r.sync()
V, W = r.accumulators()
Z = o.Z
dW = np.abs(W - Wo); dW[:, :, 0] = 0
print("relW", np.linalg.norm(dW) / np.linalg.norm(Wo[:, :, 1:]))
idx = np.argsort(dW.ravel())[::-1][:12]
for k in idx:
    z, y, x = np.unravel_index(k, W.shape)
    uy = y if y <= Z//2 else y - Z; uz = z if z <= Z//2 else z - Z
    print("u=(%d,%d,%d) |u|=%.2f W=%.6f Wo=%.6f diff=%.3e" % (x, uy, uz, np.sqrt(x*x+uy*uy+uz*uz), W[z,y,x], Wo[z,y,x], dW[z,y,x]))
# per-image check of the m channel against the oracle's ctf weights
P = o.P
bad = 0
for k in range(min(n, 512), n)[:0]:
    pass
# find images whose slice m-channel disagrees with the oracle rule
cnt = 0
for k in list(range(0, n))[-88:]:
    pass
print("n>1e-3:", (dW > 1e-3).sum())
