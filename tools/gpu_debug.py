import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from xmipp3_b200 import synth
from xmipp3_b200._lib import Reconstructor, make_particles

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
maxres = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
d = synth.make_dataset(n, N, seed=0)
cols = dict(rot=d['rot'], tilt=d['tilt'], psi=d['psi'])
o = O.Oracle(N, max_resolution=maxres)
op = O.make_particles(n, **cols)
o.insert(d['images'], op, threads=1)
Vo, Wo = o.accumulators()
r = Reconstructor(N, max_resolution=maxres)
r.insert(d['images'], make_particles(n, **cols)); r.sync()
V, W = r.accumulators()
Z = o.Z; P = o.P
# slice check for image 0
S, Rp = r.debug_slice(0)
F, A = o.preprocess(d['images'][0], op[0])
side = S.shape[0]
err = 0; cnt = 0
for ip in range(-P//2+1, P//2+1):
    for j in range(0, P//2+1):
        fx = j / P; fy = ip / P
        if fx*fx + fy*fy > maxres*maxres: continue
        val = F[ip % P, j]
        if j > 0:
            got = S[ip+Rp, j+Rp]
            err = max(err, abs(got[0] + 1j*got[1] - val)); cnt += 1
            got = S[-ip+Rp, -j+Rp]
            err = max(err, abs(got[0] + 1j*got[1] - np.conj(val)))
print("slice max err", err, "of", np.abs(F).max(), "checked", cnt, "nonzero slice entries", (S[:,:,2] != 0).sum())
dW = np.abs(W - Wo)
dW[:, :, 0] = 0
idx = np.argsort(dW.ravel())[::-1][:15]
print("max W", Wo.max())
for k in idx:
    z, y, x = np.unravel_index(k, W.shape)
    uy = y if y <= Z//2 else y - Z; uz = z if z <= Z//2 else z - Z
    print("u=(%d,%d,%d) |u|=%.2f W=%.6f Wo=%.6f diff=%.2e" % (x, uy, uz, np.sqrt(x*x+uy*uy+uz*uz), W[z,y,x], Wo[z,y,x], dW[z,y,x]))
print("count of voxels with err>1e-5:", (dW > 1e-5).sum(), "of nonzero", (Wo > 0).sum())
# missing vs extra
print("gpu zero where oracle nonzero:", ((W == 0) & (Wo > 1e-6))[:, :, 1:].sum(), " gpu nonzero where oracle zero:", ((W != 0) & (Wo == 0))[:, :, 1:].sum())
# radial profile of the W error
zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Z), np.arange(Z//2+1), indexing='ij')
uz = np.where(zz <= Z//2, zz, zz - Z); uy = np.where(yy <= Z//2, yy, yy - Z)
rad = np.sqrt(xx**2 + uy**2 + uz**2)
for r0 in range(0, int(rad.max())+1, 2):
    m = (rad >= r0) & (rad < r0+2) & (xx > 0)
    if m.sum() == 0: continue
    print("r=[%d,%d) n=%d  maxdiff=%.2e  sumW=%.4f sumWo=%.4f  nbad=%d" % (r0, r0+2, m.sum(), dW[m].max(), W[m].sum(), Wo[m].sum(), (dW[m] > 1e-4).sum()))
