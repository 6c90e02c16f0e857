"""Pure-numpy mini oracle for tiny sizes (TEST INFRASTRUCTURE ONLY).

An independent restatement of ProgRecFourier (RF.cpp) written against numpy.fft
and scipy.special instead of our own FFT / polynomial Bessels, used to cross-check
oracle/recfourier_oracle.cpp.  Pure-Python loops: use box <= 16.
No CTF here (the C++ oracle's CTF is checked against a separate numpy formula).
"""
import numpy as np
from scipy import special


def _kaiser_value(r, a, alpha, m=0):
    if r / a > 1:
        return 0.0
    arg = alpha * np.sqrt(1 - (r / a) ** 2)
    assert m == 0
    return special.i0(arg) / special.i0(alpha)


def _kaiser_fourier(w, a, alpha, m=0):
    assert m == 0
    t = 2 * np.pi * a * w
    sigma = np.sqrt(abs(alpha * alpha - t * t))
    if t > alpha:
        b = special.jv(1.5, sigma)
    else:
        b = special.iv(1.5, sigma)
    return (2 * np.pi) ** 1.5 * a ** 3 * b / (special.i0(alpha) * sigma ** 1.5)


def euler(rot, tilt, psi):
    # ZYZ: A = Rz(psi) * Ry(tilt) * Rz(rot) in xmipp's passive convention
    a, b, g = np.radians([rot, tilt, psi])

    def rz(t):
        return np.array([[np.cos(t), np.sin(t), 0], [-np.sin(t), np.cos(t), 0], [0, 0, 1]])

    def ry(t):
        return np.array([[np.cos(t), 0, -np.sin(t)], [0, 1, 0], [np.sin(t), 0, np.cos(t)]])

    return rz(g) @ ry(b) @ rz(a)


def reconstruct(images, rot, tilt, psi, pad=2.0, max_res=0.5, blob=(1.9, 0, 15.0), sym=(), weights=None,
                return_accumulators=False):
    n, N, _ = images.shape
    P = int(N * pad)
    Z = int(N * pad)
    X = Z // 2 + 1
    r, m, alpha = blob
    T = 10000
    iw0 = 1.0 / _kaiser_fourier(0.0, r, alpha, m)            # pad_proj == pad_vol
    tab = np.array([_kaiser_value(r * np.sqrt(i / (T - 1)), r, alpha, m) * iw0 for i in range(T)])
    idelta = (T - 1) / (r * r)
    dF = (np.sqrt(3.0) * N / 2) / (T - 1)
    ftab = np.array([_kaiser_fourier(dF * i, r / (pad * N), alpha, m) * (pad * N) ** 3 * iw0 for i in range(T)])
    V = np.zeros((Z, Z, X), dtype=np.complex128)
    W = np.zeros((Z, Z, X))
    Rs = [np.eye(3)] + [np.asarray(s) for s in sym]
    for k in range(n):
        padded = np.zeros((P, P))
        lo = -(N // 2)
        idx = (np.arange(N) + lo) % P
        padded[np.ix_(idx, idx)] = images[k]
        F = np.fft.rfft2(padded) / (P * P)
        Ainv = euler(rot[k], tilt[k], psi[k]).T
        wgt = 1.0 if weights is None else weights[k]
        for Rm in Rs:
            M = Rm @ Ainv
            for i in range(P):
                fy = (i if i <= P // 2 else i - P) / P
                for j in range(P // 2 + 1):
                    fx = j / P
                    if fx * fx + fy * fy > max_res * max_res:
                        continue
                    p = Z * (M @ np.array([fx, fy, 0.0]))
                    c1 = np.ceil(p - r).astype(int)
                    c2 = np.floor(p + r).astype(int)
                    for iz in range(c1[2], c2[2] + 1):
                        for iy in range(c1[1], c2[1] + 1):
                            for ix in range(c1[0], c2[0] + 1):
                                d2 = (ix - p[0]) ** 2 + (iy - p[1]) ** 2 + (iz - p[2]) ** 2
                                if d2 > r * r:
                                    continue
                                w = tab[int(d2 * idelta + 0.5)] * wgt
                                wx = ix % Z
                                if wx > Z // 2:
                                    V[(-iz) % Z, (-iy) % Z, (-wx) % Z] += w * np.conj(F[i, j])
                                    W[(-iz) % Z, (-iy) % Z, (-wx) % Z] += w
                                else:
                                    V[iz % Z, iy % Z, wx] += w * F[i, j]
                                    W[iz % Z, iy % Z, wx] += w
    if return_accumulators:
        return V, W
    # weights
    Ws = W.copy()
    Vs = V.copy()
    yh = Z // 2 - 1 if Z % 2 == 0 else Z // 2
    for kk in range(Z):
        for ii in range(1, yh + 1):
            mw = 0.5 * (Ws[kk, ii, 0] + Ws[-kk % Z, -ii % Z, 0])
            Ws[kk, ii, 0] = Ws[-kk % Z, -ii % Z, 0] = mw
            mv = 0.5 * (Vs[kk, ii, 0] + np.conj(Vs[-kk % Z, -ii % Z, 0]))
            Vs[kk, ii, 0] = mv
            Vs[-kk % Z, -ii % Z, 0] = np.conj(mv)
    for kk in range(1, yh + 1):
        mw = 0.5 * (Ws[kk, 0, 0] + Ws[-kk % Z, 0, 0])
        Ws[kk, 0, 0] = Ws[-kk % Z, 0, 0] = mw
        mv = 0.5 * (Vs[kk, 0, 0] + np.conj(Vs[-kk % Z, 0, 0]))
        Vs[kk, 0, 0] = mv
        Vs[-kk % Z, 0, 0] = np.conj(mv)
    with np.errstate(divide="ignore"):
        Winv = np.where(np.abs(Ws) > 1e-3, 1.0 / np.where(Ws == 0, 1, Ws), V.real)
        keep = (1.0 / Winv) > 1e-3
    corr = pad ** 2 / (N * pad ** 3)
    G = np.where(keep, Vs * corr * Winv, 0)
    vol = np.fft.irfftn(G, s=(Z, Z, Z), axes=(0, 1, 2)) * Z ** 3          # unnormalised backward transform
    g = np.arange(N) + (-(N // 2))
    sub = vol[np.ix_(g % Z, g % Z, g % Z)]
    kk, ii, jj = np.meshgrid(g, g, g, indexing="ij")
    R = np.sqrt(kk * kk + ii * ii + jj * jj)
    fac = ftab[np.rint(R / dF).astype(int)]
    f2 = np.sinc(R / (2 * N)) ** 2
    out = sub / (f2 * fac)
    out *= f2.mean()
    return out
