"""Synthetic inputs for the direct-Fourier reconstruction path.

Analytic Gaussian phantoms, their exact projections (closed form), optional CTF
multiplication in Fourier space and integer shifts.  Pure numpy: this is data
preparation for tests and bench.py, not part of the hot path.

Geometry follows the reference: projection coordinates = A·(x,y,z)^T with
A = Euler_angles2matrix(rot, tilt, psi) (data/fourier_projection.h:66-73), image
pixel [i][j] <-> (x' = j - N/2, y' = i - N/2).
"""
import numpy as np

from .geometry import euler_matrix, point_group_matrices  # noqa: F401  (re-export)


def make_phantom(n_gauss=30, box=64, seed=0, sym=None):
    """Random isotropic Gaussians inside a ball of radius 0.30*box.

    Returns (centres[n,3] as (x,y,z), sigmas[n], amps[n]).  With ``sym`` (e.g. 'd7')
    every seed Gaussian is replicated under the whole point group, so the phantom is
    exactly symmetric.
    """
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(n_gauss, 3))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    c *= (0.30 * box) * rng.uniform(0, 1, size=(n_gauss, 1)) ** (1.0 / 3.0)
    s = rng.uniform(0.02 * box, 0.06 * box, size=n_gauss)
    a = rng.uniform(0.5, 1.5, size=n_gauss)
    if sym is not None and sym.lower() != "c1":
        mats = [np.eye(3)] + list(point_group_matrices(sym))
        c = np.concatenate([c @ m.T for m in mats], axis=0)
        s = np.tile(s, len(mats))
        a = np.tile(a, len(mats))
    return c, s, a


def phantom_volume(phantom, box):
    """Sample the phantom on the box^3 grid (logical origin at box//2); [z][y][x]."""
    c, s, a = phantom
    g = np.arange(box) - box // 2
    vol = np.zeros((box, box, box))
    for (cx, cy, cz), sg, am in zip(c, s, a):
        ex = np.exp(-0.5 * ((g - cx) / sg) ** 2)
        ey = np.exp(-0.5 * ((g - cy) / sg) ** 2)
        ez = np.exp(-0.5 * ((g - cz) / sg) ** 2)
        vol += am * ez[:, None, None] * ey[None, :, None] * ex[None, None, :]
    return vol


def random_orientations(n, seed=1):
    rng = np.random.default_rng(seed)
    rot = rng.uniform(0, 360, n)
    tilt = np.degrees(np.arccos(rng.uniform(-1, 1, n)))
    psi = rng.uniform(0, 360, n)
    return rot, tilt, psi


def project(phantom, box, rot, tilt, psi, dtype=np.float32, chunk=256):
    """Exact projections of the Gaussian phantom: images[n, box, box]."""
    c, s, a = phantom
    n = len(rot)
    g = (np.arange(box) - box // 2).astype(np.float64)
    out = np.empty((n, box, box), dtype=dtype)
    amp2d = a * s * np.sqrt(2 * np.pi)
    for b0 in range(0, n, chunk):
        b1 = min(n, b0 + chunk)
        A = np.stack([euler_matrix(rot[k], tilt[k], psi[k]) for k in range(b0, b1)])  # [b,3,3]
        pc = np.einsum("bij,gj->bgi", A, c)           # projected centres (x', y', z')
        gx = np.exp(-0.5 * ((g[None, None, :] - pc[:, :, 0:1]) / s[None, :, None]) ** 2)
        gy = np.exp(-0.5 * ((g[None, None, :] - pc[:, :, 1:2]) / s[None, :, None]) ** 2)
        out[b0:b1] = np.einsum("bgi,bgj->bij", gy * amp2d[None, :, None], gx).astype(dtype)
    return out


def ctf_2d(box, sampling, kV, defocusU, defocusV, defocus_angle, Cs, Q0):
    """Pure CTF (K=1, no envelope) on the box x box FFT grid; formula of
    data/ctf.h:452-496 / ctf.cpp:645-680 restated in numpy for data synthesis."""
    f = np.fft.fftfreq(box)
    X = f[None, :] / sampling
    Y = f[:, None] / sampling
    lam = 12.2643247 / np.sqrt(kV * 1e3 * (1.0 + 0.978466e-6 * kV * 1e3))
    K1 = np.pi * lam
    K2 = np.pi / 2 * Cs * 1e7 * lam ** 3
    u2 = X * X + Y * Y
    ang = np.arctan2(Y, X)
    davg = -(defocusU + defocusV) * 0.5
    ddev = -(defocusU - defocusV) * 0.5
    deltaf = davg + ddev * np.cos(2 * (ang - np.radians(defocus_angle)))
    deltaf = np.where((np.abs(X) < 1e-6) & (np.abs(Y) < 1e-6), 0.0, deltaf)
    arg = K1 * deltaf * u2 + K2 * u2 * u2
    return -(np.sqrt(1 - Q0 * Q0) * np.sin(arg) - Q0 * np.cos(arg))


def apply_ctf(images, sampling, kV, defocusU, defocusV, defocus_angle, Cs, Q0):
    """Multiply every image by its CTF in box x box Fourier space."""
    n, box, _ = images.shape
    out = np.empty_like(images)
    for k in range(n):
        c = ctf_2d(box, sampling, kV[k], defocusU[k], defocusV[k], defocus_angle[k], Cs[k], Q0[k])
        out[k] = np.fft.ifft2(np.fft.fft2(images[k].astype(np.float64)) * c).real.astype(images.dtype)
    return out


def random_ctf_params(n, seed=3):
    rng = np.random.default_rng(seed)
    dU = rng.uniform(10000, 30000, n)
    dV = dU + rng.uniform(-500, 500, n)
    ang = rng.uniform(0, 180, n)
    return dict(kV=np.full(n, 300.0), defocusU=dU, defocusV=dV, defocus_angle=ang,
                Cs=np.full(n, 2.7), Q0=np.full(n, 0.07))


def preshift(images, shift_x, shift_y):
    """Shift image content by -shift so that the metadata shift re-centres it."""
    out = np.empty_like(images)
    for k in range(images.shape[0]):
        out[k] = np.roll(images[k], (-int(shift_y[k]), -int(shift_x[k])), axis=(0, 1))
    return out


def make_dataset(n, box, seed=0, sym=None, ctf=False, shifts=False, sampling=1.5, n_gauss=30):
    """Images + per-particle parameter columns for one synthetic config.

    Returns dict(images, rot, tilt, psi, shift_x, shift_y, ctf=dict|None, phantom).
    """
    ph = make_phantom(n_gauss=n_gauss if sym in (None, "c1") else 3, box=box, seed=seed, sym=sym)
    rot, tilt, psi = random_orientations(n, seed + 1)
    img = project(ph, box, rot, tilt, psi)
    d = dict(rot=rot, tilt=tilt, psi=psi, phantom=ph, shift_x=np.zeros(n), shift_y=np.zeros(n), ctf=None,
             sampling=sampling)
    if ctf:
        cp = random_ctf_params(n, seed + 3)
        img = apply_ctf(img, sampling, **cp)
        d["ctf"] = cp
    if shifts:
        rng = np.random.default_rng(seed + 2)
        sx = rng.integers(-5, 6, n).astype(np.float64)
        sy = rng.integers(-5, 6, n).astype(np.float64)
        img = preshift(img, sx, sy)
        d["shift_x"], d["shift_y"] = sx, sy
    d["images"] = np.ascontiguousarray(img, dtype=np.float32)
    return d


def fsc(a, b):
    """Fourier shell correlation between two cubic maps, shells 1..box//2 (numpy restatement;
    the reference's implementation is frc_dpr in xmippCore, used at resolution_fsc.cpp:188)."""
    n = a.shape[0]
    fa = np.fft.fftn(a)
    fb = np.fft.fftn(b)
    f = np.fft.fftfreq(n) * n
    r = np.sqrt(f[:, None, None] ** 2 + f[None, :, None] ** 2 + f[None, None, :] ** 2)
    shell = np.rint(r).astype(np.int64).ravel()
    nsh = n // 2 + 1
    num = np.bincount(shell, weights=(fa * np.conj(fb)).real.ravel(), minlength=nsh)[:nsh]
    da = np.bincount(shell, weights=(np.abs(fa) ** 2).ravel(), minlength=nsh)[:nsh]
    db = np.bincount(shell, weights=(np.abs(fb) ** 2).ravel(), minlength=nsh)[:nsh]
    with np.errstate(invalid="ignore", divide="ignore"):
        return num / np.sqrt(da * db)


def rel_l2(a, b):
    """||a-b|| / ||b||"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
