"""ctypes wrapper of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package xmipp3_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_recfourier.so")
_SRC = os.path.join(_HERE, "recfourier_oracle.cpp")


_SRC_FAST = os.path.join(_HERE, "recfourier_fast_oracle.cpp")


def build(force=False):
    """Compile the C++ restatements with g++ (no CUDA, no external libraries).  The --fast restatement is single
    precision and decision-sensitive, so its translation unit is built without a*b+c contraction."""
    deps = [_SRC, _SRC_FAST, os.path.join(_HERE, "oracle_abi.h")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps if os.path.exists(d))):
        return _SO
    if not os.path.exists(_SRC):
        if os.path.exists(_SO):
            return _SO
        raise RuntimeError("oracle source missing")
    base = ["g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-fPIC", "-pthread"]
    obj = [os.path.join(_HERE, "_oracle_main.o"), os.path.join(_HERE, "_oracle_fast.o")]
    # no a*b+c contraction in the exact restatement either: the reference is built for baseline x86-64 (no FMA), and the
    # race-free "slabs" mode is then bit-identical to the single-thread result
    subprocess.check_call(base + ["-ffp-contract=off", "-c", _SRC, "-o", obj[0]], cwd=_HERE)
    subprocess.check_call(base + ["-ffp-contract=off", "-c", _SRC_FAST, "-o", obj[1]], cwd=_HERE)
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", _SO] + obj, cwd=_HERE)
    for o in obj:
        os.remove(o)
    return _SO


class Config(C.Structure):
    _fields_ = [
        ("img_size", C.c_int32), ("n_sym", C.c_int32),
        ("pad_proj", C.c_double), ("pad_vol", C.c_double),
        ("max_resolution", C.c_double),
        ("blob_radius", C.c_double), ("blob_alpha", C.c_double),
        ("blob_order", C.c_int32), ("use_ctf", C.c_int32),
        ("sampling", C.c_double), ("min_ctf", C.c_double),
        ("phase_flipped", C.c_int32), ("use_weights", C.c_int32),
        ("n_iter_weight", C.c_int32), ("reserved", C.c_int32),
        ("sym_matrices", C.POINTER(C.c_double)),
    ]


PARTICLE_FIELDS = ["rot", "tilt", "psi", "shift_x", "shift_y", "weight",
                   "kV", "defocusU", "defocusV", "defocus_angle", "Cs", "Ca", "espr", "ispr", "alpha",
                   "DeltaF", "DeltaR", "Q0", "K", "envR0", "envR1", "envR2", "phase_shift", "vpp_radius"]
PARTICLE_DTYPE = np.dtype([(f, np.float64) for f in PARTICLE_FIELDS])


def make_particles(n, **cols):
    """Structured array of per-particle parameters with the reference's defaults
    (data/ctf.cpp:365-419: kV 100, K 1, everything else 0; weight 1)."""
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    p["weight"] = 1.0
    p["kV"] = 100.0
    p["K"] = 1.0
    for k, v in cols.items():
        p[k] = v
    if "defocusV" not in cols and "defocusU" in cols:
        p["defocusV"] = p["defocusU"]
    return p


SPACE_DTYPE = np.dtype([("minX", np.int32), ("minY", np.int32), ("minZ", np.int32), ("maxX", np.int32), ("maxY", np.int32),
                        ("maxZ", np.int32), ("dir", np.int32), ("projectionIndex", np.int32), ("maxDistanceSqr", np.float32),
                        ("unitNormal", np.float32, 3), ("topOrigin", np.float32, 3), ("bottomOrigin", np.float32, 3),
                        ("transformInv", np.float32, 9), ("weight", np.float32)])     # orf_space / refk_space

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orf_create.restype = C.c_void_p
        L.orf_create.argtypes = [C.POINTER(Config)]
        L.orf_destroy.argtypes = [C.c_void_p]
        L.orf_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
        L.orf_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orf_insert_slabs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orf_get_accumulators.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_add_accumulators.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_finalize.argtypes = [C.c_void_p, C.c_void_p]
        L.orf_tables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orf_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_apply_shift.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orf_ctf_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orf_bspline_coeffs_2d.argtypes = [C.c_void_p, C.c_int]
        L.orf_bspline_interp_2d.restype = C.c_double
        L.orf_bspline_interp_2d.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orf_ctf_value.restype = C.c_double
        L.orf_ctf_value.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orf_ctf_argument.restype = C.c_double
        L.orf_ctf_argument.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orf_ctf_K1.restype = C.c_double
        L.orf_ctf_K1.argtypes = [C.c_void_p]
        L.orf_ctf_grid.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        L.orf_euler.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p]
        L.orf_idx2digfreq.restype = C.c_double
        L.orf_idx2digfreq.argtypes = [C.c_int, C.c_int]
        for f in ("orf_kaiser_value", "orf_kaiser_fourier_value"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        for f in ("orf_bessi0", "orf_bessj0"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double]
        L.orf_fft2_r2c.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orf_fft1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orf_finish_fourier.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_projector_create.restype = C.c_void_p
        L.orf_projector_create.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
        L.orf_projector_destroy.argtypes = [C.c_void_p]
        L.orf_projector_project.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orf_fast_create.restype = C.c_void_p
        L.orf_fast_create.argtypes = [C.POINTER(Config)]
        L.orf_fast_create2.restype = C.c_void_p
        L.orf_fast_create2.argtypes = [C.POINTER(Config), C.c_int]
        L.orf_fast_export_buffer.restype = C.c_int
        L.orf_fast_export_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_fast_tables.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orf_fast_destroy.argtypes = [C.c_void_p]
        L.orf_fast_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.orf_fast_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orf_fast_get_temp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_fast_finalize.argtypes = [C.c_void_p, C.c_void_p]
        L.orf_fast_set_temp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orf_fast_fourier.argtypes = [C.c_void_p, C.c_void_p]
        L.orf_fast_half_spaces.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _make_config(img_size, padding, max_resolution, blob, sym_matrices, use_ctf, sampling, min_ctf, phase_flipped,
                 use_weights, n_iter_weight):
    sm = np.zeros((0, 9)) if sym_matrices is None else np.ascontiguousarray(sym_matrices, dtype=np.float64).reshape(-1, 9)
    cfg = Config()
    cfg.img_size = int(img_size)
    cfg.n_sym = sm.shape[0]
    cfg.pad_proj, cfg.pad_vol = float(padding[0]), float(padding[1])
    cfg.max_resolution = float(max_resolution)
    cfg.blob_radius, cfg.blob_order, cfg.blob_alpha = float(blob[0]), int(blob[1]), float(blob[2])
    cfg.use_ctf = int(bool(use_ctf))
    cfg.sampling = float(sampling)
    cfg.min_ctf = float(min_ctf)
    cfg.phase_flipped = int(bool(phase_flipped))
    cfg.use_weights = int(bool(use_weights))
    cfg.n_iter_weight = int(n_iter_weight)
    cfg.sym_matrices = sm.ctypes.data_as(C.POINTER(C.c_double)) if sm.size else None
    return cfg, sm


class ProjectorOracle:
    """CPU restatement of FourierProjector (data/fourier_projection.cpp): padded 3-D transform of a volume, central
    slices by NEAREST (0) / LINEAR (1) / BSPLINE3 (3) interpolation, inverse 2-D transform."""

    def __init__(self, volume, padding=2.0, max_freq=0.5, degree=3):
        self._L = lib()
        v = np.ascontiguousarray(volume, dtype=np.float32)
        assert v.ndim == 3 and v.shape[0] == v.shape[1] == v.shape[2]
        self.N = v.shape[0]
        self._h = self._L.orf_projector_create(_ptr(v), self.N, float(padding), float(max_freq), int(degree))
        if not self._h:
            raise RuntimeError("orf_projector_create failed")

    def __del__(self):
        try:
            if self._h:
                self._L.orf_projector_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def project(self, rot, tilt, psi, ctf=None):
        out = np.empty((self.N, self.N), dtype=np.float64)
        c = None
        if ctf is not None:
            c = np.ascontiguousarray(ctf, dtype=np.float64)
            assert c.shape == (self.N, self.N // 2 + 1)
        self._L.orf_projector_project(self._h, float(rot), float(tilt), float(psi), _ptr(c) if c is not None else None, _ptr(out))
        return out


class FastOracle:
    """CPU restatement of the --fast arithmetic (ProgRecFourierGPU with useFast: nearest-pixel insertion in single
    precision + one final 3-D blob convolution); see recfourier_fast_oracle.cpp."""

    def __init__(self, img_size, padding=(2.0, 2.0), max_resolution=0.5, blob=(1.9, 0, 15.0),
                 sym_matrices=None, use_ctf=False, sampling=1.0, min_ctf=0.01, phase_flipped=False, use_weights=False,
                 use_fast=True):
        """use_fast=False: host side of the GPU program WITHOUT --fast (buffers with blob-thick traverse spaces, no final
        blob convolution); the device arithmetic of that mode is not restated, insert() is then unavailable and the temporary
        spaces come from the compiled reference kernel (oracle/ref_kernel.py) through set_temp_spaces()."""
        self._L = lib()
        self._cfg, self._sym = _make_config(img_size, padding, max_resolution, blob, sym_matrices, use_ctf, sampling,
                                            min_ctf, phase_flipped, use_weights, 1)
        self.use_fast = bool(use_fast)
        self.use_ctf = bool(use_ctf)
        self.n_sym = 1 + (0 if sym_matrices is None else len(np.asarray(sym_matrices).reshape(-1, 9)))
        self.blob = tuple(blob)
        self.max_resolution = float(max_resolution)
        self._h = self._L.orf_fast_create2(C.byref(self._cfg), int(self.use_fast))
        if not self._h:
            raise RuntimeError("orf_fast_create failed")
        v = [C.c_int() for _ in range(4)]
        self._L.orf_fast_dims(self._h, *[C.byref(x) for x in v])
        self.S, self.sx, self.sy, self.Pv = [x.value for x in v]
        self.N = int(img_size)

    def __del__(self):
        try:
            if self._h:
                self._L.orf_fast_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def insert(self, images, particles):
        assert self.use_fast, "only the --fast device arithmetic is restated on the CPU"
        images = np.ascontiguousarray(images, dtype=np.float32)
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        assert images.shape == (len(particles), self.N, self.N)
        self._L.orf_fast_insert(self._h, _ptr(images), _ptr(particles), len(particles))

    def export_buffer(self, images, particles):
        """The RecFourierBufferData contents ProgRecFourierGPU::prepareBuffer builds for these images
        (reconstruct_fourier_gpu.cpp:323-415): (FFTs [k, sy, sx] complex64, CTFs, modulators [k, sy, sx] float32 or None,
        spaces [k * n_sym] SPACE_DTYPE), k = images kept."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        n = len(particles)
        assert images.shape == (n, self.N, self.N)
        F = np.zeros((n, self.sy, self.sx), dtype=np.complex64)
        Cc = np.zeros((n, self.sy, self.sx), dtype=np.float32)
        M = np.zeros((n, self.sy, self.sx), dtype=np.float32)
        sp = np.zeros(n * self.n_sym, dtype=SPACE_DTYPE)
        k = self._L.orf_fast_export_buffer(self._h, _ptr(images), _ptr(particles), n, _ptr(F), _ptr(Cc), _ptr(M), _ptr(sp))
        if not self.use_ctf:
            Cc = M = None
        else:
            Cc, M = Cc[:k], M[:k]
        return F[:k], Cc, M, sp[:k * self.n_sym]

    def tables(self):
        t = np.empty(10000, dtype=np.float32)
        ids, iw0 = C.c_float(), C.c_float()
        self._L.orf_fast_tables(self._h, _ptr(t), C.byref(ids), C.byref(iw0))
        return t, ids.value, iw0.value

    def temp_spaces(self):
        n = self.S + 1
        V = np.empty((n, n, n), dtype=np.complex64)
        W = np.empty((n, n, n), dtype=np.float32)
        self._L.orf_fast_get_temp(self._h, _ptr(V), _ptr(W))
        return V, W

    def set_temp_spaces(self, V, W):
        """Replace the temporary spaces (e.g. by the ones a GPU run produced) before finalize()."""
        n = self.S + 1
        V = np.ascontiguousarray(V, dtype=np.complex64)
        W = np.ascontiguousarray(W, dtype=np.float32)
        assert V.shape == (n, n, n) and W.shape == (n, n, n)
        self._L.orf_fast_set_temp(self._h, _ptr(V), _ptr(W))

    def half_spaces(self):
        """mirrorAndCrop of the temporary spaces (reconstruct_fourier_gpu.cpp:697-730): V, W on [S+1, S+1, S/2+1], last axis =
        centred x >= 0, before symmetrisation and weighting."""
        n, X = self.S + 1, self.S // 2 + 1
        V = np.empty((n, n, X), dtype=np.complex64)
        W = np.empty((n, n, X), dtype=np.float32)
        self._L.orf_fast_half_spaces(self._h, _ptr(V), _ptr(W))
        return V, W

    def fourier(self):
        """The Fourier volume handed to the inverse transform (after blob convolution, symmetrisation and weights)."""
        out = np.empty((self.Pv, self.Pv, self.Pv // 2 + 1), dtype=np.complex128)
        self._L.orf_fast_fourier(self._h, _ptr(out))
        return out

    def finalize(self):
        out = np.empty((self.N,) * 3, dtype=np.float64)
        self._L.orf_fast_finalize(self._h, _ptr(out))
        return out


class Oracle:
    """CPU ProgRecFourier restatement.  Parameters mirror the reference CLI (RF.cpp:42-58)."""

    def __init__(self, img_size, padding=(2.0, 2.0), max_resolution=0.5, blob=(1.9, 0, 15.0),
                 sym_matrices=None, use_ctf=False, sampling=1.0, min_ctf=0.01, phase_flipped=False,
                 use_weights=False, n_iter_weight=1):
        self._L = lib()
        sm = np.zeros((0, 9)) if sym_matrices is None else np.ascontiguousarray(sym_matrices, dtype=np.float64).reshape(-1, 9)
        self._sym = sm
        cfg = Config()
        cfg.img_size = int(img_size)
        cfg.n_sym = sm.shape[0]
        cfg.pad_proj, cfg.pad_vol = float(padding[0]), float(padding[1])
        cfg.max_resolution = float(max_resolution)
        cfg.blob_radius, cfg.blob_order, cfg.blob_alpha = float(blob[0]), int(blob[1]), float(blob[2])
        cfg.use_ctf = int(bool(use_ctf))
        cfg.sampling = float(sampling)
        cfg.min_ctf = float(min_ctf)
        cfg.phase_flipped = int(bool(phase_flipped))
        cfg.use_weights = int(bool(use_weights))
        cfg.n_iter_weight = int(n_iter_weight)
        cfg.sym_matrices = sm.ctypes.data_as(C.POINTER(C.c_double)) if sm.size else None
        self._cfg = cfg
        self._h = self._L.orf_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("orf_create failed")
        n, p, z = C.c_int(), C.c_int(), C.c_int()
        self._L.orf_dims(self._h, C.byref(n), C.byref(p), C.byref(z))
        self.N, self.P, self.Z = n.value, p.value, z.value

    def __del__(self):
        try:
            if self._h:
                self._L.orf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def insert(self, images, particles, threads=1, scheme="reference"):
        """scheme "reference": the reference's thread scheme (rows handed out under a mutex; exact for threads = 1, loses
        updates with >= 3 threads like the reference).  scheme "slabs": race-free parallel mode, every slab of the volume
        owned by one task; bit-identical to threads = 1 for any thread count (used for the large parity cases)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        assert images.shape == (len(particles), self.N, self.N)
        if scheme == "slabs":
            self._L.orf_insert_slabs(self._h, _ptr(images), _ptr(particles), len(particles), int(threads))
        else:
            assert scheme == "reference"
            self._L.orf_insert(self._h, _ptr(images), _ptr(particles), len(particles), int(threads))

    def accumulators(self):
        Z, X = self.Z, self.Z // 2 + 1
        V = np.empty((Z, Z, X), dtype=np.complex128)
        W = np.empty((Z, Z, X), dtype=np.float64)
        self._L.orf_get_accumulators(self._h, _ptr(V), _ptr(W))
        return V, W

    def add_accumulators(self, V, W):
        V = np.ascontiguousarray(V, dtype=np.complex128)
        W = np.ascontiguousarray(W, dtype=np.float64)
        self._L.orf_add_accumulators(self._h, _ptr(V), _ptr(W))

    def finalize(self):
        out = np.empty((self.N,) * 3, dtype=np.float64)
        self._L.orf_finalize(self._h, _ptr(out))
        return out

    def tables(self):
        a = np.empty(10000)
        b = np.empty(10000)
        d1, d2 = C.c_double(), C.c_double()
        self._L.orf_tables(self._h, _ptr(a), _ptr(b), C.byref(d1), C.byref(d2))
        return a, b, d1.value, d2.value

    def preprocess(self, image, particle):
        image = np.ascontiguousarray(image, dtype=np.float32)
        particle = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
        F = np.empty((self.P, self.P // 2 + 1), dtype=np.complex128)
        A = np.empty((3, 3))
        self._L.orf_preprocess(self._h, _ptr(image), _ptr(particle), _ptr(F), _ptr(A))
        return F, A

    def apply_shift(self, image, sx, sy):
        image = np.ascontiguousarray(image, dtype=np.float32)
        out = np.empty((self.N, self.N))
        self._L.orf_apply_shift(self._h, _ptr(image), float(sx), float(sy), _ptr(out))
        return out

    def ctf_weights(self, particle, i, j):
        particle = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
        a, b = C.c_double(), C.c_double()
        self._L.orf_ctf_weights(self._h, _ptr(particle), int(i), int(j), C.byref(a), C.byref(b))
        return a.value, b.value


def ctf_value(particle, X, Y):
    particle = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
    return lib().orf_ctf_value(_ptr(particle), float(X), float(Y))


def bspline_coeffs_2d(a):
    c = np.array(a, dtype=np.float64, order="C")
    assert c.ndim == 2 and c.shape[0] == c.shape[1]
    lib().orf_bspline_coeffs_2d(_ptr(c), c.shape[0])
    return c


def bspline_interp_2d(c, x, y):
    c = np.ascontiguousarray(c, dtype=np.float64)
    return lib().orf_bspline_interp_2d(_ptr(c), c.shape[0], float(x), float(y))


def ctf_grid(particle, n, Tm, what="value"):
    """CTF values ("value") or sine arguments ("argument") on the n x n FFT grid at sampling Tm."""
    particle = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
    out = np.empty((n, n))
    lib().orf_ctf_grid(_ptr(particle), int(n), float(Tm), 1 if what == "argument" else 0, _ptr(out))
    return out


def ctf_K1(particle):
    particle = np.ascontiguousarray(particle, dtype=PARTICLE_DTYPE).reshape(1)
    return lib().orf_ctf_K1(_ptr(particle))


def euler(rot, tilt, psi):
    m = np.empty((3, 3))
    lib().orf_euler(float(rot), float(tilt), float(psi), _ptr(m))
    return m


def fft2_r2c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty((a.shape[0], a.shape[1] // 2 + 1), dtype=np.complex128)
    lib().orf_fft2_r2c(_ptr(a), a.shape[0], a.shape[1], _ptr(out))
    return out


def fft1(a, sign=-1):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    out = np.empty_like(a)
    lib().orf_fft1(_ptr(a), a.shape[0], int(sign), _ptr(out))
    return out
