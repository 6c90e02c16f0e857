"""ctypes binding of include/recfourier_b200.h (the CUDA library).

There is no CPU fallback: if the shared object is missing it is built with nvcc,
and if no CUDA device is present `Reconstructor()` raises with the library's own
error message.
"""
import ctypes as C
import os

import numpy as np

from . import _build

ABI_VERSION = 1

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5

EXPORTED_SYMBOLS = [
    "rfb200_create", "rfb200_destroy", "rfb200_last_error", "rfb200_get_info",
    "rfb200_insert_batch", "rfb200_insert_batch_device", "rfb200_sync", "rfb200_reset",
    "rfb200_nccl_unique_id", "rfb200_nccl_init", "rfb200_reduce_nccl", "rfb200_accumulator_ptrs",
    "rfb200_ipc_export", "rfb200_ipc_import", "rfb200_reduce_p2p", "rfb200_ipc_release",
    "rfb200_set_ranks", "rfb200_reduce_p2p_prepare", "rfb200_reduce_p2p_run",
    "rfb200_export_accumulators", "rfb200_finalize", "rfb200_get_timings",
    "rfb200_halfset_push", "rfb200_halfset_merge", "rfb200_timer_start", "rfb200_timer_stop", "rfb200_weight_sum", "rfb200_get_streams",
    "rfb200_debug_slice_dims", "rfb200_debug_get_slice", "rfb200_weight_sum_begin", "rfb200_weight_sum_end",
    "rfb200_host_alloc", "rfb200_host_free", "rfb200_device_count", "rfb200_measure_fp32_peak", "rfb200_warmup", "rfb200_debug_fast_fourier",
    "rfb200_projector_create", "rfb200_projector_project", "rfb200_projector_project_device", "rfb200_projector_last_error",
    "rfb200_projector_destroy",
]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("img_size", C.c_int32),
        ("pad_proj", C.c_double), ("pad_vol", C.c_double), ("max_resolution", C.c_double),
        ("blob_radius", C.c_double), ("blob_alpha", C.c_double),
        ("blob_order", C.c_int32), ("n_sym", C.c_int32),
        ("sym_matrices", C.POINTER(C.c_double)),
        ("use_ctf", C.c_int32), ("phase_flipped", C.c_int32),
        ("sampling", C.c_double), ("min_ctf", C.c_double),
        ("use_weights", C.c_int32), ("n_iter_weight", C.c_int32),
        ("fast", C.c_int32), ("device", C.c_int32),
        ("max_batch", C.c_int32), ("reserved0", C.c_int32),
    ]


class Timings(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("h2d_ms", "preprocess_ms", "fft2d_ms", "slice_ms", "gather_ms", "edge_ms",
                                           "finalize_ms", "reduce_ms")] + \
               [(n, C.c_int64) for n in ("images", "planes", "gather_launches", "kernel_launches")]


class Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "P", "Z", "X", "tiles_x", "tiles_y", "tiles_z", "tile")] + \
               [("n_blocked", C.c_int64)] + \
               [(n, C.c_int32) for n in ("chunk_images", "n_tiles_active", "n_edge_items")]


PARTICLE_FIELDS = ["rot", "tilt", "psi", "shift_x", "shift_y", "weight",
                   "kV", "defocusU", "defocusV", "defocus_angle", "Cs", "Ca", "espr", "ispr", "alpha",
                   "DeltaF", "DeltaR", "Q0", "K", "envR0", "envR1", "envR2", "phase_shift", "vpp_radius"]
PARTICLE_DTYPE = np.dtype([(f, np.float64) for f in PARTICLE_FIELDS])


def make_particles(n, **cols):
    """Structured array of metadata rows with the reference's column defaults
    (data/ctf.cpp:365-419: kV 100, K 1, defocusV = defocusU, rest 0; weight 1)."""
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    p["weight"] = 1.0
    p["kV"] = 100.0
    p["K"] = 1.0
    for k, v in cols.items():
        p[k] = v
    if "defocusV" not in cols and "defocusU" in cols:
        p["defocusV"] = p["defocusU"]
    return p


_lib = None


def load(build=True):
    """Load librecfourier_b200.so (building it first if needed) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_CUDA
    if os.environ.get("RFB200_LIB"):          # developer override: A/B-test an alternative build of the same sources
        path, build = os.environ["RFB200_LIB"], False
    if build:
        path = _build.build_cuda()
    if not os.path.exists(path):
        raise RuntimeError("librecfourier_b200.so is missing: the CUDA extension must be built (xmipp3_b200._build.build_cuda)")
    L = C.CDLL(path)
    H = C.c_void_p
    L.rfb200_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.rfb200_destroy.argtypes = [H]
    L.rfb200_destroy.restype = None
    L.rfb200_last_error.argtypes = [H]
    L.rfb200_last_error.restype = C.c_char_p
    L.rfb200_get_info.argtypes = [H, C.POINTER(Info)]
    L.rfb200_insert_batch.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32]
    L.rfb200_insert_batch_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32]
    L.rfb200_sync.argtypes = [H]
    L.rfb200_reset.argtypes = [H]
    L.rfb200_nccl_unique_id.argtypes = [C.c_void_p]
    L.rfb200_nccl_init.argtypes = [H, C.c_void_p, C.c_int32, C.c_int32]
    L.rfb200_reduce_nccl.argtypes = [H, C.c_int32]
    L.rfb200_ipc_export.argtypes = [H, C.c_void_p]
    L.rfb200_ipc_import.argtypes = [H, C.c_int32, C.c_void_p]
    L.rfb200_reduce_p2p.argtypes = [H, C.c_int32]
    L.rfb200_ipc_release.argtypes = [H]
    L.rfb200_set_ranks.argtypes = [H, C.c_int32, C.c_int32]
    L.rfb200_reduce_p2p_prepare.argtypes = [H]
    L.rfb200_reduce_p2p_run.argtypes = [H, C.c_int32]
    L.rfb200_accumulator_ptrs.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.rfb200_export_accumulators.argtypes = [H, C.c_void_p, C.c_void_p]
    L.rfb200_finalize.argtypes = [H, C.c_void_p]
    L.rfb200_warmup.argtypes = [H]
    L.rfb200_get_timings.argtypes = [H, C.POINTER(Timings)]
    L.rfb200_halfset_push.argtypes = [H]
    L.rfb200_halfset_merge.argtypes = [H]
    L.rfb200_timer_start.argtypes = [H]
    L.rfb200_timer_stop.argtypes = [H, C.POINTER(C.c_double)]
    L.rfb200_weight_sum.argtypes = [H, C.POINTER(C.c_double)]
    L.rfb200_weight_sum_begin.argtypes = [H]
    L.rfb200_weight_sum_end.argtypes = [H, C.POINTER(C.c_double)]
    L.rfb200_get_streams.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.rfb200_debug_slice_dims.argtypes = [H, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.rfb200_debug_get_slice.argtypes = [H, C.c_int32, C.c_void_p]
    L.rfb200_debug_fast_fourier.argtypes = [H, C.c_void_p]
    L.rfb200_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    L.rfb200_projector_create.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.POINTER(H)]
    L.rfb200_projector_project.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.rfb200_projector_project_device.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.rfb200_projector_last_error.argtypes = [H]
    L.rfb200_projector_last_error.restype = C.c_char_p
    L.rfb200_projector_destroy.argtypes = [H]
    L.rfb200_projector_destroy.restype = None
    _lib = L
    return L


def measure_fp32_peak(device=0):
    """Measured FP32 SIMT peak (TFLOP/s) of a device: FFMA chain micro-benchmark inside the library."""
    v = C.c_double()
    rc = load().rfb200_measure_fp32_peak(int(device), C.byref(v))
    if rc != OK:
        raise RecFourierError(rc, "rfb200_measure_fp32_peak failed")
    return v.value


class RecFourierError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rfb200 error %d: %s" % (code, msg))
        self.code = code


class Reconstructor:
    """Thin object wrapper over one rfb200 handle (one GPU, one private V/W accumulator pair).

    Parameters mirror the reference CLI (reconstruct_fourier.cpp:42-58).
    """

    def __init__(self, img_size, padding=(2.0, 2.0), max_resolution=0.5, blob=(1.9, 0, 15.0), sym_matrices=None,
                 use_ctf=False, sampling=1.0, min_ctf=0.01, phase_flipped=False, use_weights=False, n_iter_weight=1,
                 fast=False, device=0, max_batch=0):
        self._L = load()
        sm = np.zeros((0, 9)) if sym_matrices is None else np.ascontiguousarray(sym_matrices, dtype=np.float64).reshape(-1, 9)
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.img_size = int(img_size)
        cfg.pad_proj, cfg.pad_vol = float(padding[0]), float(padding[1])
        cfg.max_resolution = float(max_resolution)
        cfg.blob_radius, cfg.blob_order, cfg.blob_alpha = float(blob[0]), int(blob[1]), float(blob[2])
        cfg.n_sym = sm.shape[0]
        cfg.sym_matrices = sm.ctypes.data_as(C.POINTER(C.c_double)) if sm.size else None
        cfg.use_ctf = int(bool(use_ctf))
        cfg.phase_flipped = int(bool(phase_flipped))
        cfg.sampling = float(sampling)
        cfg.min_ctf = float(min_ctf)
        cfg.use_weights = int(bool(use_weights))
        cfg.n_iter_weight = int(n_iter_weight)
        cfg.fast = int(bool(fast))
        cfg.device = int(device)
        cfg.max_batch = int(max_batch)
        self._h = C.c_void_p()
        rc = self._L.rfb200_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            msg = self._L.rfb200_last_error(None)
            self._h = None
            raise RecFourierError(rc, (msg or b"").decode())
        info = Info()
        self._check(self._L.rfb200_get_info(self._h, C.byref(info)))
        self.info = info
        self.N, self.P, self.Z = info.N, info.P, info.Z
        self.fast = bool(fast)

    def _check(self, rc):
        if rc != OK:
            raise RecFourierError(rc, (self._L.rfb200_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.rfb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- insertion
    def insert(self, images, particles):
        """images: [n, N, N] float32 numpy array (host) ; particles: make_particles(...) rows."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        assert images.shape == (len(particles), self.N, self.N), (images.shape, len(particles), self.N)
        self._check(self._L.rfb200_insert_batch(self._h, images.ctypes.data_as(C.c_void_p),
                                                particles.ctypes.data_as(C.c_void_p), len(particles)))

    def insert_host_ptr(self, ptr, particles):
        """Host images given by raw address (e.g. a pinned torch tensor's data_ptr())."""
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        self._check(self._L.rfb200_insert_batch(self._h, C.c_void_p(int(ptr)), particles.ctypes.data_as(C.c_void_p), len(particles)))

    def insert_device_ptr(self, ptr, particles):
        """Images already resident on the GPU (raw device address of n*N*N float32)."""
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        self._check(self._L.rfb200_insert_batch_device(self._h, C.c_void_p(int(ptr)), particles.ctypes.data_as(C.c_void_p), len(particles)))

    def sync(self):
        self._check(self._L.rfb200_sync(self._h))

    def reset(self):
        self._check(self._L.rfb200_reset(self._h))

    # -- multi GPU
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_ubyte * 128)()
        rc = load().rfb200_nccl_unique_id(buf)
        if rc != OK:
            raise RecFourierError(rc, "ncclGetUniqueId failed / NCCL not loadable")
        return bytes(buf)

    def nccl_init(self, unique_id, n_ranks, rank):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        self._check(self._L.rfb200_nccl_init(self._h, buf, int(n_ranks), int(rank)))

    def reduce(self, root=0):
        self._check(self._L.rfb200_reduce_nccl(self._h, int(root)))

    # peer-memory reduce inside one node: exchange the blobs of ipc_export between the ranks, import every peer's
    def ipc_export(self):
        buf = (C.c_ubyte * 256)()
        self._check(self._L.rfb200_ipc_export(self._h, buf))
        return bytes(buf)

    def ipc_import(self, rank, blob):
        buf = (C.c_ubyte * 256).from_buffer_copy(blob)
        self._check(self._L.rfb200_ipc_import(self._h, int(rank), buf))

    def reduce_p2p(self, root=0):
        self._check(self._L.rfb200_reduce_p2p(self._h, int(root)))

    # host-ordered form (no NCCL): set_ranks; per reduce: reduce_p2p_prepare, BARRIER, reduce_p2p_run, BARRIER
    def set_ranks(self, n_ranks, rank):
        self._check(self._L.rfb200_set_ranks(self._h, int(n_ranks), int(rank)))

    def reduce_p2p_prepare(self):
        self._check(self._L.rfb200_reduce_p2p_prepare(self._h))

    def reduce_p2p_run(self, root=0):
        self._check(self._L.rfb200_reduce_p2p_run(self._h, int(root)))

    def ipc_release(self):
        """unmap the peers' accumulators; the ranks must wait for each other between this call and close()"""
        self._check(self._L.rfb200_ipc_release(self._h))

    def accumulator_ptrs(self):
        v, w, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._check(self._L.rfb200_accumulator_ptrs(self._h, C.byref(v), C.byref(w), C.byref(n)))
        return v.value, w.value, n.value

    # -- results
    def accumulators(self):
        if self.fast:       # --fast: the (S+1)^3 temporary volume and weights, S + 1 = info.tile
            n = self.info.tile
            V = np.empty((n, n, n), dtype=np.complex64)
            W = np.empty((n, n, n), dtype=np.float32)
            self._check(self._L.rfb200_export_accumulators(self._h, V.ctypes.data_as(C.c_void_p), W.ctypes.data_as(C.c_void_p)))
            return V, W
        Z, X = self.Z, self.Z // 2 + 1
        V = np.empty((Z, Z, X), dtype=np.complex64)
        W = np.empty((Z, Z, X), dtype=np.float32)
        self._check(self._L.rfb200_export_accumulators(self._h, V.ctypes.data_as(C.c_void_p), W.ctypes.data_as(C.c_void_p)))
        return V, W

    def fast_fourier(self):
        """cfg.fast diagnostics: the transform handed to the inverse FFT."""
        out = np.empty((self.Z, self.Z, self.Z // 2 + 1), dtype=np.complex64)
        self._check(self._L.rfb200_debug_fast_fourier(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def finalize(self):
        out = np.empty((self.N,) * 3, dtype=np.float32)
        self._check(self._L.rfb200_finalize(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def warmup(self):
        """Create the finalisation plan / buffers (and run NCCL's connection set-up) ahead of time; may run on a second
        thread while another one inserts."""
        self._check(self._L.rfb200_warmup(self._h))

    def finalize_into(self, host_ptr):
        """finalize() into caller-owned host memory of N^3 float32 (e.g. page-locked: the 4 N^3-byte read-back then runs
        at full PCIe speed instead of being staged through a bounce buffer)."""
        self._check(self._L.rfb200_finalize(self._h, C.c_void_p(int(host_ptr))))

    def timings(self):
        t = Timings()
        self._check(self._L.rfb200_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in Timings._fields_}

    def halfset_push(self):
        self._check(self._L.rfb200_halfset_push(self._h))

    def halfset_merge(self):
        self._check(self._L.rfb200_halfset_merge(self._h))

    def timer_start(self):
        self._check(self._L.rfb200_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self._L.rfb200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def weight_sum(self):
        s = C.c_double()
        self._check(self._L.rfb200_weight_sum(self._h, C.byref(s)))
        return s.value

    def weight_sum_begin(self):
        self._check(self._L.rfb200_weight_sum_begin(self._h))

    def weight_sum_end(self):
        s = C.c_double()
        self._check(self._L.rfb200_weight_sum_end(self._h, C.byref(s)))
        return s.value

    def streams(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._L.rfb200_get_streams(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_slice(self, idx):
        side, rp = C.c_int32(), C.c_int32()
        self._check(self._L.rfb200_debug_slice_dims(self._h, C.byref(side), C.byref(rp)))
        out = np.empty((side.value, side.value, 4), dtype=np.float32)
        self._check(self._L.rfb200_debug_get_slice(self._h, int(idx), out.ctypes.data_as(C.c_void_p)))
        return out, rp.value


class FourierProjector:
    """GPU central-slice projector with the interface of the reference's FourierProjector
    (data/fourier_projection.h:91-175): FourierProjector(V, paddFactor, maxFreq, degree) then project(...)."""

    NEAREST, LINEAR, BSPLINE3 = 0, 1, 3

    def __init__(self, volume, padding=2.0, max_freq=0.5, degree=3, device=0):
        self._L = load()
        v = np.ascontiguousarray(volume, dtype=np.float32)
        if v.ndim != 3 or not (v.shape[0] == v.shape[1] == v.shape[2]):
            raise ValueError("volume must be a cube")
        self.N = v.shape[0]
        self._h = C.c_void_p()
        rc = self._L.rfb200_projector_create(v.ctypes.data_as(C.c_void_p), self.N, float(padding), float(max_freq), int(degree),
                                             int(device), C.byref(self._h))
        if rc != OK:
            self._h = None
            raise RecFourierError(rc, "rfb200_projector_create failed" + (" (no CUDA device; there is no CPU fallback)" if rc == ERR_CUDA else ""))

    def _check(self, rc):
        if rc != OK:
            raise RecFourierError(rc, (self._L.rfb200_projector_last_error(self._h) or b"").decode())

    def project(self, rot, tilt, psi, ctf=None):
        """rot, tilt, psi: scalars or arrays (degrees); ctf: optional [n, N, N/2+1] multipliers.  Returns [n, N, N] float32."""
        ang = np.ascontiguousarray(np.stack(np.broadcast_arrays(np.atleast_1d(rot), np.atleast_1d(tilt), np.atleast_1d(psi)), axis=1),
                                   dtype=np.float64)
        n = ang.shape[0]
        out = np.empty((n, self.N, self.N), dtype=np.float32)
        c = None
        if ctf is not None:
            c = np.ascontiguousarray(ctf, dtype=np.float32)
            assert c.shape == (n, self.N, self.N // 2 + 1), c.shape
        self._check(self._L.rfb200_projector_project(self._h, ang.ctypes.data_as(C.c_void_p),
                                                     c.ctypes.data_as(C.c_void_p) if c is not None else None, n,
                                                     out.ctypes.data_as(C.c_void_p)))
        return out

    def project_device_ptr(self, rot, tilt, psi, d_images, d_ctf=None):
        """Same, writing n*N*N float32 at the raw device address d_images (e.g. a torch tensor's data_ptr())."""
        ang = np.ascontiguousarray(np.stack(np.broadcast_arrays(np.atleast_1d(rot), np.atleast_1d(tilt), np.atleast_1d(psi)), axis=1),
                                   dtype=np.float64)
        self._check(self._L.rfb200_projector_project_device(self._h, ang.ctypes.data_as(C.c_void_p),
                                                            C.c_void_p(int(d_ctf)) if d_ctf else None, ang.shape[0],
                                                            C.c_void_p(int(d_images))))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rfb200_projector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
