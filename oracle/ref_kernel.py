"""ctypes wrapper of oracle/_ref/librefkernel.so: the REFERENCE's own CUDA kernel of this path (processBufferKernel,
reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:898-951) compiled from /root/reference by oracle/build_ref.py.
TEST / BENCH INFRASTRUCTURE ONLY — never imported by the product package."""
import ctypes as C
import os

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "librefkernel.so")


def available():
    return os.path.exists(SO)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefKernel:
    """One instance per process (the reference keeps streams, wrappers and the blob table in module globals).
    Built from a FastOracle (use_fast True or False), which supplies the host side of ProgRecFourierGPU."""

    def __init__(self, fast_oracle, max_images=64):
        self._L = C.CDLL(SO)
        L = self._L
        L.refk_create.argtypes = [C.c_int] * 6 + [C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
        L.refk_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
        L.refk_time.argtypes = [C.c_int, C.c_float, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.refk_download.argtypes = [C.c_void_p, C.c_void_p]
        fo = self.fo = fast_oracle
        table, ids, iw0 = fo.tables()
        r, order, alpha = fo.blob
        # oneOverBessiOrderAlpha (reconstruct_fourier_gpu.cpp:245-262): 1 / bessi<order>(alpha); only order 0 is driven here
        assert int(order) == 0
        one_over = 1.0 / O.lib().orf_bessi0(float(alpha))
        self.S, self.max_images = fo.S, int(max_images)
        rc = L.refk_create(fo.S, fo.sx, fo.sy, self.max_images, fo.n_sym, int(fo.use_ctf), float(r), float(alpha), int(order),
                           float(ids), float(iw0), float(one_over), _ptr(table))
        if rc != 0:
            raise RuntimeError("refk_create failed (%d): no CUDA device, or a second instance in this process" % rc)
        self._open = True
        self.max_res_sqr = float(np.float32(fo.max_resolution) ** 2)

    def process(self, images, particles):
        """prepareBuffer (restated host side) + the reference's processBufferGPU, max_images images per call."""
        n = len(particles)
        for i0 in range(0, n, self.max_images):
            F, Cc, M, sp = self.fo.export_buffer(images[i0:i0 + self.max_images], particles[i0:i0 + self.max_images])
            if len(F) == 0:
                continue
            rc = self._L.refk_process(_ptr(F), _ptr(Cc), _ptr(M), _ptr(sp), len(F), int(self.fo.use_fast), self.max_res_sqr)
            assert rc == 0, rc
        return self

    def time(self, reps=5):
        """(ms per processBufferGPU call, ms per processBufferKernel launch) on the buffer of the last process() call."""
        a, b = C.c_float(), C.c_float()
        rc = self._L.refk_time(int(self.fo.use_fast), self.max_res_sqr, int(reps), C.byref(a), C.byref(b))
        assert rc == 0, rc
        return a.value, b.value

    def temp_spaces(self):
        n = self.S + 1
        V = np.empty((n, n, n), dtype=np.complex64)
        W = np.empty((n, n, n), dtype=np.float32)
        assert self._L.refk_download(_ptr(V), _ptr(W)) == 0
        return V, W

    def clear(self):
        self._L.refk_clear()

    def profile(self):
        v = [C.c_int() for _ in range(6)]
        self._L.refk_profile(*[C.byref(x) for x in v])
        return dict(zip(["BLOCK_DIM", "SHARED_BLOB_TABLE", "SHARED_IMG", "PRECOMPUTE_BLOB_VAL", "TILE", "GRID_DIM_Z"], [x.value for x in v]))

    def close(self):
        if getattr(self, "_open", False):
            self._L.refk_destroy()
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
