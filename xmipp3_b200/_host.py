"""ctypes binding of the host-side helper library (librecfourier_host.so): metadata, image I/O,
symmetry lists and the CLI parser of the C++ program.  No CUDA involved."""
import ctypes as C
import os

import numpy as np

from . import _build
from ._lib import PARTICLE_DTYPE

_lib = None


def load():
    global _lib
    if _lib is None:
        path = _build.build_host()
        if path is None or not os.path.exists(path):
            raise RuntimeError("librecfourier_host.so could not be built")
        L = C.CDLL(path)
        L.rfh_last_error.restype = C.c_char_p
        L.rfh_symmetry_matrices.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
        L.rfh_read_particles.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.rfh_image_info.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 3 + [C.POINTER(C.c_long)]
        L.rfh_read_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.rfh_write_volume.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.rfh_write_stack.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_long]
        L.rfh_parse_cli.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_char_p, C.c_int]
        L.rfh_gather_plan_check.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_longlong)]
        _lib = L
    return _lib


class HostError(RuntimeError):
    pass


def _err():
    return HostError((load().rfh_last_error() or b"").decode())


def symmetry_matrices(name):
    L = load()
    n = L.rfh_symmetry_matrices(name.encode(), None, 0)
    if n < 0:
        raise _err()
    out = np.zeros((n, 3, 3))
    if n:
        L.rfh_symmetry_matrices(name.encode(), out.ctypes.data_as(C.c_void_p), n)
    return out


def read_particles(md_file, use_ctf=False):
    """Rows of the metadata as the C++ program sees them: (particles, image names, has_ctf)."""
    L = load()
    has = C.c_int()
    n = L.rfh_read_particles(md_file.encode(), int(use_ctf), None, None, 0, 0, C.byref(has))
    if n < 0:
        raise _err()
    p = np.zeros(n, dtype=PARTICLE_DTYPE)
    names = C.create_string_buffer(n * 512)
    if L.rfh_read_particles(md_file.encode(), int(use_ctf), p.ctypes.data_as(C.c_void_p), names, 512, n, C.byref(has)) < 0:
        raise _err()
    nm = [names.raw[i * 512:(i + 1) * 512].split(b"\0", 1)[0].decode() for i in range(n)]
    return p, nm, bool(has.value)


def image_info(spec):
    L = load()
    nx, ny, nz, n = C.c_int(), C.c_int(), C.c_int(), C.c_long()
    if L.rfh_image_info(spec.encode(), C.byref(nx), C.byref(ny), C.byref(nz), C.byref(n)) < 0:
        raise _err()
    return nx.value, ny.value, nz.value, n.value


def read_image(spec, nx, ny):
    out = np.empty((ny, nx), dtype=np.float32)
    if load().rfh_read_image(spec.encode(), out.ctypes.data_as(C.c_void_p), nx, ny) < 0:
        raise _err()
    return out


def write_volume(spec, vol):
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nz, ny, nx = vol.shape
    if load().rfh_write_volume(spec.encode(), vol.ctypes.data_as(C.c_void_p), nx, ny, nz) < 0:
        raise _err()


def write_stack(spec, imgs):
    imgs = np.ascontiguousarray(imgs, dtype=np.float32)
    n, ny, nx = imgs.shape
    if load().rfh_write_stack(spec.encode(), imgs.ctypes.data_as(C.c_void_p), nx, ny, n) < 0:
        raise _err()


def parse_cli(argv):
    L = load()
    arr = (C.c_char_p * (len(argv) + 1))(b"prog", *[a.encode() for a in argv])
    out = C.create_string_buffer(4096)
    if L.rfh_parse_cli(len(argv) + 1, arr, out, 4096) < 0:
        raise _err()
    return dict(l.split("=", 1) for l in out.value.decode().strip().split("\n"))


def gather_plan_check(N, pad_proj=2.0, pad_vol=2.0, max_res=0.5, r=1.9):
    """Host-side plan of the stick gather for one geometry, checked for consistency (see host_abi.cpp)."""
    out = (C.c_longlong * 9)()
    if load().rfh_gather_plan_check(int(N), float(pad_proj), float(pad_vol), float(max_res), float(r), out) != 0:
        raise _err()
    keys = ("units_x", "units_y", "units_z", "edge_items", "uncovered", "double_covered", "bad_rim_entries", "bad_inner_pixels", "lost_points")
    return dict(zip(keys, [int(v) for v in out]))
