"""In-tree build of the native libraries (nvcc for sm_100a, g++ for the host side).

The shared objects are written next to this file so that they travel with the
repository snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_CUDA = os.path.join(_HERE, "librecfourier_b200.so")
LIB_HOST = os.path.join(_HERE, "librecfourier_host.so")
CLI_BIN = os.path.join(_HERE, "xmipp_reconstruct_fourier_b200")
PROJECT_BIN = os.path.join(_HERE, "xmipp_phantom_project_b200")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _sources(*dirs, exts=(".cu", ".cuh", ".h", ".hpp", ".cpp")):
    out = []
    for d in dirs:
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if f.endswith(exts)]
    return out


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return p if os.path.exists(p) else None


def build_cuda(force=False, verbose=False):
    """librecfourier_b200.so: the CUDA kernels + the C ABI of include/recfourier_b200.h."""
    srcs = _sources(CSRC) + [os.path.join(_ROOT, "include", "recfourier_b200.h")]
    srcs = [s for s in srcs if os.sep + "host" + os.sep not in s]
    if not force and _newer(LIB_CUDA, srcs):
        return LIB_CUDA
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB_CUDA):
            return LIB_CUDA
        raise RuntimeError("nvcc not found and no prebuilt librecfourier_b200.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB_CUDA, os.path.join(CSRC, "rf_api.cu"), "-lcufft", "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd, cwd=_ROOT)
    return LIB_CUDA


def build_host(force=False):
    """librecfourier_host.so (metadata / image I/O / symmetry / program class, no CUDA) and the CLI."""
    hdir = os.path.join(CSRC, "host")
    cpps = sorted(f for f in _sources(hdir, exts=(".cpp",)))
    if not cpps:
        return None
    deps = _sources(hdir) + _sources(CSRC, exts=(".h", ".hpp")) + [os.path.join(_ROOT, "include", "recfourier_b200.h")]
    lib_srcs = [c for c in cpps if not c.endswith("_main.cpp")]
    if force or not _newer(LIB_HOST, deps):
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-I", os.path.join(_ROOT, "include"),
               "-I", CSRC, "-o", LIB_HOST] + lib_srcs + ["-ldl"]
        subprocess.check_call(cmd, cwd=_ROOT)
    main = os.path.join(hdir, "reconstruct_fourier_main.cpp")
    if os.path.exists(main) and (force or not _newer(CLI_BIN, deps)):
        cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(_ROOT, "include"), "-I", CSRC, "-o", CLI_BIN,
               main, "-L", _HERE, "-lrecfourier_host", "-Wl,-rpath,$ORIGIN", "-ldl"]
        subprocess.check_call(cmd, cwd=_ROOT)
    pmain = os.path.join(hdir, "phantom_project_main.cpp")
    if os.path.exists(pmain) and (force or not _newer(PROJECT_BIN, deps)):
        cmd = ["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(_ROOT, "include"), "-I", CSRC, "-o", PROJECT_BIN,
               pmain, "-L", _HERE, "-lrecfourier_host", "-Wl,-rpath,$ORIGIN", "-ldl"]
        subprocess.check_call(cmd, cwd=_ROOT)
    return LIB_HOST


def build_all(force=False):
    build_cuda(force=force)
    build_host(force=force)
