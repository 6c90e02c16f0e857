#!/usr/bin/env python
"""Build oracle/_ref/librefkernel.so: the REFERENCE's own CUDA translation unit of this path
(/root/reference/src/xmipp/libraries/reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp), compiled by nvcc for sm_100a
from where it lies, behind the small C ABI of oracle/ref_harness.cu.  TEST / BENCH INFRASTRUCTURE ONLY.

Needs /root/reference (this container); the GPU box only uses the prebuilt library, which travels with the snapshot
(oracle/_ref/ is git-ignored, not gpurun-ignored).  xmippCore and cuFFTAdvisor are absent from the reference tree
(SURVEY F1), so four stand-in headers (oracle/ref_shims/) take their place; the C++ CPU program itself cannot be built
here.  The recipe the reference uses for this file: `.cpp` under reconstruction_cuda/ compiled as
CUDA, C++17, --expt-extended-lambda (src/xmipp/CMakeLists.txt:197-201, CMakeLists.txt:89-106)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIBS = "/root/reference/src/xmipp/libraries"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "librefkernel.so")


def build(force=False):
    src = os.path.join(HERE, "ref_harness.cu")
    ref_cpp = os.path.join(REF_LIBS, "reconstruction_cuda", "cuda_gpu_reconstruct_fourier.cpp")
    if not os.path.exists(ref_cpp):
        return OUT if os.path.exists(OUT) else None
    deps = [src, ref_cpp] + [os.path.join(r, f) for r, _, fs in os.walk(os.path.join(HERE, "ref_shims")) for f in fs]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-extended-lambda",
           "-Xcompiler", "-fPIC", "-shared", "-w",
           "-I", os.path.join(HERE, "ref_shims"),          # first: shadows core/*, cudaAsserts.h, cuda_xmipp_utils.h
           "-I", REF_LIBS, "-I", os.path.join(REF_LIBS, "reconstruction_cuda"),
           "-o", OUT, src]
    subprocess.check_call(cmd, cwd=HERE)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("oracle/_ref:", p if p else "reference tree not present and no prebuilt library")
