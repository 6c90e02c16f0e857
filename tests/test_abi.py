"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/recfourier_b200.h declares, its POD structs match the ctypes mirror, and it fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "recfourier_b200.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from xmipp3_b200 import _lib
    L = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(L, s), "missing export " + s
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_struct_layout_matches_header():
    from xmipp3_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "recfourier_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(rfb200_config), sizeof(rfb200_particle), sizeof(rfb200_timings), sizeof(rfb200_info));
  printf("%zu %zu %zu %zu\n", offsetof(rfb200_config, sym_matrices), offsetof(rfb200_config, sampling),
         offsetof(rfb200_config, max_batch), offsetof(rfb200_info, n_blocked));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])   # header is plain C
        out = subprocess.check_output([exe]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[0] == C.sizeof(_lib.Config)
    assert sizes[1] == C.sizeof(C.c_double) * len(_lib.PARTICLE_FIELDS) == _lib.PARTICLE_DTYPE.itemsize
    assert sizes[2] == C.sizeof(_lib.Timings)
    assert sizes[3] == C.sizeof(_lib.Info)
    assert sizes[4] == _lib.Config.sym_matrices.offset
    assert sizes[5] == _lib.Config.sampling.offset
    assert sizes[6] == _lib.Config.max_batch.offset
    assert sizes[7] == _lib.Info.n_blocked.offset


def test_oracle_and_product_particle_layouts_agree():
    from oracle import oracle as O
    from xmipp3_b200 import _lib
    assert O.PARTICLE_FIELDS == _lib.PARTICLE_FIELDS


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_argument_validation_needs_no_gpu():
    from xmipp3_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    assert L.rfb200_create(None, C.byref(h)) == _lib.ERR_ARG
    cfg = _lib.Config()
    cfg.abi_version = 999
    assert L.rfb200_create(C.byref(cfg), C.byref(h)) == _lib.ERR_ARG
    assert b"abi_version" in L.rfb200_last_error(None)
    with pytest.raises(_lib.RecFourierError) as e:
        _lib.Reconstructor(32, max_resolution=0.7)
    assert e.value.code == _lib.ERR_ARG
    with pytest.raises(_lib.RecFourierError) as e:
        _lib.Reconstructor(32, blob=(1.9, 1, 15.0))
    assert e.value.code == _lib.ERR_ARG
    with pytest.raises(_lib.RecFourierError) as e:
        _lib.Reconstructor(32, blob=(-1.0, 0, 15.0), fast=True)
    assert e.value.code == _lib.ERR_ARG


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    from xmipp3_b200 import _lib
    with pytest.raises(_lib.RecFourierError) as e:
        _lib.Reconstructor(32)
    assert e.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    # the product package must never import, link or execute anything under oracle/
    pkg = os.path.join(ROOT, "xmipp3_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
