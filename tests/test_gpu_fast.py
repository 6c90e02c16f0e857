"""GPU parity of the --fast arithmetic (cfg.fast: nearest-pixel insertion + final blob convolution, rf_fast.cuh)
against its CPU restatement (oracle/recfourier_fast_oracle.cpp).  The plane / voxel / pixel decisions are single
precision in the reference; both sides evaluate them with individually rounded operations, so the temporary spaces
must agree voxel for voxel (same support) and to FP32 accumulation order in value."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, io, synth
from xmipp3_b200._lib import Reconstructor, make_particles
from xmipp3_b200.reconstruct_fourier import ProgRecFourier

pytestmark = pytest.mark.gpu


def _cols(d, ctf):
    c = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        c.update(d["ctf"])
    return c


CASES = [
    dict(N=32, n=150, ctf=False, shifts=False),
    dict(N=32, n=120, ctf=True, shifts=True),
    dict(N=32, n=100, ctf=True, shifts=True, min_ctf=0.2, phase_flipped=True),
    dict(N=25, n=80, ctf=False, shifts=True),                                  # odd box
    dict(N=24, n=40, ctf=True, shifts=False, sym="d7"),
    dict(N=32, n=80, ctf=False, shifts=False, padding=(1.0, 1.5), max_resolution=0.3),
    dict(N=32, n=80, ctf=False, shifts=False, padding=(2.0, 1.0), blob=(2.4, 2, 10.0)),
    dict(N=64, n=200, ctf=True, shifts=True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_fast_matches_the_restatement(case, oracle_mod):
    O = oracle_mod
    N, n = case["N"], case["n"]
    ctf = case["ctf"]
    d = synth.make_dataset(n, N, seed=13, ctf=ctf, shifts=case["shifts"], sym=case.get("sym"))
    mats = geometry.point_group_matrices(case["sym"]) if case.get("sym") else None
    kw = dict(padding=case.get("padding", (2.0, 2.0)), max_resolution=case.get("max_resolution", 0.5),
              blob=case.get("blob", (1.9, 0, 15.0)), sym_matrices=mats, use_ctf=ctf, sampling=d["sampling"],
              min_ctf=case.get("min_ctf", 0.01), phase_flipped=case.get("phase_flipped", False))
    cols = _cols(d, ctf)
    r = Reconstructor(N, fast=True, max_batch=64, **kw)
    r.insert(d["images"], make_particles(n, **cols))
    V, W = r.accumulators()
    F = r.fast_fourier()
    vol = r.finalize()
    t = r.timings()
    r.close()
    f = O.FastOracle(N, **kw)
    f.insert(d["images"], O.make_particles(n, **cols))
    Vo, Wo = f.temp_spaces()
    Fo = f.fourier()
    ref = f.finalize()
    # the transform handed to the inverse FFT (blob convolution, symmetrisation, weights), away from the x = Pv/2 plane,
    # which the library stores as its Hermitian part (the only part a c2r transform sees)
    assert synth.rel_l2(F[:, :, :-1], Fo[:, :, :-1]) <= 3e-5
    assert V.shape == Vo.shape == (f.S + 1,) * 3
    # same voxels touched: the nearest-voxel decisions are identical
    assert np.array_equal(W != 0, Wo != 0)
    assert synth.rel_l2(W, Wo) <= 2e-6
    # V: the GPU transform is single precision (the reference's is double, cast to float), sums are atomics
    assert synth.rel_l2(V, Vo) <= 2e-5
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999
    assert t["kernel_launches"] > 0 and t["gather_launches"] > 0


def test_fast_weights_batches_and_reset(oracle_mod):
    """--weight (zero-weight images skipped), several batches, reset, and half-set push / merge on the temporary spaces."""
    O = oracle_mod
    N, n = 32, 90
    d = synth.make_dataset(n, N, seed=17, ctf=False, shifts=False)
    w = np.linspace(0.0, 2.0, n)
    w[5] = 0.0
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], weight=w)
    p = make_particles(n, **cols)
    r = Reconstructor(N, fast=True, use_weights=True, max_batch=32)
    r.insert(d["images"][:40], p[:40])
    r.halfset_push()
    r.insert(d["images"][40:], p[40:])
    r.halfset_merge()
    vol = r.finalize()
    r.reset()
    r.insert(d["images"], p)
    vol2 = r.finalize()
    r.close()
    f = O.FastOracle(N, use_weights=True)
    f.insert(d["images"], O.make_particles(n, **cols))
    ref = f.finalize()
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert synth.rel_l2(vol2, ref) <= 1e-4


def test_cli_fast(tmp_path, oracle_mod):
    """--fast through the drop-in CLI."""
    O = oracle_mod
    N, n = 32, 70
    d = synth.make_dataset(n, N, seed=19, ctf=True, shifts=True)
    stack = str(tmp_path / "particles.stk")
    io.write_spider_stack(stack, d["images"])
    c = d["ctf"]
    cols = {"image": ["%06d@particles.stk" % (k + 1) for k in range(n)], "angleRot": d["rot"], "angleTilt": d["tilt"],
            "anglePsi": d["psi"], "shiftX": d["shift_x"], "shiftY": d["shift_y"], "ctfVoltage": c["kV"],
            "ctfDefocusU": c["defocusU"], "ctfDefocusV": c["defocusV"], "ctfDefocusAngle": c["defocus_angle"],
            "ctfSphericalAberration": c["Cs"], "ctfQ0": c["Q0"]}
    md = str(tmp_path / "input.xmd")
    io.write_xmd(md, cols)
    out = str(tmp_path / "fast.vol")
    prog = ProgRecFourier(useCTF=True, Ts=d["sampling"], fast=True, bufferSize=32, minCTF=0.1)
    prog.setIO(md, out)
    prog.run()
    f = O.FastOracle(N, use_ctf=True, sampling=d["sampling"], min_ctf=0.1)
    f.insert(d["images"], O.make_particles(n, **_cols(d, True)))
    ref = f.finalize()
    vol = io.read_volume(out)
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999
