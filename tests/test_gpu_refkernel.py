"""GPU tests against the REFERENCE's own CUDA kernel (run with -m gpu).

oracle/_ref/librefkernel.so is the reference's translation unit reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp
compiled for sm_100a from /root/reference by oracle/build_ref.py (the C++ CPU program cannot be built here: xmippCore is
absent).  It pins two things to reference CODE rather than to a restatement:

  * --fast: the temporary volume / weights of the reference's processBufferKernel<useFast> on the buffers the restated
    host side prepares == oracle/recfourier_fast_oracle.cpp's CPU restatement of that device code (and therefore the
    product's k_fast_insert, which tests/test_gpu_fast.py holds against the restatement);
  * the exact path: the reference kernel WITHOUT --fast (per-voxel Kaiser-Bessel blob gather, processVoxelBlob,
    :510-652) is a second, independent implementation of the interpolation the CPU program performs by scattering
    (reconstruct_fourier.cpp:586-792).  Its temporary spaces, folded by the restated mirrorAndCrop, are compared VOXEL BY
    VOXEL with the accumulators V and W of the product and of the oracle: a shared misreading of the blob table, its
    scaling, the Hermitian fold, the double-counted column 0 or the CTF weighting in oracle/ and xmipp3_b200/ would show
    up here (measured agreement: 2e-5 on V, 4e-5 on W)."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refk():
    from oracle import ref_kernel
    if not ref_kernel.available():
        pytest.skip("oracle/_ref/librefkernel.so not built (needs /root/reference at build time)")
    return ref_kernel


def _cols(d, ctf):
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        cols.update(d["ctf"])
    return cols


@pytest.mark.parametrize("case", [dict(N=32, n=60, ctf=True), dict(N=64, n=40, ctf=False, shifts=True), dict(N=32, n=10, sym="c3")],
                         ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_reference_kernel_fast_equals_the_fast_restatement(oracle_mod, refk, case):
    case = dict(case)
    N, n, ctf, sym = case.pop("N"), case.pop("n"), case.pop("ctf", False), case.pop("sym", None)
    d = synth.make_dataset(n, N, seed=11, ctf=ctf, shifts=case.pop("shifts", False), sym=sym)
    p = oracle_mod.make_particles(n, **_cols(d, ctf))
    kw = dict(use_ctf=ctf, sampling=d["sampling"], sym_matrices=geometry.point_group_matrices(sym) if sym else None)
    cpu = oracle_mod.FastOracle(N, **kw)
    cpu.insert(d["images"], p)
    Vc, Wc = cpu.temp_spaces()
    k = refk.RefKernel(oracle_mod.FastOracle(N, **kw), max_images=25)      # --bufferSize default (reconstruct_fourier_gpu.cpp:68)
    try:
        k.process(d["images"], p)
        Vr, Wr = k.temp_spaces()
    finally:
        k.close()
    assert np.array_equal(Wr != 0, Wc != 0)                  # the same voxels are touched
    assert np.abs(Wr - Wc).max() <= 2e-6 * np.abs(Wc).max()  # float atomics: the order of the additions differs
    assert np.linalg.norm(Vr - Vc) <= 2e-6 * np.linalg.norm(Vc)


@pytest.mark.parametrize("case", [dict(N=64, n=300, ctf=True), dict(N=64, n=200, ctf=False), dict(N=32, n=40, sym="d2")],
                         ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_reference_blob_kernel_agrees_with_the_exact_path(oracle_mod, refk, case):
    """ACCUMULATORS, voxel by voxel: the reference kernel's temporary spaces, folded by the restated mirrorAndCrop
    (reconstruct_fourier_gpu.cpp:697-730), against the product's V and W and the oracle's.  Compared inside 80 % of the
    resolution sphere: ProgRecFourierGPU zeroes the pixels beyond --max_resolution but still counts their blob weights
    (cropAndShift :292-321 + processVoxelBlob :560-650), the CPU program skips them (reconstruct_fourier.cpp:597), so the
    two programs differ by design in the outermost shells — the reference's own words for its GPU program are "gives
    slightly different results".  The final maps are compared with that caveat (loose gate)."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    case = dict(case)
    N, n, ctf, sym = case.pop("N"), case.pop("n"), case.pop("ctf", False), case.pop("sym", None)
    d = synth.make_dataset(n, N, seed=12, ctf=ctf, sym=sym)
    cols = _cols(d, ctf)
    p = oracle_mod.make_particles(n, **cols)
    mats = geometry.point_group_matrices(sym) if sym else None
    kw = dict(use_ctf=ctf, sampling=d["sampling"], sym_matrices=mats)
    host = oracle_mod.FastOracle(N, use_fast=False, **kw)      # host side of ProgRecFourierGPU, no --fast
    k = refk.RefKernel(host, max_images=25)
    try:
        k.process(d["images"], p)
        Vr, Wr = k.temp_spaces()
    finally:
        k.close()
    # (the reference kernel leaves a handful of non-finite voxels — 2 to 11 of 2.1 M at box 64 on B200 — which its
    # processWeights turns into zeros, reconstruct_fourier_gpu.cpp:751-767; they are excluded from the comparison)
    bad = ~(np.isfinite(Wr) & np.isfinite(Vr.real) & np.isfinite(Vr.imag))
    if bad.any():
        zz0, yy0, xx0 = np.nonzero(bad)
        rad = np.sqrt((zz0 - host.S / 2.0) ** 2 + (yy0 - host.S / 2.0) ** 2 + (xx0 - host.S / 2.0) ** 2)
        print("reference kernel: %d non-finite voxels, radius %.1f .. %.1f of %d" % (bad.sum(), rad.min(), rad.max(), host.S // 2))
    host.set_temp_spaces(Vr, Wr)
    Vh, Wh = host.half_spaces()                                # [S+1, S+1, S/2+1], centred (z, y), x >= 0
    ref_map = host.finalize()
    o = oracle_mod.Oracle(N, **kw)
    o.insert(d["images"], p, threads=4, scheme="slabs")
    Vo, Wo = o.accumulators()                                  # [Z, Z, Z/2+1], FFT index order
    oracle_map = o.finalize()
    r = Reconstructor(N, **kw)
    r.insert(d["images"], make_particles(n, **cols))
    Vg, Wg = r.accumulators()
    gpu_map = r.finalize()
    r.close()
    S, Z = host.S, o.Z
    c = np.arange(-(S // 2), S // 2 + 1)
    zz, yy, xx = np.meshgrid(c, c, np.arange(0, S // 2 + 1), indexing="ij")
    inner = (zz ** 2 + yy ** 2 + xx ** 2 <= (0.8 * S / 2) ** 2) & (xx >= 1) & np.isfinite(Wh) & np.isfinite(Vh.real) & np.isfinite(Vh.imag)
    iz, iy, ix = zz[inner] % Z, yy[inner] % Z, xx[inner]
    res = {}
    for name, (V, W) in (("oracle", (Vo, Wo)), ("gpu", (Vg, Wg))):
        res["acc_W_%s_vs_refkernel" % name] = float(np.linalg.norm(W[iz, iy, ix] - Wh[inner]) / np.linalg.norm(Wh[inner]))
        res["acc_V_%s_vs_refkernel" % name] = float(np.linalg.norm(V[iz, iy, ix] - Vh[inner]) / np.linalg.norm(Vh[inner]))
    res.update({"map_gpu_vs_refkernel": float(synth.rel_l2(gpu_map, ref_map)), "map_oracle_vs_refkernel": float(synth.rel_l2(oracle_map, ref_map)),
                "map_gpu_vs_oracle": float(synth.rel_l2(gpu_map, oracle_map)), "voxels_compared": int(inner.sum())})
    print(res)
    # single-precision gather with float atomics and a single-precision table index (reference kernel) vs FP64 scatter
    # (oracle) vs FP32 gather with FP64 stick origins (product).  Measured on B200: 2e-5 (V), 4e-5 .. 5e-5 (W).
    for key in ("acc_W_oracle_vs_refkernel", "acc_V_oracle_vs_refkernel", "acc_W_gpu_vs_refkernel", "acc_V_gpu_vs_refkernel"):
        assert res[key] <= 1e-4, res
    assert res["map_gpu_vs_refkernel"] <= 5e-2 and res["map_gpu_vs_oracle"] <= 1e-4, res
