"""CPU tests of the --fast restatement (oracle/recfourier_fast_oracle.cpp): geometry of the temporary spaces,
nearest-pixel insertion against an independent numpy evaluation of the same single-precision formulas, the final
blob convolution as a linear operator, and agreement with the exact path at the level the reference promises
("slightly different results", reconstruct_fourier_gpu.cpp:71-72)."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, synth


def _particles(O, d, n, ctf=False):
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        cols.update(d["ctf"])
    return O.make_particles(n, **cols)


def test_fast_dims_follow_the_reference(oracle_mod):
    # paddedImgSize = N*pad_vol; maxVolumeIndexYZ = 2*ceil(ceil(P*maxRes*2)/2); fftSizeX = S/2 (reconstruct_fourier_gpu.cpp:229-232, 434)
    f = oracle_mod.FastOracle(32)
    assert (f.Pv, f.S, f.sx, f.sy) == (64, 64, 32, 64)
    f = oracle_mod.FastOracle(32, padding=(1.0, 1.5), max_resolution=0.3)
    assert (f.Pv, f.S, f.sx, f.sy) == (48, 30, 15, 30)
    f = oracle_mod.FastOracle(25, padding=(2.0, 2.0), max_resolution=0.5)
    assert (f.Pv, f.S, f.sx, f.sy) == (50, 50, 25, 50)


def test_fast_single_untilted_image_lands_on_the_central_plane(oracle_mod):
    """rot = tilt = psi = 0: the normal is z, every (x, y) column is hit at z = S/2 and takes pixel (x - S/2, y): the
    temporary volume's central plane is the cropped transform itself, weights are 1 inside the traverse space."""
    O = oracle_mod
    N = 16
    rng = np.random.default_rng(3)
    img = rng.standard_normal((1, N, N)).astype(np.float32)
    f = O.FastOracle(N)
    f.insert(img, O.make_particles(1, rot=0.0, tilt=0.0, psi=0.0))
    V, W = f.temp_spaces()
    S, half = f.S, f.S // 2
    assert np.count_nonzero(W[:half]) == 0 and np.count_nonzero(W[half + 1:]) == 0
    # expected: 2-D transform of the padded, centred image, scaled 1/P^2, cut at 0.5
    P = f.Pv
    pad = np.zeros((P, P))
    first = -(N // 2)
    idx = (np.arange(N) + first) % P
    pad[np.ix_(idx, idx)] = img[0]
    F = np.fft.rfft2(pad) / (P * P)
    plane = V[half]
    ok = 0
    for y in range(S + 1):
        for x in range(half, S + 1):
            if W[half, y, x] == 0:
                continue
            ix, iy = x - half, y          # nearest pixel: image x = lattice x - S/2, image row y (centred at S/2)
            if ix > half - 1 or iy > S - 1:
                continue                  # clamped pixels
            i = iy - half if iy >= half else iy + P - half
            fx, fy = ix / P, (i if i <= P // 2 else i - P) / P
            exp = F[i, ix] if fx * fx + fy * fy <= np.float32(0.25) else 0.0
            assert abs(plane[y, x] - exp) <= 1e-6 * max(1.0, abs(exp))
            ok += 1
    assert ok > 0.7 * (np.pi / 2) * half * half


def test_fast_is_linear_in_the_images_and_close_to_the_exact_path(oracle_mod):
    O = oracle_mod
    N, n = 24, 300
    d = synth.make_dataset(n, N, seed=4, ctf=False, shifts=True)
    p = _particles(O, d, n)
    f = O.FastOracle(N)
    f.insert(d["images"], p)
    v1 = f.finalize()
    f2 = O.FastOracle(N)
    f2.insert(3.0 * d["images"], p)
    assert synth.rel_l2(f2.finalize(), 3.0 * v1) <= 1e-6          # weights do not depend on the data
    e = O.Oracle(N)
    e.insert(d["images"], p, threads=1)
    ve = e.finalize()
    assert np.corrcoef(v1.ravel(), ve.ravel())[0, 1] >= 0.995      # "slightly different results"
    assert synth.rel_l2(v1, ve) <= 0.15


def test_fast_symmetry_and_ctf_run_and_correlate(oracle_mod):
    """Nearest-pixel insertion picks single 1/CTF-amplified pixels where the blob path averages ~29 of them, so with CTF
    correction the two modes only agree loosely near the CTF zeros (corr 0.86 at --minCTF 0.01, 0.98 at 0.2)."""
    O = oracle_mod
    N, n = 24, 60
    d = synth.make_dataset(n, N, seed=6, ctf=True, shifts=False, sym="d7")
    p = _particles(O, d, n, ctf=True)
    mats = geometry.point_group_matrices("d7")
    f = O.FastOracle(N, sym_matrices=mats, use_ctf=True, sampling=d["sampling"], min_ctf=0.2)
    f.insert(d["images"], p)
    vf = f.finalize()
    e = O.Oracle(N, sym_matrices=mats, use_ctf=True, sampling=d["sampling"], min_ctf=0.2)
    e.insert(d["images"], p, threads=1)
    ve = e.finalize()
    assert np.all(np.isfinite(vf))
    assert np.corrcoef(vf.ravel(), ve.ravel())[0, 1] >= 0.97


def test_fast_golden_fixture(oracle_mod):
    """Regression pin of the restatement itself (tests/golden/fast_oracle_box16.npz, made by tests/golden/make_fast_golden.py):
    the decisions are single precision, so any change of operation order shows up here."""
    import os
    O = oracle_mod
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fast_oracle_box16.npz"))
    N, n = 16, int(g["n"])
    d = synth.make_dataset(n, N, seed=21, ctf=True, shifts=True)
    p = _particles(O, d, n, ctf=True)
    f = O.FastOracle(N, use_ctf=True, sampling=d["sampling"], min_ctf=0.1)
    f.insert(d["images"], p)
    V, W = f.temp_spaces()
    assert int(np.count_nonzero(W)) == int(g["nnz_w"])
    np.testing.assert_allclose(W.sum(dtype=np.float64), float(g["sum_w"]), rtol=1e-6)
    np.testing.assert_allclose(np.abs(V).sum(dtype=np.float64), float(g["sum_abs_v"]), rtol=1e-5)
    np.testing.assert_allclose(f.finalize(), g["vol"], rtol=0, atol=1e-6 * np.abs(g["vol"]).max())
