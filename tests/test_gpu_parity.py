"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Gate (BASELINE.json north_star): relative L2 error <= 1e-4 on the real-space map and FSC >= 0.999
in every shell to Nyquist.  The accumulators are additionally compared at 2e-5 (FP32 vs FP64)."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, synth

pytestmark = pytest.mark.gpu

REL_L2_GATE = 1e-4     # north_star tolerance on the map
FSC_GATE = 0.999       # north_star tolerance per shell
ACC_TOL = 2e-5         # accumulators, relative L2 (x > 0 part; the x = 0 plane is stored symmetrised)


def _cols(d, ctf):
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        cols.update(d["ctf"])
    return cols


def _run_pair(oracle_mod, N, n, ctf=False, shifts=False, sym=None, seed=0, weights=None, **kw):
    from xmipp3_b200._lib import Reconstructor, make_particles
    d = synth.make_dataset(n, N, seed=seed, ctf=ctf, shifts=shifts, sym=sym)
    cols = _cols(d, ctf)
    if weights is not None:
        cols["weight"] = weights
    mats = geometry.point_group_matrices(sym) if sym else None
    args = dict(sym_matrices=mats, use_ctf=ctf, sampling=d["sampling"], **kw)
    o = oracle_mod.Oracle(N, **args)
    o.insert(d["images"], oracle_mod.make_particles(n, **cols), threads=1)
    r = Reconstructor(N, **args)
    r.insert(d["images"], make_particles(n, **cols))
    return o, r, d


def _check(o, r):
    Vo, Wo = o.accumulators()
    V, W = r.accumulators()
    assert np.linalg.norm(V[:, :, 1:] - Vo[:, :, 1:]) <= ACC_TOL * np.linalg.norm(Vo[:, :, 1:])
    assert np.linalg.norm(W[:, :, 1:] - Wo[:, :, 1:]) <= ACC_TOL * np.linalg.norm(Wo[:, :, 1:])
    vo = o.finalize()
    v = r.finalize()
    rel = synth.rel_l2(v, vo)
    f = synth.fsc(v, vo)
    assert rel <= REL_L2_GATE, rel
    assert np.nanmin(f[1:]) >= FSC_GATE, np.nanmin(f[1:])
    return rel


CASES = [
    dict(N=16, n=20),
    dict(N=32, n=100),
    dict(N=32, n=100, ctf=True, shifts=True),
    dict(N=32, n=30, sym="d7"),
    dict(N=32, n=12, sym="o"),
    dict(N=32, n=100, max_resolution=0.3),
    dict(N=24, n=50),
    dict(N=25, n=50),                               # odd box, Z = 50
    dict(N=27, n=40, padding=(1.0, 1.0)),           # odd Z, no padding
    dict(N=32, n=50, padding=(1.0, 2.0)),           # interpolation window 2
    dict(N=32, n=50, padding=(2.0, 1.5)),           # pixel pitch 0.75 voxel, window 6
    dict(N=32, n=50, blob=(1.5, 0, 10.0)),
    dict(N=32, n=50, blob=(2.4, 2, 12.0)),          # blob order 2
    dict(N=32, n=50, n_iter_weight=0),              # --iter 0: no weight correction
    dict(N=32, n=60, ctf=True, phase_flipped=True),
    dict(N=32, n=60, ctf=True, min_ctf=0.2),
    dict(N=32, n=40, ctf=True, min_ctf=0.3, n_iter_weight=2),    # --iter 2: weight refinement pass
    dict(N=32, n=40, ctf=True, min_ctf=0.3, n_iter_weight=3),
    dict(N=24, n=30, n_iter_weight=2, sym="c3"),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_parity_small(oracle_mod, case):
    case = dict(case)
    o, r, _ = _run_pair(oracle_mod, case.pop("N"), case.pop("n"), **case)
    _check(o, r)
    r.close()


def test_parity_config1_full(oracle_mod):
    """BASELINE config 1: 1,000 synthetic 64x64 phantom projections, padding 2, C1."""
    o, r, _ = _run_pair(oracle_mod, 64, 1000)
    rel = _check(o, r)
    assert rel < 2e-5
    r.close()


def test_parity_config2_subset(oracle_mod):
    """BASELINE config 2 (128x128, CTF + random shifts, C1) on its first 600 particles."""
    o, r, _ = _run_pair(oracle_mod, 128, 600, ctf=True, shifts=True)
    _check(o, r)
    r.close()


def test_parity_config4_subset(oracle_mod):
    """BASELINE config 4 geometry (D7, 14 insertions per image) at box 64."""
    o, r, _ = _run_pair(oracle_mod, 64, 100, sym="d7")
    _check(o, r)
    r.close()


def test_weights_and_zero_weight_images(oracle_mod):
    w = np.linspace(0.0, 2.0, 40)
    w[7] = 0.0
    o, r, _ = _run_pair(oracle_mod, 32, 40, weights=w, use_weights=True)
    _check(o, r)
    r.close()


def test_ctf_envelope_parameters(oracle_mod):
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 32, 30
    d = synth.make_dataset(n, N, seed=4, ctf=True)
    cols = _cols(d, True)
    cols.update(Ca=2.0, espr=1.0, ispr=0.5, alpha=0.1, DeltaF=50.0, DeltaR=1.0, K=0.9, envR0=0.01, envR1=0.001)
    o = oracle_mod.Oracle(N, use_ctf=True, sampling=1.5)
    o.insert(d["images"], oracle_mod.make_particles(n, **cols), threads=1)
    r = Reconstructor(N, use_ctf=True, sampling=1.5)
    r.insert(d["images"], make_particles(n, **cols))
    _check(o, r)
    r.close()


def test_slices_match_oracle_preprocess(oracle_mod):
    """K1 (pad + CenterFFT + R2C + crop + mirror) against the oracle's per-image transform."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    N = 32
    d = synth.make_dataset(4, N, seed=9, shifts=True)
    cols = _cols(d, False)
    o = oracle_mod.Oracle(N)
    r = Reconstructor(N)
    r.insert(d["images"], make_particles(4, **cols))
    r.sync()
    P = o.P
    for k in range(4):
        F, _ = o.preprocess(d["images"][k], oracle_mod.make_particles(4, **cols)[k])
        S, Rp = r.debug_slice(k)
        scale = np.abs(F).max()
        for ip in range(-P // 2 + 1, P // 2 + 1):
            for j in range(1, P // 2 + 1):
                if (j / P) ** 2 + (ip / P) ** 2 > 0.25:
                    assert S[ip + Rp, j + Rp, 2] == 0
                    continue
                got = S[ip + Rp, j + Rp]
                assert abs(got[0] + 1j * got[1] - F[ip % P, j]) < 2e-6 * scale
                mir = S[-ip + Rp, -j + Rp]
                assert abs(mir[0] + 1j * mir[1] - np.conj(F[ip % P, j])) < 2e-6 * scale
                assert got[2] == 1.0 and mir[2] == 1.0
    r.close()


def test_fractional_shifts_bspline(oracle_mod):
    """shiftX/shiftY with fractional parts go through cubic B-spline interpolation with wrap
    (xmippCore readApplyGeo) on both sides."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 32, 60
    d = synth.make_dataset(n, N, seed=13)
    rng = np.random.default_rng(5)
    sx = rng.uniform(-3, 3, n)
    sy = rng.uniform(-3, 3, n)
    sx[:5] = np.round(sx[:5])            # mixed: integer x with fractional y ...
    sy[5:10] = np.round(sy[5:10])
    sx[10:15] = np.round(sx[10:15])      # ... and fully integer rows inside the same chunk
    sy[10:15] = np.round(sy[10:15])
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=sx, shift_y=sy)
    o = oracle_mod.Oracle(N)
    o.insert(d["images"], oracle_mod.make_particles(n, **cols), threads=1)
    r = Reconstructor(N)
    r.insert(d["images"], make_particles(n, **cols))
    _check(o, r)
    r.close()


def test_chunking_batching_and_determinism():
    """Size-independent properties: the result does not depend on how the particles are batched,
    two runs are bit-identical (no atomics), and reset() returns to zero."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 32, 150
    d = synth.make_dataset(n, N, seed=11, ctf=True)
    p = make_particles(n, **_cols(d, True))
    kw = dict(use_ctf=True, sampling=1.5)
    a = Reconstructor(N, **kw)
    a.insert(d["images"], p)
    Va, Wa = a.accumulators()
    b = Reconstructor(N, max_batch=32, **kw)          # 5 chunks
    b.insert(d["images"], p)
    Vb, Wb = b.accumulators()
    assert np.linalg.norm(Va - Vb) <= 2e-6 * np.linalg.norm(Va)
    assert np.linalg.norm(Wa - Wb) <= 2e-6 * np.linalg.norm(Wa)
    c = Reconstructor(N, **kw)
    c.insert(d["images"], p)
    Vc, Wc = c.accumulators()
    assert np.array_equal(Va, Vc) and np.array_equal(Wa, Wc)     # deterministic
    # linearity: two halves into one handle == all at once
    e = Reconstructor(N, **kw)
    e.insert(d["images"][:70], p[:70])
    e.insert(d["images"][70:], p[70:])
    Ve, We = e.accumulators()
    assert np.linalg.norm(Va - Ve) <= 2e-6 * np.linalg.norm(Va)
    e.reset()
    Ve, We = e.accumulators()
    assert not Ve.any() and not We.any()
    # empty batch is a no-op
    e.insert(d["images"][:0], p[:0])
    assert not e.accumulators()[1].any()
    for x in (a, b, c, e):
        x.close()


def test_hermitian_and_weight_symmetry_properties():
    """Domain properties at a size the oracle is not needed for: W >= 0, the exported x = 0 plane is
    Hermitian (forceWeightSymmetry / enforceHermitianSymmetry, RF.cpp:1188-1221)."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 300
    d = synth.make_dataset(n, N, seed=21)
    r = Reconstructor(N)
    r.insert(d["images"], make_particles(n, **_cols(d, False)))
    V, W = r.accumulators()
    Z = r.Z
    assert W.min() >= 0
    idx = (-np.arange(Z)) % Z
    W0 = W[:, :, 0]
    V0 = V[:, :, 0]
    yh = Z // 2 - 1
    Wm = W0[idx][:, idx]
    Vm = np.conj(V0[idx][:, idx])
    assert np.abs(W0[:, 1:yh + 1] - Wm[:, 1:yh + 1]).max() <= 1e-5 * W0.max()
    assert np.abs(V0[:, 1:yh + 1] - Vm[:, 1:yh + 1]).max() <= 1e-5 * np.abs(V0).max()
    vol = r.finalize()
    ph = synth.phantom_volume(d["phantom"], N)
    assert np.corrcoef(vol.ravel(), ph.ravel())[0, 1] > 0.999
    r.close()


def test_error_behaviour():
    from xmipp3_b200 import _lib
    r = _lib.Reconstructor(16)
    img = np.zeros((1, 16, 16), np.float32)
    with pytest.raises(AssertionError):
        r.insert(np.zeros((1, 8, 8), np.float32), _lib.make_particles(1))     # wrong image size is caught before the call
    with pytest.raises(_lib.RecFourierError) as e:
        r.reduce(0)                                  # no communicator yet
    assert e.value.code == _lib.ERR_STATE
    r.close()


def test_fused_fft_chain_matches_cufft_chain(monkeypatch):
    """The fused preprocessing chain (own shared-memory FFTs, rf_fft.cuh) against the general chain
    (k_pad_images -> cuFFT R2C -> k_make_slices2) on the same particles: same slices, same accumulators."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    for N, pad in ((32, 2.0), (64, 2.0), (128, 1.0), (100, 1.28)):
        d = synth.make_dataset(24, N, seed=17, ctf=True, shifts=True)
        p = make_particles(24, **_cols(d, True))
        kw = dict(use_ctf=True, sampling=d["sampling"], padding=(pad, pad))
        monkeypatch.delenv("RFB200_FFT", raising=False)
        a = Reconstructor(N, **kw)
        a.insert(d["images"], p)
        Sa, _ = a.debug_slice(3)
        Va, Wa = a.accumulators()
        monkeypatch.setenv("RFB200_FFT", "cufft")
        b = Reconstructor(N, **kw)
        b.insert(d["images"], p)
        Sb, _ = b.debug_slice(3)
        Vb, Wb = b.accumulators()
        monkeypatch.delenv("RFB200_FFT", raising=False)
        assert np.abs(Sa[..., :2] - Sb[..., :2]).max() <= 3e-6 * np.abs(Sb[..., :2]).max()
        assert np.array_equal(Sa[..., 2], Sb[..., 2])
        assert np.linalg.norm(Va - Vb) <= 3e-6 * np.linalg.norm(Vb)
        assert np.linalg.norm(Wa - Wb) <= 3e-6 * np.linalg.norm(Wb)
        a.close()
        b.close()


def test_two_handles_on_one_device_run_concurrently(oracle_mod):
    """The plane tables of a gather launch travel as kernel parameters, so handles share no device state: two
    reconstructions interleaved on one GPU (the reference runs N host threads on N streams, reconstruct_fourier_gpu.cpp:
    417-473) give the results of two separate runs, bit for bit."""
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 300
    da = synth.make_dataset(n, N, seed=71, ctf=True)
    db = synth.make_dataset(n, N, seed=72, sym="c3")
    pa, pb = make_particles(n, **_cols(da, True)), make_particles(n, **_cols(db, False))
    ref = []
    for d, p, kw in ((da, pa, dict(use_ctf=True, sampling=da["sampling"])), (db, pb, dict(sym_matrices=geometry.point_group_matrices("c3")))):
        r = Reconstructor(N, max_batch=50, **kw)
        r.insert(d["images"], p)
        ref.append(r.accumulators())
        r.close()
    a = Reconstructor(N, max_batch=50, use_ctf=True, sampling=da["sampling"])
    b = Reconstructor(N, max_batch=50, sym_matrices=geometry.point_group_matrices("c3"))
    for i0 in range(0, n, 100):          # asynchronous inserts, interleaved: the kernels of a and b overlap on the device
        a.insert(da["images"][i0:i0 + 100], pa[i0:i0 + 100])
        b.insert(db["images"][i0:i0 + 100], pb[i0:i0 + 100])
    Va, Wa = a.accumulators()
    Vb, Wb = b.accumulators()
    a.close()
    b.close()
    assert np.array_equal(Va, ref[0][0]) and np.array_equal(Wa, ref[0][1])
    assert np.array_equal(Vb, ref[1][0]) and np.array_equal(Wb, ref[1][1])


def test_warmup_on_a_second_thread_while_inserting():
    """rfb200_warmup (finalisation plan + buffers, NCCL connection set-up) is the one entry point that may run beside an
    insert on the same handle (the CLI hides ~1.5 s of one-off set-up behind the particle loop that way)."""
    import threading
    from xmipp3_b200._lib import Reconstructor, make_particles
    N, n = 64, 400
    d = synth.make_dataset(n, N, seed=81, ctf=True)
    p = make_particles(n, **_cols(d, True))
    a = Reconstructor(N, use_ctf=True, sampling=d["sampling"], max_batch=40)
    a.insert(d["images"], p)
    ref = a.finalize()
    a.close()
    b = Reconstructor(N, use_ctf=True, sampling=d["sampling"], max_batch=40)
    t = threading.Thread(target=b.warmup)
    t.start()
    b.insert(d["images"], p)
    t.join()
    vol = b.finalize()
    b.close()
    assert np.array_equal(vol, ref)
