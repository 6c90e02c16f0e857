"""Full-size GPU tests (BASELINE configs 3/4 geometry: box 256, pad 2, Z = 512), where the CPU oracle would take
minutes: size-independent properties of the domain instead of element-wise comparison —
  * linearity / batching invariance (the accumulators are a sum over particles),
  * determinism (two runs are bit-identical: exclusive voxel ownership, no atomics),
  * Hermitian symmetry of the x = 0 plane and W >= 0,
  * reconstruction quality against the analytic phantom (FSC and correlation),
  * agreement with the CPU oracle on the SAME box-256 geometry for a small particle subset."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, synth

pytestmark = pytest.mark.gpu

N = 256


def _particles(d, ctf=False):
    from xmipp3_b200._lib import make_particles
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        cols.update(d["ctf"])
    return make_particles(len(d["rot"]), **cols), cols


def test_box256_oracle_subset(oracle_mod):
    """40 particles at the benchmark geometry against the oracle (the oracle needs ~10 s for these)."""
    from xmipp3_b200._lib import Reconstructor
    n = 40
    d = synth.make_dataset(n, N, seed=31, ctf=True, shifts=True)
    p, cols = _particles(d, True)
    o = oracle_mod.Oracle(N, use_ctf=True, sampling=d["sampling"])
    o.insert(d["images"], oracle_mod.make_particles(n, **cols), threads=1)
    r = Reconstructor(N, use_ctf=True, sampling=d["sampling"])
    r.insert(d["images"], p)
    Vo, Wo = o.accumulators()
    V, W = r.accumulators()
    assert np.linalg.norm(W[:, :, 1:] - Wo[:, :, 1:]) <= 2e-5 * np.linalg.norm(Wo[:, :, 1:])
    assert np.linalg.norm(V[:, :, 1:] - Vo[:, :, 1:]) <= 2e-5 * np.linalg.norm(Vo[:, :, 1:])
    vol, ref = r.finalize(), o.finalize()
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999
    r.close()


def test_box256_properties():
    from xmipp3_b200._lib import Reconstructor
    n = 1536                                    # three gather launches per insert
    d = synth.make_dataset(n, N, seed=32, ctf=False)
    p, _ = _particles(d)
    a = Reconstructor(N)
    a.insert(d["images"], p)
    sa = a.weight_sum()
    b = Reconstructor(N, max_batch=500)         # different chunking and two calls
    b.insert(d["images"][:700], p[:700])
    b.insert(d["images"][700:], p[700:])
    sb = b.weight_sum()
    assert abs(sa - sb) <= 2e-6 * abs(sa)       # checksum of the weights (FP64 reduction on the device)
    Va, Wa = a.accumulators()
    Vb, Wb = b.accumulators()
    assert np.linalg.norm(Va - Vb) <= 3e-6 * np.linalg.norm(Va)
    assert np.linalg.norm(Wa - Wb) <= 3e-6 * np.linalg.norm(Wa)
    b.close()
    c = Reconstructor(N)
    c.insert(d["images"], p)
    assert c.weight_sum() == sa                 # bit-identical rerun
    Vc, Wc = c.accumulators()
    assert np.array_equal(Vc, Va) and np.array_equal(Wc, Wa)
    c.close()
    # expected total weight: every cut-off pixel spreads (approximately) unit mass, twice on the mirrored half
    assert Wa.min() >= 0
    Z = a.Z
    idx = (-np.arange(Z)) % Z
    yh = Z // 2 - 1
    W0, V0 = Wa[:, :, 0], Va[:, :, 0]
    assert np.abs(W0[:, 1:yh + 1] - W0[idx][:, idx][:, 1:yh + 1]).max() <= 1e-5 * W0.max()
    assert np.abs(V0[:, 1:yh + 1] - np.conj(V0[idx][:, idx])[:, 1:yh + 1]).max() <= 1e-5 * np.abs(V0).max()
    vol = a.finalize()
    ph = synth.phantom_volume(d["phantom"], N)
    assert np.corrcoef(vol.ravel(), ph.ravel())[0, 1] > 0.999
    f = synth.fsc(vol, ph)
    assert np.nanmin(f[1:10]) > 0.999           # the Gaussians (sigma >= 5 px) only occupy the lowest shells
    a.close()


def test_box256_d7_symmetry_property():
    """Config 4 geometry: with --sym d7 the map must be invariant under the group (up to interpolation)."""
    from xmipp3_b200._lib import Reconstructor
    n = 60
    d = synth.make_dataset(n, N, seed=33, sym="d7")
    p, _ = _particles(d)
    r = Reconstructor(N, sym_matrices=geometry.point_group_matrices("d7"))
    r.insert(d["images"], p)
    assert r.timings()["planes"] == 14 * n      # 14 insertions per image
    vol = r.finalize()
    # 2-fold about X: (x, y, z) -> (x, -y, -z) in logical coordinates (index i <-> N - i for i >= 1)
    rot = vol[1:, 1:, 1:][::-1, ::-1, :]
    a, b = vol[1:, 1:, 1:], rot
    assert np.corrcoef(a.ravel(), b.ravel())[0, 1] > 0.9999
    ph = synth.phantom_volume(d["phantom"], N)
    assert np.corrcoef(vol.ravel(), ph.ravel())[0, 1] > 0.999
    r.close()


def test_box512_config5_geometry(monkeypatch):
    """Config 5 geometry (box 512, pad 2, Z = 1024: 6.5 GB of accumulators, 1024-point transforms): the fused
    preprocessing chain against the cuFFT chain, determinism, and quality against the analytic phantom."""
    from xmipp3_b200._lib import Reconstructor
    n, box = 96, 512
    d = synth.make_dataset(n, box, seed=41, ctf=True)
    p, _ = _particles(d, True)
    kw = dict(use_ctf=True, sampling=d["sampling"], max_batch=64)
    a = Reconstructor(box, **kw)
    a.insert(d["images"], p)
    sa = a.weight_sum()
    Va, Wa = a.accumulators()
    vol = a.finalize()
    a.close()
    monkeypatch.setenv("RFB200_FFT", "cufft")
    b = Reconstructor(box, **kw)
    monkeypatch.delenv("RFB200_FFT", raising=False)
    b.insert(d["images"], p)
    sb = b.weight_sum()
    Vb, Wb = b.accumulators()
    b.close()
    assert abs(sa - sb) <= 2e-6 * abs(sb)
    assert np.linalg.norm(Va - Vb) <= 3e-6 * np.linalg.norm(Vb)
    assert np.linalg.norm(Wa - Wb) <= 3e-6 * np.linalg.norm(Wb)
    del Vb, Wb
    assert np.isfinite(vol).all() and Wa.min() >= 0
    ph = synth.phantom_volume(d["phantom"], box)
    assert np.corrcoef(vol.ravel(), ph.ravel())[0, 1] > 0.98      # 96 views only
