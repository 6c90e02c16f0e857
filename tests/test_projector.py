"""Central-slice projector: CPU tests of the restatement of FourierProjector
(data/fourier_projection.cpp) and GPU parity of rfb200_projector_* against it.  The reference holds no golden
projection in-tree (xmippCore's FourierTransformer / ShiftFFT / CenterFFT are absent), so the restatement is anchored on
closed-form projections of a Gaussian phantom, which pin the Euler convention, the centring and the scale, and on the
round trip through the reconstruction path."""
import numpy as np
import pytest

from xmipp3_b200 import synth


def _phantom(N, seed=1, n_gauss=10):
    ph = synth.make_phantom(n_gauss=n_gauss, box=N, seed=seed)
    return ph, synth.phantom_volume(ph, N).astype(np.float32)


def test_projector_oracle_matches_closed_form_projections(oracle_mod):
    O = oracle_mod
    N = 32
    ph, vol = _phantom(N)
    rot, tilt, psi = synth.random_orientations(5, 5)
    ana = synth.project(ph, N, rot, tilt, psi)
    tol = {0: 0.25, 1: 0.06, 3: 0.008}      # interpolation error of each degree on this phantom (measured 0.18 / 0.044 / 0.0047)
    for deg in (0, 1, 3):
        pr = O.ProjectorOracle(vol, 2.0, 0.5, deg)
        for k in range(5):
            p = pr.project(rot[k], tilt[k], psi[k])
            assert synth.rel_l2(p, ana[k]) <= tol[deg], (deg, k)


def test_projector_oracle_untilted_is_the_sum_along_z_and_ctf_is_a_fourier_multiplier(oracle_mod):
    O = oracle_mod
    N = 24
    ph, vol = _phantom(N, seed=2)
    pr = O.ProjectorOracle(vol, 2.0, 0.5, 3)
    p0 = pr.project(0, 0, 0)
    # exact lattice samples at rot = tilt = psi = 0, except beyond maxFrequency (cut) and on the Nyquist row / column, whose
    # samples at index P/2 fall outside the coefficient window [-(P/2-1), P/2-1] and are mirrored (FP:297-301, 189-215)
    assert synth.rel_l2(p0, vol.sum(0)) <= 0.02
    rng = np.random.default_rng(0)
    ctf = rng.uniform(-1, 1, (N, N // 2 + 1))
    ctf[:, 0] = 1.0
    if N % 2 == 0:
        ctf[:, N // 2] = 1.0
    # build the full-plane multiplier by Hermitian symmetry of a real multiplier: M(-i, -j) = M(i, j)
    pc = pr.project(12.0, 47.0, -80.0, ctf)
    p = pr.project(12.0, 47.0, -80.0)
    F = np.fft.rfft2(p) * ctf
    # rows i and -i of columns 0 / N/2 carry equal multipliers (1), so the product stays a valid half-plane transform
    assert synth.rel_l2(pc, np.fft.irfft2(F, s=(N, N))) <= 1e-9


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("deg", [0, 1, 3])
@pytest.mark.parametrize("N,pad,maxf", [(32, 2.0, 0.5), (25, 2.0, 0.4), (48, 1.5, 0.5)])
def test_gpu_projector_matches_the_restatement(oracle_mod, deg, N, pad, maxf):
    from xmipp3_b200._lib import FourierProjector
    O = oracle_mod
    ph, vol = _phantom(N, seed=3)
    rng = np.random.default_rng(N + deg)
    vol = (vol + 0.05 * rng.standard_normal(vol.shape)).astype(np.float32)       # broadband content: exercises the high frequencies
    n = 6
    rot, tilt, psi = synth.random_orientations(n, 11)
    rot[0] = tilt[0] = psi[0] = 0.0
    pr = O.ProjectorOracle(vol, pad, maxf, deg)
    g = FourierProjector(vol, pad, maxf, deg)
    ctf = rng.uniform(0.2, 1.0, (n, N, N // 2 + 1)).astype(np.float32)
    out = g.project(rot, tilt, psi)
    outc = g.project(rot, tilt, psi, ctf)
    g.close()
    for k in range(n):
        ref = pr.project(rot[k], tilt[k], psi[k])
        refc = pr.project(rot[k], tilt[k], psi[k], ctf[k].astype(np.float64))
        # NEAREST: a coordinate that rounds differently in FP32-rounded angles picks another voxel; none observed, but allow a few
        tol = 2e-5 if deg else 2e-3
        assert synth.rel_l2(out[k], ref) <= tol, (k, synth.rel_l2(out[k], ref))
        assert synth.rel_l2(outc[k], refc) <= tol, (k, synth.rel_l2(outc[k], refc))


@gpu
def test_gpu_project_then_reconstruct_round_trip():
    """Size-independent property tying the two ends of the path together: projections of a volume, made on the GPU and
    inserted by the reconstruction path with the same angles, give the volume back."""
    from xmipp3_b200._lib import FourierProjector, Reconstructor, make_particles
    N, n = 64, 3000
    ph, vol = _phantom(N, seed=4, n_gauss=20)
    rot, tilt, psi = synth.random_orientations(n, 21)
    g = FourierProjector(vol, 2.0, 0.5, 3)
    imgs = g.project(rot, tilt, psi)
    g.close()
    ana = synth.project(ph, N, rot[:8], tilt[:8], psi[:8])
    for k in range(8):
        assert synth.rel_l2(imgs[k], ana[k]) <= 0.01
    r = Reconstructor(N)
    r.insert(imgs, make_particles(n, rot=rot, tilt=tilt, psi=psi))
    rec = r.finalize()
    r.close()
    assert np.corrcoef(rec.ravel(), vol.ravel())[0, 1] >= 0.99
    f = synth.fsc(rec, vol)
    # the Gaussian phantom has no power near Nyquist (the last shells compare interpolation error with nothing); measured:
    # >= 0.9995 in the first 14 shells, 0.988 at shell 21, falling to 0.14 at the last one
    assert np.nanmin(f[1:N // 2 - 10]) >= 0.97


# ---- the drop-in CLI (Fourier mode of xmipp_phantom_project, reconstruction/project.cpp:35-86)
def _project_bin():
    from xmipp3_b200 import _build
    _build.build_host()
    return _build.PROJECT_BIN


def test_phantom_project_cli_argument_errors(tmp_path):
    import subprocess
    exe = _project_bin()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "-i <volume_file>" in r.stderr
    r = subprocess.run([exe, "-i", "v.vol", "-o", "p.xmp", "--angles", "0", "0", "0"], capture_output=True, text=True)
    assert r.returncode == 2 and "--method fourier" in r.stderr                       # real_space (the reference default) is not here
    r = subprocess.run([exe, "-i", "v.vol", "-o", "p.xmp", "--method", "fourier", "2", "0.5", "cubic", "--angles", "0", "0", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "nearest, linear, bspline" in r.stderr               # message of project.cpp:59
    r = subprocess.run([exe, "-i", str(tmp_path / "nope.vol"), "-o", "p.xmp", "--method", "fourier", "--angles", "0", "0", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "XMIPP_ERROR" in r.stderr
    r = subprocess.run([exe, "-i", "v.vol", "-o", "p.xmp", "--method", "fourier", "--params", "x.param", "--angles", "0", "0", "0"],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "mutually exclusive" in r.stderr                     # message of project.cpp:65-66
    # parameter files (project.cpp:220-388): random ranges and noise are refused, a missing file is an error
    par = tmp_path / "noise.param"
    par.write_text("# XMIPP_STAR_1 *\n#\ndata_block1\n_dimensions2D '32 32'\n_projRotRange '0 90 4'\n_projTiltRange '0'\n"
                   "_projPsiRange '0'\n_noisePixelLevel '0.5 0'\n")
    r = subprocess.run([exe, "-i", "v.vol", "-o", "p.stk", "--method", "fourier", "--params", str(par)], capture_output=True, text=True)
    assert r.returncode == 1 and "noisePixelLevel" in r.stderr


@gpu
def test_phantom_project_cli_matches_the_restatement(tmp_path, oracle_mod):
    import subprocess
    from xmipp3_b200 import io
    O = oracle_mod
    exe = _project_bin()
    N = 32
    ph, vol = _phantom(N, seed=6)
    io.write_spider(str(tmp_path / "vol.vol"), vol)
    pr = O.ProjectorOracle(vol, 2.0, 0.5, 3)
    # single projection (PhantomProject.test_case1 of the reference runs exactly this shape of command)
    out = str(tmp_path / "image.xmp")
    r = subprocess.run([exe, "-i", str(tmp_path / "vol.vol"), "-o", out, "--method", "fourier", "2", "0.5", "bspline",
                        "--angles", "30", "60", "-45"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    img = io.read_spider(out).reshape(N, N)
    assert synth.rel_l2(img, pr.project(30.0, 60.0, -45.0)) <= 2e-5
    # a set of orientations from a metadata file -> stack + metadata that the reconstruction program reads
    rot, tilt, psi = synth.random_orientations(5, 3)
    io.write_xmd(str(tmp_path / "angles.xmd"), {"angleRot": rot, "angleTilt": tilt, "anglePsi": psi})
    stack = str(tmp_path / "proj.stk")
    r = subprocess.run([exe, "-i", str(tmp_path / "vol.vol"), "-o", stack, "--method", "fourier", "2", "0.5", "linear",
                        "--angles_md", str(tmp_path / "angles.xmd")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    pl = O.ProjectorOracle(vol, 2.0, 0.5, 1)
    from xmipp3_b200 import _host
    p, names, _ = _host.read_particles(str(tmp_path / "proj.xmd"))
    assert len(names) == 5 and abs(p["rot"][2] - rot[2]) < 1e-4
    for k in range(5):
        got = _host.read_image(names[k], N, N)
        assert synth.rel_l2(got, pl.project(rot[k], tilt[k], psi[k])) <= 2e-5
    # --params (the reference's usual way to make a projection set, project.cpp:62-72, 220-388): deterministic ranges,
    # projection number (i_rot * Ntilt + i_tilt) * Npsi + i_psi
    par = tmp_path / "proj.param"
    par.write_text("# XMIPP_STAR_1 *\n#\ndata_block1\n_dimensions2D '%d %d'\n_projRotRange '0 90 3'\n_projTiltRange '20 60 2'\n"
                   "_projPsiRange '15'\n_noisePixelLevel '0 0'\n" % (N, N))
    stack2 = str(tmp_path / "grid.stk")
    r = subprocess.run([exe, "-i", str(tmp_path / "vol.vol"), "-o", stack2, "--method", "fourier", "2", "0.5", "linear",
                        "--params", str(par)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p, names, _ = _host.read_particles(str(tmp_path / "grid.xmd"))
    expect = [(ro, ti, 15.0) for ro in (0.0, 45.0, 90.0) for ti in (20.0, 60.0)]
    assert len(names) == 6
    for k, (ro, ti, ps) in enumerate(expect):
        assert abs(p["rot"][k] - ro) < 1e-6 and abs(p["tilt"][k] - ti) < 1e-6 and abs(p["psi"][k] - ps) < 1e-6
        assert synth.rel_l2(_host.read_image(names[k], N, N), pl.project(ro, ti, ps)) <= 2e-5
    # ... and an angle file named in the parameter file
    par.write_text("# XMIPP_STAR_1 *\n#\ndata_block1\n_dimensions2D '%d %d'\n_projAngleFile angles.xmd\n" % (N, N))
    r = subprocess.run([exe, "-i", str(tmp_path / "vol.vol"), "-o", str(tmp_path / "fromfile.stk"), "--method", "fourier", "2", "0.5",
                        "linear", "--params", str(par)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p, names, _ = _host.read_particles(str(tmp_path / "fromfile.xmd"))
    assert len(names) == 5 and synth.rel_l2(_host.read_image(names[4], N, N), pl.project(rot[4], tilt[4], psi[4])) <= 2e-5
