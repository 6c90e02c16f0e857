"""CPU tests of the C++ host side (csrc/host): metadata reader, image readers/writers, symmetry lists,
CLI parsing — the pieces of the reference program the path needs around the CUDA library."""
import os
import subprocess

import numpy as np
import pytest

from xmipp3_b200 import _build, _host, geometry, io

REF_IMG = "/root/reference/src/xmipp/resources/test/image"
REF_XMD = "/root/reference/src/xmipp/resources/test/sampling/experimental_images.xmd"


def test_symmetry_lists_match_python_twin():
    for name in ("c1", "c4", "d7", "t", "o", "i1", "i3h", "c3v", "d2h", "s4"):
        a = _host.symmetry_matrices(name)
        b = geometry.point_group_matrices(name)
        assert a.shape == b.shape, name
        # same set of matrices (order may differ)
        for m in a:
            assert any(np.abs(m - x).max() < 1e-6 for x in b), name
    assert _host.symmetry_matrices("i3h").shape[0] == 119     # test_symmetries_main.cpp:44-51
    with pytest.raises(_host.HostError):
        _host.symmetry_matrices("q9")


def test_symmetry_file(tmp_path):
    f = tmp_path / "d3.sym"
    f.write_text("rot_axis 3 0 0 1\nrot_axis 2 1 0 0\n")
    a = _host.symmetry_matrices(str(f))
    b = geometry.point_group_matrices("d3")
    assert a.shape == b.shape == (5, 3, 3)
    for m in a:
        assert any(np.abs(m - x).max() < 1e-6 for x in b)


def test_metadata_rows_defaults_and_disabled(tmp_path):
    md = tmp_path / "in.xmd"
    io.write_xmd(str(md), {
        "image": ["000001@s.stk", "000002@s.stk", "000003@s.stk"],
        "enabled": [1, -1, 1],
        "angleRot": [10.5, 20.0, 30.0], "angleTilt": [45.0, 50.0, 55.0], "anglePsi": [-3.0, 0.0, 7.25],
        "shiftX": [1.0, 0.0, -2.0], "shiftY": [0.0, 0.0, 3.0],
        "ctfVoltage": [300.0] * 3, "ctfDefocusU": [15000.0, 16000.0, 17000.0], "ctfSphericalAberration": [2.7] * 3,
        "ctfQ0": [0.07] * 3,
    })
    p, names, has = _host.read_particles(str(md), use_ctf=True)
    assert has and len(p) == 2                                  # removeDisabled drops the second row
    assert names[0].endswith("000001@s.stk") or names[0].endswith("s.stk")
    assert p["rot"][1] == 30.0 and p["psi"][1] == 7.25 and p["shift_y"][1] == 3.0
    assert p["weight"][0] == 1.0                                # default
    assert p["defocusV"][0] == 15000.0                          # defaults to defocusU (ctf.cpp:1181)
    assert p["kV"][0] == 300.0 and p["K"][0] == 1.0 and p["Cs"][0] == 2.7 and p["Q0"][0] == 0.07
    p2, _, has2 = _host.read_particles(str(md), use_ctf=False)
    assert not has2 and p2["defocusU"][0] == 0.0 and p2["kV"][0] == 100.0
    with pytest.raises(_host.HostError):
        _host.read_particles(str(tmp_path / "missing.xmd"))


def test_metadata_ctfmodel_indirection(tmp_path):
    io.write_ctfparam(str(tmp_path / "m.ctfparam"), ctfVoltage=200.0, ctfDefocusU=21000.0, ctfDefocusV=20000.0,
                      ctfDefocusAngle=33.0, ctfSphericalAberration=2.0, ctfQ0=0.1)
    io.write_xmd(str(tmp_path / "in.xmd"), {"image": ["1@s.stk"], "angleRot": [0.0], "angleTilt": [0.0], "anglePsi": [0.0],
                                             "ctfModel": ["m.ctfparam"]})
    p, _, has = _host.read_particles(str(tmp_path / "in.xmd"), use_ctf=True)
    assert has and p["kV"][0] == 200.0 and p["defocusV"][0] == 20000.0 and p["defocus_angle"][0] == 33.0 and p["Q0"][0] == 0.1


@pytest.mark.skipif(not os.path.exists(REF_XMD), reason="reference tree not present")
def test_reads_reference_sample_metadata():
    p, names, _ = _host.read_particles(REF_XMD)
    assert len(p) == 3
    assert abs(p["rot"][0] - 2.5645) < 1e-9 and abs(p["tilt"][0] - 39.456) < 1e-9 and abs(p["shift_x"][0] - 262.8) < 1e-9
    assert names[0].endswith("images/proj_sh000001.spi")


@pytest.mark.skipif(not os.path.exists(REF_XMD), reason="reference tree not present")
def test_every_reference_xmd_fixture_parses_or_fails_cleanly():
    """All .xmd fixtures of the reference's test resources: particle sets (an `image` column) are read, the others
    (sampling tables, block files, a row pointing at a missing ctfModel) are refused with a message, never a crash."""
    import glob
    root = os.path.dirname(os.path.dirname(REF_XMD))
    files = sorted(glob.glob(os.path.join(root, "**", "*.xmd"), recursive=True))
    assert len(files) >= 15
    read = 0
    for f in files:
        try:
            p, names, _ = _host.read_particles(f, use_ctf=True)
            assert len(p) == len(names) > 0
            read += 1
        except _host.HostError as e:
            assert "image" in str(e) or "cannot open" in str(e), (f, str(e))
    assert read >= 4


def test_image_roundtrips(tmp_path):
    rng = np.random.default_rng(0)
    imgs = rng.normal(size=(5, 12, 12)).astype(np.float32)
    for ext, writer in ((".stk", io.write_spider_stack), (".mrcs", io.write_mrc)):
        path = str(tmp_path / ("s" + ext))
        writer(path, imgs)
        nx, ny, nz, n = _host.image_info(path)
        assert (nx, ny, n) == (12, 12, 5)
        for k in range(5):
            got = _host.read_image("%06d@%s" % (k + 1, path), 12, 12)
            assert np.array_equal(got, imgs[k])
        with pytest.raises(_host.HostError):
            _host.read_image("7@" + path, 12, 12)
        with pytest.raises(_host.HostError):
            _host.read_image("1@" + path, 16, 16)
    # C++ writers read back by the Python twins
    for ext in (".stk", ".mrcs"):
        path = str(tmp_path / ("w" + ext))
        _host.write_stack(path, imgs)
        back = io.read_spider(path) if ext == ".stk" else io.read_mrc(path)
        assert np.array_equal(back, imgs)
    vol = rng.normal(size=(6, 6, 6)).astype(np.float32)
    for ext in (".vol", ".mrc"):
        path = str(tmp_path / ("v" + ext))
        _host.write_volume(path, vol)
        assert np.array_equal(io.read_volume(path), vol)
    # single Spider image without index, and the ":fmt" override
    io.write_spider(str(tmp_path / "one.xmp"), imgs[0])
    assert np.array_equal(_host.read_image(str(tmp_path / "one.xmp"), 12, 12), imgs[0])
    io.write_mrc(str(tmp_path / "stack.dat"), imgs)
    assert np.array_equal(_host.read_image("3@" + str(tmp_path / "stack.dat") + ":mrcs", 12, 12), imgs[2])


@pytest.mark.skipif(not os.path.isdir(REF_IMG), reason="reference tree not present")
def test_reads_reference_image_fixtures():
    # the same 3x3 image stored as Spider (both byte orders) and MRC
    a = _host.read_image(REF_IMG + "/singleImage.spi", 3, 3)
    b = _host.read_image(REF_IMG + "/singleImage_swap.spi", 3, 3)
    c = _host.read_image(REF_IMG + "/singleImage.mrc", 3, 3)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert _host.image_info(REF_IMG + "/smallStack.stk") == (64, 64, 1, 4)
    assert _host.image_info(REF_IMG + "/smallStack.mrcs") == (64, 64, 1, 4)
    for k in range(1, 5):
        s = _host.read_image("%d@%s/smallStack.stk" % (k, REF_IMG), 64, 64)
        m = _host.read_image("%d@%s/smallStack.mrcs" % (k, REF_IMG), 64, 64)
        assert np.array_equal(s, m)
    assert np.array_equal(io.read_spider(REF_IMG + "/smallStack.stk"), np.stack(
        [_host.read_image("%d@%s/smallStack.stk" % (k, REF_IMG), 64, 64) for k in range(1, 5)]))
    # ... and as IMAGIC (.hed + .img), addressed through either file
    assert np.array_equal(_host.read_image(REF_IMG + "/singleImage.hed", 3, 3), a)
    assert np.array_equal(_host.read_image(REF_IMG + "/singleImage.img", 3, 3), a)
    assert _host.image_info(REF_IMG + "/smallStack.hed") == (64, 64, 1, 4)
    for k in range(1, 5):
        assert np.array_equal(_host.read_image("%d@%s/smallStack.hed" % (k, REF_IMG), 64, 64),
                              _host.read_image("%d@%s/smallStack.stk" % (k, REF_IMG), 64, 64))


def test_imagic_stack_types_and_byte_order(tmp_path):
    """IMAGIC stacks written here: REAL, INTG (16-bit) and PACK (unsigned bytes), little and big endian headers."""
    rng = np.random.default_rng(5)
    n, ny, nx = 3, 6, 5
    data = {"REAL": rng.normal(size=(n, ny, nx)).astype(np.float32),
            "INTG": rng.integers(-3000, 3000, size=(n, ny, nx)).astype(np.int16),
            "PACK": rng.integers(0, 256, size=(n, ny, nx)).astype(np.uint8)}
    for typ, arr in data.items():
        for order in ("<", ">"):
            base = str(tmp_path / ("s_%s_%s" % (typ, "le" if order == "<" else "be")))
            hed = np.zeros((n, 256), dtype=order + "i4")
            for k in range(n):
                hed[k, 0], hed[k, 1], hed[k, 12], hed[k, 13] = k + 1, n - 1, ny, nx
            raw = hed.tobytes()
            raw = b"".join(raw[1024 * k:1024 * k + 56] + typ.encode() + raw[1024 * k + 60:1024 * (k + 1)] for k in range(n))
            open(base + ".hed", "wb").write(raw)
            arr.astype(arr.dtype.newbyteorder(order)).tofile(base + ".img")
            assert _host.image_info(base + ".hed") == (nx, ny, 1, n)
            for k in range(n):
                got = _host.read_image("%d@%s.img" % (k + 1, base), nx, ny)
                assert np.array_equal(got, arr[k].astype(np.float32)), (typ, order, k)
    with pytest.raises(Exception):
        _host.read_image(str(tmp_path / "missing.hed"), nx, ny)


def test_cli_parsing_defaults_and_errors():
    d = _host.parse_cli(["-i", "in.xmd"])
    # defaults of reconstruct_fourier.cpp:42-58
    assert d["fn_out"] == "rec_fourier.vol" and d["fn_sym"] == "c1" and d["iter"] == "1"
    assert float(d["pad_proj"]) == 2.0 and float(d["pad_vol"]) == 2.0 and float(d["max_resolution"]) == 0.5
    assert float(d["blob_radius"]) == 1.9 and d["blob_order"] == "0" and float(d["blob_alpha"]) == 15.0
    assert d["useCTF"] == "0" and float(d["minCTF"]) == 0.01 and d["do_weights"] == "0"
    d = _host.parse_cli(["-i", "a.xmd", "-o", "out.mrc", "--sym", "d7", "--padding", "1.5", "3", "--blob", "2.1", "2", "10.4",
                         "--max_resolution", "0.4", "--useCTF", "--sampling", "1.34", "--minCTF", "0.05", "--phaseFlipped",
                         "--weight", "--thr", "6", "2", "--iter", "0", "--device", "3", "--bufferSize", "256", "--fftOnGPU"])
    assert d["fn_out"] == "out.mrc" and d["fn_sym"] == "d7" and float(d["pad_proj"]) == 1.5 and float(d["pad_vol"]) == 3.0
    assert float(d["blob_radius"]) == 2.1 and d["blob_order"] == "2" and float(d["blob_alpha"]) == 10.4
    assert d["useCTF"] == "1" and float(d["sampling"]) == 1.34 and float(d["minCTF"]) == 0.05 and d["phaseFlipped"] == "1"
    assert d["do_weights"] == "1" and d["threads"] == "6" and d["iter"] == "0" and d["device"] == "3" and d["bufferSize"] == "256"
    assert d["gpus"] == "1" and d["worldSize"] == "1"
    assert _host.parse_cli(["-i", "a.xmd", "--gpus", "4"])["gpus"] == "4"
    assert _host.parse_cli(["-i", "a.xmd", "--gpus", "all"])["gpus"] == "-1"
    # flags of xmipp_mpi_cuda_reconstruct_fourier (mpi_reconstruct_fourier_gpu.cpp:52-65): one node, static sharding
    d = _host.parse_cli(["-i", "a.xmd", "--mpi_job_size", "500", "-gpusPerNode", "4", "-threadsPerGPU", "3"])
    assert d["gpus"] == "4" and d["threads"] == "12"
    assert _host.parse_cli(["-i", "a.xmd", "-gpusPerNode", "4", "--gpus", "2"])["gpus"] == "2"
    for bad in (["-o", "x.vol"], ["-i", "a.xmd", "--bogus"], ["-i", "a.xmd", "--padding", "two"], ["-i", "a.xmd", "--gpus", "0"],
                ["-i", "a.xmd", "--gpus"]):
        with pytest.raises(_host.HostError):
            _host.parse_cli(bad)


def test_cli_binary_reports_errors(tmp_path):
    _build.build_host()
    exe = _build.CLI_BIN
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "-i <md_file>" in r.stderr
    r = subprocess.run([exe, "-i", str(tmp_path / "nope.xmd"), "-v", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "XMIPP_ERROR" in r.stderr
    # --gpus 2 without a GPU: both ranks fail, the launcher reports it and leaves no rendezvous file behind
    r = subprocess.run([exe, "-i", str(tmp_path / "nope.xmd"), "-v", "0", "--gpus", "2"], capture_output=True, text=True)
    assert r.returncode == 1 and "a GPU rank failed" in r.stderr
    # ranks started by an external launcher need the rendezvous file
    r = subprocess.run([exe, "-i", "x.xmd"], capture_output=True, text=True, env=dict(os.environ, RFB200_WORLD_SIZE="2", RFB200_RANK="1"))
    assert r.returncode == 2 and "RFB200_ID_FILE" in r.stderr


@pytest.mark.parametrize("geo", [dict(N=16), dict(N=25), dict(N=27, pad_proj=1.0, pad_vol=1.0), dict(N=32, pad_proj=1.0, pad_vol=2.0),
                                 dict(N=32, pad_proj=2.0, pad_vol=1.5), dict(N=32, max_res=0.3), dict(N=24, r=2.4), dict(N=40)],
                         ids=lambda g: "-".join("%s=%s" % kv for kv in g.items()))
def test_gather_plan_is_consistent(geo):
    """Host-side plan of the stick gather (no GPU): every voxel the gather owns is covered by exactly one stick of
    each plane class, the pixel validity table equals a brute-force evaluation of the resolution cut-off
    (RF.cpp:594-598), the fast-path radius really is all-valid, and every other reachable lattice point is an
    edge item."""
    from xmipp3_b200 import _host
    res = _host.gather_plan_check(**geo)
    assert res["units_x"] > 0 and res["units_y"] > 0 and res["units_z"] > 0
    for k in ("uncovered", "double_covered", "bad_rim_entries", "bad_inner_pixels", "lost_points"):
        assert res[k] == 0, (k, res)
