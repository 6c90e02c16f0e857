"""GPU test of the drop-in CLI: the C++ program reads an .xmd + image stack from disk, reconstructs on the
GPU through the C ABI and writes the volume; the result is compared with the CPU oracle fed the same data."""
import numpy as np
import pytest

from xmipp3_b200 import geometry, io, synth
from xmipp3_b200.reconstruct_fourier import ProgRecFourier

pytestmark = pytest.mark.gpu


def _dataset(tmp_path, N, n, ctf, shifts, stack_ext, sym=None, seed=0):
    d = synth.make_dataset(n, N, seed=seed, ctf=ctf, shifts=shifts, sym=sym)
    stack = str(tmp_path / ("particles" + stack_ext))
    (io.write_spider_stack if stack_ext == ".stk" else io.write_mrc)(stack, d["images"])
    cols = {"image": ["%06d@%s" % (k + 1, "particles" + stack_ext) for k in range(n)],      # relative to the .xmd
            "enabled": [1] * n,
            "angleRot": d["rot"], "angleTilt": d["tilt"], "anglePsi": d["psi"],
            "shiftX": d["shift_x"], "shiftY": d["shift_y"]}
    if ctf:
        c = d["ctf"]
        cols.update({"ctfVoltage": c["kV"], "ctfDefocusU": c["defocusU"], "ctfDefocusV": c["defocusV"],
                     "ctfDefocusAngle": c["defocus_angle"], "ctfSphericalAberration": c["Cs"], "ctfQ0": c["Q0"]})
    md = str(tmp_path / "input.xmd")
    io.write_xmd(md, cols)
    return d, md


def _oracle(oracle_mod, d, N, ctf, sym=None):
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"])
    if ctf:
        cols.update(d["ctf"])
    mats = geometry.point_group_matrices(sym) if sym else None
    o = oracle_mod.Oracle(N, sym_matrices=mats, use_ctf=ctf, sampling=d["sampling"])
    o.insert(d["images"], oracle_mod.make_particles(len(d["rot"]), **cols), threads=1)
    return o.finalize()


def test_cli_end_to_end_spider_ctf(tmp_path, oracle_mod):
    N, n = 32, 80
    d, md = _dataset(tmp_path, N, n, True, True, ".stk")
    out = str(tmp_path / "rec.vol")
    prog = ProgRecFourier(useCTF=True, Ts=d["sampling"], numThreads=4, bufferSize=32)
    prog.setIO(md, out)
    log = prog.run(verbose=1)
    assert "images inserted" in log
    vol = io.read_volume(out)
    ref = _oracle(oracle_mod, d, N, True)
    assert vol.shape == (N, N, N)
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999
    # the in-process route gives the same volume
    prog.fn_out = str(tmp_path / "rec2.mrc")
    vol2 = prog.run_in_process()
    assert synth.rel_l2(vol2, vol) <= 2e-6
    assert np.array_equal(io.read_volume(prog.fn_out), vol2)


def test_cli_end_to_end_mrcs_symmetry(tmp_path, oracle_mod):
    N, n = 32, 24
    d, md = _dataset(tmp_path, N, n, False, False, ".mrcs", sym="d7", seed=3)
    out = str(tmp_path / "rec.mrc")
    prog = ProgRecFourier(fn_sym="d7")
    prog.setIO(md, out)
    prog.run()
    vol = io.read_volume(out)
    ref = _oracle(oracle_mod, d, N, False, sym="d7")
    assert synth.rel_l2(vol, ref) <= 1e-4
    assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999


def test_cli_prepare_fsc_and_iter(tmp_path, oracle_mod):
    """--prepare_fsc writes the two half-set maps (images [0, (n-1)/2] and the rest, RF.cpp:846, 991-1053) and
    the full map; --iter 2 runs the weight refinement (RF.cpp:1080-1092)."""
    N, n = 32, 51
    d, md = _dataset(tmp_path, N, n, True, False, ".stk", seed=5)
    out = str(tmp_path / "full.vol")
    root = str(tmp_path / "fsc")
    prog = ProgRecFourier(useCTF=True, Ts=d["sampling"], fn_fsc=root, NiterWeight=2, minCTF=0.2, bufferSize=16)
    prog.setIO(md, out)
    prog.run()
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], **d["ctf"])
    p = oracle_mod.make_particles(n, **cols)
    split = (n - 1) // 2 + 1
    kw = dict(use_ctf=True, sampling=d["sampling"], min_ctf=0.2, n_iter_weight=2)
    for name, sl in ((root + "_1_recons.vol", slice(0, split)), (root + "_2_recons.vol", slice(split, n)), (out, slice(0, n))):
        o = oracle_mod.Oracle(N, **kw)
        o.insert(d["images"][sl], p[sl], threads=1)
        ref = o.finalize()
        vol = io.read_volume(name)
        assert synth.rel_l2(vol, ref) <= 1e-4, name
        assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999, name


def test_cli_errors(tmp_path):
    N, n = 16, 4
    d, md = _dataset(tmp_path, N, n, False, False, ".stk")
    prog = ProgRecFourier(maxResolution=0.9)
    prog.setIO(md, str(tmp_path / "x.vol"))
    with pytest.raises(RuntimeError) as e:
        prog.run()
    assert "max_resolution" in str(e.value)


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("collective", ["nccl", "p2p", "auto"])
def test_cli_two_gpus_prepare_fsc(tmp_path, oracle_mod, monkeypatch, collective):
    """--gpus 2: the program forks one rank per GPU, each inserts a contiguous shard, one reduce per map (ncclReduce, or
    with RFB200_REDUCE=p2p / auto the peer-memory kernel over CUDA IPC mappings, no NCCL); full map and both half-set maps must equal the
    oracle's (and hence the single-GPU program's)."""
    monkeypatch.setenv("RFB200_REDUCE", collective)
    N, n = 32, 61
    d, md = _dataset(tmp_path, N, n, True, True, ".stk", seed=11)
    out = str(tmp_path / "full.vol")
    root = str(tmp_path / "fsc")
    prog = ProgRecFourier(useCTF=True, Ts=d["sampling"], fn_fsc=root, bufferSize=8, gpus=2)
    prog.setIO(md, out)
    log = prog.run(verbose=1)
    assert "2 GPUs" in log
    cols = dict(rot=d["rot"], tilt=d["tilt"], psi=d["psi"], shift_x=d["shift_x"], shift_y=d["shift_y"], **d["ctf"])
    p = oracle_mod.make_particles(n, **cols)
    split = (n - 1) // 2 + 1
    for name, sl in ((root + "_1_recons.vol", slice(0, split)), (root + "_2_recons.vol", slice(split, n)), (out, slice(0, n))):
        o = oracle_mod.Oracle(N, use_ctf=True, sampling=d["sampling"])
        o.insert(d["images"][sl], p[sl], threads=1)
        ref = o.finalize()
        vol = io.read_volume(name)
        assert synth.rel_l2(vol, ref) <= 1e-4, name
        assert np.nanmin(synth.fsc(vol, ref)[1:]) >= 0.999, name


def test_cli_gpus_all_single_box(tmp_path, oracle_mod):
    """--gpus all on whatever the box has (1 GPU: runs in the launcher's own process)."""
    N, n = 16, 20
    d, md = _dataset(tmp_path, N, n, False, False, ".stk", seed=2)
    out = str(tmp_path / "rec.vol")
    prog = ProgRecFourier(gpus="all", bufferSize=4)
    prog.setIO(md, out)
    prog.run()
    ref = _oracle(oracle_mod, d, N, False)
    assert synth.rel_l2(io.read_volume(out), ref) <= 1e-4
