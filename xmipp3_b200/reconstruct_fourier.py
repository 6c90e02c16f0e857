"""Python front door of the path, in the shape the reference is used: every in-tree caller of the
Fourier reconstruction shells out to the CLI (e.g. reconstruct_significant.cpp:606,794), so this
module does the same with the B200 binary and also offers the in-process route over the C ABI.
"""
import subprocess

import numpy as np

from . import _build, _host, io
from ._lib import Reconstructor


class ProgRecFourier:
    """Mirror of ProgRecFourier / ProgRecFourierGPU (reconstruct_fourier.h:80-118): same parameter
    names and defaults, `setIO()` + `run()` as in ProgReconsBase (recons.h:36-44)."""

    def __init__(self, **kw):
        self.fn_sel = None
        self.fn_out = "rec_fourier.vol"
        self.fn_sym = "c1"
        self.fn_fsc = ""
        self.do_weights = False
        self.padding_factor_proj = 2.0
        self.padding_factor_vol = 2.0
        self.blob = (1.9, 0, 15.0)
        self.maxResolution = 0.5
        self.numThreads = 1
        self.NiterWeight = 1
        self.useCTF = False
        self.phaseFlipped = False
        self.minCTF = 0.01
        self.Ts = 1.0
        self.device = 0
        self.bufferSize = 1024
        self.fast = False      # --fast: nearest-pixel insertion + final blob convolution
        self.gpus = 1          # --gpus <n | "all">: one process per GPU, NCCL reduce onto rank 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown parameter " + k)
            setattr(self, k, v)

    def setIO(self, fn_in, fn_out):
        self.fn_sel, self.fn_out = fn_in, fn_out

    def argv(self):
        a = ["-i", self.fn_sel, "-o", self.fn_out, "--sym", self.fn_sym,
             "--padding", repr(float(self.padding_factor_proj)), repr(float(self.padding_factor_vol)),
             "--blob", repr(float(self.blob[0])), str(int(self.blob[1])), repr(float(self.blob[2])),
             "--max_resolution", repr(float(self.maxResolution)), "--thr", str(int(self.numThreads)),
             "--iter", str(int(self.NiterWeight)), "--minCTF", repr(float(self.minCTF)),
             "--device", str(int(self.device)), "--bufferSize", str(int(self.bufferSize))]
        if self.fn_fsc:
            a += ["--prepare_fsc", self.fn_fsc]
        if self.do_weights:
            a.append("--weight")
        if self.useCTF:
            a += ["--useCTF", "--sampling", repr(float(self.Ts))]
        if self.phaseFlipped:
            a.append("--phaseFlipped")
        if self.fast:
            a.append("--fast")
        if self.gpus != 1:
            a += ["--gpus", str(self.gpus)]
        return a

    def run(self, verbose=0):
        """Run the C++ program (xmipp_reconstruct_fourier_b200); raises on a non-zero exit code."""
        _build.build_host()
        _build.build_cuda()
        r = subprocess.run([_build.CLI_BIN] + self.argv() + ["-v", str(int(verbose))], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("xmipp_reconstruct_fourier_b200 failed (%d): %s" % (r.returncode, r.stderr.strip()))
        return r.stdout

    def run_in_process(self):
        """Same result through the C ABI from Python (metadata and images read by the C++ host helpers)."""
        p, names, has_ctf = _host.read_particles(self.fn_sel, use_ctf=self.useCTF)
        nx, ny, _, _ = _host.image_info(names[0])
        imgs = np.stack([_host.read_image(n, nx, ny) for n in names])
        sym = _host.symmetry_matrices(self.fn_sym) if self.fn_sym else None
        r = Reconstructor(nx, padding=(self.padding_factor_proj, self.padding_factor_vol), max_resolution=self.maxResolution,
                          blob=self.blob, sym_matrices=sym, use_ctf=has_ctf, sampling=self.Ts, min_ctf=self.minCTF,
                          phase_flipped=self.phaseFlipped, use_weights=self.do_weights, n_iter_weight=self.NiterWeight,
                          fast=self.fast, device=self.device, max_batch=self.bufferSize)
        r.insert(imgs, p)
        vol = r.finalize()
        r.close()
        _host.write_volume(self.fn_out, vol)
        return vol
