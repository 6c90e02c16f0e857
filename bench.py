#!/usr/bin/env python
"""bench.py — particles/s inserted by the B200 direct-Fourier reconstruction path.

Workload (BASELINE.json `metric`, config[2]): synthetic 256x256 phantom projections with CTF,
padding 2, C1, blob 1.9/0/15, max_resolution 0.5.  One *step* = one pass of the hot path
(pad/shift -> cuFFT R2C -> CTF-weighted slices -> voxel-centric gather) over one batch of
`--batch` particles.  `value` is measured with the batch already resident in HBM; `e2e` goes
through the C-ABI call with pinned HOST buffers (H2D inside the timed region), reads one 8-byte
result back per step, and includes the final NCCL reduce (N>1), normalisation, 3-D inverse FFT
and the D2H copy of the volume.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                     (CPU oracle port on the host cores)
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particles/sec inserted (box 256, pad 2)"
UNIT = "particles/s"
SAMPLING = 1.5


# ----------------------------------------------------------------------------- work model
def work_model(box, pad=2.0, max_res=0.5, r=1.9):
    P = int(box * pad)
    Z = int(box * pad)
    i = np.arange(-(P - 1 - P // 2), P // 2 + 1)[:, None] / P
    j = np.arange(0, P // 2 + 1)[None, :] / P
    npix = int(((i * i + j * j) <= max_res * max_res).sum())
    pairs = npix * (4.0 / 3.0) * math.pi * r ** 3
    nsphere = (2.0 / 3.0) * math.pi * (0.5 * Z + r) ** 3
    return dict(P=P, Z=Z, npix=npix, pairs=pairs, nsphere=nsphere,
                k2_bytes_per_particle=npix * 8.0, k2_bytes_per_launch_fixed=24.0 * nsphere,
                k2_flops_per_particle=12.0 * pairs)


# ----------------------------------------------------------------------------- synthetic data
def euler_batch(rot, tilt, psi):
    a, b, g = np.radians(rot), np.radians(tilt), np.radians(psi)
    ca, cb, cg, sa, sb, sg = np.cos(a), np.cos(b), np.cos(g), np.sin(a), np.sin(b), np.sin(g)
    cc, cs, sc, ss = cb * ca, cb * sa, sb * ca, sb * sa
    A = np.empty((len(rot), 3, 3))
    A[:, 0, 0] = cg * cc - sg * sa; A[:, 0, 1] = cg * cs + sg * ca; A[:, 0, 2] = -cg * sb
    A[:, 1, 0] = -sg * cc - cg * sa; A[:, 1, 1] = -sg * cs + cg * ca; A[:, 1, 2] = sg * sb
    A[:, 2, 0] = sc; A[:, 2, 1] = ss; A[:, 2, 2] = cb
    return A


def synth_batch_torch(n, box, seed, device, ctf=True):
    """Gaussian-phantom projections with CTF, generated on `device` with torch (data preparation only)."""
    import torch
    from xmipp3_b200 import synth
    c, s, a = synth.make_phantom(n_gauss=30, box=box, seed=0)
    rot, tilt, psi = synth.random_orientations(n, seed + 1)
    A = euler_batch(rot, tilt, psi)
    pc = np.einsum("bij,gj->bgi", A, c)
    g = torch.arange(box, device=device, dtype=torch.float32) - box // 2
    out = torch.empty((n, box, box), device=device, dtype=torch.float32)
    st = torch.tensor(s, device=device, dtype=torch.float32)
    amp = torch.tensor(a * s * np.sqrt(2 * np.pi), device=device, dtype=torch.float32)
    cp = synth.random_ctf_params(n, seed + 3) if ctf else None
    fr = torch.fft.fftfreq(box, device=device, dtype=torch.float64) / SAMPLING
    X = fr[None, None, :]
    Y = fr[None, :, None]
    for b0 in range(0, n, 256):
        b1 = min(n, b0 + 256)
        px = torch.tensor(pc[b0:b1, :, 0], device=device, dtype=torch.float32)
        py = torch.tensor(pc[b0:b1, :, 1], device=device, dtype=torch.float32)
        gx = torch.exp(-0.5 * ((g[None, None, :] - px[:, :, None]) / st[None, :, None]) ** 2)
        gy = torch.exp(-0.5 * ((g[None, None, :] - py[:, :, None]) / st[None, :, None]) ** 2) * amp[None, :, None]
        img = torch.einsum("bgi,bgj->bij", gy, gx)
        if ctf:
            kV = 300.0
            lam = 12.2643247 / math.sqrt(kV * 1e3 * (1.0 + 0.978466e-6 * kV * 1e3))
            K1 = math.pi * lam
            K2 = math.pi / 2 * 2.7e7 * lam ** 3
            dU = torch.tensor(cp["defocusU"][b0:b1], device=device)[:, None, None]
            dV = torch.tensor(cp["defocusV"][b0:b1], device=device)[:, None, None]
            az = torch.deg2rad(torch.tensor(cp["defocus_angle"][b0:b1], device=device))[:, None, None]
            u2 = X * X + Y * Y
            ang = torch.atan2(Y, X).expand(1, box, box)
            deltaf = -(dU + dV) * 0.5 - (dU - dV) * 0.5 * torch.cos(2 * (ang - az))
            arg = K1 * deltaf * u2 + K2 * u2 * u2
            q0 = 0.07
            cval = -(math.sqrt(1 - q0 * q0) * torch.sin(arg) - q0 * torch.cos(arg))
            img = torch.fft.ifft2(torch.fft.fft2(img.to(torch.float64)) * cval).real.to(torch.float32)
        out[b0:b1] = img
    cols = dict(rot=rot, tilt=tilt, psi=psi)
    if ctf:
        cols.update(cp)
    return out, cols


def synth_batch_numpy(n, box, seed, ctf=True):
    from xmipp3_b200 import synth
    ph = synth.make_phantom(n_gauss=30, box=box, seed=0)
    rot, tilt, psi = synth.random_orientations(n, seed + 1)
    img = synth.project(ph, box, rot, tilt, psi)
    cols = dict(rot=rot, tilt=tilt, psi=psi)
    if ctf:
        cp = synth.random_ctf_params(n, seed + 3)
        img = synth.apply_ctf(img, SAMPLING, **cp)
        cols.update(cp)
    return np.ascontiguousarray(img, dtype=np.float32), cols


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm
def time_oracle(box, n_sample, threads, ctf=True, seed=100, sym="c1"):
    """Time the CPU oracle (port of ProgRecFourier, reference thread scheme) on n_sample particles."""
    from oracle import oracle as O
    img, cols = synth_batch_numpy(n_sample, box, seed, ctf)
    p = O.make_particles(n_sample, **cols)
    mats = None
    if sym != "c1":
        from xmipp3_b200 import geometry
        mats = geometry.point_group_matrices(sym)
    o = O.Oracle(box, use_ctf=ctf, sampling=SAMPLING, sym_matrices=mats)
    o.insert(img[: min(threads, n_sample)], p[: min(threads, n_sample)], threads=threads)   # touch pages, spin up
    t = time.perf_counter()
    o.insert(img, p, threads=threads)
    dt = time.perf_counter() - t
    return n_sample / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    box = args.box
    sample = args.ref_sample
    img, cols = synth_batch_numpy(sample, box, 100, True)
    p = O.make_particles(sample, **cols)
    o = O.Oracle(box, use_ctf=True, sampling=SAMPLING)
    for _ in range(args.warmup):
        o.insert(img, p, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        o.insert(img, p, threads=cores)
    dt = time.perf_counter() - t
    value = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config[2]: %dx%d phantom projections with CTF, padding 2, C1 (CPU arm: %d particles per step)" % (box, box, sample),
                   "box": box, "padding": 2, "sym": "c1", "ctf": True, "particles_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d particles per step x %d steps, oracle/recfourier_oracle.cpp (double, reference row-scheduler threading, %d threads)" % (sample, args.steps, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- other configurations, scaling, checks
OTHER_CONFIGS = [
    # BASELINE.json configs 1, 2, 4, 5 (config 3 is the headline).  batch = particles per step and GPU
    dict(name="config[0]: 64x64 phantom projections, padding 2, C1", box=64, sym="c1", ctf=False, shifts=False, batch=1000, steps=3, chunk=1024),
    dict(name="config[1]: 128x128 particles with CTF + integer shifts, C1", box=128, sym="c1", ctf=True, shifts=True, batch=2048, steps=3, chunk=1024),
    dict(name="config[3]: 256x256 particles, D7 (14 insertions per image)", box=256, sym="d7", ctf=False, shifts=False, batch=512, steps=2, chunk=512),
    dict(name="config[4]: 512x512 particles, padding 2 (1024^3 volume)", box=512, sym="c1", ctf=False, shifts=False, batch=512, steps=2, chunk=256),
]


def measure_other_config(cfg, dev, local, rank, world, max_over_ranks, barrier_all):
    """Short device-resident measurement of one of BASELINE's other configurations (same metric, same timing rules:
    1 warm-up pass over each of the two batches, `steps` timed passes, CUDA events on the library's stream, max over
    ranks).  Every batch is larger than L2 except config[0] (16 MB per batch; two batches alternate)."""
    import torch
    from xmipp3_b200 import geometry
    from xmipp3_b200._lib import Reconstructor, make_particles
    box, B, K = cfg["box"], cfg["batch"], cfg["steps"]
    mats = geometry.point_group_matrices(cfg["sym"]) if cfg["sym"] != "c1" else None
    n_ops = 1 + (len(mats) if mats is not None else 0)
    batches = []
    for s in range(2):
        img, cols = synth_batch_torch(B, box, 50000 + 1000 * rank + 10 * s, dev, ctf=cfg["ctf"])
        if cfg["shifts"]:
            rng = np.random.default_rng(7 + s)
            cols["shift_x"] = rng.integers(-5, 6, B).astype(np.float64)
            cols["shift_y"] = rng.integers(-5, 6, B).astype(np.float64)
        batches.append((img, make_particles(B, **cols)))
    torch.cuda.synchronize()
    r = Reconstructor(box, use_ctf=cfg["ctf"], sampling=SAMPLING, device=local, max_batch=cfg["chunk"], sym_matrices=mats)
    for img, p in batches:
        r.insert_device_ptr(img.data_ptr(), p)
    r.sync()
    r.reset()
    barrier_all()
    r.timer_start()
    for i in range(K):
        img, p = batches[i & 1]
        r.insert_device_ptr(img.data_ptr(), p)
    ms = max_over_ranks(r.timer_stop())
    tm = r.timings()
    r.close()
    del batches
    torch.cuda.empty_cache()
    value = world * K * B / (ms * 1e-3)
    return {"workload": cfg["name"], "box": box, "sym": cfg["sym"], "ctf": cfg["ctf"], "insertions_per_particle": n_ops,
            "particles_per_step_per_gpu": B, "steps": K, "value": value, "unit": UNIT, "insertions_per_s": value * n_ops,
            "ms_per_step": ms / K, "gpu_launches": int(tm["kernel_launches"]),
            "stage_ms_per_step": {k: tm[k] / K for k in ("preprocess_ms", "fft2d_ms", "slice_ms", "gather_ms", "edge_ms")}}


def measure_strong_scaling(r, host, box, world, rank, total, barrier_all, max_over_ranks, vol_ptr, p2p=False):
    """Config 3 AS STATED: `total` particles in all, sharded over the ranks, end to end (pinned host batches -> H2D ->
    insertion -> one NCCL reduce -> normalisation + 3-D inverse FFT + D2H of the map on rank 0).  The two pinned batches
    are cycled: the bytes moved and the work done are those of `total` distinct particles."""
    from xmipp3_b200.sharding import shard_range
    lo, hi = shard_range(total, world, rank)
    mine = hi - lo
    r.reset()
    barrier_all()
    t0 = time.perf_counter()
    done, i = 0, 0
    while done < mine:
        hbuf, p = host[i & 1]
        n = min(len(p), mine - done)
        r.insert_host_ptr(hbuf.data_ptr(), p[:n])
        done += n
        i += 1
    if world > 1:
        do_reduce(r, p2p)
    if rank == 0:
        r.finalize_into(vol_ptr)
    else:
        r.sync()
    barrier_all()
    dt = max_over_ranks(time.perf_counter() - t0)
    return {"workload": "config[2] as stated: %d particles %dx%d with CTF in total, sharded over %d GPU(s)" % (total, box, box, world),
            "scaling": "strong", "particles_total": total, "particles_per_gpu": int(mine), "seconds": dt, "value": total / dt, "unit": UNIT,
            "includes": "H2D from pinned host memory, insertion, reduce onto rank 0 (N>1: %s), normalise + 3-D IFFT + D2H of the map" % ("peer-memory kernel over NVLink" if (p2p and os.environ.get("RFB200_BENCH_REDUCE") == "p2p") else "ncclReduce")}


def bind_to_gpu_numa_node(local):
    """Run this rank (and therefore first-touch its pinned batches) on the CPUs of the NUMA node its GPU hangs off, as NCCL
    and the multi-GPU CLI do for their threads: with eight ranks pulling 25-55 GB/s each out of host memory, buffers
    that sit on the other socket halve the H2D rate.  Best effort: any missing sysfs entry leaves the affinity alone."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": open(dev + "/numa_node").read().strip(), "cpus": len(cpus)}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}
    return None


def setup_p2p(r, dev, world, rank):
    """Peer-memory reduce: every rank exports the IPC handles of its accumulators, all ranks import all others.  True when
    every rank succeeded.  Opt-in: RFB200_BENCH_REDUCE=p2p uses rfb200_reduce_p2p wherever the bench reduces, =both keeps
    ncclReduce for the measured paths and reports the two collectives side by side (`reduce_ms`, `reduce_check`).  The
    default is NCCL alone: measured on 2 and 4 B200s the NVLS reduce of NCCL is as fast or faster (profiles/
    r2c_bench_n*_p2p_reduce.json: 1.32 vs 1.36 ms at N = 2, 1.43 vs 2.00 ms at N = 4 for 856 MB per rank)."""
    import torch
    import torch.distributed as dist
    if world <= 1 or os.environ.get("RFB200_BENCH_REDUCE", "nccl") == "nccl":
        return False
    ok = 1
    try:
        blob = r.ipc_export()
    except Exception:
        blob, ok = None, 0
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    if ok and all(b is not None for b in blobs):
        try:
            for k, b in enumerate(blobs):
                if k != rank:
                    r.ipc_import(k, b)
        except Exception as e:
            print("bench.py: rank %d cannot map peer memory (%s); using the NCCL reduce" % (rank, e), file=sys.stderr)
            ok = 0
    else:
        ok = 0
    t = torch.tensor([ok], device=dev, dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def do_reduce(r, p2p):
    if p2p and os.environ.get("RFB200_BENCH_REDUCE") == "p2p":
        r.reduce_p2p(0)
    else:
        r.reduce(0)


def reduce_check(dev, local, rank, world):
    """N > 1: is the volume rank 0 finalises after rfb200_reduce_nccl the sum over the ranks?  Every rank inserts its own
    96 particles (box 64, CTF); (a) the FP64 checksum of W on rank 0 after the reduce against the sum of the per-rank
    checksums, (b) the finalised map against rank 0 inserting all world x 96 particles by itself."""
    import torch
    import torch.distributed as dist
    from xmipp3_b200._lib import Reconstructor, make_particles
    box, n = 64, 96
    def data(rk):
        img, cols = synth_batch_torch(n, box, 777000 + 13 * rk, dev, ctf=True)
        return img, make_particles(n, **cols)
    r = Reconstructor(box, use_ctf=True, sampling=SAMPLING, device=local)
    ids = [Reconstructor.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    r.nccl_init(ids[0], world, rank)
    p2p = setup_p2p(r, dev, world, rank)
    img, p = data(rank)
    r.insert_device_ptr(img.data_ptr(), p)
    mine = r.weight_sum()
    t = torch.tensor([mine], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    after_nccl = None
    if p2p:
        # both collectives on the same partial volumes: the NCCL reduce first (rank 0's accumulators then hold the sum, the
        # others are unchanged), the result noted, rank 0 re-inserts its own particles into zeroed accumulators, then the
        # peer-memory reduce, whose result is the one checked against the single-rank reconstruction below
        r.reduce(0)
        if rank == 0:
            after_nccl = r.weight_sum()
            r.reset()
            r.insert_device_ptr(img.data_ptr(), p)
        r.sync()
        dist.barrier()
        r.reduce_p2p(0)
    else:
        r.reduce(0)
    out = None
    if rank == 0:
        after = r.weight_sum()
        vol = r.finalize()
        solo = Reconstructor(box, use_ctf=True, sampling=SAMPLING, device=local)
        for rk in range(world):
            im2, p2 = data(rk)
            solo.insert_device_ptr(im2.data_ptr(), p2)
        solo_sum = solo.weight_sum()
        ref = solo.finalize()
        solo.close()
        out = {"ranks": world, "particles_per_rank": n, "collective": "rfb200_reduce_nccl, then rfb200_reduce_p2p (peer memory over NVLink) on re-inserted particles" if p2p else "rfb200_reduce_nccl",
               "weight_sum_after_nccl_reduce": after_nccl,
               "weight_sum_after_reduce": after, "sum_of_rank_weight_sums": float(t.item()),
               "weight_sum_rel_err": abs(after - float(t.item())) / abs(float(t.item())),
               "weight_sum_one_rank_all_particles": solo_sum,
               "map_rel_l2_vs_one_rank": float(np.linalg.norm(vol - ref) / np.linalg.norm(ref))}
        out["ok"] = bool(out["weight_sum_rel_err"] <= 1e-6 and out["map_rel_l2_vs_one_rank"] <= 1e-5)
    else:
        r.sync()
    if p2p:
        r.ipc_release()
        dist.barrier()      # every rank has unmapped its peers before anybody frees
    r.close()
    dist.barrier()
    return out


def measure_ref_gpu_kernel(box, n_images=50):
    """Same-box GPU baseline: the REFERENCE's own CUDA kernel for this path (processBufferKernel<false, hasCTF, 0, ...>,
    reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:898-951, blob path :510-652), compiled for sm_100a from the
    reference tree by oracle/build_ref.py, timed on buffers prepared as ProgRecFourierGPU::prepareBuffer does (restated
    host side, oracle/recfourier_fast_oracle.cpp).  Kernel only (CUDA events on its stream), and the whole
    processBufferGPU call (host copy of the buffer + H2D + kernel)."""
    from oracle import oracle as O
    from oracle import ref_kernel
    if not ref_kernel.available():
        return {"unavailable": "oracle/_ref/librefkernel.so not built (needs /root/reference at build time)"}
    img, cols = synth_batch_numpy(n_images, box, 4242, True)
    p = O.make_particles(n_images, **cols)
    host = O.FastOracle(box, use_ctf=True, sampling=SAMPLING, use_fast=False)
    out = {"kernel": "processBufferKernel<useFast=false, hasCTF=true, blobOrder=0> (reference, recompiled for sm_100a)",
           "source": "reconstruction_cuda/cuda_gpu_reconstruct_fourier.cpp:898-951 via oracle/ref_harness.cu", "unit": UNIT, "runs": []}
    for buf in (25, n_images):       # --bufferSize default of the reference (25) and a larger buffer
        k = ref_kernel.RefKernel(host, max_images=buf)
        try:
            out["profile"] = k.profile()
            k.process(img[:buf], p[:buf])
            call_ms, kernel_ms = k.time(reps=3)
        finally:
            k.close()
        out["runs"].append({"images_per_buffer": buf, "kernel_ms": kernel_ms, "call_ms": call_ms,
                            "kernel_particles_per_s": buf / (kernel_ms * 1e-3), "call_particles_per_s": buf / (call_ms * 1e-3)})
    best = max(out["runs"], key=lambda x: x["kernel_particles_per_s"])
    out["value"] = best["kernel_particles_per_s"]
    out["value_whole_call"] = max(x["call_particles_per_s"] for x in out["runs"])
    return out


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--box", type=int, default=256)
    ap.add_argument("--batch", type=int, default=4096, help="particles per step and per GPU")
    ap.add_argument("--ref-sample", type=int, default=192, help="particles per step of the CPU arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs, strong_scaling, reduce_check and ref_gpu_kernel")
    ap.add_argument("--strong-total", type=int, default=100000, help="particles of the strong-scaling block (config 3 as stated)")
    ap.add_argument("--chunk", type=int, default=1024, help="particles per preprocessing chunk inside the library (max_batch)")
    ap.add_argument("--sym", default="c1", help="point group (e.g. d7 = BASELINE config 4: 14 insertions per image); "
                    "not the headline configuration")
    ap.add_argument("--no-ctf", action="store_true", help="particles without CTF columns (BASELINE configs 1, 4, 5 name none)")
    ap.add_argument("--fast", action="store_true", help="measure the --fast arithmetic (nearest-pixel insertion + final blob "
                    "convolution) instead of the exact blob insertion; not the headline configuration")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner does) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("RFB200_BENCH_NUMA", "1") != "0" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from xmipp3_b200._lib import Reconstructor, make_particles
    box, B, K, W = args.box, args.batch, args.steps, args.warmup
    wm = work_model(box)

    # two distinct device-resident batches (each >> L2) used alternately
    batches = []
    for s in range(2):
        img, cols = synth_batch_torch(B, box, 1000 * rank + 10 * s, dev, ctf=not args.no_ctf)
        batches.append((img, make_particles(B, **cols)))
    torch.cuda.synchronize()

    sym_mats = None
    n_ops = 1
    if args.sym.lower() != "c1":
        from xmipp3_b200 import geometry
        sym_mats = geometry.point_group_matrices(args.sym)
        n_ops = len(sym_mats) + 1
    r = Reconstructor(box, use_ctf=not args.no_ctf, sampling=SAMPLING, device=local, max_batch=args.chunk, fast=args.fast, sym_matrices=sym_mats)
    if world > 1:
        ids = [Reconstructor.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        r.nccl_init(ids[0], world, rank)
    p2p = setup_p2p(r, dev, world, rank) if world > 1 else False

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        r.sync()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local)

    # ---------------- device-resident throughput ("value")
    for i in range(W):
        img, p = batches[i & 1]
        r.insert_device_ptr(img.data_ptr(), p)
    barrier()
    r.reset()
    barrier()
    clocks.start()
    r.timer_start()
    for i in range(K):
        img, p = batches[i & 1]
        r.insert_device_ptr(img.data_ptr(), p)
    ms = r.timer_stop()
    barrier()
    clocks.stop()
    ms = max_over_ranks(ms)
    tm = r.timings()
    value = world * K * B / (ms * 1e-3)
    launches = int(tm["kernel_launches"])

    # per-kernel roofline of the dominant kernel (k_gather_sticks: one launch per plane class and per sub-range of
    # <= 512 planes), from CUDA events on its stream
    g_launches = max(1, int(tm["gather_launches"]))
    g_ms = tm["gather_ms"] / g_launches
    imgs_per_launch = K * B / g_launches
    alg_bytes = imgs_per_launch * wm["k2_bytes_per_particle"] + wm["k2_bytes_per_launch_fixed"]
    alg_flops = imgs_per_launch * wm["k2_flops_per_particle"] * n_ops
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    prof = {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "gather_ncu.json")))
    except Exception:
        pass
    traffic = prof.get("dram_bytes_per_launch")
    cs = clocks.summary()
    sm_mhz = cs.get("sm_mhz") or 1965.0
    fp32_nominal = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    try:
        from xmipp3_b200._lib import measure_fp32_peak
        fp32_peak = measure_fp32_peak(local)
        fp32_src = "measured: FFMA-chain micro-benchmark on this GPU (rfb200_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x %.0f MHz = %.1f" % (sm_mhz, fp32_nominal)
    except Exception as e:          # older library build
        fp32_peak = fp32_nominal
        fp32_src = "nominal 148 SM x 128 lanes x 2 x %.0f MHz (micro-benchmark unavailable: %s)" % (sm_mhz, e)
    roofline = {"kernel": "k_gather_sticks<4,cls>", "bound": "hbm", "achieved": alg_bytes / (g_ms * 1e-3) / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": alg_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": prof.get("source", "none: no ncu summary in profiles/gather_ncu.json") +
                                  " (one ncu --set full capture of this kernel, committed; NOT measured in this run)",
                "ms_per_launch": g_ms, "particles_per_launch": imgs_per_launch,
                "us_per_particle": 1e3 * g_ms / imgs_per_launch,
                "algorithmic_bytes_per_launch": alg_bytes,
                "frac_note": "algorithmic bytes per launch = particles x Npix x 8 B (slices) + 24 B x Nsphere (every launch reads and "
                             "writes the touched volume once, SURVEY 8d); launches with more planes amortise the second term, so "
                             "`achieved` falls when the kernel gets faster per particle by batching more planes per launch",
                "bound_actual": "L1/shared-memory data pipe: about 75 wavefronts per warp step-iteration (42 tag requests of the window loads, 27 of "
                                "the blob-table lookups, 5.5 of the stick accumulators) against about 60 issue cycles (profiles/r2b_gather_hot_loop.txt); "
                                "the contract's `bound` only admits hbm|tensor",
                "note": "the gather is bound by the L1/shared-memory data pipe (per-pair pixel and blob-table fetches), "
                        "not by HBM or FP32: see l1_data_pipe (ncu) and fp32"}
    fp32 = {"achieved": alg_flops / (g_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": alg_flops / (g_ms * 1e-3) / 1e12 / fp32_peak,
            "peak_source": fp32_src}
    l1_pipe = {"frac": prof.get("l1_data_pipe_frac"), "issue_frac": prof.get("issue_active_frac"),
               "source": prof.get("source", "no ncu summary in profiles/gather_ncu.json"),
               "note": "ncu l1tex__data_pipe_lsu_wavefronts / smsp__issue_active of the same kernel (cold, serialised)"}
    stage_ms = {k: tm[k] / K for k in ("preprocess_ms", "fft2d_ms", "slice_ms", "gather_ms", "edge_ms")}
    # roofline fraction of every kernel of the step against the measured HBM peak, from the stage timers (CUDA events on
    # the compute stream).  algorithmic bytes = SURVEY section 8d: K1 (whole preprocessing) = N^2*4 (raw image) +
    # P(P/2+1)*8 (half-plane transform) + Npix*8 (compact slice) per particle; format bytes = what this implementation's
    # buffers move (intermediate T of the fused chain written + read once, half-plane pair-format slice write)
    P = wm["P"]
    Xh = P // 2 + 1
    Rp, col_off = P // 2 + 5, 4                    # Geometry::Rp / colOff at pad 2, r 1.9 (K = 4)
    side, pitch = 2 * Rp + 1, (Rp + col_off + 3) & ~1
    k1_alg = box * box * 4 + P * Xh * 8 + wm["npix"] * 8
    k_bytes = {"k_fft_rows (K1r)": (None, box * box * 4 + Xh * box * 8, tm["fft2d_ms"]),
               "k_fft_cols_slices (K1c)": (None, Xh * box * 8 + side * pitch * 16, tm["slice_ms"]),
               "K1 chain (K1r + K1c)": (k1_alg, box * box * 4 + 2 * Xh * box * 8 + side * pitch * 16, tm["fft2d_ms"] + tm["slice_ms"]),
               "k_gather_sticks (K2')": (wm["k2_bytes_per_particle"] + wm["k2_bytes_per_launch_fixed"] / imgs_per_launch, None, tm["gather_ms"]),
               "k_edge2 + k_damped_scatter (K2e', K2r)": (None, None, tm["edge_ms"])}
    per_kernel = []
    for name, (alg, fmt, t_ms) in k_bytes.items():
        e = {"kernel": name, "ms_per_step": t_ms / K, "us_per_particle": 1e3 * t_ms / (K * B)}
        if alg is not None and t_ms > 0:
            gbs = alg * K * B / (t_ms * 1e-3) / 1e9
            e.update({"algorithmic_bytes_per_particle": alg, "achieved_GBps": gbs, "hbm_frac": gbs / hbm_peak})
        if fmt is not None and t_ms > 0:
            e.update({"format_bytes_per_particle": fmt, "format_GBps": fmt * K * B / (t_ms * 1e-3) / 1e9})
        per_kernel.append(e)
    if args.fast:
        per_kernel = None
        # k_fast_insert: one launch per chunk; per lattice column crossing the half-disc of a plane one 16-byte folded pixel
        # is read and V (8 B) + W (4 B) are read-modified-written
        S = 2 * box
        hits = 0.5 * np.pi * (S / 2.0) ** 2
        alg_bytes = imgs_per_launch * hits * (16 + 24)
        roofline = {"kernel": "k_fast_insert", "bound": "hbm", "achieved": alg_bytes / (g_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "frac": alg_bytes / (g_ms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_src,
                    "ms_per_launch": g_ms, "particles_per_launch": imgs_per_launch,
                    "note": "scatter of one nearest pixel per lattice column and plane with FP32 atomics (L2 atomic throughput bound)"}
        fp32 = l1_pipe = None

    # ---------------- end to end through the C ABI with pinned host buffers
    e2e = None
    extra = {}
    if not args.no_e2e:
        host = []
        for img, p in batches:
            hbuf = torch.empty(img.shape, dtype=torch.float32, pin_memory=True)
            hbuf.copy_(img)
            host.append((hbuf, p))
        torch.cuda.synchronize()
        r.reset()
        r.insert_host_ptr(host[0][0].data_ptr(), host[0][1])
        r.weight_sum()
        if world > 1:
            r.reduce(0)          # warm-up of the communicator (connection set-up happens on the first collective)
            if p2p:
                r.reduce_p2p(0)  # and of the peer mappings
        vol_pinned = torch.empty((box, box, box), dtype=torch.float32, pin_memory=True) if rank == 0 else None
        if rank == 0:
            r.finalize_into(vol_pinned.data_ptr())      # creates the 3-D FFT plan and the finalisation buffers
        r.reset()
        barrier()
        t0 = time.perf_counter()
        t_call_ins = t_call_sum = 0.0
        # pipelined host loop: the per-step result (8-byte D2H) of step i-1 is collected after step i has been
        # handed over, so the PCIe transfer of step i overlaps the kernels of step i-1; every step's H2D and D2H
        # are inside the timed region
        results = []
        for i in range(K):
            hbuf, p = host[i & 1]
            ta = time.perf_counter()
            r.insert_host_ptr(hbuf.data_ptr(), p)
            tb = time.perf_counter()
            if i > 0:
                results.append(r.weight_sum_end())
            r.weight_sum_begin()
            tc = time.perf_counter()
            t_call_ins += tb - ta
            t_call_sum += tc - tb
        results.append(r.weight_sum_end())
        assert len(results) == K and all(b > a for a, b in zip(results, results[1:])), "per-step results must grow"
        t_ins = time.perf_counter()
        if world > 1:
            do_reduce(r, p2p)
            r.sync()
        t_red = time.perf_counter()
        vol = None
        if rank == 0:           # the caller owns the output buffer (page-locked here, like the input batches)
            r.finalize_into(vol_pinned.data_ptr())
            vol = vol_pinned.numpy()
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        dt = max_over_ranks(t1 - t0)
        e2e = {"value": world * K * B / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(B * box * box * 4 + B * 24 * 8),
               "d2h_bytes_per_step": int(8 + (box ** 3 * 4) // K),
               "includes": "H2D from pinned host memory, per-step 8-byte read-back (collected one step late: pipelined host loop), final reduce (N>1), normalise + 3-D IFFT + D2H of the volume"}
        extra = {"numa_binding": numa, "e2e_insert_call_ms_per_step": 1e3 * t_call_ins / K, "e2e_result_call_ms_per_step": 1e3 * t_call_sum / K,
                 "e2e_insert_s": t_ins - t0, "e2e_reduce_s": t_red - t_ins, "e2e_finalize_s": t1 - t_red}
        if rank == 0 and vol is not None:
            extra["volume_finite"] = bool(np.isfinite(vol).all())
        tm2 = r.timings()
        extra["gpu_launches_e2e"] = int(tm2["kernel_launches"])
        extra["e2e_stage_ms_per_step"] = {k: tm2[k] / K for k in ("h2d_ms", "preprocess_ms", "fft2d_ms", "slice_ms", "gather_ms", "edge_ms")}
        extra["e2e_h2d_ms_per_step"] = tm2["h2d_ms"] / K
        extra["e2e_h2d_GBps"] = (B * box * box * 4 / 1e9) / (tm2["h2d_ms"] / K * 1e-3) if tm2["h2d_ms"] > 0 else None

    # ---------------- strong scaling, the other BASELINE configurations, reduce check, the reference's own GPU kernel
    headline = box == 256 and args.sym.lower() == "c1" and not args.no_ctf and not args.fast
    if not args.no_extras and not args.no_e2e and headline:
        vol_ptr = vol_pinned.data_ptr() if rank == 0 else 0
        if world > 1:
            # the two collectives side by side on the headline accumulators (content does not matter for the time)
            cmp_ms = {}
            for name, fn in (("nccl", lambda: r.reduce(0)),) + ((("p2p", lambda: r.reduce_p2p(0)),) if p2p else ()):
                ts = []
                for _ in range(3):
                    barrier()
                    tq = time.perf_counter()
                    fn()
                    r.sync()
                    ts.append(max_over_ranks(time.perf_counter() - tq))
                cmp_ms[name] = 1e3 * min(ts)
            extra["reduce_ms"] = dict(cmp_ms, used="p2p" if (p2p and os.environ.get("RFB200_BENCH_REDUCE") == "p2p") else "nccl",
                                      bytes_per_rank=int(12 * r.accumulator_ptrs()[2]))
        extra["strong_scaling"] = measure_strong_scaling(r, host, box, world, rank, args.strong_total, barrier, max_over_ranks, vol_ptr, p2p)
    if p2p:
        r.ipc_release()
        barrier()           # every rank has unmapped its peers before anybody frees
    r.close()
    r = None
    del batches
    if not args.no_e2e:
        del host
    torch.cuda.empty_cache()
    if not args.no_extras and headline:
        def barrier_all():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
        oc = []
        for cfg in OTHER_CONFIGS:
            try:
                oc.append(measure_other_config(cfg, dev, local, rank, world, max_over_ranks, barrier_all))
                if rank == 0 and world == 1 and not args.no_cpu_baseline:
                    # the CPU arm on a small sample of the same configuration (about 2-4 s each)
                    cores = os.cpu_count() or 1
                    n_s = {64: 1000, 128: 400, 256: 32, 512: 32}[cfg["box"]]
                    rate, dt = time_oracle(cfg["box"], n_s, cores, ctf=cfg["ctf"], seed=300, sym=cfg["sym"])
                    oc[-1]["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                              "sample": "%d particles in %.1f s (oracle port, reference thread scheme)" % (n_s, dt)}
            except Exception as e:      # one configuration failing must not take the headline line with it
                oc.append({"workload": cfg["name"], "error": "%s: %s" % (type(e).__name__, e)})
        extra["other_configs"] = oc
        if world > 1:
            extra["reduce_check"] = reduce_check(dev, local, rank, world)
        if rank == 0 and world == 1:
            try:
                extra["ref_gpu_kernel"] = measure_ref_gpu_kernel(box)
                extra["ref_gpu_kernel"]["ours_over_reference_kernel"] = value / extra["ref_gpu_kernel"]["value"] if extra["ref_gpu_kernel"].get("value") else None
            except Exception as e:
                extra["ref_gpu_kernel"] = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        rate0, _ = time_oracle(box, max(2 * cores, 32), cores, ctf=not args.no_ctf)
        n_s = int(min(4000, max(64, rate0 * 15)))
        rate, dt = time_oracle(box, n_s, cores, ctf=not args.no_ctf)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d particles of the same workload in %.1f s (oracle/recfourier_oracle.cpp, double, reference thread scheme)" % (n_s, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "config[2]: %dx%d Gaussian-phantom projections %s, padding 2, %s, blob 1.9/0/15, max_resolution 0.5%s" % (box, box, "without CTF" if args.no_ctf else "with CTF", args.sym.upper(), ", --fast arithmetic" if args.fast else ""),
                       "box": box, "padding": 2, "sym": args.sym.lower(), "insertions_per_particle": n_ops, "ctf": not args.no_ctf, "particles_per_step_per_gpu": B,
                       "l2": "inputs larger than L2: each step reads a %.2f GB batch, two batches alternate" % (B * box * box * 4 / 1e9),
                       "parallelism": "particle sharding, %d rank(s), one reduce of V and W onto rank 0 before normalisation" % world,
                       "stage_ms_per_step": stage_ms},
            "clocks": cs, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "per_kernel_roofline": per_kernel, "fp32": fp32, "l1_data_pipe": l1_pipe, "cpu_baseline": cpu_baseline,
        }
        line.update(extra)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if r is not None:
        r.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
