#!/usr/bin/env python
"""Benchmark of the spot-analysis hot path (BASELINE.json metric: Gvoxel/s for
filter + spot detect + EDT, and achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|small]

One "step" = one pass of the hot path over one synthetic stack: min/max normalise -> LoG
(sigma 2) -> 26-neighbour NMS + threshold + compaction -> overlap prune -> fibre-mask filter ->
P25 baseline -> Gaussian blur at spots -> SNR filter, then the anisotropic EDT of the
NOT-nucleus mask and the per-spot distance gather.

* ``value``  : voxels processed / device time with the stack already resident in HBM.
* ``e2e``    : the same step from HOST buffers through the drop-in host API
               (muscle_fish.find_spots_snrfilter_with_distances: pinned float32 (x,y,z) image + uint8
               masks in, host float64 spots / table / distances out), H2D and D2H inside the timed region.
               ``e2e_variants`` holds the other ways a caller can hand the data over: the two
               reference-signature calls one after the other, raw uint16 counts, and what
               general_pipeline/fish_analysis.py:196-199 actually passes (pageable float64 + int64 masks).
* ``batch``  : BASELINE configs[2] -- 64 stacks of config-1 geometry through parallel.run_batch, every
               N-th stack per rank, double-buffered pinned staging (host buffers in, host results out).
* ``slab``   : (N > 1) BASELINE configs[3] -- one mosaic 16384x2048x200 split into z-slabs, and into x-slabs
               beside them (parallel.slab_analyze: halo exchange and the all-to-all before the EDT's envelope
               passes over NVLink peer memory, NCCL for the small collectives), strong scaling, per-phase
               times, plus a check of a quarter mosaic against a single-GPU run of the same data.
* N > 1      : one process per GPU (torchrun); ``value`` / ``e2e`` run independent stacks per rank, no
               data-path collective (the reference's batch_process.py model) -> "scaling": "weak".
* ``--impl reference``: the reference's CPU arithmetic (oracle: scipy.ndimage float64) on the
               box's host cores, one bounded sample stack per worker process.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Gvoxel/s (filter+spot detect+EDT)"
UNIT = "Gvoxel/s"
SIGMA, T_SPOT = 2.0, 0.025
SAMPLING = (0.0577, 0.0577, 0.5)

WORKLOADS = {
    # name: (shape_xyz, n_spots, n_nuclei, seed, description)
    "C2": ((2048, 2048, 100), 20000, 120, 2, "configs[1]: single fiber stack 2048x2048x100 fp32, LoG spot detect "
                                             "+ anisotropic EDT"),
    "C1": ((1024, 1024, 60), 5000, 30, 1234, "configs[0]: 1024x1024x60 stack, 5k spots, 30 nuclei"),
    "small": ((256, 192, 40), 300, 6, 11, "debug stack 256x192x40"),
}
MOSAIC = ((16384, 2048, 200), 640000, 4000, 4)  # configs[3]

# algorithmic bytes per voxel of each kernel (fp32 volumes, u8 masks)
ALGO_BYTES = {
    "minmax": 4, "conv_y_gd": 12, "conv_z_mid": 16, "conv_x_last": 12, "nms_compact": 4,
    "masked_quantile_3pass": 15, "masked_quantile_bracketed": 5, "masked_quantile_minmax": 5, "edt_scan_y": 3,
    "edt_envelope_x": 8, "edt_envelope_z": 10, "ingest_xyz": 8, "ingest_flat": 8,
}


def workload_config(name):
    """The ``config`` object of the JSON line -- identical for both arms."""
    shape, n_spots, n_nuc, _, desc = WORKLOADS[name]
    return {"workload": desc, "shape_xyz": list(shape), "spots": n_spots, "nuclei": n_nuc, "sigma": SIGMA,
            "t_spot": T_SPOT, "sampling_um": list(SAMPLING), "stacks_per_gpu": 1,
            "l2": "inputs (1.7 GB/stack at C2) larger than the 126 MB L2; no explicit flush"}


_JSON_FD = None


def _capture_stdout():
    """Everything libraries print to stdout (NCCL banners, ...) goes to stderr; the one JSON line is
    written to the real stdout by _emit()."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
            return
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def _loop(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU (reference) side
def _cpu_pipeline(st):
    import oracle
    spots, _ = oracle.find_spots_snrfilter(st["img"], sigma=SIGMA, t_spot=T_SPOT, mask=st["fiber_mask"],
                                           pair_order="set")
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=SAMPLING)
    return len(spots), float(oracle.edt_gather(grass, spots, SAMPLING).sum())


_WORKER_STACK = None


def _worker_init(shape, n_spots, n_nuc, seed_base):
    global _WORKER_STACK
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    ident = os.getpid()
    _WORKER_STACK = synth.make_stack(shape=shape, n_spots=n_spots, n_nuclei=n_nuc, seed=seed_base + ident % 1000)


def _worker_step(_):
    return _cpu_pipeline(_WORKER_STACK)


def _sample_geometry(workload, sample_xy):
    shape, n_spots, n_nuc, seed, _ = WORKLOADS[workload]
    sx = min(sample_xy, shape[0])
    sy = min(sample_xy, shape[1])
    frac = (sx * sy) / float(shape[0] * shape[1])
    return (sx, sy, shape[2]), max(8, int(round(n_spots * frac))), max(1, int(round(n_nuc * frac))), seed


def cpu_baseline_single(workload="C1"):
    """The oracle (scipy float64, what the reference's libraries compute), 1 process / 1 thread, on the whole
    configs[0] stack -- the CPU-runnable case BASELINE.md section 4 plans the CPU arm on."""
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    shape, n_spots, n_nuc, seed, desc = WORKLOADS[workload]
    st = synth.make_stack(shape=shape, n_spots=n_spots, n_nuclei=n_nuc, seed=seed)
    t0 = time.perf_counter()
    found, _ = _cpu_pipeline(st)
    dt = time.perf_counter() - t0
    nvox = float(np.prod(shape))
    return {"value": nvox / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{desc}: the whole {shape[0]}x{shape[1]}x{shape[2]} stack once, {found} spots kept, "
                      f"{dt:.1f} s of scipy.ndimage float64 on one core (spot path + EDT + gather)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    ncores = len(os.sched_getaffinity(0))
    workers = max(1, min(ncores, args.ref_workers or 64))
    shape, n_spots, n_nuc, seed = _sample_geometry(args.workload, args.ref_sample)
    ctx = mp.get_context("fork")
    with ctx.Pool(workers, initializer=_worker_init, initargs=(shape, n_spots, n_nuc, seed)) as pool:
        pool.map(_worker_step, range(workers))  # make sure every worker is up (stack generated)
        for _ in range(args.warmup):
            pool.map(_worker_step, range(workers), chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_step, range(workers), chunksize=1)
        dt = time.perf_counter() - t0
    nvox = float(np.prod(shape)) * workers * args.steps
    value = nvox / dt / 1e9
    sample = (f"each step: {workers} worker processes (the reference's only parallel mode is one image per "
              f"process, general_pipeline/batch_process.py) x one {shape[0]}x{shape[1]}x{shape[2]} crop-sized stack "
              f"at the workload's spot / nucleus density ({n_spots} spots, {n_nuc} nuclei); oracle = "
              f"scipy.ndimage float64; throughput is per voxel")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------------- GPU side
class _Ctx:
    """rank / world / collectives of one bench process"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup):
        """wall clock of `steps` calls, barrier + synchronize on both sides, max over ranks (seconds)"""
        out = None
        for _ in range(warmup):
            out = fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        self.barrier()
        return self.max_over_ranks(time.perf_counter() - t0), out


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def bench_e2e(ctx, args, img, fiber, nuc, shape, nvox):
    """host buffers in, host results out.  Returns (e2e object, variants)"""
    import torch
    from muscle_fish_b200 import muscle_fish as mf, ndimage as nd
    h_img = torch.empty(shape, dtype=torch.float32).pin_memory()
    h_fib = torch.empty(shape, dtype=torch.uint8).pin_memory()
    h_nuc = torch.empty(shape, dtype=torch.uint8).pin_memory()
    h_img.copy_(img.permute(1, 2, 0))
    h_fib.copy_(fiber.permute(1, 2, 0))
    h_nuc.copy_(nuc.permute(1, 2, 0))
    torch.cuda.synchronize()
    a_img, a_fib, a_nuc = h_img.numpy(), h_fib.numpy(), h_nuc.numpy()

    def fused(ai=a_img, af=a_fib, an=a_nuc):
        return _quiet(mf.find_spots_snrfilter_with_distances, ai, an, SAMPLING, sigma=SIGMA, t_spot=T_SPOT, mask=af)

    def two_calls():
        spots, data = _quiet(mf.find_spots_snrfilter, a_img, sigma=SIGMA, t_spot=T_SPOT, mask=a_fib)
        return spots, data, nd.spot_distances(a_nuc, spots, SAMPLING)

    def d2h_bytes(res):
        spots, data, d = res
        return int(spots.size * 8 + d.size * 8 + (len(data) - 1) * 6 * 8)

    steps = max(1, min(args.steps, args.e2e_steps))
    warm = max(1, min(args.warmup, 3))
    from muscle_fish_b200 import device as D

    def counted(fn, nsteps, nwarm):
        """ctx.timed + the bytes host_to_device_many queued per call (masks cross PCIe as bits: device.PackedMask)"""
        b0 = D.H2D_BYTES
        dt_, res_ = ctx.timed(fn, nsteps, nwarm)
        return dt_, res_, int((D.H2D_BYTES - b0) // (nsteps + nwarm))

    dt, res, h2d = counted(fused, steps, warm)
    h2d_bytes_whole = int(h_img.numel() * 4 + h_fib.numel() + h_nuc.numel())
    # the ceiling the host side sets: the same three pinned buffers copied to the device and nothing else, every
    # rank at once (with N ranks on one host they share its memory system and PCIe root complexes)
    d_img = torch.empty(shape, dtype=torch.float32, device="cuda")
    d_fib = torch.empty(shape, dtype=torch.uint8, device="cuda")
    d_nuc = torch.empty(shape, dtype=torch.uint8, device="cuda")

    def copies_only():
        d_nuc.copy_(h_nuc, non_blocking=True)
        d_img.copy_(h_img, non_blocking=True)
        d_fib.copy_(h_fib, non_blocking=True)
        torch.cuda.synchronize()

    dtc, _ = ctx.timed(copies_only, 3, 1)
    del d_img, d_fib, d_nuc
    h2d_only_ms = dtc / 3 * 1e3
    e2e = {"value": ctx.world * nvox * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h_bytes(res), "steps": steps, "ms_per_step": dt / steps * 1e3,
           "pcie_floor_ms": h2d / 55e9 * 1e3, "masks": "bit-packed by host threads inside the timed region"
           if D.PACK_MASKS else "whole bytes",
           "h2d_bytes_whole_byte_masks": h2d_bytes_whole,
           "h2d_only_ms_all_ranks_at_once": h2d_only_ms,
           "h2d_only_gbs_per_gpu": h2d_bytes_whole / h2d_only_ms / 1e6,
           "api": "muscle_fish.find_spots_snrfilter_with_distances (= find_spots_snrfilter + the EDT/interpn pair of "
                  "nocodazole/fish_analysis_noco.py:218-221,297-298 in one call): pinned host float32 (x,y,z) image "
                  "+ uint8 masks in, host float64 spots / table / distances out"}
    variants = {}
    vsteps = max(1, min(steps, 3))
    dt2, res2, hb2 = counted(two_calls, vsteps, 1)
    variants["two_reference_calls_f32_pinned"] = {
        "value": ctx.world * nvox * vsteps / dt2 / 1e9, "ms_per_step": dt2 / vsteps * 1e3, "h2d_bytes_per_step": hb2,
        "api": "muscle_fish.find_spots_snrfilter, then ndimage.spot_distances (round-1 e2e)",
        "same_spots_as_fused": bool(np.array_equal(res2[0], res[0]) and np.array_equal(res2[2], res[2]))}
    if not args.quick:
        # raw detector counts: uint16 over PCIe, no widening on the host
        u_img = torch.empty(shape, dtype=torch.uint16).pin_memory()
        u_img.copy_((img.permute(1, 2, 0) * 16.0).round().clamp(0, 65535).to(torch.uint16))
        torch.cuda.synchronize()
        au = u_img.numpy()
        dt3, res3, hb = counted(lambda: fused(au), vsteps, 1)
        variants["uint16_pinned"] = {"value": ctx.world * nvox * vsteps / dt3 / 1e9, "ms_per_step": dt3 / vsteps * 1e3,
                                     "h2d_bytes_per_step": hb, "pcie_floor_ms": hb / 55e9 * 1e3,
                                     "found_spots": int(len(res3[0])),
                                     "api": "same call, the image as uint16 counts (x16 gain, rounded)"}
        del u_img, au
        # what general_pipeline/fish_analysis.py:196-199 hands over: pageable float64 (fix_bleaching's output)
        # and int64 0/1 masks (np.where(..., 1, 0))
        p_img = a_img.astype(np.float64)
        p_fib = a_fib.astype(np.int64)
        p_nuc = a_nuc.astype(np.int64)
        dt4, res4, hb = counted(lambda: fused(p_img, p_fib, p_nuc), 2, 1)
        variants["float64_int64_pageable"] = {
            "value": ctx.world * nvox * 2 / dt4 / 1e9, "ms_per_step": dt4 / 2 * 1e3, "h2d_bytes_per_step": hb,
            "same_spots_as_fused": bool(np.array_equal(res4[0], res[0])),
            "api": "same call, pageable float64 image + int64 masks: the dtypes "
                   "general_pipeline/fish_analysis.py:196-199 passes"}
        del p_img, p_fib, p_nuc
    return e2e, variants


def bench_batch(ctx, args):
    """configs[2]: 64 stacks of configs[0] geometry, one process per GPU taking every N-th stack"""
    import torch
    from muscle_fish_b200 import parallel as P, synth
    shape, n_spots, n_nuc, seed, _ = WORKLOADS["C1" if not args.quick else "small"]
    n_stacks = args.batch_stacks
    pool = []
    for k in range(4):  # four distinct stacks in pinned host memory stand for the image files
        st = synth.make_stack_device(shape, n_spots, n_nuc, 1000 + k, z_first=False)
        u = (st["img"] * 16.0).round().clamp(0, 65535).to(torch.uint16)
        pool.append(tuple(t.cpu().pin_memory() for t in (u, st["fiber_mask"], st["nuc_mask"])))
        del st, u
    torch.cuda.synchronize()

    def loader(i, img, fiber, nuc):  # "reading file i": the pinned pool buffer is handed over as it is
        a, b, c = pool[i % len(pool)]
        return a, b, c

    found = []

    def once():
        found.clear()
        res = P.run_batch(n_stacks, loader, shape, torch.uint16, SIGMA, T_SPOT, SAMPLING, rank=ctx.rank,
                          world=ctx.world)
        found.extend(len(r["spots"]) for _, r in res)
        return res

    dt, res = ctx.timed(once, 1, 1)
    nvox = float(np.prod(shape))
    from muscle_fish_b200 import _lib as L, device as D
    packed = D.PACK_MASKS
    per_stack = int(nvox) * 2 + (2 * int(L.lib().mfish_mask_bits_bytes(int(nvox))) if packed else 2 * int(nvox))
    return {"workload": f"configs[2]: {n_stacks} stacks {shape[0]}x{shape[1]}x{shape[2]} (uint16 image + two uint8 masks "
                        f"per stack in pinned host memory{'; the masks cross PCIe as bits, packed by host threads inside the timed region' if packed else ''}), "
                        f"every {ctx.world}-th stack per GPU, parallel.run_batch",
            "n_stacks": n_stacks, "stacks_this_rank": len(res), "value": n_stacks * nvox / dt / 1e9, "unit": UNIT,
            "ms_total": dt * 1e3, "ms_per_stack_per_gpu": dt * 1e3 / max(1, len(res)),
            "h2d_bytes_per_stack": per_stack, "pcie_floor_ms_per_stack": per_stack / 55e9 * 1e3,
            "spots_first_stack": found[0] if found else None, "scaling": "strong (fixed batch)"}


def _mosaic_for(world, quick):
    (X, Y, Z), ns, nn, seed = MOSAIC
    if quick:
        return (1024, 512, 64), 2500, 16, seed
    # ~63 bytes of device memory per voxel per rank (extended slab, LoG scratch, EDT workspace, re-partition
    # buffers): keep a rank below ~150 GB of the 180 GB
    while X * Y * Z / world * 63 > 150e9 and X > 2048:
        X //= 2
        ns //= 2
        nn //= 2
    return (X, Y, Z), ns, nn, seed


def bench_slab(ctx, args):
    """configs[3]: one mosaic split across the ranks (strong scaling) through parallel.slab_analyze, in the
    mandated z-slab decomposition and in the x-slab decomposition beside it; per-phase times; and a check of
    each decomposition against one GPU on a mosaic that fits one GPU"""
    import torch
    import torch.distributed as dist
    from muscle_fish_b200 import device as D, parallel as P, synth
    world, rank = ctx.world, ctx.rank
    edt_group = dist.new_group()

    def slabs(kind, shape, ns, nn, seed):
        X, Y, Z = shape
        if kind == "z":
            lay = P.SlabLayout(Z, X, Y)
            st = synth.make_stack_device(shape, ns, nn, seed, z_range=(lay.z0, lay.z1))
        else:
            lay = P.XSlabLayout(Z, X, Y)
            st = synth.make_stack_device(shape, ns, nn, seed, x_range=(lay.x0, lay.x1))
        notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
        img = st["img"]
        if kind == "z":  # the resident slab sits where the halo exchange extends it (no per-step copy of it)
            home = P.slab_input_buffer(lay, SIGMA, img.dtype, img.device)
            if home is not None:
                home.copy_(img)
                img = home
        return lay, img, st["fiber_mask"], notnuc

    def cleanup():
        P.clear_buffers()
        D.WS.clear()
        torch.cuda.empty_cache()

    # -- 1. same result as one GPU (a quarter-length mosaic fits one B200)
    cshape, cns, cnn, cseed = ((4096, 2048, 200), 160000, 1000, 4) if not args.quick else ((512, 256, 48), 600, 8, 4)
    got = {}
    for kind in ("z", "x"):
        lay, img, fib, notnuc = slabs(kind, cshape, cns, cnn, cseed)
        got[kind] = P.slab_analyze(img, fib, notnuc, lay, sigma=SIGMA, t_spot=T_SPOT, sampling_xyz=SAMPLING,
                                   edt_group=edt_group)
        got[kind].pop("dist_slab", None)
        got[kind].pop("dist_xslab", None)
        del img, fib, notnuc
        cleanup()
    verdict = torch.zeros(6, dtype=torch.float64, device="cuda")
    if rank == 0:
        st = synth.make_stack_device(cshape, cns, cnn, cseed)
        nn1 = (st["nuc_mask"] == 0).to(torch.uint8)
        one = D.analyze_stack(st["img"], st["fiber_mask"], nn1, SIGMA, T_SPOT, SAMPLING)
        flags = []
        for kind in ("z", "x"):
            same_spots = np.array_equal(one["spots"], got[kind]["spots"])
            flags += [float(same_spots), float(same_spots and np.array_equal(one["spot_dist"], got[kind]["spot_dist"]))]
        verdict[:] = torch.tensor(flags + [len(one["spots"]), len(got["z"]["spots"])])
        del st, nn1, one
        cleanup()
    dist.broadcast(verdict, 0)
    v = verdict.tolist()
    matches = {k: {"spots_identical": bool(v[2 * i]), "distances_identical": bool(v[2 * i + 1])}
               for i, k in enumerate(("z_slabs", "x_slabs"))}
    matches.update({"mosaic": list(cshape), "spots_one_gpu": int(v[4]), "spots_slabs": int(v[5]),
                    "how": f"spot list (coordinates, order) and per-spot distances of parallel.slab_analyze on {world} "
                           "ranks == device.analyze_stack on rank 0 over the same voxels, bit for bit"})
    ctx.barrier()

    # -- 2. timing on the mosaic itself, both decompositions
    shape, ns, nn, seed = _mosaic_for(world, args.quick)
    X, Y, Z = shape
    nvox = float(X) * Y * Z
    steps = max(1, min(args.steps, args.slab_steps))
    out = {"matches_n1": matches, "mosaic": [X, Y, Z], "n_gpus": world, "unit": UNIT, "steps": steps,
           "scaling": "strong", "nvlink_peer_copy_reference_gbs": 770.0,
           "exchange": "peer memory (mfish_copy2d into symmetric buffers)" if P._PEER_STATE["ok"] is not False and
           os.environ.get("MFISH_P2P", "1") != "0" else "NCCL send/recv"}
    radius = 8
    for kind in ("z", "x"):
        lay, img, fib, notnuc = slabs(kind, shape, ns, nn, seed)

        def step():
            return P.slab_analyze(img, fib, notnuc, lay, sigma=SIGMA, t_spot=T_SPOT, sampling_xyz=SAMPLING,
                                  edt_group=edt_group, full_table=False)

        torch.cuda.reset_peak_memory_stats()
        dt, res = ctx.timed(step, steps, 2)
        ms = dt / steps * 1e3
        # single steps as well: side by side, the two chains' cross-rank waits make a step vary by several ms
        singles = [ctx.timed(step, 1, 0)[0] * 1e3 for _ in range(6)]
        # one more step with per-phase wall clocks (synchronising: the two chains then run one after the other)
        # (an untimed one first: without the second communicator the EDT's exchange makes new connections)
        P.slab_analyze(img, fib, notnuc, lay, sigma=SIGMA, t_spot=T_SPOT, sampling_xyz=SAMPLING, full_table=False)
        P.PHASES = {}
        P.slab_analyze(img, fib, notnuc, lay, sigma=SIGMA, t_spot=T_SPOT, sampling_xyz=SAMPLING, full_table=False)
        phases = {k: round(val * 1e3, 3) for k, val in P.PHASES.items()}
        P.PHASES = None
        if kind == "z":
            halo_bytes = 2 * (radius + 1) * X * Y * 4  # received by an interior rank
            a2a_bytes = (Z / world) * X * Y * 2 * (world - 1) / world  # sent by every rank (uint16 row distances)
        else:
            halo_bytes = 2 * 12 * Z * Y * 4
            a2a_bytes = Z * (X / world) * Y * 2 * (world - 1) / world  # uint16 row distances
        gbs = lambda nbytes, key: (nbytes / (phases[key] * 1e-3) / 1e9) if phases.get(key, 0) > 0 else None
        out[kind + "_slabs"] = {
            "workload": f"configs[3]: mosaic {X}x{Y}x{Z} in {world} {kind}-slabs, parallel.slab_analyze (spot path + "
                        "EDT + per-spot distances), device-resident slabs",
            "ms_per_step": ms, "value": nvox / ms / 1e6, "found_spots": int(len(res["spots"])),
            "ms_single_steps": [round(v, 2) for v in singles], "ms_median_single": float(np.median(singles)),
            "phases_ms_serialised": phases,
            "halo_bytes_in_per_rank": int(halo_bytes), "halo_gbs": gbs(halo_bytes, "halo"),
            "all_to_all_bytes_out_per_rank": int(a2a_bytes), "all_to_all_gbs": gbs(a2a_bytes, "edt_all_to_all"),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
        }
        del img, fib, notnuc, res
        cleanup()
        ctx.barrier()
    # the headline of the object: the mandated decomposition; the comparison beside it
    out["ms_per_step"] = out["z_slabs"]["ms_per_step"]
    out["value"] = out["z_slabs"]["value"]
    out["workload"] = out["z_slabs"]["workload"]
    return out


def run_ours(args):
    ctx = _Ctx()
    torch, dist = ctx.torch, ctx.dist
    rank, world, local = ctx.rank, ctx.world, ctx.local

    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import _lib, device as D, parallel as P, synth
    P.bind_to_gpu_numa(local)

    shape, n_spots, n_nuc, seed, desc = WORKLOADS[args.workload]
    nvox = float(np.prod(shape))
    st = synth.make_stack_device(shape, n_spots, n_nuc, seed + rank, device="cuda", z_first=True)
    img, fiber, nuc = st["img"], st["fiber_mask"], st["nuc_mask"]
    notnuc = (nuc == 0).to(torch.uint8)  # bench input preparation (not timed)
    torch.cuda.synchronize()

    def step():
        return D.analyze_stack(img, fiber, notnuc, SIGMA, T_SPOT, SAMPLING)

    # ---- device-resident throughput
    res = None
    for _ in range(args.warmup):
        res = step()
    ctx.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.profile_enable(True)
    _lib.profile_collect()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    ctx.barrier()
    launches = _lib.launch_count() - n0
    kern = _lib.profile_collect()
    _lib.profile_enable(False)
    clocks = sampler.stop()
    ms_step = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * nvox / (ms_step * 1e-3) / 1e9
    found = int(len(res["spots"]))

    # ---- end to end through the drop-in host API
    D.WS.clear()
    torch.cuda.empty_cache()
    e2e, variants = bench_e2e(ctx, args, img, fiber, nuc, shape, nvox)
    del st, img, fiber, nuc, notnuc
    D.WS.clear()
    torch.cuda.empty_cache()

    batch = bench_batch(ctx, args) if not args.no_batch else None
    D.WS.clear()
    torch.cuda.empty_cache()
    slab = bench_slab(ctx, args) if (world > 1 and not args.no_slab) else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    kernels = {}
    for name, (calls, ms) in kern.items():
        per = ms / max(calls, 1)
        entry = {"calls_per_step": calls / args.steps, "ms": per}
        if name in ALGO_BYTES:
            gbs = ALGO_BYTES[name] * nvox / (per * 1e-3) / 1e9
            entry.update({"bytes_per_voxel": ALGO_BYTES[name], "gbs": gbs, "frac": gbs / peak})
        kernels[name] = entry
    share = {n: k["ms"] * k["calls_per_step"] for n, k in kernels.items() if n in ALGO_BYTES}
    dom = max(share, key=share.get)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES[dom] * nvox,
                "share_of_step": share[dom] / ms_step}
    cpu = cpu_baseline_single("C1" if not args.quick else "small") if (world == 1 and not args.no_cpu_baseline) else None
    cfg = workload_config(args.workload)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg, "found_spots": found, "clocks": clocks, "e2e": e2e,
        "e2e_variants": variants, "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
        "cpu_baseline": cpu, "batch": batch, "slab": slab,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--slab-steps", type=int, default=10)
    ap.add_argument("--batch-stacks", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--quick", action="store_true", help="small shapes for the batch / slab parts, no e2e variants "
                                                         "(debugging the harness)")
    ap.add_argument("--ref-sample", type=int, default=320, help="x/y edge of the per-worker CPU sample stack")
    ap.add_argument("--ref-workers", type=int, default=0)
    args = ap.parse_args()
    _capture_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
