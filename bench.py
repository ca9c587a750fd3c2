#!/usr/bin/env python
"""Benchmark of the spot-analysis hot path (BASELINE.json metric: Gvoxel/s for
filter + spot detect + EDT, and achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|small]

One "step" = one pass of the hot path over one synthetic stack: min/max normalise -> LoG
(sigma 2) -> 26-neighbour NMS + threshold + compaction -> overlap prune -> fibre-mask filter ->
P25 baseline -> Gaussian blur at spots -> SNR filter, then the anisotropic EDT of the
NOT-nucleus mask and the per-spot distance gather.

* ``value``  : voxels processed / device time with the stack already resident in HBM.
* ``e2e``    : same step through the drop-in host API (muscle_fish.find_spots_snrfilter +
               ndimage.spot_distances) with pinned HOST buffers in, host results out.
* N > 1      : one process per GPU (torchrun), independent stacks per rank, no data-path
               collective (the reference's batch_process.py model) -> "scaling": "weak".
* ``--impl reference``: the reference's CPU arithmetic (oracle: scipy.ndimage float64) on the
               box's host cores, one stack per worker process.
"""
from __future__ import annotations

import argparse
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

# algorithmic bytes per voxel of each kernel (fp32 volumes, u8 masks)
ALGO_BYTES = {
    "minmax": 4, "conv_y_gd": 12, "conv_z_mid": 16, "conv_x_last": 12, "nms_compact": 4,
    "masked_quantile_3pass": 15, "masked_quantile_bracketed": 5, "masked_quantile_minmax": 5, "edt_scan_y": 3, "edt_envelope_x": 8, "edt_envelope_z": 10,
    "ingest_xyz": 8, "ingest_flat": 8,
}


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


def cpu_baseline_single(workload, sample_xy=448):
    """oracle, 1 process / 1 thread, on a crop-sized sample of the workload (same density)."""
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    shape, n_spots, n_nuc, seed = _sample_geometry(workload, sample_xy)
    st = synth.make_stack(shape=shape, n_spots=n_spots, n_nuclei=n_nuc, seed=seed)
    t0 = time.perf_counter()
    _cpu_pipeline(st)
    dt = time.perf_counter() - t0
    nvox = float(np.prod(shape))
    return {"value": nvox / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{shape[0]}x{shape[1]}x{shape[2]} stack at the workload's spot/nucleus density, "
                      f"{n_spots} spots, {n_nuc} nuclei, {dt:.1f} s of scipy float64 on one core"}


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
    sample = (f"each step: {workers} worker processes x one {shape[0]}x{shape[1]}x{shape[2]} stack "
              f"({n_spots} spots, {n_nuc} nuclei; the workload's density), oracle = scipy.ndimage float64")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "sigma": SIGMA, "t_spot": T_SPOT,
                   "sampling_um": SAMPLING, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------------- GPU side
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import _lib, device as D, synth
    from muscle_fish_b200 import muscle_fish as mf, ndimage as nd

    shape, n_spots, n_nuc, seed, desc = WORKLOADS[args.workload]
    nvox = float(np.prod(shape))
    st = synth.make_stack_device(shape, n_spots, n_nuc, seed + rank, device="cuda", z_first=True)
    img, fiber, nuc = st["img"], st["fiber_mask"], st["nuc_mask"]
    notnuc = (nuc == 0).to(torch.uint8)  # bench input preparation (not timed)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return D.analyze_stack(img, fiber, notnuc, SIGMA, T_SPOT, SAMPLING)

    # ---- device-resident throughput
    res = None
    for _ in range(args.warmup):
        res = step()
    barrier()
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
    barrier()
    launches = _lib.launch_count() - n0
    kern = _lib.profile_collect()
    _lib.profile_enable(False)
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps
    value = world * nvox / (ms_step * 1e-3) / 1e9

    # ---- end to end through the drop-in host API, pinned host buffers
    h_img = torch.empty(shape, dtype=torch.float32).pin_memory()
    h_fib = torch.empty(shape, dtype=torch.uint8).pin_memory()
    h_nuc = torch.empty(shape, dtype=torch.uint8).pin_memory()
    h_img.copy_(img.permute(1, 2, 0))
    h_fib.copy_(fiber.permute(1, 2, 0))
    h_nuc.copy_(nuc.permute(1, 2, 0))
    torch.cuda.synchronize()
    a_img, a_fib, a_nuc = h_img.numpy(), h_fib.numpy(), h_nuc.numpy()
    del st
    D.WS.clear()
    torch.cuda.empty_cache()

    def step_e2e():
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            spots, _ = mf.find_spots_snrfilter(a_img, sigma=SIGMA, t_spot=T_SPOT, mask=a_fib)
        d = nd.spot_distances(a_nuc, spots, SAMPLING)
        return spots, d

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(max(1, min(args.warmup, 3))):
        spots_e, d_e = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        spots_e, d_e = step_e2e()
    barrier()
    dt = time.perf_counter() - t0
    tm = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    e2e_value = world * nvox * e2e_steps / float(tm.item()) / 1e9
    h2d = int(h_img.numel() * 4 + h_fib.numel() + h_nuc.numel())
    d2h = int(spots_e.size * 8 + d_e.size * 8 + len(spots_e) * 16)

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
    cpu = cpu_baseline_single(args.workload) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "shape_xyz": list(shape), "spots": n_spots, "nuclei": n_nuc, "sigma": SIGMA,
                   "t_spot": T_SPOT, "sampling_um": list(SAMPLING), "stacks_per_gpu": 1,
                   "l2": "inputs (1.7 GB/stack at C2) larger than the 126 MB L2; no explicit flush",
                   "found_spots": int(len(res["spots"]))},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "muscle_fish.find_spots_snrfilter + ndimage.spot_distances, pinned host "
                                           "numpy (x,y,z) in, host float64 out"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": cpu,
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
