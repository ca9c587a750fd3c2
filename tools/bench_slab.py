"""z-slab decomposed mosaic (BASELINE configs[3]): spot detection + SNR filter + EDT + spot distances on
one volume split across the ranks.  torchrun --nproc-per-node N tools/bench_slab.py [--shape X Y Z] ...
Prints one JSON line from rank 0 (strong scaling: the mosaic is fixed, N varies)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import muscle_fish_b200
from muscle_fish_b200 import parallel as P, synth, device as D

ap = argparse.ArgumentParser()
ap.add_argument("--shape", type=int, nargs=3, default=[16384, 2048, 200])
ap.add_argument("--spots", type=int, default=640000)
ap.add_argument("--nuclei", type=int, default=4000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--phases", action="store_true")
ap.add_argument("--cprofile", action="store_true")
ap.add_argument("--no-overlap", action="store_true", help="skip the second measurement (EDT chain on a side "
                "stream with its own communicator, beside the spot chain)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
edt_group = dist.new_group()  # second communicator: the EDT's all-to-all runs beside the spot chain's collectives
X, Y, Z = args.shape
lay = P.SlabLayout(Z, X, Y)
st = synth.make_stack_device((X, Y, Z), args.spots, args.nuclei, 4, z_range=(lay.z0, lay.z1))
img, fib = st["img"], st["fiber_mask"]
notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
del st
SAMPLING = (0.0577, 0.0577, 0.5)

def step():
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = P.slab_find_spots_snrfilter(img, fib, lay, sigma=2.0, t_spot=0.025, full_table=False)
    torch.cuda.synchronize(); t["spots"] = time.perf_counter() - t0; t0 = time.perf_counter()
    dmap = P.slab_edt(notnuc, lay, SAMPLING)
    torch.cuda.synchronize(); t["edt"] = time.perf_counter() - t0; t0 = time.perf_counter()
    sd = P.slab_spot_distances(dmap, res["spots"], lay)
    torch.cuda.synchronize(); t["gather"] = time.perf_counter() - t0
    return res, sd, t

def step_overlap():
    """the distance transform does not depend on the spots: queue its chain (no host round trips) on a
    side stream / second communicator first, then run the spot chain; join before the gather"""
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = P.slab_analyze(img, fib, notnuc, lay, sigma=2.0, t_spot=0.025, sampling_xyz=SAMPLING, edt_group=edt_group,
                         full_table=False)
    torch.cuda.synchronize(); t["all"] = time.perf_counter() - t0
    return res, res["spot_dist"], t

for _ in range(args.warmup):
    step()
dist.barrier(); torch.cuda.synchronize()
if args.phases:
    P.PHASES = {}
t0 = time.perf_counter()
parts = {}
if args.cprofile and rank == 0:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
for _ in range(args.steps):
    res, sd, t = step()
    for k, v in t.items():
        parts[k] = parts.get(k, 0.0) + v
if args.cprofile and rank == 0:
    pr.disable(); pstats.Stats(pr, stream=sys.stderr).sort_stats("cumtime").print_stats(30)
dist.barrier(); torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(dt.item()) / args.steps * 1e3
    print(json.dumps({"workload": f"z-slab mosaic {X}x{Y}x{Z}", "n_gpus": world, "ms_per_step": ms,
                      "gvox_per_s": X * Y * Z / ms / 1e6, "spots": int(len(res["spots"])),
                      "stage_ms": {k: v / args.steps * 1e3 for k, v in parts.items()},
                      "phases_ms": {k: round(v / args.steps * 1e3, 2) for k, v in (P.PHASES or {}).items()},
                      "halo_planes": 9, "slab_planes": lay.z1 - lay.z0,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
if not args.no_overlap and not args.phases:
    sd_serial = sd
    for _ in range(args.warmup):
        step_overlap()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    parts = {}
    for _ in range(args.steps):
        res2, sd2, t = step_overlap()
        for k, v in t.items():
            parts[k] = parts.get(k, 0.0) + v
    dist.barrier(); torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    same = bool(np.array_equal(res2["spots"], res["spots"]) and np.array_equal(sd2, sd_serial))
    if rank == 0:
        ms = float(dt.item()) / args.steps * 1e3
        print(json.dumps({"workload": f"z-slab mosaic {X}x{Y}x{Z}, EDT chain beside the spot chain", "n_gpus": world,
                          "ms_per_step": ms, "gvox_per_s": X * Y * Z / ms / 1e6, "same_result_as_serial": same,
                          "stage_ms": {k: v / args.steps * 1e3 for k, v in parts.items()}}), flush=True)
dist.destroy_process_group()
