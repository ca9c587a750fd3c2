"""Mosaic (configs[3]) under torchrun: the two chains of parallel.slab_analyze one after the other, side by side,
side by side with the spot chain on the high-priority stream, and with the transform's envelope passes gated behind
the LoG -- both decompositions.  One JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import muscle_fish_b200  # noqa: F401
from muscle_fish_b200 import parallel as P, synth, device as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
edt_group = dist.new_group()
X, Y, Z = (16384, 2048, 200) if world >= 4 else (8192, 2048, 200)
ns, nn = (640000, 4000) if world >= 4 else (320000, 2000)
SAMPLING = (0.0577, 0.0577, 0.5)
out = {"world": world, "mosaic": [X, Y, Z]}
for kind in ("z", "x"):
    if kind == "z":
        lay = P.SlabLayout(Z, X, Y)
        st = synth.make_stack_device((X, Y, Z), ns, nn, 4, z_range=(lay.z0, lay.z1))
    else:
        lay = P.XSlabLayout(Z, X, Y)
        st = synth.make_stack_device((X, Y, Z), ns, nn, 4, x_range=(lay.x0, lay.x1))
    img, fib = st["img"], st["fiber_mask"]
    notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
    del st
    if kind == "z":
        home = P.slab_input_buffer(lay, 2.0, img.dtype, img.device)
        if home is not None:
            home.copy_(img); img = home
    for mode in ("serial", "overlap", "overlap_prio", "gated"):
        os.environ["MFISH_SLAB_PRIO"] = "1" if mode == "overlap_prio" else "0"
        os.environ["MFISH_SLAB_GATE"] = "1" if mode == "gated" else "0"
        def step():
            return P.slab_analyze(img, fib, notnuc, lay, sigma=2.0, t_spot=0.025, sampling_xyz=SAMPLING,
                                  edt_group=None if mode == "serial" else edt_group, full_table=False)
        for _ in range(2):
            step()
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            res = step()
            torch.cuda.synchronize(); dist.barrier()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor(ts, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{kind}_{mode}"] = {"ms_median": float(t.median()) * 1e3, "ms_min": float(t.min()) * 1e3,
                                 "ms_all": [round(v * 1e3, 2) for v in t.tolist()], "spots": int(len(res["spots"]))}
    del img, fib, notnuc
    P.clear_buffers(); D.WS.clear(); torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
