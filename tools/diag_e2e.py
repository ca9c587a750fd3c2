import os, sys, time, io, contextlib, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import muscle_fish_b200
from muscle_fish_b200 import device as D, synth, muscle_fish as mf, ndimage as nd
shape = (2048, 2048, 100)
st = synth.make_stack_device(shape, 20000, 120, 2)
h_img = torch.empty(shape, dtype=torch.float32).pin_memory(); h_img.copy_(st["img"].permute(1, 2, 0))
h_fib = torch.empty(shape, dtype=torch.uint8).pin_memory(); h_fib.copy_(st["fiber_mask"].permute(1, 2, 0))
h_nuc = torch.empty(shape, dtype=torch.uint8).pin_memory(); h_nuc.copy_(st["nuc_mask"].permute(1, 2, 0))
torch.cuda.synchronize()
a_img, a_fib, a_nuc = h_img.numpy(), h_fib.numpy(), h_nuc.numpy()
del st; torch.cuda.empty_cache()
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        s2, _ = mf.find_spots_snrfilter(a_img, sigma=2.0, t_spot=0.025, mask=a_fib)
    return s2, nd.spot_distances(a_nuc, s2, (0.0577, 0.0577, 0.5))
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
print("GPU timeline (us from first GPU op): name, start, dur")
last_end = t0
for e in evs:
    gap = e.time_range.start - last_end
    if e.time_range.end - e.time_range.start > 300 or gap > 300:
        print(f"  {e.name[:60]:60s} start {e.time_range.start - t0:9.0f} dur {e.time_range.end - e.time_range.start:8.0f} gap_before {gap:8.0f}")
    last_end = max(last_end, e.time_range.end)
print("span us", last_end - t0)
