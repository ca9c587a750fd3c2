"""Diagnostic: where the host-buffer step (device.analyze_host through the drop-in call) spends its time
after the last byte has crossed PCIe.  Run under gpurun."""
import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import muscle_fish_b200
from muscle_fish_b200 import device as D, synth, muscle_fish as mf

shape = (2048, 2048, 100)
st = synth.make_stack_device(shape, 20000, 120, 2)
h_img = torch.empty(shape, dtype=torch.float32).pin_memory(); h_img.copy_(st["img"].permute(1, 2, 0))
h_fib = torch.empty(shape, dtype=torch.uint8).pin_memory(); h_fib.copy_(st["fiber_mask"].permute(1, 2, 0))
h_nuc = torch.empty(shape, dtype=torch.uint8).pin_memory(); h_nuc.copy_(st["nuc_mask"].permute(1, 2, 0))
del st
torch.cuda.synchronize()
a_img, a_fib, a_nuc = h_img.numpy(), h_fib.numpy(), h_nuc.numpy()
S = (0.0577, 0.0577, 0.5)

# raw H2D of the three buffers, back to back
d = [torch.empty_like(t, device="cuda") for t in (h_nuc, h_img, h_fib)]
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for dst, src in zip(d, (h_nuc, h_img, h_fib)):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); print("raw H2D of 2.52 GB: %.2f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
del d

import cProfile, pstats
for it in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        r = mf.find_spots_snrfilter_with_distances(a_img, a_nuc, S, sigma=2.0, t_spot=0.025, mask=a_fib)
    torch.cuda.synchronize(); print("fused call: %.2f ms (%d spots)" % ((time.perf_counter() - t0) * 1e3, len(r[0])), flush=True)
pr = cProfile.Profile(); pr.enable()
with contextlib.redirect_stdout(io.StringIO()):
    r = mf.find_spots_snrfilter_with_distances(a_img, a_nuc, S, sigma=2.0, t_spot=0.025, mask=a_fib)
pr.disable()
pstats.Stats(pr, stream=sys.stdout).sort_stats("cumtime").print_stats(28)
