"""Diagnostic: the critical path of one device-resident step (device.analyze_stack): host time stamps at every
read-back (the host blocks there) and CUDA-event times of the device stages.  Run under gpurun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import muscle_fish_b200
from muscle_fish_b200 import device as D, synth

shape = (2048, 2048, 100)
st = synth.make_stack_device(shape, 20000, 120, 2)
img, fiber, nuc = st["img"], st["fiber_mask"], st["nuc_mask"]
notnuc = (nuc == 0).to(torch.uint8)
torch.cuda.synchronize()
for it in range(6):
    tm = D.StageTimer()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tm.mark("start")
    res = D.analyze_stack(img, fiber, notnuc, 2.0, 0.025, (0.0577, 0.0577, 0.5), timer=tm)
    t_ret = time.perf_counter()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if it >= 3:
        marks = " ".join(f"{n}={1e3 * (t - t0):.2f}" for n, t in tm.marks)
        stages = " ".join(f"{n}={sum(v):.2f}" for n, v in tm.totals_ms().items())
        print(f"step {it}: returned {1e3 * (t_ret - t0):.2f} ms, drained {1e3 * (t1 - t0):.2f} ms | host marks: {marks} | "
              f"device stages (ms): {stages}", flush=True)
# back-to-back steps, as bench.py times them
for _ in range(3):
    D.analyze_stack(img, fiber, notnuc, 2.0, 0.025, (0.0577, 0.0577, 0.5))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    D.analyze_stack(img, fiber, notnuc, 2.0, 0.025, (0.0577, 0.0577, 0.5))
torch.cuda.synchronize(); print("10 back-to-back steps: %.3f ms each" % ((time.perf_counter() - t0) * 100), flush=True)
