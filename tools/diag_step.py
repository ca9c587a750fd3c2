"""Diagnostic: wall-clock breakdown of analyze_stack stages and H2D bandwidth (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import muscle_fish_b200
from muscle_fish_b200 import device as D, synth, _lib

shape = (2048, 2048, 100)
st = synth.make_stack_device(shape, 20000, 120, 2)
img, fiber, nuc = st["img"], st["fiber_mask"], st["nuc_mask"]
notnuc = (nuc == 0).to(torch.uint8)
torch.cuda.synchronize()


class WallTimer:
    def __init__(self): self.t = {}
    def run(self, name, fn, *a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); self.t[name] = self.t.get(name, 0) + (time.perf_counter() - t0) * 1e3
        return out

for it in range(4):
    wt = WallTimer()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = D.analyze_stack(img, fiber, notnuc, 2.0, 0.025, (0.0577, 0.0577, 0.5), timer=wt)
    torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 1e3
    print(it, "total %.2f ms" % tot, {k: round(v, 2) for k, v in wt.t.items()}, "sum %.2f" % sum(wt.t.values()), flush=True)
print("raw peaks:", len(res["spots_masked"]), "kept:", len(res["spots"]))

h = torch.empty(shape, dtype=torch.float32).pin_memory()
d = torch.empty(shape, dtype=torch.float32, device="cuda")
for name, src in (("pinned tensor", h), ("from_numpy(pinned)", torch.from_numpy(h.numpy()))):
    print(name, "is_pinned", src.is_pinned())
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("  H2D %.1f GB/s" % (h.numel() * 4 / dt / 1e9))
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("  D2H %.1f GB/s" % (h.numel() * 4 / dt / 1e9))
a = h.numpy()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    v = D.ingest(a)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("  ingest(a) %.1f ms" % (dt * 1e3))
