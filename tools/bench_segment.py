"""Segmentation blurs at configs[1] size: per-kernel times of gaussian (10,10,1) / (20,20,2) and the achieved
fraction of the FP32 FMA pipe (2R+1 FMAs per voxel and pass; 128 FMA/clk/SM x 148 SMs x the SM clock)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_fish_b200
from muscle_fish_b200 import _lib, device as D, taps as T

shape = (100, 2048, 2048)
img = torch.rand(shape, device="cuda")
vol = D.Volume(img, None)
clk = 1.965e9
for sig in ((10, 10, 1), (20, 20, 2)):
    for _ in range(2):
        out = D.gaussian(vol, sig)
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = D.gaussian(vol, sig)
    e1.record(); torch.cuda.synchronize()
    k = _lib.profile_collect(); _lib.profile_enable(False)
    R = T.gaussian_radius(float(sig[0]))
    line = []
    for n, (c, ms) in sorted(k.items()):
        per = ms / c
        fma = img.numel() * (2 * R + 1) if "big" in n else 0
        frac = fma / (per * 1e-3) / (128 * 148 * clk) if fma else 0
        line.append(f"{n}={per:.3f} ms" + (f" ({frac:.2f} of the FMA pipe)" if fma else ""))
    print(f"gaussian{sig}: total {e0.elapsed_time(e1) / 3:.2f} ms | " + " ".join(line), flush=True)
