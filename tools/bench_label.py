"""Connected-component labelling at configs[1] size (2048x2048x100): the device path of segmentation.label_device /
threshold_fiber / threshold_nuclei against scipy.ndimage.label on the host (one core), same masks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import muscle_fish_b200  # noqa: F401
from muscle_fish_b200 import segmentation as sg, synth
from scipy import ndimage as ndi

shape = (2048, 2048, 100) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
st = synth.make_stack_device(shape, 20000, 120, 2, z_first=False)  # (x, y, z) device tensors
nuc = (st["nuc_mask"] != 0)
noise = torch.rand(nuc.shape, device="cuda") < 0.002
mask = (nuc | noise).contiguous()
for full in (True, False):
    for _ in range(2):
        lab, n = sg.label_device(mask, full=full)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lab, n = sg.label_device(mask, full=full)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"label_device full={full}: {dt * 1e3:.1f} ms, {n} components, {mask.numel() / dt / 1e9:.1f} Gvoxel/s", flush=True)
t0 = time.perf_counter()
clean = sg.remove_small_holes_device(sg.remove_small_objects_device(mask, 2048), 2048)
torch.cuda.synchronize()
print(f"remove_small_objects + remove_small_holes (2048): {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
h = mask.cpu().numpy()
t0 = time.perf_counter()
ref, n_ref = ndi.label(h, structure=np.ones((3, 3, 3), bool))
dt = time.perf_counter() - t0
print(f"scipy.ndimage.label (26-neighbourhood, one core): {dt:.2f} s, {n_ref} components", flush=True)
lab, n = sg.label_device(mask, full=True)
print("identical labels:", bool(n == n_ref and np.array_equal(lab.cpu().numpy(), ref)), flush=True)
