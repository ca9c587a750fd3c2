"""Two device-resident steps of the C2 workload (the second one is what ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_fish_b200
from muscle_fish_b200 import device as D, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
shape, ns, nn, seed = synth.CONFIGS[wl]
st = synth.make_stack_device(shape, ns, nn, seed)
notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
for _ in range(2):
    res = D.analyze_stack(st["img"], st["fiber_mask"], notnuc, 2.0, 0.025, (0.0577, 0.0577, 0.5))
torch.cuda.synchronize()
print("spots", len(res["spots"]))
