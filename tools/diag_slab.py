import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import muscle_fish_b200
from muscle_fish_b200 import parallel as P, synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X, Y, Z = 4096, 2048, 200
lay = P.SlabLayout(Z, X, Y)
st = synth.make_stack_device((X, Y, Z), 160000, 1000, 4, z_range=(lay.z0, lay.z1))
img, fib = st["img"], st["fiber_mask"]
for _ in range(2):
    res = P.slab_find_spots_snrfilter(img, fib, lay, sigma=2.0, t_spot=0.025)
torch.cuda.synchronize(); dist.barrier()
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    res = P.slab_find_spots_snrfilter(img, fib, lay, sigma=2.0, t_spot=0.025)
torch.cuda.synchronize(); pr.disable()
if rank == 0:
    pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
dist.destroy_process_group()
