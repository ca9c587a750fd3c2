"""Diagnostic (torchrun, N ranks): slab path against one GPU on the same voxels -- where do they differ?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import muscle_fish_b200
from muscle_fish_b200 import device as D, parallel as P, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = (0.0577, 0.0577, 0.5)
shape, ns, nn, seed = (2048, 1024, 96), 20000, 120, 4
X, Y, Z = shape
lay = P.SlabLayout(Z, X, Y)
st = synth.make_stack_device(shape, ns, nn, seed, z_range=(lay.z0, lay.z1))
notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
res = P.slab_analyze(st["img"], st["fiber_mask"], notnuc, lay, sigma=2.0, t_spot=0.025, sampling_xyz=S)
if rank == 0:
    full = synth.make_stack_device(shape, ns, nn, seed)
    print("slab voxels identical to the full stack's planes:",
          bool(torch.equal(full["img"][lay.z0:lay.z1], st["img"])), bool(torch.equal(full["nuc_mask"][lay.z0:lay.z1], st["nuc_mask"])))
    nn1 = (full["nuc_mask"] == 0).to(torch.uint8)
    one = D.analyze_stack(full["img"], full["fiber_mask"], nn1, 2.0, 0.025, S)
    a, b = one["spots"], res["spots"]
    print("counts", len(a), len(b), "baseline", one["baseline"], res["baseline"], "snr", one["snr"], res["snr"])
    sa = {tuple(r) for r in a[:, :3].astype(int).tolist()}
    sb = {tuple(r) for r in b[:, :3].astype(int).tolist()}
    print("set equal", sa == sb, "only one", len(sa - sb), "only slab", len(sb - sa))
    if len(a) == len(b):
        diff = np.nonzero((a != b).any(axis=1))[0]
        print("rows differing in order", len(diff), diff[:10])
        if len(diff):
            i = diff[0]
            print(a[i - 1:i + 3], b[i - 1:i + 3])
    # masked tables
    print("masked counts", len(one["spots_masked"]), len(res["spots_masked"]))
    ia = np.lexsort(one["spots_masked"][:, :3].T[::-1]); ib = np.lexsort(res["spots_masked"][:, :3].T[::-1])
    if len(ia) == len(ib):
        print("masked sets equal", np.array_equal(one["spots_masked"][ia], res["spots_masked"][ib]),
              "intensity max abs diff", np.abs(one["intensity"][ia] - res["intensity"][ib]).max())
    da = dict(zip(map(tuple, a[:, :3].astype(int).tolist()), one["spot_dist"]))
    db = dict(zip(map(tuple, b[:, :3].astype(int).tolist()), res["spot_dist"]))
    common = sa & sb
    dd = np.array([abs(da[k] - db[k]) for k in common])
    print("distance max abs diff over common spots", dd.max() if len(dd) else None, "nonzero", int((dd != 0).sum()))
dist.barrier()
dist.destroy_process_group()
