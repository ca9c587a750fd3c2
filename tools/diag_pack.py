"""Where the host-buffer step with bit-packed masks spends its time (C2 shapes): packing alone / beside an image
upload, the return time of host_to_device_many, arrival times of the three buffers, and the whole call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import muscle_fish_b200  # noqa: E402,F401
from muscle_fish_b200 import device as D, muscle_fish as mf  # noqa: E402

shape = (2048, 2048, 100)
rng = np.random.default_rng(0)
h_img = torch.empty(shape, dtype=torch.float32).pin_memory()
h_img.normal_(100.0, 10.0)
h_fib = torch.zeros(shape, dtype=torch.uint8).pin_memory()
h_fib[:, 400:1650, 8:92] = 1
h_nuc = torch.zeros(shape, dtype=torch.uint8).pin_memory()
h_nuc[100:200, 500:600, 20:40] = 1
a_img, a_fib, a_nuc = h_img.numpy(), h_fib.numpy(), h_nuc.numpy()
print("cpus", len(os.sched_getaffinity(0)), flush=True)
pool = D._stage_resources()[0]
print("pool threads", pool._max_workers)
for rep in range(3):
    t0 = time.perf_counter()
    futs, buf, _, _ = D._start_pack(a_nuc)
    t1 = time.perf_counter()
    for f in futs:
        f.result()
    t2 = time.perf_counter()
    print(f"pack alone: submit {1e3 * (t1 - t0):.2f} ms, done {1e3 * (t2 - t0):.2f} ms, {len(futs)} jobs")
d_img = torch.empty(shape, dtype=torch.float32, device="cuda")
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_img.copy_(h_img, non_blocking=True)
    futs, buf, _, _ = D._start_pack(a_nuc)
    for f in futs:
        f.result()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"pack beside the image upload: pack done {1e3 * (t2 - t0):.2f} ms, upload done {1e3 * (t3 - t0):.2f} ms")
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = D.host_to_device_many([a_nuc, a_img, a_fib], kinds=("mask", "image", "mask"), pack_masks=True)
    t1 = time.perf_counter()
    stamps = []
    for (t, e) in res:
        e.synchronize()
        stamps.append(time.perf_counter())
    print(f"host_to_device_many returns after {1e3 * (t1 - t0):.2f} ms; arrivals (nuc, img, fibre) "
          + ", ".join(f"{1e3 * (s - t0):.2f}" for s in stamps))
    del res
from muscle_fish_b200 import synth  # noqa: E402
del d_img
st = synth.make_stack_device(shape, 20000, 120, 2)
h_img.copy_(st["img"].permute(1, 2, 0))
h_fib.copy_(st["fiber_mask"].permute(1, 2, 0))
h_nuc.copy_(st["nuc_mask"].permute(1, 2, 0))
del st
torch.cuda.synchronize()
for pack in (True, False, True):
    D.PACK_MASKS = pack
    for rep in range(4):
        tm = D.StageTimer()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tm.mark("start")
        r = D.analyze_host(a_img, a_fib, a_nuc, sigma=2.0, t_spot=0.025, sampling_xyz=(0.0577, 0.0577, 0.5), timer=tm)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if rep >= 2:
            print(f"pack={pack}: analyze_host {1e3 * (t1 - t0):.2f} ms, {len(r['spots'])} spots | marks: "
                  + " ".join(f"{n}={1e3 * (ts - t0):.1f}" for n, ts in tm.marks)
                  + " | stages: " + " ".join(f"{k}={v[0]:.2f}" for k, v in tm.totals_ms().items()), flush=True)
