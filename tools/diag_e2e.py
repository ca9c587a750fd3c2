"""Diagnostic: wall-clock breakdown of the e2e (host buffers) step."""
import os, sys, time, io, contextlib
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
del st
torch.cuda.empty_cache()

def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"   {label:28s} {(time.perf_counter()-t0)*1e3:8.2f} ms", flush=True); return r

for it in range(3):
    print("iter", it)
    t0 = time.perf_counter()
    vol = T("ingest img", lambda: D.ingest(a_img))
    m = T("ingest fiber", lambda: D.ingest(a_fib, as_mask=True))
    spots = T("blob_log", lambda: D.blob_log(vol, 2.0, 2.0, 1, 0.025))
    sm = T("gather mask", lambda: spots[D.gather(m, spots) != 0])
    base = T("quantile", lambda: D.masked_percentile(vol, m, 25))
    inten = T("blur pts", lambda: D.blur_at_points(vol, sm, (3, 3, 0.3)))
    nn = T("ingest nuc (negate)", lambda: D.ingest(a_nuc, as_mask=True, negate=True))
    dist = T("edt", lambda: D.edt(nn, (0.0577, 0.0577, 0.5)))
    sd = T("gather dist", lambda: D.gather(dist, sm))
    print("   staged total %.2f ms" % ((time.perf_counter() - t0) * 1e3))
    del vol, m, nn, dist
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        s2, _ = mf.find_spots_snrfilter(a_img, sigma=2.0, t_spot=0.025, mask=a_fib)
    t1 = time.perf_counter()
    d2 = nd.spot_distances(a_nuc, s2, (0.0577, 0.0577, 0.5))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("   API: find_spots_snrfilter %.2f ms, spot_distances %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
