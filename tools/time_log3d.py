"""Time mfish_log3d_f32 per kernel for a given build of the library (variant exploration)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import muscle_fish_b200
from muscle_fish_b200 import _lib
for path in sys.argv[1:]:
    _lib._lib = None
    _lib.LIB_PATH = os.path.abspath(path)
    lib = _lib.lib()
    vol = torch.rand(100, 2048, 2048, device="cuda")
    bufs = [torch.empty_like(vol) for _ in range(4)]
    mm = torch.tensor([0.0, 1.0], device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.mfish_log3d_f32(vol.data_ptr(), *[b.data_ptr() for b in bufs], 100, 2048, 2048, 2.0, mm.data_ptr(), st)
    torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_collect()
    for _ in range(10):
        lib.mfish_log3d_f32(vol.data_ptr(), *[b.data_ptr() for b in bufs], 100, 2048, 2048, 2.0, mm.data_ptr(), st)
    torch.cuda.synchronize()
    k = _lib.profile_collect(); _lib.profile_enable(False)
    print(os.path.basename(path), {n: round(ms / c, 3) for n, (c, ms) in k.items()}, flush=True)
    del vol, bufs
