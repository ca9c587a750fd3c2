"""Probe (torchrun, N >= 2): does torch's symmetric memory rendezvous work on this box, and how fast is a strided
block written straight into a peer's buffer (the slab all-to-all without NCCL) against NCCL send/recv?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"world": world}
nz, nx, ny = 100, 8192 // world * 2 // 2, 2048
nyl = ny // world
src = torch.randint(0, 60000, (nz, nx, ny), dtype=torch.int32, device="cuda").to(torch.uint16)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3
# NCCL grouped send/recv (what parallel.repartition_z_to_y does)
dst = torch.empty((nz * world, nx, nyl), dtype=torch.uint16, device="cuda")
snd = [torch.empty((nz, nx, nyl), dtype=torch.uint16, device="cuda") for _ in range(world)]
def nccl_a2a():
    ops = []
    for s in range(world):
        if s == rank:
            continue
        snd[s].copy_(src[:, :, s * nyl:(s + 1) * nyl])
        ops.append(dist.P2POp(dist.isend, snd[s].view(torch.uint8), s))
        ops.append(dist.P2POp(dist.irecv, dst[s * nz:(s + 1) * nz].view(torch.uint8), s))
    dst[rank * nz:(rank + 1) * nz] = src[:, :, rank * nyl:(rank + 1) * nyl]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
bytes_out = (world - 1) * nz * nx * nyl * 2
ms = timeit(nccl_a2a)
out["nccl_ms"] = ms; out["nccl_out_gbs"] = bytes_out / ms / 1e6
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty((nz * world, nx, nyl), dtype=torch.uint16, device="cuda")
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    peers = [hdl.get_buffer(s, (nz * world, nx, nyl), torch.uint16) for s in range(world)]
    def p2p_a2a():
        hdl.barrier(channel=0)
        for k in range(1, world):
            s = (rank + k) % world
            peers[s][rank * nz:(rank + 1) * nz].copy_(src[:, :, s * nyl:(s + 1) * nyl])
        buf[rank * nz:(rank + 1) * nz] = src[:, :, rank * nyl:(rank + 1) * nyl]
        hdl.barrier(channel=1)
    ms = timeit(p2p_a2a)
    out["symm_ms"] = ms; out["symm_out_gbs"] = bytes_out / ms / 1e6
    nccl_a2a(); p2p_a2a(); torch.cuda.synchronize()
    out["same"] = bool(torch.equal(buf, dst))
    # pull variant: read the peer's strided block, write locally
    srcs = symm.empty((nz, nx, ny), dtype=torch.uint16, device="cuda"); srcs.copy_(src)
    h2 = symm.rendezvous(srcs, dist.group.WORLD)
    psrc = [h2.get_buffer(s, (nz, nx, ny), torch.uint16) for s in range(world)]
    def p2p_pull():
        h2.barrier(channel=0)
        for k in range(1, world):
            s = (rank + k) % world
            dst[s * nz:(s + 1) * nz].copy_(psrc[s][:, :, rank * nyl:(rank + 1) * nyl])
        dst[rank * nz:(rank + 1) * nz] = srcs[:, :, rank * nyl:(rank + 1) * nyl]
        h2.barrier(channel=1)
    ms = timeit(p2p_pull)
    out["pull_ms"] = ms; out["pull_in_gbs"] = bytes_out / ms / 1e6
    # contiguous halo-like block: 9 planes of float32
    hb = symm.empty((18, 4096, 2048), dtype=torch.float32, device="cuda")
    h3 = symm.rendezvous(hb, dist.group.WORLD)
    ph = [h3.get_buffer(s, (18, 4096, 2048), torch.float32) for s in range(world)]
    own = torch.randn((9, 4096, 2048), device="cuda")
    def halo():
        h3.barrier(channel=0)
        ph[(rank + 1) % world][:9].copy_(own)
        ph[(rank - 1) % world][9:].copy_(own)
        h3.barrier(channel=1)
    ms = timeit(halo)
    out["halo_ms"] = ms; out["halo_out_gbs"] = 2 * own.numel() * 4 / ms / 1e6
except Exception as e:  # noqa: BLE001
    out["symm_error"] = repr(e)[:400]
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
