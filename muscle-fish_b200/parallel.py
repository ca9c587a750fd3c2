"""Multi-GPU drivers: one process per GPU, ``torch.distributed`` for the plumbing.

Two partitionings:

* **batch** -- independent stacks, one per rank round-robin, no data-path collective (``run_batch``: pinned
  double-buffered staging so that PCIe never waits for the GPU).  This is the reference's own model:
  ``general_pipeline/batch_process.py:28-52`` writes one job per image.
* **mosaic** -- one large volume ``[nz][nx][ny]`` split into contiguous z-slabs (``SlabLayout``) or x-slabs
  (``XSlabLayout``), ``slab_analyze``:
    - min/max + P25 baseline:   ONE pass over the raw slab; sample / candidate histograms and counters all-reduced
    - LoG / NMS / blur:         radius+1 input planes (12 rows) from each neighbour; every rank reports only the
                                peaks whose centre lies in its own planes (rows)
    - overlap prune:            all-gather of the (small) peak lists, same prune on every rank
    - EDT:                      row scan slab-local; the uint16 row distances are re-partitioned into y-slabs
                                (all-to-all), both envelope passes run there on whole lines
    - exchanges:                on CUDA tensors blocks are written straight into the receiver's symmetric buffer
                                over NVLink (``mfish_copy2d`` / ``mfish_copy3d``), else grouped isend / irecv

The communication code only moves torch tensors; the voxel work is done by an ``ops`` object.
``DeviceOps`` (default) calls the CUDA library.  The CPU tests inject an oracle-backed ``ops`` to
check the partition / halo / all-to-all logic on ``gloo`` -- the product has no CPU compute path.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist

from . import taps as T


# optional wall-clock phase accounting (tools/bench_slab.py --phases)
PHASES = None


class _phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PHASES is not None:
            torch.cuda.synchronize()
            import time
            self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if PHASES is not None:
            import time
            torch.cuda.synchronize()
            PHASES[self.name] = PHASES.get(self.name, 0.0) + time.perf_counter() - self.t0


# Large per-step temporaries (extended slabs, re-partition buffers) are cached by (tag, shape): a
# slab step alternates multi-GB buffers of different sizes, and letting them churn through the
# caching allocator costs cudaMalloc/cudaFree round trips every step.  A cached buffer is valid
# until the next call that uses the same tag.
_BUFS = {}


def _buf(tag, shape, dtype, device):
    k = (tag, tuple(shape), dtype, str(device))
    b = _BUFS.get(k)
    if b is None:
        b = _BUFS[k] = torch.empty(shape, dtype=dtype, device=device)
    return b


def clear_buffers():
    _BUFS.clear()
    _PEER_BUFS.clear()
    _STAGING.clear()


# ------------------------------------------------------------------------ NVLink peer memory
# On CUDA tensors the block moves between slabs (halo planes / rows, the all-to-all in front of the EDT's
# envelope passes) do not go through send/recv pairs: every rank holds a SYMMETRIC buffer (torch symmetric
# memory: same shape on every rank, mapped into every peer), and a sender writes its block straight into the
# receiver's buffer with one pack-and-transfer kernel (``mfish_copy2d``: 16-byte loads from local HBM, 16-byte
# stores through NVSwitch), bracketed by two device-side barriers of the buffer's handle.  No staging copies,
# nothing to unpack, measured ~2x the NCCL send/recv rate on these block shapes (tools/probe_symm.py).
# MFISH_P2P=0, CPU tensors (the gloo tests) or a failed rendezvous fall back to grouped isend / irecv.
_PEER_BUFS = {}
_PEER_STATE = {"ok": None}


def _peer_ok(t: torch.Tensor) -> bool:
    import os
    return t.is_cuda and _PEER_STATE["ok"] is not False and os.environ.get("MFISH_P2P", "1") != "0"


def _peer_buf(tag, shape, dtype, device, group):
    """(local tensor, handle, [view of every rank's buffer]) of the symmetric buffer `tag`; the first use of a
    (tag, shape) is collective.  None when symmetric memory is unavailable."""
    k = (tag, tuple(shape), dtype, str(device), id(group))
    ent = _PEER_BUFS.get(k)
    if ent is None:
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(tuple(shape), dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            peers = [hdl.get_buffer(r, tuple(shape), dtype) for r in range(hdl.world_size)]
            _PEER_STATE["ok"] = True
        except Exception as e:  # noqa: BLE001  (systemic: every rank fails the same way)
            if _PEER_STATE["ok"]:
                raise
            import warnings
            warnings.warn(f"symmetric memory unavailable ({e!r}): slab exchanges use NCCL send/recv")
            _PEER_STATE["ok"] = False
            return None
        ent = _PEER_BUFS[k] = (buf, hdl, peers)
    return ent


def _block_copy(dst: torch.Tensor, src: torch.Tensor):
    """dst[...] = src[...] for 3-D views with a contiguous last axis; dst may live on a peer GPU.  One
    ``mfish_copy2d`` / ``mfish_copy3d`` launch (16-byte words), torch's copy for anything unaligned."""
    assert dst.shape == src.shape and dst.dtype == src.dtype and dst.dim() == 3
    n2, n1, w = src.shape
    if n2 == 0 or n1 == 0 or w == 0:
        return
    es = src.element_size()
    rb = w * es
    ok = src.stride(2) == 1 and dst.stride(2) == 1 and rb % 16 == 0 and src.data_ptr() % 16 == 0 \
        and dst.data_ptr() % 16 == 0 and all((v.stride(0) * es) % 16 == 0 and (v.stride(1) * es) % 16 == 0
                                             for v in (src, dst))
    if not ok:
        dst.copy_(src)
        return
    from . import _lib as L
    from . import device as D
    lib, st = L.lib(), D._stream()
    sp1, sp2, dp1, dp2 = src.stride(1) * es, src.stride(0) * es, dst.stride(1) * es, dst.stride(0) * es
    if sp1 == rb and dp1 == rb:  # contiguous n1*w runs, n2 of them
        L.check(lib.mfish_copy2d(dst.data_ptr(), dp2, src.data_ptr(), sp2, rb * n1, n2, st))
    else:
        L.check(lib.mfish_copy3d(dst.data_ptr(), dp1, dp2, src.data_ptr(), sp1, sp2, rb, n1, n2, st))


# ------------------------------------------------------------------------------ partitioning
def shard_stacks(n_stacks: int, rank: int, world: int):
    """Stack indices of this rank: round-robin, like one SLURM job per image."""
    return list(range(rank, n_stacks, world))


def split_bounds(n: int, parts: int):
    """Balanced contiguous split of range(n) into `parts` pieces: list of (begin, end)."""
    base, rem = divmod(n, parts)
    out, b = [], 0
    for i in range(parts):
        e = b + base + (1 if i < rem else 0)
        out.append((b, e))
        b = e
    return out


def bind_to_gpu_numa(device_index: int):
    """Restrict this process to the CPUs the driver reports as local to its GPU, BEFORE pinned staging
    buffers are allocated, so that first-touch places them on the GPU's NUMA node.  Best effort: returns
    the CPU list used, or None when NVML has no affinity information."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


_STAGING = {}


class _StagingSlot:
    """One stack's worth of pinned host staging + device landing buffers (reference (x, y, z) order)."""

    def __init__(self, shape_xyz, img_dtype, device):
        self.img_h = torch.empty(shape_xyz, dtype=img_dtype).pin_memory()
        self.fib_h = torch.empty(shape_xyz, dtype=torch.uint8).pin_memory()
        self.nuc_h = torch.empty(shape_xyz, dtype=torch.uint8).pin_memory()
        self.img_d = torch.empty(shape_xyz, dtype=img_dtype, device=device)
        self.fib_d = torch.empty(shape_xyz, dtype=torch.uint8, device=device)
        self.nuc_d = torch.empty(shape_xyz, dtype=torch.uint8, device=device)
        self.copied = torch.cuda.Event()
        self.bytes = sum(t.numel() * t.element_size() for t in (self.img_h, self.fib_h, self.nuc_h))
        # the masks cross PCIe as bits (device.PackedMask): pinned bit streams + their landing buffers
        from . import _lib as L
        nbits = int(L.lib().mfish_mask_bits_bytes(int(np.prod(shape_xyz))))
        self.fib_bits_h = torch.zeros(nbits, dtype=torch.uint8).pin_memory()
        self.nuc_bits_h = torch.zeros(nbits, dtype=torch.uint8).pin_memory()
        self.fib_bits_d = torch.empty(nbits, dtype=torch.uint8, device=device)
        self.nuc_bits_d = torch.empty(nbits, dtype=torch.uint8, device=device)
        self.packed = False

    def host_views(self):
        return self.img_h.numpy(), self.fib_h.numpy(), self.nuc_h.numpy()


def run_batch(n_stacks: int, loader, shape_xyz, img_dtype=torch.uint16, sigma=2.0, t_spot=0.025,
              sampling_xyz=(1.0, 1.0, 1.0), snr=None, rank=None, world=None, on_result=None):
    """The batch mode of the reference (general_pipeline/batch_process.py:28-52 writes one job per image; here
    one process per GPU takes every ``world``-th stack) with the slow link kept busy: two pinned staging
    slots per GPU, so that while stack i is analysed, stack i+1 crosses PCIe and stack i+2 is being read.

    ``loader(index, img, fiber_mask, nuc_mask)`` fills three pinned host arrays in the reference's (x, y, z)
    order (``img`` of ``img_dtype`` -- raw uint16 counts need no widening on the host -- masks uint8 0/1), or
    returns its own three (pinned) buffers instead; it runs in a worker thread and stands for reading one
    image file.  No collective is involved.
    Returns ``[(index, result dict of device.analyze_host)]`` for this rank's stacks; ``on_result(index, res)``
    is called as each stack finishes."""
    import threading
    from . import device as D
    D._require_cuda()
    rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
    world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
    mine = shard_stacks(n_stacks, rank, world)
    if not mine:
        return []
    dev = torch.device("cuda", torch.cuda.current_device())
    # pinning half a gigabyte costs ~0.1 s (more with eight ranks doing it at once): the two slots outlive the call
    skey = (tuple(shape_xyz), img_dtype, str(dev))
    slots = _STAGING.get(skey)
    if slots is None:
        slots = _STAGING[skey] = [_StagingSlot(tuple(shape_xyz), img_dtype, dev) for _ in range(2)]
    copy_stream = _STAGING.setdefault(("stream", str(dev)), torch.cuda.Stream())
    errors = []
    pack_masks = D.PACK_MASKS  # bits with at most two ranks per host, whole bytes beyond (see device.PACK_MASKS)

    def fill(slot, index):
        try:
            slot.copied.synchronize()  # the previous upload out of this slot has left the host buffer
            ret = loader(index, *slot.host_views())
            # a loader that already holds the stack in pinned memory returns (img, fiber, nuc) instead of
            # filling the staging views; those buffers are then uploaded as they are
            slot.src = (slot.img_h, slot.fib_h, slot.nuc_h) if ret is None else tuple(
                t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t)) for t in ret)
            # one bit per mask voxel over PCIe: packed here, by the worker threads of the upload path, while the
            # previous stack is crossing PCIe / being analysed
            slot.packed = False
            if pack_masks:
                pairs = [(slot.src[1], slot.fib_bits_h), (slot.src[2], slot.nuc_bits_h)]
                arrs = [D._packable_any_size(m) for m, _ in pairs]
                if all(a is not None and a.size == slot.fib_h.numel() for a in arrs):
                    futs = []
                    for a, (_, bits) in zip(arrs, pairs):
                        futs += D.pack_into(a, bits)
                    for f in futs:
                        f.result()
                    slot.packed = True
        except BaseException as e:  # surfaced in the main thread
            errors.append(e)

    def start_fill(k):
        th = threading.Thread(target=fill, args=(slots[k % 2], mine[k]), daemon=True)
        th.start()
        return th

    def upload(k):
        s = slots[k % 2]
        # the landing buffers of this slot were last read by the analysis of stack k-2 on the compute stream
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            img_h, fib_h, nuc_h = s.src
            if s.packed:
                s.nuc_bits_d.copy_(s.nuc_bits_h, non_blocking=True)  # the distance transform starts first
                s.img_d.copy_(img_h, non_blocking=True)
                s.fib_bits_d.copy_(s.fib_bits_h, non_blocking=True)
            else:
                s.nuc_d.copy_(nuc_h, non_blocking=True)
                s.img_d.copy_(img_h, non_blocking=True)
                s.fib_d.copy_(fib_h, non_blocking=True)
            s.uploaded_packed = s.packed
            s.copied.record(copy_stream)

    results = []
    start_fill(0).join()
    if errors:
        raise errors[0]
    upload(0)
    pending = start_fill(1) if len(mine) > 1 else None  # thread reading stack k+1
    for k, index in enumerate(mine):
        if pending is not None:  # stack k+1 has been read: queue its upload behind stack k's
            pending.join()
            pending = None
            if errors:
                raise errors[0]
            upload(k + 1)
        if k + 2 < len(mine):
            # stack k+2 goes into the pinned slot stack k is leaving; fill() waits for that DMA itself, so the
            # read overlaps the analysis of stack k below
            pending = start_fill(k + 2)
        s = slots[k % 2]
        torch.cuda.current_stream().wait_event(s.copied)
        if s.uploaded_packed:
            fib_in = D.PackedMask(s.fib_bits_d, tuple(shape_xyz))
            nuc_in = D.PackedMask(s.nuc_bits_d, tuple(shape_xyz))
        else:
            fib_in, nuc_in = s.fib_d, s.nuc_d
        res = D.analyze_host(s.img_d, fib_in, nuc_in, sigma=sigma, t_spot=t_spot, sampling_xyz=sampling_xyz,
                             snr=snr)
        res.pop("dist", None)  # the dense map stays a per-stack temporary
        results.append((index, res))
        if on_result is not None:
            on_result(index, res)
    if errors:
        raise errors[0]
    return results


class SlabLayout:
    """z-slab ownership (filters, NMS, row scan), y-slab ownership (EDT envelope passes; or x-slab ownership for
    the z pass alone) of a [nz][nx][ny] mosaic."""

    def __init__(self, nz, nx, ny, rank=None, world=None):
        self.nz, self.nx, self.ny = int(nz), int(nx), int(ny)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.zb = split_bounds(self.nz, self.world)
        self.xb = split_bounds(self.nx, self.world)
        self.yb = split_bounds(self.ny, self.world)

    y0 = property(lambda self: self.yb[self.rank][0])
    y1 = property(lambda self: self.yb[self.rank][1])

    @property
    def z0(self):
        return self.zb[self.rank][0]

    @property
    def z1(self):
        return self.zb[self.rank][1]

    @property
    def x0(self):
        return self.xb[self.rank][0]

    @property
    def x1(self):
        return self.xb[self.rank][1]

    def min_slab(self):
        return min(e - b for b, e in self.zb)


class XSlabLayout:
    """x-slab ownership (filters, NMS, row scan) and y-slab ownership (EDT envelope passes) of a [nz][nx][ny]
    mosaic.  x is the fibre axis of a stitched mosaic (16384 of 16384 x 2048 x 200): slabs along it pay a
    12-row halo on 2048 own rows (1 %), where 25-plane z-slabs pay 9 + 9 planes (72 %)."""

    def __init__(self, nz, nx, ny, rank=None, world=None):
        self.nz, self.nx, self.ny = int(nz), int(nx), int(ny)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.xb = split_bounds(self.nx, self.world)
        self.yb = split_bounds(self.ny, self.world)

    x0 = property(lambda self: self.xb[self.rank][0])
    x1 = property(lambda self: self.xb[self.rank][1])
    y0 = property(lambda self: self.yb[self.rank][0])
    y1 = property(lambda self: self.yb[self.rank][1])

    def min_slab(self):
        return min(e - b for b, e in self.xb)


# ----------------------------------------------------------------------------- communication
def exchange_halo(slab: torch.Tensor, halo: int, lay: SlabLayout, group=None):
    """Append up to `halo` planes of each z-neighbour.  Returns (ext, lo, hi): the extended slab and
    the number of planes added below / above (0 at the global faces)."""
    r, P = lay.rank, lay.world
    if halo <= 0 or P == 1:
        return slab, 0, 0
    if lay.min_slab() < halo:
        raise ValueError(f"slab thickness {lay.min_slab()} < halo {halo}: use fewer ranks for this mosaic")
    lo = halo if r > 0 else 0
    hi = halo if r < P - 1 else 0
    nzl = slab.shape[0]
    if _peer_ok(slab):
        nz_max = max(e - b for b, e in lay.zb)
        pm = _peer_buf("halo_ext", (2 * halo + nz_max,) + tuple(slab.shape[1:]), slab.dtype, slab.device, group)
        if pm is not None:
            buf, hdl, peers = pm
            hdl.barrier(channel=0)  # every rank is done with the previous step's extended slab
            if slab.data_ptr() != buf[lo:lo + nzl].data_ptr():  # (slab_input_buffer: the slab already lives there)
                _block_copy(buf[lo:lo + nzl], slab)
            if r > 0:  # my lowest planes are the upper halo of the rank below
                b_, e_ = lay.zb[r - 1]
                lo_n = halo if r - 1 > 0 else 0
                _block_copy(peers[r - 1][lo_n + (e_ - b_):lo_n + (e_ - b_) + halo], slab[:halo])
            if r < P - 1:  # my highest planes are the lower halo of the rank above
                _block_copy(peers[r + 1][:halo], slab[nzl - halo:])
            hdl.barrier(channel=1)  # every block has landed
            return buf[:lo + nzl + hi], lo, hi
    ext = _buf("halo_ext", (lo + nzl + hi,) + tuple(slab.shape[1:]), slab.dtype, slab.device)
    ext[lo:lo + nzl] = slab
    ops = []
    send_lo = send_hi = None
    if r > 0:
        send_lo = slab[:halo].contiguous()
        ops.append(dist.P2POp(dist.isend, send_lo, r - 1, group))
        ops.append(dist.P2POp(dist.irecv, ext[:lo], r - 1, group))
    if r < P - 1:
        send_hi = slab[nzl - halo:].contiguous()
        ops.append(dist.P2POp(dist.isend, send_hi, r + 1, group))
        ops.append(dist.P2POp(dist.irecv, ext[lo + nzl:], r + 1, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    return ext, lo, hi


def slab_input_buffer(lay: SlabLayout, sigma: float, dtype=torch.float32, device=None, group=None):
    """The rank's own z-slab ``[nz_local][nx][ny]`` as a view INSIDE the extended buffer the halo exchange of
    ``slab_find_spots_snrfilter(sigma)`` works in: a caller that puts its slab there (H2D copy, generator) saves the
    pipeline the copy of the slab into the extended buffer every step.  None when peer memory is not in use
    (then any tensor is as good as another)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    halo = T.gaussian_radius(sigma) + 1
    r, P = lay.rank, lay.world
    probe = torch.empty(0, device=device)
    if P == 1 or not _peer_ok(probe):
        return None
    nz_max = max(e - b for b, e in lay.zb)
    pm = _peer_buf("halo_ext", (2 * halo + nz_max, lay.nx, lay.ny), dtype, device, group)
    if pm is None:
        return None
    lo = halo if r > 0 else 0
    return pm[0][lo:lo + (lay.z1 - lay.z0)]


def exchange_halo_x(slab: torch.Tensor, halo: int, lay: "XSlabLayout", group=None):
    """x-slab [nz][nx_local][ny]: append up to `halo` x-rows of each x-neighbour.  Returns (ext, lo, hi)."""
    r, P = lay.rank, lay.world
    if halo <= 0 or P == 1:
        return slab, 0, 0
    if lay.min_slab() < halo:
        raise ValueError(f"slab width {lay.min_slab()} < halo {halo}: use fewer ranks for this mosaic")
    lo = halo if r > 0 else 0
    hi = halo if r < P - 1 else 0
    nz, nxl, ny = slab.shape
    widths = [e - b for b, e in lay.xb]
    if _peer_ok(slab):
        # one flat symmetric allocation; every rank's extended slab is a contiguous [nz][lo + nxl + hi][ny]
        # view of its own buffer, and a neighbour addresses it in THAT rank's shape
        pm = _peer_buf("halo_ext_x", (nz * (2 * halo + max(widths)) * ny,), slab.dtype, slab.device, group)
        if pm is not None:
            _, hdl, _ = pm

            def ext_of(q):
                lo_q = halo if q > 0 else 0
                hi_q = halo if q < P - 1 else 0
                return hdl.get_buffer(q, (nz, lo_q + widths[q] + hi_q, ny), slab.dtype), lo_q

            mine, _ = ext_of(r)
            hdl.barrier(channel=0)
            _block_copy(mine[:, lo:lo + nxl], slab)
            if r > 0:
                below, lo_b = ext_of(r - 1)
                _block_copy(below[:, lo_b + widths[r - 1]:], slab[:, :halo])
            if r < P - 1:
                above, _ = ext_of(r + 1)
                _block_copy(above[:, :halo], slab[:, nxl - halo:])
            hdl.barrier(channel=1)
            return mine, lo, hi
    ext = _buf("halo_ext_x", (nz, lo + nxl + hi, ny), slab.dtype, slab.device)
    ext[:, lo:lo + nxl] = slab
    ops, keep = [], []
    if r > 0:
        snd = slab[:, :halo].contiguous()
        rcv = _buf("halo_rx_lo", (nz, halo, ny), slab.dtype, slab.device)
        keep.append((rcv, slice(0, lo)))
        ops += [dist.P2POp(dist.isend, snd, r - 1, group), dist.P2POp(dist.irecv, rcv, r - 1, group)]
    if r < P - 1:
        snd2 = slab[:, nxl - halo:].contiguous()
        rcv2 = _buf("halo_rx_hi", (nz, halo, ny), slab.dtype, slab.device)
        keep.append((rcv2, slice(lo + nxl, lo + nxl + hi)))
        ops += [dist.P2POp(dist.isend, snd2, r + 1, group), dist.P2POp(dist.irecv, rcv2, r + 1, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for buf, sl in keep:
        ext[:, sl] = buf
    return ext, lo, hi


def _wire(t: torch.Tensor) -> torch.Tensor:
    """A contiguous buffer as NCCL can carry it: 16-bit integers travel as bytes (NCCL has no Short)."""
    if t.dtype in (torch.int16, torch.uint16):
        return t.view(torch.uint8)
    return t


def repartition_x_to_y(t: torch.Tensor, lay: "XSlabLayout", group=None) -> torch.Tensor:
    """[nz][nx_local][ny] (x-slab) -> [nz][nx][ny_local] (y-slab): the all-to-all between the EDT's row scan and
    its two envelope passes (2 bytes per voxel on the wire: the row distances are uint16)."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    nz = lay.nz
    if _peer_ok(t) and len({e - b for b, e in lay.yb}) == 1:
        pm = _peer_buf("x2y_out", (nz, lay.nx, lay.y1 - lay.y0), t.dtype, t.device, group)
        if pm is not None:
            buf, hdl, peers = pm
            hdl.barrier(channel=0)
            for k in range(P):
                s_ = (r + k) % P  # own block first, then the peers in ring order
                yb, ye = lay.yb[s_]
                _block_copy(peers[s_][:, lay.x0:lay.x1] if s_ != r else buf[:, lay.x0:lay.x1], t[:, :, yb:ye])
            hdl.barrier(channel=1)
            return buf
    out = _buf("x2y_out", (nz, lay.nx, lay.y1 - lay.y0), t.dtype, t.device)
    ops, recvs = [], {}
    for s in range(P):
        if s == r:
            continue
        yb, ye = lay.yb[s]
        snd = _buf(f"x2y_send{s}", (nz, t.shape[1], ye - yb), t.dtype, t.device)
        snd.copy_(t[:, :, yb:ye])
        xb, xe = lay.xb[s]
        rcv = _buf(f"x2y_recv{s}", (nz, xe - xb, lay.y1 - lay.y0), t.dtype, t.device)
        recvs[s] = rcv
        ops.append(dist.P2POp(dist.isend, _wire(snd), s, group))
        ops.append(dist.P2POp(dist.irecv, _wire(rcv), s, group))
    out[:, lay.x0:lay.x1] = t[:, :, lay.y0:lay.y1]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for s, rcv in recvs.items():
        out[:, lay.xb[s][0]:lay.xb[s][1]] = rcv
    return out


def repartition_z_to_y(t: torch.Tensor, lay: SlabLayout, group=None) -> torch.Tensor:
    """[nz_local][nx][ny] (z-slab) -> [nz][nx][ny_local] (y-slab): the all-to-all between the EDT's row scan and
    its two envelope passes when the mosaic is cut along z (2 bytes per voxel on the wire instead of the 4 of the
    in-plane squares; a peer's block arrives where it belongs, nothing to unpack)."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    if _peer_ok(t) and len({e - b for b, e in lay.yb}) == 1:
        pm = _peer_buf("z2y_out", (lay.nz, lay.nx, lay.y1 - lay.y0), t.dtype, t.device, group)
        if pm is not None:
            buf, hdl, peers = pm
            hdl.barrier(channel=0)
            for k in range(P):
                s_ = (r + k) % P
                yb, ye = lay.yb[s_]
                _block_copy(peers[s_][lay.z0:lay.z1] if s_ != r else buf[lay.z0:lay.z1], t[:, :, yb:ye])
            hdl.barrier(channel=1)
            return buf
    out = _buf("z2y_out", (lay.nz, lay.nx, lay.y1 - lay.y0), t.dtype, t.device)
    ops = []
    for s in range(P):
        if s == r:
            continue
        yb, ye = lay.yb[s]
        snd = _buf(f"z2y_send{s}", (t.shape[0], lay.nx, ye - yb), t.dtype, t.device)
        snd.copy_(t[:, :, yb:ye])
        ops.append(dist.P2POp(dist.isend, _wire(snd), s, group))
        ops.append(dist.P2POp(dist.irecv, _wire(out[lay.zb[s][0]:lay.zb[s][1]]), s, group))
    out[lay.z0:lay.z1] = t[:, :, lay.y0:lay.y1]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def repartition_z_to_x(t: torch.Tensor, lay: SlabLayout, group=None) -> torch.Tensor:
    """[nz_local][nx][ny] (z-slab) -> [nz][nx_local][ny] (x-slab): the all-to-all before the EDT z pass."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    ny = lay.ny
    out = _buf("z2x_out", (lay.nz, lay.x1 - lay.x0, ny), t.dtype, t.device)
    ops = []
    for s in range(P):
        if s == r:
            continue
        b, e = lay.xb[s]
        snd = _buf(f"z2x_send{s}", (t.shape[0], e - b, ny), t.dtype, t.device)
        snd.copy_(t[:, b:e, :])
        ops.append(dist.P2POp(dist.isend, snd, s, group))
        ops.append(dist.P2POp(dist.irecv, out[lay.zb[s][0]:lay.zb[s][1]], s, group))
    out[lay.z0:lay.z1] = t[:, lay.x0:lay.x1, :]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def repartition_x_to_z(t: torch.Tensor, lay: SlabLayout, group=None) -> torch.Tensor:
    """[nz][nx_local][ny] (x-slab) -> [nz_local][nx][ny] (z-slab)."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    out = _buf("x2z_out", (lay.z1 - lay.z0, lay.nx, lay.ny), t.dtype, t.device)
    recvs = [_buf(f"x2z_recv{i}", (lay.z1 - lay.z0, e - b, lay.ny), t.dtype, t.device)
             for i, (b, e) in enumerate(lay.xb)]
    ops = []
    for s in range(P):
        if s == r:
            continue
        ops.append(dist.P2POp(dist.isend, t[lay.zb[s][0]:lay.zb[s][1]].contiguous(), s, group))
        ops.append(dist.P2POp(dist.irecv, recvs[s], s, group))
    recvs[r] = t[lay.z0:lay.z1]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for s, (b, e) in enumerate(lay.xb):
        out[:, b:e, :] = recvs[s]
    return out


def _allreduce_np(arr: np.ndarray, device, op=dist.ReduceOp.SUM, group=None) -> np.ndarray:
    t = torch.as_tensor(np.ascontiguousarray(arr)).to(device)
    dist.all_reduce(t, op=op, group=group)
    return t.cpu().numpy()


# --------------------------------------------------------------------------- device compute
class DeviceOps:
    """Voxel work on the GPU through libmfish_sm100 (see ``device.py``)."""

    def __init__(self):
        from . import device as D  # raises without CUDA / without the built extension
        D._require_cuda()
        self.D = D

    fused_range_peaks = True

    def minmax(self, slab):
        return self.D.minmax(slab)

    def log(self, ext, mm, sigma, z_range=None):
        out = _buf("log_ext", ext.shape, torch.float32, ext.device)
        return self.D.log_volume(self.D.Volume(ext, mm), sigma, out=out, z_range=z_range)

    # The list-handling ops take and return tensors on the volume's device, so a slab step makes
    # one device->host copy at the very end.  Points are int64 [n, 3] rows (z, x, y).
    def peaks(self, log_ext, thr):
        """-> (raster int64 [n] over the (x, y, z) order of `log_ext`, value float32 [n])"""
        return self.D.detect_peaks(log_ext, thr)

    def rank_prune(self, raster, val, dims_zxy, cutoff_sq):
        """global peaks -> (raster in rank order, alive bool): descending value, raster ties"""
        nz, nx, ny = dims_zxy
        idx_r, _, alive = self.D.rank_prune_equal(raster.contiguous(), val.contiguous(), (1, nx, ny, nz), cutoff_sq)
        return idx_r, alive.bool()

    def gather_u8(self, vol, zxy):
        return self.D.gather_points(vol, zxy.to(torch.int32).contiguous())

    def gather_f32(self, vol, zxy):
        return self.D.gather_points(vol, zxy.to(torch.int32).contiguous())

    def blur_at_points(self, ext, mm, zxy, sigma_xyz):
        return self.D.blur_at_points_dev(self.D.Volume(ext, mm), zxy.to(torch.int32).contiguous(), sigma_xyz,
                                         "nearest")

    # masked select of the two order statistics around quantile q; `reduce` all-reduces (sum) the integer
    # tensors it is handed (sample / candidate histograms, counters)
    def masked_quantile(self, vol, mask, q, reduce):
        return self.D.masked_quantile_distributed(vol, mask, q, reduce)

    def masked_quantile_minmax(self, vol, mask, q, reduce):
        """the same, and the slab's own {min, max} over all voxels from the same pass -> (v0, v1, gamma, n, mm)"""
        mm = torch.empty(2, dtype=torch.float32, device=vol.device)
        return self.D.masked_quantile_distributed(vol, mask, q, reduce, minmax_out=mm) + (mm,)

    def log_peaks(self, ext, mm, sigma, thr, z_range=None, between=None):
        """LoG + thresholded NMS of an (extended) slab in one library call, on its planes ``z_range`` only if
        given -> (raster (x*ny + y)*nzs + z int64 [n] over those nzs planes, value float32 [n]), unordered"""
        out = _buf("log_ext", ext.shape, torch.float32, ext.device)
        _, finish = self.D.log_peaks_async(self.D.Volume(ext, mm), sigma, thr, out=out, z_range=z_range)
        if between is not None:
            between()  # every heavy kernel of the spot chain is queued; the host has not waited for any yet
        return finish()

    def edt_rowscan(self, mask):
        return self.D.edt_rowscan(mask, out=_buf("edt_dy", mask.shape, torch.uint16, mask.device))

    def edt_xz(self, dy, sampling_xyz, ny_full):
        """both envelope passes on row distances [nz][nx][ny_local] -> float32 distances, same layout"""
        f2, kind = self.D.edt_xpass(dy, sampling_xyz, ny_full, out=_buf("edt_f2y", dy.shape, torch.int32, dy.device))
        return self.D.edt_zpass(f2, kind, sampling_xyz, out=_buf("edt_disty", dy.shape, torch.float32, dy.device),
                                plane_xy=(dy.shape[1], ny_full))

    def edt_inplane(self, mask, sampling_xyz):
        return self.D.edt_inplane(mask, sampling_xyz, out=_buf("edt_f2", mask.shape, torch.int32, mask.device))

    def edt_zpass(self, f2, kind, sampling_xyz, plane_xy=None):
        return self.D.edt_zpass(f2, kind, sampling_xyz, out=_buf("edt_dist", f2.shape, torch.float32, f2.device),
                                plane_xy=plane_xy)


# ------------------------------------------------------------------------ slab-mode pipeline
class _ZDecomp:
    """z-slabs [nz_local][nx][ny]: what differs between the decompositions of the spot pipeline"""

    def __init__(self, lay: SlabLayout):
        self.lay = lay

    def detect(self, ops, img_slab, mm, sigma, t_spot, radius, group, after_log=None):
        """halo exchange + LoG + NMS; -> (extended slab, GLOBAL raster of the own peaks, their values)"""
        lay = self.lay
        nz = lay.nz
        # r planes for the z filter pass + 1 plane of LoG for the NMS (the blur at the spots reaches 1 plane)
        with _phase("halo"):
            ext, lo, hi = exchange_halo(img_slab, radius + 1, lay, group)
        self.lo = lo
        # LoG on the extended slab (planes within `radius` of an interior cut are contaminated by the boundary
        # rule and are not used); reflect applies only at the global faces
        a = radius if lo else 0
        b = ext.shape[0] - (radius if hi else 0)
        nzs = b - a
        with _phase("log"):
            if getattr(ops, "fused_range_peaks", False):  # candidate test inside the last LoG pass
                ras_l, val = ops.log_peaks(ext, mm, sigma, t_spot, z_range=(a, b), between=after_log)
            else:
                log_ext = ops.log(ext, mm, sigma, z_range=(a, b))
                if after_log is not None:
                    after_log()
                ras_l, val = ops.peaks(log_ext[a:b], t_spot)
        with _phase("peaks"):
            # local raster ((x*ny + y)*nzs + zs) -> global raster ((x*ny + y)*nz + zg); keep own planes
            zs = ras_l % nzs
            xy = ras_l // nzs
            zg = zs + (a - lo + lay.z0)
            own = (zg >= lay.z0) & (zg < lay.z1)
            return ext, (xy * nz + zg)[own], val[own]

    def own(self, z, x, y):
        return (z >= self.lay.z0) & (z < self.lay.z1)

    def mask_pts(self, z, x, y):
        return torch.stack([z - self.lay.z0, x, y], dim=1)

    def ext_pts(self, z, x, y):
        return torch.stack([z - self.lay.z0 + self.lo, x, y], dim=1)


class _XDecomp:
    """x-slabs [nz][nx_local][ny]"""

    def __init__(self, lay: XSlabLayout):
        self.lay = lay

    def detect(self, ops, img_slab, mm, sigma, t_spot, radius, group, after_log=None):
        lay = self.lay
        nz, ny = lay.nz, lay.ny
        # radius rows for the x filter pass + 1 row of LoG for the NMS; the blur at the spots (sigma 3) reaches 12
        with _phase("halo"):
            ext, lo, hi = exchange_halo_x(img_slab, max(radius + 1, 12), lay, group)
        self.lo = lo
        with _phase("log"):
            if getattr(ops, "fused_range_peaks", False):
                ras_l, val = ops.log_peaks(ext, mm, sigma, t_spot, between=after_log)
            else:
                if after_log is not None:
                    after_log()
                ras_l, val = ops.log_peaks(ext, mm, sigma, t_spot)
            # raster (x_e*ny + y)*nz + z over the extended slab
        with _phase("peaks"):
            xe = ras_l // (ny * nz)
            own = (xe >= lo) & (xe < lo + img_slab.shape[1])
            return ext, ras_l[own] + (lay.x0 - lo) * ny * nz, val[own]

    def own(self, z, x, y):
        return (x >= self.lay.x0) & (x < self.lay.x1)

    def mask_pts(self, z, x, y):
        return torch.stack([z, x - self.lay.x0, y], dim=1)

    def ext_pts(self, z, x, y):
        return torch.stack([z, x - self.lay.x0 + self.lo, y], dim=1)


def slab_find_spots_snrfilter(img_slab: torch.Tensor, mask_slab: torch.Tensor, lay, sigma=2.0,
                              t_spot=0.025, snr=None, ops=None, group=None, full_table=True, after_log=None):
    """``find_spots_snrfilter`` (modules/muscle_fish.py:281-402) on a slab-decomposed mosaic: ``lay`` a
    ``SlabLayout`` (z-slabs ``[nz_local][nx][ny]``) or an ``XSlabLayout`` (x-slabs ``[nz][nx_local][ny]``);
    ``img_slab`` float32 / ``mask_slab`` uint8.  Every rank returns the same global result:
    dict(spots (N,4) float64 [x,y,z,sigma], spots_masked, intensity, baseline, snr).
    ``full_table=False`` skips the host copy of the per-spot table of ALL mask-passing spots
    (``spots_masked`` / ``intensity``, the reference's ``spot_data``) and returns only the kept spots.
    ``after_log()`` is called once the chain's heavy kernels (min/max + P25 pass, LoG) are queued and before the
    host first waits for the peak count: the place to queue other heavy work behind them."""
    with _phase("setup"):
        ops = ops or DeviceOps()
        dec = _XDecomp(lay) if isinstance(lay, XSlabLayout) else _ZDecomp(lay)
    dev = img_slab.device
    nz, nx, ny = lay.nz, lay.nx, lay.ny
    def reduce_hist(h):
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    # 1. global min / max -- and, where the ops offer it, the P25 baseline's order statistics (step 6) from the
    #    same single pass over the raw slab (the percentile is selected on raw values and mapped afterwards)
    quant = None
    with _phase("minmax"):
        if hasattr(ops, "masked_quantile_minmax"):
            *quant, mm = ops.masked_quantile_minmax(img_slab, mask_slab, 0.25, reduce_hist)
        else:
            mm = ops.minmax(img_slab).clone()
        lo_hi = mm.to(torch.float32)
        lo_t, hi_t = lo_hi[0:1].clone(), lo_hi[1:2].clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX, group=group)
        mm = torch.cat([lo_t, hi_t])
        lo_v, hi_v = float(lo_t.item()), float(hi_t.item())
    # 2./3. halo of raw input, LoG and NMS on the extended slab; every rank keeps the peaks it owns
    radius = T.gaussian_radius(sigma)
    ext, ras_g, val = dec.detect(ops, img_slab, mm, sigma, t_spot, radius, group, after_log=after_log)
    # 4. all-gather the peak lists (padded to the longest), prune identically everywhere
    with _phase("gather_peaks"):
        n_loc = torch.tensor([ras_g.numel()], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(n_loc) for _ in range(lay.world)]
        dist.all_gather(counts, n_loc, group=group)
        counts = [int(c.item()) for c in counts]
        nmax = max(counts)
        empty = {"spots": np.empty((0, 4)), "spots_masked": np.empty((0, 4)), "intensity": np.empty(0),
                 "baseline": None, "snr": snr}
        if nmax == 0:
            return empty
        pad_r = torch.zeros(nmax, dtype=torch.int64, device=dev)
        pad_v = torch.zeros(nmax, dtype=torch.float32, device=dev)
        pad_r[:ras_g.numel()] = ras_g
        pad_v[:val.numel()] = val
        all_r = [torch.empty_like(pad_r) for _ in range(lay.world)]
        all_v = [torch.empty_like(pad_v) for _ in range(lay.world)]
        dist.all_gather(all_r, pad_r, group=group)
        dist.all_gather(all_v, pad_v, group=group)
        all_r = torch.cat([t[:c] for t, c in zip(all_r, counts)])
        all_v = torch.cat([t[:c] for t, c in zip(all_v, counts)])
    cutoff = max(T.prune_cutoff_sq(float(sigma), None, 0.5), 0)
    with _phase("prune"):
        ranked, alive = ops.rank_prune(all_r, all_v, (nz, nx, ny), cutoff)
        ranked = ranked[alive]
    with _phase("decode"):
        z = ranked % nz
        t = ranked // nz
        y = t % ny
        x = t // ny
    # 5. mask filter + 7. blurred intensity: the owner of each spot (its slab) evaluates both on its
    #    mask slab / extended raw slab; one all-reduce carries them to every rank
    ownz = dec.own(z, x, y)
    with _phase("mask_blur"):
        both = torch.zeros((2, ranked.numel()), dtype=torch.float64, device=dev)
        if bool(ownz.any()):
            zo, xo, yo = z[ownz], x[ownz], y[ownz]
            inm = ops.gather_u8(mask_slab, dec.mask_pts(zo, xo, yo))
            both[0, ownz] = (inm != 0).to(torch.float64)
            both[1, ownz] = ops.blur_at_points(ext, mm, dec.ext_pts(zo, xo, yo), (3, 3, 0.3))
        dist.all_reduce(both, op=dist.ReduceOp.SUM, group=group)
    # 6. P25 baseline over the (global) mask
    with _phase("quantile"):
        v0, v1, gamma, n = quant if quant is not None else ops.masked_quantile(img_slab, mask_slab, 0.25, reduce_hist)
    if n == 0:
        raise IndexError("percentile of an empty selection (mask selects no voxel)")
    a01 = T.interp01(np.array([v0, v1], dtype=np.float64), lo_v, hi_v)
    baseline = T.lerp(float(a01[0]), float(a01[1]), float(gamma))
    # mask filter, P90 of the intensities and the SNR cut stay on the device; only compact results
    # (int32 coordinates + float64 intensities of the masked spots) cross to the host
    with _phase("select"):
        sel = both[0] > 0
        xyz = torch.stack([x[sel], y[sel], z[sel]], dim=1).to(torch.int32)
        inten_d = both[1][sel]
        nm = int(inten_d.numel())
        if nm == 0:
            return empty
        if snr is None:
            # np.percentile(inten, 90), method 'linear': two order statistics + lerp on the host
            srt = torch.sort(inten_d).values
            _, k, g = T.percentile_virtual_index(nm, 90)
            pair = srt[[k, min(k + 1, nm - 1)]].tolist()
            thresh = T.lerp(pair[0], pair[1], g) / (baseline * np.sqrt(10))
            snr = max(thresh, np.sqrt(10))
        keep_d = inten_d / baseline > snr
        kept_d = xyz[keep_d]

    def table(xyz_host):
        out = np.empty((len(xyz_host), 4), dtype=np.float64)
        out[:, :3] = xyz_host
        out[:, 3] = float(sigma)
        return out

    with _phase("d2h"):
        kept_h = kept_d.cpu().numpy()
        if full_table:
            xyz_h = xyz.cpu().numpy()
            inten = inten_d.cpu().numpy()
    with _phase("host_tail"):
        res = {"spots": table(kept_h), "baseline": baseline, "snr": snr, "n_masked": nm}
        if full_table:
            res["spots_masked"] = table(xyz_h)
            res["intensity"] = inten
    return res


def slab_edt(notnuc_slab: torch.Tensor, lay: SlabLayout, sampling_xyz=(1.0, 1.0, 1.0), ops=None, group=None,
             back_to_z=False):
    """``distance_transform_edt(~region_nuc, sampling)`` (nocodazole/fish_analysis_noco.py:218) on a
    z-slab decomposed mask.  Returns the distance map in x-slab layout ``[nz][nx_local][ny]`` (or
    back in z-slab layout when ``back_to_z``)."""
    ops = ops or DeviceOps()
    with _phase("edt_inplane"):
        f2, kind = ops.edt_inplane(notnuc_slab, sampling_xyz)
    with _phase("edt_all_to_all"):
        f2x = repartition_z_to_x(f2, lay, group)
    with _phase("edt_zpass"):
        dx = ops.edt_zpass(f2x, kind, sampling_xyz, plane_xy=(lay.nx, lay.ny))
    return repartition_x_to_z(dx, lay, group) if back_to_z else dx


def slab_edt_rows(notnuc_slab: torch.Tensor, lay: SlabLayout, sampling_xyz=(1.0, 1.0, 1.0), ops=None, group=None):
    """The same transform with the cut moved in front of BOTH envelope passes: row scan on the z-slab, one
    all-to-all of the uint16 row distances into y-slabs (half the bytes of ``slab_edt``'s 4-byte in-plane squares),
    x and z passes there.  Returns the distance map in y-slab layout ``[nz][nx][ny_local]``
    (``yslab_spot_distances`` looks spots up in it)."""
    ops = ops or DeviceOps()
    with _phase("edt_rowscan"):
        dy = ops.edt_rowscan(notnuc_slab)
    with _phase("edt_all_to_all"):
        dyy = repartition_z_to_y(dy, lay, group)
    with _phase("edt_xz"):
        return ops.edt_xz(dyy, sampling_xyz, lay.ny)


def slab_spot_distances(dist_xslab: torch.Tensor, spots_xyz: np.ndarray, lay: SlabLayout, ops=None, group=None):
    """EDT value at each spot (the interpn gather of nocodazole/fish_analysis_noco.py:297-298) from the
    x-slab partitioned distance map; every rank gets the full vector."""
    ops = ops or DeviceOps()
    dev = dist_xslab.device
    s = torch.as_tensor(np.ascontiguousarray(np.asarray(spots_xyz)[:, :3]).astype(np.int64)).to(dev)
    out = torch.zeros(s.shape[0], dtype=torch.float64, device=dev)
    if s.shape[0]:
        x = s[:, 0]
        own = (x >= lay.x0) & (x < lay.x1)
        if bool(own.any()):
            so = s[own]
            loc = torch.stack([so[:, 2], so[:, 0] - lay.x0, so[:, 1]], dim=1)
            out[own] = ops.gather_f32(dist_xslab, loc).to(torch.float64)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.cpu().numpy()


def xslab_edt(notnuc_slab: torch.Tensor, lay: XSlabLayout, sampling_xyz=(1.0, 1.0, 1.0), ops=None, group=None):
    """The distance transform on an x-slab decomposed mask ``[nz][nx_local][ny]``: row scan on the x-slab, ONE
    all-to-all of the uint16 row distances into y-slabs, both envelope passes there.  Returns the distance
    map in y-slab layout ``[nz][nx][ny_local]``."""
    ops = ops or DeviceOps()
    with _phase("edt_rowscan"):
        dy = ops.edt_rowscan(notnuc_slab)
    with _phase("edt_all_to_all"):
        dyy = repartition_x_to_y(dy, lay, group)
    with _phase("edt_xz"):
        return ops.edt_xz(dyy, sampling_xyz, lay.ny)


def yslab_spot_distances(dist_yslab: torch.Tensor, spots_xyz: np.ndarray, lay, ops=None, group=None):
    """EDT value at each spot from the y-slab partitioned distance map (``lay``: either layout; both carry the
    y bounds); every rank gets the full vector."""
    ops = ops or DeviceOps()
    dev = dist_yslab.device
    s = torch.as_tensor(np.ascontiguousarray(np.asarray(spots_xyz)[:, :3]).astype(np.int64)).to(dev)
    out = torch.zeros(s.shape[0], dtype=torch.float64, device=dev)
    if s.shape[0]:
        y = s[:, 1]
        own = (y >= lay.y0) & (y < lay.y1)
        if bool(own.any()):
            so = s[own]
            loc = torch.stack([so[:, 2], so[:, 0], so[:, 1] - lay.y0], dim=1)
            out[own] = ops.gather_f32(dist_yslab, loc).to(torch.float64)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out.cpu().numpy()


xslab_spot_distances = yslab_spot_distances


def slab_analyze(img_slab: torch.Tensor, mask_slab: torch.Tensor, notnuc_slab: torch.Tensor, lay,
                 sigma=2.0, t_spot=0.025, sampling_xyz=(1.0, 1.0, 1.0), snr=None, ops=None, group=None,
                 edt_group=None, full_table=True, edt_wire="rows"):
    """The whole path on a slab-decomposed mosaic (``lay``: ``SlabLayout`` = z-slabs, ``XSlabLayout`` = x-slabs):
    ``slab_find_spots_snrfilter`` + the distance transform + the per-spot distances.  The distance transform
    does not depend on the spots and its chain has no host round trip: given ``edt_group`` (a second
    communicator over the same ranks, ``dist.new_group()``) on CUDA tensors it is queued first on a side
    stream, so its all-to-all and kernels run beside the spot chain's halo exchange, small collectives and
    host round trips.  ``edt_wire`` (z-slabs): "rows" = uint16 row distances cross the cut, both envelope passes
    run on y-slabs (``slab_edt_rows``); "squares" = the 4-byte in-plane squares cross it, z pass on x-slabs
    (``slab_edt``).  Returns the spot dict plus ``spot_dist`` (float64 per kept spot) and ``dist_slab`` (the
    distance map: y-slab partitioned, or x-slab partitioned for z-slabs with ``edt_wire="squares"``)."""
    xs = isinstance(lay, XSlabLayout)
    squares = (not xs) and edt_wire == "squares"
    edt_fn = xslab_edt if xs else (slab_edt if squares else slab_edt_rows)
    dist_fn = slab_spot_distances if squares else yslab_spot_distances
    if edt_group is not None and img_slab.is_cuda:
        from . import device as D
        import os
        cur = torch.cuda.current_stream()
        side = D._side_stream()
        side.wait_stream(cur)
        gated = (not squares) and os.environ.get("MFISH_SLAB_GATE", "0") == "1"
        if gated:
            # MFISH_SLAB_GATE=1 (experiment, off): the transform's light front (row scan + re-partition) beside the
            # spot chain's heavy front, its envelope passes queued BEHIND the LoG so that they run while the spot
            # chain's end waits on its round trips.  Measured (tools/slab_modes.py) no better than one chain after
            # the other (2 GPUs 38.9 vs 38.3 ms side by side; 8 GPUs x-slabs 32.2 vs 26.6): the small kernels and
            # collectives of the chain's end queue behind the envelope passes' blocks.
            ops_ = ops or DeviceOps()
            with torch.cuda.stream(side):
                with _phase("edt_rowscan"):
                    dy = ops_.edt_rowscan(notnuc_slab)
                with _phase("edt_all_to_all"):
                    dyy = (repartition_x_to_y if xs else repartition_z_to_y)(dy, lay, edt_group)
            late = {}

            def after_log():
                if "dmap" in late:
                    return
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    with _phase("edt_xz"):
                        late["dmap"] = ops_.edt_xz(dyy, sampling_xyz, lay.ny)
        else:
            after_log = None
            with torch.cuda.stream(side):
                dmap = edt_fn(notnuc_slab, lay, sampling_xyz, ops=ops, group=edt_group)
        # The spot chain on the high-priority stream (its LoG passes dispatched ahead of the transform's blocks, its
        # latency-bound end overlapping the envelope passes).  Measured (tools/slab_modes.py, median ms, z- / x-slabs,
        # without -> with): 2 GPUs 38.3 -> 35.8 / 39.5 -> 36.7; 4 GPUs 48.8 -> 49.4 / 45.1 -> 40.9; 8 GPUs
        # 35.6 -> 36.8 / 26.6 -> 30.4 (the overlap is lost: as slow as one chain after the other).  Hence: on up to
        # 4 ranks, off beyond; MFISH_SLAB_PRIO=0/1 overrides.
        import os
        prio = os.environ.get("MFISH_SLAB_PRIO", "auto")
        use_prio = prio == "1" or (prio not in ("0", "1") and lay.world <= 4)
        spot = D._list_stream() if use_prio else cur
        spot.wait_stream(cur)
        with torch.cuda.stream(spot):
            res = slab_find_spots_snrfilter(img_slab, mask_slab, lay, sigma=sigma, t_spot=t_spot, snr=snr, ops=ops,
                                            group=group, full_table=full_table, after_log=after_log)
        if gated:
            after_log()  # (a chain that returned before its LoG)
            dmap = late["dmap"]
        cur.wait_stream(spot)
        cur.wait_stream(side)
        dmap.record_stream(cur)
    else:
        res = slab_find_spots_snrfilter(img_slab, mask_slab, lay, sigma=sigma, t_spot=t_spot, snr=snr, ops=ops,
                                        group=group, full_table=full_table)
        dmap = edt_fn(notnuc_slab, lay, sampling_xyz, ops=ops, group=group)
    res["spot_dist"] = dist_fn(dmap, res["spots"], lay, ops=ops, group=group)
    res["dist_slab"] = dmap
    if squares:
        res["dist_xslab"] = dmap
    return res
