"""Device-side pipeline: torch CUDA tensors as buffer carriers, every voxel
operation a call into ``libmfish_sm100.so`` on torch's current stream.

Device volumes are ``[nz, nx, ny]`` (y contiguous); the reference's logical
order is ``(x, y, z)`` (``modules/muscle_fish.py:60-71``).  ``Volume`` keeps the
raw float32 voxels plus the global (min, max) pair; ``su.normalize_image`` is
never materialised -- the filters map ``(v - min) / (max - min)`` on load.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L
from . import taps as T

_DTYPES = {
    np.dtype(np.uint8): L.U8, np.dtype(np.bool_): L.U8, np.dtype(np.uint16): L.U16,
    np.dtype(np.int16): L.I16, np.dtype(np.int32): L.I32, np.dtype(np.int64): L.I64,
    np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64,
}
_TORCH_DTYPES = {
    torch.uint8: L.U8, torch.bool: L.U8, torch.uint16: L.U16, torch.int16: L.I16, torch.int32: L.I32,
    torch.int64: L.I64, torch.float32: L.F32, torch.float64: L.F64,
}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("muscle-fish_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def host_to_device(arr) -> torch.Tensor:
    """Copy a host array to the device in its own dtype (pinned sources go async)."""
    _require_cuda()
    if isinstance(arr, torch.Tensor):
        if arr.is_cuda:
            return arr.contiguous()
        t = arr.contiguous()
        if t.dtype == torch.bool:
            t = t.view(torch.uint8)
        return t.to("cuda", non_blocking=t.is_pinned())
    a = np.asarray(arr)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if a.dtype not in _DTYPES:
        a = a.astype(np.float64)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    if not a.flags.writeable:  # torch.from_numpy warns on read-only arrays
        a = a.copy()
    t = torch.from_numpy(a)
    return t.to("cuda", non_blocking=t.is_pinned())


STAGED_UPLOADS = os.environ.get("MFISH_STAGED_UPLOADS", "1") != "0"
_COPY_STREAMS = {}
_SIDE_STREAMS = {}
_LIST_STREAMS = {}


def _side_stream():
    dev = torch.cuda.current_device()
    st = _SIDE_STREAMS.get(dev)
    if st is None:
        st = _SIDE_STREAMS[dev] = torch.cuda.Stream()
    return st


def _list_stream():
    """High-priority stream for the spot-list handling (tens of tiny kernels and a few host round
    trips): its CTAs are dispatched ahead of the remaining CTAs of the distance transform running
    on the side stream, instead of queueing behind each full-grid kernel."""
    dev = torch.cuda.current_device()
    st = _LIST_STREAMS.get(dev)
    if st is None:
        st = _LIST_STREAMS[dev] = torch.cuda.Stream(priority=-1)
    return st


# ---- pageable host arrays: narrowed on the host by worker threads, staged through pinned memory
# The reference's callers hand over pageable float64 images (fix_bleaching's output) and int64 0/1 masks
# (np.where(..., 1, 0)): general_pipeline/fish_analysis.py:196-199.  Copying those as they are moves 8 bytes per
# voxel per array through the driver's own synchronous staging (~11 GB/s).  The device converts an image to
# float32 and a mask to "voxel != 0" on ingest anyway, so the same conversion is done on the host instead --
# numpy ufuncs on chunks in a few threads (they release the GIL) writing into a ring of pinned buffers -- and
# 4 (image) or 1 (mask) bytes per voxel cross PCIe asynchronously while later chunks are being converted.
_STAGE_MIN_BYTES = 64 << 20
_STAGE_CHUNK_ELEMS = 8 << 20
_STAGE = {}


def _is_pinned(a: np.ndarray) -> bool:
    """is this numpy array backed by page-locked memory (then the plain asynchronous copy is the fast path)?"""
    try:
        if not a.flags.c_contiguous:
            return False
        b = a if a.dtype != np.bool_ else a.view(np.uint8)
        if not b.flags.writeable:
            return False
        return torch.from_numpy(b).is_pinned()
    except Exception:  # noqa: BLE001
        return False


def _stage_resources():
    ent = _STAGE.get("pool")
    if ent is None:
        from concurrent.futures import ThreadPoolExecutor
        try:
            ncpu = len(os.sched_getaffinity(0))
        except AttributeError:
            ncpu = os.cpu_count() or 1
        nthreads = max(1, min(16, ncpu // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        depth = nthreads + 2
        ring = [torch.empty(_STAGE_CHUNK_ELEMS * 4, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        ent = _STAGE["pool"] = (ThreadPoolExecutor(nthreads), ring, [None] * depth)
    return ent


# ---- masks as bits: a mask is one bit of information per voxel, and two whole-byte masks are a third of the PCIe
# bytes of a float32 stack.  Worker threads pack "voxel != 0" (mfish_host_pack_mask, plain C in the library, the
# GIL released) into pinned memory while the image is crossing PCIe; the device expands the bits straight into the
# uint8 [z][x][y] layout (mfish_ingest_mask_bits).  This pays where PCIe is the limit: one or two processes on the
# host (C2 host-buffer step 52.8 -> 44.3 ms).  With four or eight ranks the host's memory system is the limit
# (8 x 2.5 GB of pinned copies at once run at 23 GB/s per GPU), and the packing threads' reads
# of the masks add to its load: measured at 4 ranks, 66.9 ms against ~57 ms expected from whole-byte copies, the
# batch 7.22 against 6.83 ms per stack.  Default: bits with at most two ranks per host, whole bytes beyond;
# MFISH_PACK_MASKS=1 / 0 forces either.
_pack_env = os.environ.get("MFISH_PACK_MASKS")
PACK_MASKS = (_pack_env != "0") if _pack_env is not None else int(os.environ.get("LOCAL_WORLD_SIZE", "1")) <= 2
_PACK_MIN_ELEMS = 16 << 20
_PACK_CHUNK_ELEMS = 4 << 20  # a multiple of 64
_PACK_RING = 4
_PACK_SIZES_KEPT = 3
_SPLIT_MIN_BYTES = 64 << 20  # a pinned array this large is sent in pieces, a pending packed mask between them
_SPLIT_PIECES = 4
H2D_BYTES = 0  # bytes queued for upload by host_to_device_many since import (bench.py reads differences)


@dataclass
class PackedMask:
    """a mask on the device as the bit stream of its C-ordered host array (``ingest`` expands it)"""
    bits: torch.Tensor  # uint8, mfish_mask_bits_bytes(n) long
    shape: tuple


def _packable(arr, min_elems=None):
    if isinstance(arr, torch.Tensor):
        if arr.is_cuda or not arr.is_contiguous():
            return None
        arr = arr.view(torch.uint8).numpy() if arr.dtype == torch.bool else arr.numpy()
    a = np.asarray(arr)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    if a.ndim != 3 or a.size < (_PACK_MIN_ELEMS if min_elems is None else min_elems) or not a.flags.c_contiguous \
            or a.dtype not in _DTYPES:
        return None
    return a


def _packable_any_size(arr):
    """``_packable`` without the size threshold (run_batch packs every stack's masks)"""
    return _packable(arr, min_elems=0)


def pack_into(a: np.ndarray, bits: torch.Tensor):
    """queue the packing of the C-contiguous mask ``a`` into the (pinned) uint8 tensor ``bits`` on the worker
    threads -> futures"""
    pool = _stage_resources()[0]
    src, dst, code, esz = a.ctypes.data, bits.data_ptr(), _DTYPES[a.dtype], a.dtype.itemsize
    fn = L.lib().mfish_host_pack_mask

    def job(i0, n):
        L.check(fn(src + i0 * esz, code, n, dst + i0 // 8, 0))

    return [pool.submit(job, i0, min(_PACK_CHUNK_ELEMS, a.size - i0)) for i0 in range(0, a.size, _PACK_CHUNK_ELEMS)]


def _start_pack(a: np.ndarray):
    """queue the packing of ``a`` on the worker threads -> (futures, pinned bits tensor, ring slot)"""
    nbytes = int(L.lib().mfish_mask_bits_bytes(a.size))
    key = ("bits", nbytes)
    ring = _STAGE.get(key)
    if ring is None:
        ring = _STAGE[key] = {"bufs": [], "events": [], "next": 0}
        # pinned rings are kept for the few most recent stack sizes only (a long-running caller with stacks of many
        # shapes must not accumulate page-locked memory)
        sizes = _STAGE.setdefault("bits_sizes", [])
        sizes.append(key)
        while len(sizes) > _PACK_SIZES_KEPT:
            old = _STAGE.pop(sizes.pop(0), None)
            for ev in (old or {}).get("events", []):
                if ev is not None:
                    ev.synchronize()
    if len(ring["bufs"]) < _PACK_RING:
        ring["bufs"].append(torch.zeros(nbytes, dtype=torch.uint8).pin_memory())
        ring["events"].append(None)
        slot = len(ring["bufs"]) - 1
    else:
        slot = ring["next"] % _PACK_RING
        ring["next"] += 1
        if ring["events"][slot] is not None:
            ring["events"][slot].synchronize()  # the previous DMA out of this buffer has finished
    buf = ring["bufs"][slot]
    return pack_into(a, buf), buf, (ring, slot), a


def _upload_staged(a: np.ndarray, kind: str, cs, cur):
    """pageable (x, y, z) array -> (device tensor, ready event): masks arrive as uint8 0/1, float64 / int64 images
    as float32, other images in their own dtype"""
    import threading
    pool, ring, events = _stage_resources()
    depth = len(ring)
    if kind == "mask":
        out_np = np.dtype(np.uint8)
    elif a.dtype in (np.dtype(np.float64), np.dtype(np.int64)):
        out_np = np.dtype(np.float32)
    else:
        out_np = a.dtype  # 8 / 16 / 32-bit images travel as they are
    out_t = torch.from_numpy(np.empty(0, out_np)).dtype
    esz = out_np.itemsize
    n0 = a.shape[0]
    per_row = int(np.prod(a.shape[1:]))
    rows = max(1, _STAGE_CHUNK_ELEMS // per_row)
    if per_row > _STAGE_CHUNK_ELEMS:  # one leading-axis row does not fit a ring buffer: the plain path
        return None
    chunks = [(r0, min(n0, r0 + rows)) for r0 in range(0, n0, rows)]
    dst = torch.empty(a.shape, dtype=out_t, device="cuda")
    released = [threading.Event() for _ in chunks]  # chunk c's DMA has been queued (its slot's event recorded)

    def convert(c):
        r0, r1 = chunks[c]
        slot = c % depth
        if c >= depth:
            released[c - depth].wait()
        ev = events[slot]
        if ev is not None:
            ev.synchronize()  # the previous DMA out of this pinned buffer has finished
        view = ring[slot][:(r1 - r0) * per_row * esz].numpy().view(out_np).reshape((r1 - r0,) + a.shape[1:])
        if kind == "mask":
            np.not_equal(a[r0:r1], 0, out=view.view(np.bool_))
        else:
            np.copyto(view, a[r0:r1], casting="unsafe")
        return slot

    futs = [pool.submit(convert, c) for c in range(len(chunks))]
    cs.wait_stream(cur)
    for c, f in enumerate(futs):
        slot = f.result()
        r0, r1 = chunks[c]
        src = ring[slot][:(r1 - r0) * per_row * esz].view(out_t).view((r1 - r0,) + tuple(a.shape[1:]))
        with torch.cuda.stream(cs):
            dst[r0:r1].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        events[slot] = ev
        released[c].set()
    return dst, ev


def host_to_device_many(arrs, kinds=None, pack_masks=False):
    """Queue the H2D copies of several host arrays back to back on a side stream and return
    [(device tensor, ready event)], so that work on the first array overlaps the later copies
    (PCIe is the slow link of the host-buffer path).  Device / non-array inputs pass through.
    ``kinds`` (per array: "image" / "mask" / None) tells what the array is FOR: large pageable arrays of a known
    kind are narrowed on the host (float32 / 0-1 uint8: exactly what ``ingest`` makes of them) while they are
    staged through pinned memory.
    ``pack_masks``: the caller hands every "mask" result to ``ingest`` (which expands a ``PackedMask``), so large
    masks travel as bits.  All masks are packed by the worker threads from the start; a mask listed BEFORE a large
    pinned array is sent between the pieces of that array, as soon as it is packed (PCIe is never left idle),
    the others after it."""
    global H2D_BYTES
    _require_cuda()
    dev = torch.cuda.current_device()
    cur = torch.cuda.current_stream()
    cs = _COPY_STREAMS.get(dev)
    if cs is None:
        cs = _COPY_STREAMS[dev] = torch.cuda.Stream()
    kinds = list(kinds) if kinds is not None else [None] * len(arrs)
    out = [None] * len(arrs)
    started = False

    def start():
        nonlocal started
        if not started:
            cs.wait_stream(cur)  # the blocks may have been in use by earlier compute work
            started = True

    jobs = {}
    if pack_masks and PACK_MASKS:
        for i, (arr, kind) in enumerate(zip(arrs, kinds)):
            a = _packable(arr) if (kind == "mask" and arr is not None) else None
            if a is not None:
                jobs[i] = _start_pack(a)
    pending = []

    def flush(only_finished=False):
        global H2D_BYTES
        for i in list(pending):
            futs, buf, (ring, slot), a = jobs[i]
            if only_finished and not all(f.done() for f in futs):
                break  # masks keep their order
            for f in futs:
                f.result()
            start()
            dst = torch.empty(buf.shape, dtype=torch.uint8, device="cuda")
            with torch.cuda.stream(cs):
                dst.copy_(buf, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            ring["events"][slot] = ev
            H2D_BYTES += buf.numel()
            out[i] = (PackedMask(dst, tuple(a.shape)), ev)
            pending.remove(i)

    for i, (arr, kind) in enumerate(zip(arrs, kinds)):
        if i in jobs:
            pending.append(i)
            continue
        if arr is None:
            out[i] = (None, None)
            continue
        if isinstance(arr, PackedMask):  # already on the device as bits (parallel.run_batch)
            out[i] = (arr, None)
            continue
        if isinstance(arr, torch.Tensor):
            src = arr if arr.is_cuda else arr.contiguous()
        else:
            a = np.asarray(arr)
            if (kind in ("image", "mask") and a.ndim == 3 and a.nbytes >= _STAGE_MIN_BYTES and a.dtype in _DTYPES
                    and STAGED_UPLOADS and not _is_pinned(a)):
                flush()
                start()
                staged = _upload_staged(a, kind, cs, cur)
                if staged is not None:
                    H2D_BYTES += staged[0].numel() * staged[0].element_size()
                    out[i] = staged
                    continue
            if a.dtype == np.bool_:
                a = a.view(np.uint8)
            if a.dtype not in _DTYPES:
                a = a.astype(np.float64)
            if not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)
            if not a.flags.writeable:
                a = a.copy()
            src = torch.from_numpy(a)
        if src.dtype == torch.bool:
            src = src.view(torch.uint8)
        if src.is_cuda:
            out[i] = (src.contiguous(), None)
            continue
        H2D_BYTES += src.numel() * src.element_size()
        if not src.is_pinned():
            # pageable memory: the driver stages the copy synchronously anyway
            flush()
            out[i] = (src.to("cuda"), None)
            continue
        dst = torch.empty(src.shape, dtype=src.dtype, device="cuda")  # owned by the compute stream
        start()
        pieces = _SPLIT_PIECES if (pending and src.dim() >= 1 and src.shape[0] >= _SPLIT_PIECES
                                   and src.numel() * src.element_size() >= _SPLIT_MIN_BYTES) else 1
        if pieces == 1:
            with torch.cuda.stream(cs):
                dst.copy_(src, non_blocking=True)
        else:
            # The array goes in pieces, one piece queued ahead of the one in flight, so PCIe never idles; a mask
            # listed before it is queued the moment its packing is finished (checked while waiting for the DMA
            # to move on), i.e. behind at most two pieces -- its consumer (the distance transform) then runs
            # under the rest of the upload.  A mask that is still being packed when the last piece is queued
            # (a host whose memory system is shared by eight ranks) follows the array instead of stalling it.
            import time
            bounds = [src.shape[0] * k // pieces for k in range(pieces + 1)]
            done_ev = []

            def queue_piece(k):
                with torch.cuda.stream(cs):
                    dst[bounds[k]:bounds[k + 1]].copy_(src[bounds[k]:bounds[k + 1]], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(cs)
                done_ev.append(e)

            queue_piece(0)
            queue_piece(1)
            for k in range(2, pieces):
                while not done_ev[k - 2].query():
                    flush(only_finished=True)
                    time.sleep(1e-4)
                flush(only_finished=True)
                queue_piece(k)
        with torch.cuda.stream(cs):
            ev = torch.cuda.Event()
            ev.record(cs)
        out[i] = (dst, ev)
    flush()
    return out


@dataclass
class Volume:
    """One channel on the device: raw float32 voxels [nz, nx, ny] + {min, max}."""
    data: torch.Tensor
    minmax: torch.Tensor | None  # float32[2] on device, or None (use voxels as they are)

    @property
    def shape_zxy(self):
        return tuple(self.data.shape)

    @property
    def shape_xyz(self):
        nz, nx, ny = self.data.shape
        return (nx, ny, nz)

    @property
    def nvox(self):
        return self.data.numel()

    def minmax_host(self):
        lo, hi = self.minmax.tolist()
        return float(lo), float(hi)


def ingest(src, z_first=False, as_mask=False, want_minmax=True, negate=False, ready=None):
    """Host/device array in reference order (x,y,z) [or (z,x,y) when ``z_first``]
    -> device layout.  Returns ``Volume`` (float32 + minmax) or a uint8 mask tensor.
    ``ready``: event of an asynchronous copy of ``src`` (see ``host_to_device_many``)."""
    if isinstance(src, PackedMask):
        if not as_mask:
            raise TypeError("a bit-packed array is a mask: ingest(..., as_mask=True)")
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
            src.bits.record_stream(torch.cuda.current_stream())
        if len(src.shape) != 3:
            raise ValueError("expected a 3-D volume, got shape %r" % (tuple(src.shape),))
        if z_first:
            nz, nx, ny = src.shape
        else:
            nx, ny, nz = src.shape
        out = torch.empty((nz, nx, ny), dtype=torch.uint8, device=src.bits.device)
        L.check(L.lib().mfish_ingest_mask_bits(src.bits.data_ptr(), L.LAYOUT_ZXY if z_first else L.LAYOUT_XYZ, nx, ny, nz,
                                               out.data_ptr(), L.DST_ISZERO if negate else L.DST_NONZERO, _stream()))
        return out
    t = host_to_device(src)
    if ready is not None:
        torch.cuda.current_stream().wait_event(ready)
        t.record_stream(torch.cuda.current_stream())  # allocated by the copy's caller, consumed here
    if t.dim() != 3:
        raise ValueError("expected a 3-D volume, got shape %r" % (tuple(t.shape),))
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    if t.dtype not in _TORCH_DTYPES:
        raise TypeError(f"unsupported dtype {t.dtype}")
    if z_first:
        nz, nx, ny = t.shape
        layout = L.LAYOUT_ZXY
    else:
        nx, ny, nz = t.shape
        layout = L.LAYOUT_XYZ
    if as_mask:
        out = torch.empty((nz, nx, ny), dtype=torch.uint8, device=t.device)
        L.check(L.lib().mfish_ingest(t.data_ptr(), _TORCH_DTYPES[t.dtype], layout, nx, ny, nz, out.data_ptr(),
                                     L.DST_ISZERO if negate else L.DST_NONZERO, None, _stream()))
        return out
    out = torch.empty((nz, nx, ny), dtype=torch.float32, device=t.device)
    mm = torch.empty(2, dtype=torch.float32, device=t.device) if want_minmax else None
    L.check(L.lib().mfish_ingest(t.data_ptr(), _TORCH_DTYPES[t.dtype], layout, nx, ny, nz, out.data_ptr(),
                                 L.DST_F32, mm.data_ptr() if mm is not None else None, _stream()))
    return Volume(out, mm)


def ingest_as_float(src):
    """``skimage.util.img_as_float`` semantics for the inputs of blob_log / filters.gaussian: floating
    images as they are, unsigned integers divided by the type's maximum, bool as 0/1.  The scale is
    applied on load by the first filter pass (``Volume.minmax`` = {0, type max})."""
    a = src if isinstance(src, torch.Tensor) else np.asarray(src)
    dt = np.dtype(str(a.dtype).replace("torch.", "")) if isinstance(a, torch.Tensor) else a.dtype
    vol = ingest(a, want_minmax=False)
    if dt.kind == "u":
        vol.minmax = torch.tensor([0.0, float(np.iinfo(dt).max)], dtype=torch.float32, device=vol.data.device)
    elif dt.kind == "i":
        raise NotImplementedError("signed integer images: convert with skimage.util.img_as_float first")
    return vol


class Handoff:
    """Device-resident hand-off between the reference's consecutive calls.

    ``su.normalize_image`` -> ``mf.fix_bleaching`` -> ``mf.find_spots_snrfilter``
    (general_pipeline/fish_analysis.py:132-142,196-199) pass the SAME float64 volume from one call to the
    next through host memory: 8 bytes per voxel up and down PCIe at every step.  The functions here return
    real host float64 arrays (the scripts index, clip and plot them), but remember, per returned array,
    the device volume that holds exactly what ``ingest(array)`` would produce; when that very array object
    comes back as an argument the upload is skipped.  To keep the pairing sound the returned arrays are
    read-only (a script that wants to edit one takes a ``.copy()``, which is then uploaded like any other
    array); the entry dies with the array.  ``MFISH_HANDOFF=0`` switches the mechanism off."""

    def __init__(self):
        self._entries = {}
        self.enabled = os.environ.get("MFISH_HANDOFF", "1") != "0"
        self.hits = 0

    def register(self, arr: np.ndarray, vol: "Volume", z_first: bool):
        if not self.enabled or not isinstance(arr, np.ndarray) or arr.ndim != 3:
            return arr
        import weakref
        arr.flags.writeable = False
        key = id(arr)
        self._entries[key] = (weakref.ref(arr, lambda _r, k=key: self._entries.pop(k, None)), vol, bool(z_first))
        return arr

    def lookup(self, arr, z_first: bool):
        if not self.enabled or not isinstance(arr, np.ndarray):
            return None
        e = self._entries.get(id(arr))
        if e is None or e[0]() is not arr or arr.flags.writeable or e[2] != bool(z_first):
            return None
        self.hits += 1
        return e[1]

    def clear(self):
        self._entries.clear()


HANDOFF = Handoff()
# the same for MASKS this package returned (threshold_fiber's result is the `mask=` of every later call of
# general_pipeline/fish_analysis.py:152-199): the payload is the uint8 device mask [nz][nx][ny]
MASK_HANDOFF = Handoff()


def ingest_mask(src, z_first=False, negate=False, ready=None):
    """``ingest(src, as_mask=True)``, skipping the upload for a mask array this package returned itself"""
    m = MASK_HANDOFF.lookup(src, z_first)
    if m is not None:
        return (m == 0).to(torch.uint8) if negate else m
    return ingest(src, z_first=z_first, as_mask=True, negate=negate, ready=ready)


def minmax(vol: torch.Tensor) -> torch.Tensor:
    mm = torch.empty(2, dtype=torch.float32, device=vol.device)
    L.check(L.lib().mfish_minmax_f32(vol.data_ptr(), vol.numel(), mm.data_ptr(), _stream()))
    return mm


def normalize_f64(src: torch.Tensor, vmin=0.0, vmax=1.0) -> torch.Tensor:
    """np.interp(src, (min, max), (vmin, vmax)) as float64, same element order."""
    src = src.contiguous()
    if src.dtype == torch.bool:
        src = src.view(torch.uint8)
    out = torch.empty(src.shape, dtype=torch.float64, device=src.device)
    ws = torch.empty(2, dtype=torch.int64, device=src.device)
    L.check(L.lib().mfish_normalize_f64(src.data_ptr(), _TORCH_DTYPES[src.dtype], src.numel(), float(vmin),
                                        float(vmax), out.data_ptr(), ws.data_ptr(), _stream()))
    return out


def export_f64(vol_zxy: torch.Tensor, z_first=False, out: torch.Tensor | None = None) -> torch.Tensor:
    """device [nz,nx,ny] float32 -> float64 in the reference's (x,y,z) order (or (z,x,y))."""
    nz, nx, ny = vol_zxy.shape
    shp = (nz, nx, ny) if z_first else (nx, ny, nz)
    if out is None:
        out = torch.empty(shp, dtype=torch.float64, device=vol_zxy.device)
    L.check(L.lib().mfish_export_f64(vol_zxy.data_ptr(), nx, ny, nz,
                                     L.LAYOUT_ZXY if z_first else L.LAYOUT_XYZ, out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------ filters
class Workspace:
    """Scratch volumes reused across calls (keyed by shape) so steady-state steps do not allocate."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, shape, dtype, device):
        k = (key, tuple(shape), dtype, str(device))
        b = self._bufs.get(k)
        if b is None:
            b = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[k] = b
        return b

    def clear(self):
        self._bufs.clear()


WS = Workspace()


def log_volume(vol: Volume, sigma: float, out: torch.Tensor | None = None, z_range=None) -> torch.Tensor:
    """-sigma^2 * gaussian_laplace(normalize(img), sigma): one scale of blob_log's cube.
    ``z_range=(z0, z1)``: only those planes of the result are computed (z-extended slabs)."""
    d = vol.data
    shp, dev = d.shape, d.device
    if out is None:
        out = torch.empty(shp, dtype=torch.float32, device=dev)
    t1 = WS.get("log1", shp, torch.float32, dev)
    t2 = WS.get("log2", shp, torch.float32, dev)
    t3 = WS.get("log3", shp, torch.float32, dev)
    nz, nx, ny = shp
    z0, z1 = (0, nz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    L.check(L.lib().mfish_log3d_range_f32(d.data_ptr(), out.data_ptr(), t1.data_ptr(), t2.data_ptr(), t3.data_ptr(),
                                          nz, nx, ny, float(sigma),
                                          vol.minmax.data_ptr() if vol.minmax is not None else None, z0, z1,
                                          _stream()))
    return out


def log_cube(vol: Volume, sigmas) -> torch.Tensor:
    """[ns, nz, nx, ny] scale-normalised negative LoG cube."""
    sigmas = [float(s) for s in sigmas]
    cube = torch.empty((len(sigmas),) + tuple(vol.data.shape), dtype=torch.float32, device=vol.data.device)
    for k, s in enumerate(sigmas):
        log_volume(vol, s, out=cube[k])
    return cube


def gaussian(vol: Volume, sigma_xyz, mode="nearest") -> torch.Tensor:
    """skimage.filters.gaussian(normalize(img), sigma) on the device, [nz,nx,ny] float32."""
    d = vol.data
    if np.isscalar(sigma_xyz):
        sigma_xyz = (sigma_xyz,) * 3
    sx, sy, sz = [float(s) for s in sigma_xyz]
    out = torch.empty_like(d)
    tmp = WS.get("gtmp", d.shape, torch.float32, d.device)
    nz, nx, ny = d.shape
    b = {"nearest": L.BOUNDARY_NEAREST, "reflect": L.BOUNDARY_REFLECT}[mode]
    sig = L.darray([sz, sx, sy])
    L.check(L.lib().mfish_gaussian3d_f32(d.data_ptr(), out.data_ptr(), tmp.data_ptr(), nz, nx, ny, sig, b,
                                         vol.minmax.data_ptr() if vol.minmax is not None else None, _stream()))
    return out


# ------------------------------------------------------------ spot detection
_DEFAULT_CAPACITY = 1 << 20


def _launch_peaks(cube, thr, cap):
    ns, nz, nx, ny = cube.shape
    dev = cube.device
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    idx = WS.get("pk_idx", (cap,), torch.int64, dev)
    val = WS.get("pk_val", (cap,), torch.float32, dev)
    L.check(L.lib().mfish_nms_compact_f32(cube.data_ptr(), ns, nz, nx, ny, thr, idx.data_ptr(), val.data_ptr(),
                                          count.data_ptr(), cap, _stream()))
    return idx, val, count


def detect_peaks_async(cube: torch.Tensor, threshold: float, capacity: int | None = None):
    """Launch peak_local_max on the cube; ``finish()`` of the returned handle waits for the count
    (the only host synchronisation) and returns (raster_idx int64[n], value float32[n]) on the
    device, unordered.  Raster index is over the reference's (x, y, z, s) cube."""
    if cube.dim() == 3:
        cube = cube.unsqueeze(0)
    if float(threshold) < 0:
        raise ValueError("negative thresholds are not supported (reference zero-pads the scale axis)")
    thr = max(float(threshold), 0.0)  # threshold_rel = 0.0  =>  max(t, 0 * cube.max()), cube.max() >= 0
    cap = int(capacity or max(_DEFAULT_CAPACITY, cube.numel() // 512))
    state = {"h": _launch_peaks(cube, thr, cap), "cap": cap}

    def finish():
        while True:
            idx, val, count = state["h"]
            n = int(count.item()) & 0xFFFFFFFF
            if n <= state["cap"]:
                return idx[:n], val[:n]
            state["cap"] = int(n * 1.25) + 1024
            state["h"] = _launch_peaks(cube, thr, state["cap"])

    return finish


def detect_peaks(cube: torch.Tensor, threshold: float, capacity: int | None = None):
    """peak_local_max on the cube: (raster_idx int64[n], value float32[n]) on the device, unordered."""
    return detect_peaks_async(cube, threshold, capacity)()


def log_peaks_async(vol: Volume, sigma: float, threshold: float, out: torch.Tensor | None = None,
                    capacity: int | None = None, z_range=None):
    """One scale of blob_log in a single library call: LoG volume + thresholded 26-neighbour NMS +
    compaction, the candidate test fused into the last LoG pass.  Returns ``(log_volume, finish)``;
    ``finish()`` waits for the count and returns (raster_idx, value) like ``detect_peaks``.
    ``z_range=(a, b)`` (z-extended slabs): only the planes [a, b) are produced and searched, as a cube of
    their own -- raster indices are relative to ``log_volume[a:b]``."""
    if float(threshold) < 0:
        raise ValueError("negative thresholds are not supported (reference zero-pads the scale axis)")
    d = vol.data
    shp, dev = d.shape, d.device
    nz, nx, ny = shp
    za, zb = (0, nz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    if d.numel() >= (1 << 32):  # candidate offsets are 32-bit: fall back to the two-call form
        log = log_volume(vol, sigma, out, z_range=z_range)
        return log, detect_peaks_async(log[za:zb], threshold, capacity)
    if out is None:
        out = torch.empty(shp, dtype=torch.float32, device=dev)
    t1 = WS.get("log1", shp, torch.float32, dev)
    t2 = WS.get("log2", shp, torch.float32, dev)
    t3 = WS.get("log3", shp, torch.float32, dev)
    thr = max(float(threshold), 0.0)
    cap = int(capacity or max(_DEFAULT_CAPACITY, d.numel() // 512))
    cand_cap = max(1 << 22, d.numel() // 64)
    cand_ws = WS.get("pk_cand", (cand_cap + 1,), torch.int32, dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    idx = WS.get("pk_idx", (cap,), torch.int64, dev)
    val = WS.get("pk_val", (cap,), torch.float32, dev)
    L.check(L.lib().mfish_log3d_peaks_range_f32(d.data_ptr(), out.data_ptr(), t1.data_ptr(), t2.data_ptr(),
                                                t3.data_ptr(), nz, nx, ny, float(sigma),
                                                vol.minmax.data_ptr() if vol.minmax is not None else None, za, zb,
                                                thr, cand_ws.data_ptr(), cand_cap, idx.data_ptr(), val.data_ptr(),
                                                count.data_ptr(), cap, _stream()))

    def finish():
        n = int(count.item()) & 0xFFFFFFFF
        if n == 0xFFFFFFFF or n > cap:  # candidate overflow / result overflow: stand-alone kernel
            return detect_peaks(out[za:zb], threshold, None if n == 0xFFFFFFFF else int(n * 1.25) + 1024)
        return idx[:n], val[:n]

    return out, finish


# Order in which peak_local_max hands the peaks to _prune_blobs (it decides which blob of an overlapping
# equal-sigma pair survives, ~0.1 % of the peaks at the dense-stress density and none at C1/C2 density --
# tools/measure_peak_order.py): "response" = skimage >= 0.18 (descending intensity, the frozen default),
# "reverse_raster" = skimage 0.14-0.17, the reference's era.  MFISH_PEAK_ORDER overrides the default.
PEAK_ORDERS = {"response": L.ORDER_RESPONSE, "reverse_raster": L.ORDER_REVERSE_RASTER}
DEFAULT_PEAK_ORDER = os.environ.get("MFISH_PEAK_ORDER", "response")


def _order_code(peak_order):
    po = DEFAULT_PEAK_ORDER if peak_order is None else peak_order
    if po not in PEAK_ORDERS:
        raise ValueError(f"peak_order must be one of {sorted(PEAK_ORDERS)}, got {po!r}")
    return PEAK_ORDERS[po]


def rank_prune_equal(idx, val, dims_sxyz, cutoff_sq: int, peak_order=None):
    """Order peaks (``peak_order``) and apply the equal-sigma prune.
    Returns (idx_ranked, val_ranked, alive uint8), all on the device."""
    ns, nx, ny, nz = dims_sxyz
    n = idx.numel()
    dev = idx.device
    idx_o = torch.empty(n, dtype=torch.int64, device=dev)
    val_o = torch.empty(n, dtype=torch.float32, device=dev)
    alive = torch.ones(n, dtype=torch.uint8, device=dev)
    if n:
        L.check(L.lib().mfish_rank_prune(idx.data_ptr(), val.data_ptr(), n, ns, nz, nx, ny, int(cutoff_sq),
                                         _order_code(peak_order), idx_o.data_ptr(), val_o.data_ptr(),
                                         alive.data_ptr(), _stream()))
    return idx_o, val_o, alive


def overlap_pairs(idx_ranked, dims_sxyz, cutoff_sq_matrix):
    """Candidate (i<j) rank pairs for the unequal-sigma prune.  Returns int32 [npairs, 2] on the host."""
    ns, nx, ny, nz = dims_sxyz
    n = idx_ranked.numel()
    dev = idx_ranked.device
    cut = L.i64array(np.asarray(cutoff_sq_matrix, dtype=np.int64).reshape(-1).tolist())
    cap = max(1 << 16, 4 * n)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    while True:
        pairs = torch.empty((cap, 2), dtype=torch.int32, device=dev)
        L.check(L.lib().mfish_overlap_pairs(idx_ranked.data_ptr(), n, ns, nz, nx, ny, cut, pairs.data_ptr(),
                                            count.data_ptr(), cap, _stream()))
        m = int(count.item()) & 0xFFFFFFFF
        if m <= cap:
            return pairs[:m].cpu().numpy()
        cap = int(m * 1.25) + 1024


def decode_raster(idx: np.ndarray, dims_sxyz):
    """raster index over (x, y, z, s) -> integer arrays x, y, z, s."""
    ns, nx, ny, nz = dims_sxyz
    idx = np.asarray(idx, dtype=np.int64)
    s = idx % ns
    t = idx // ns
    z = t % nz
    t = t // nz
    y = t % ny
    x = t // ny
    return x, y, z, s


def prune_unequal_host(n, s_idx, sigmas, pairs):
    """skimage _prune_blobs over candidate pairs in sorted (i, j) order; every candidate pair
    already overlaps by more than 0.5 (cut-off table), so only the sigma comparison remains."""
    sig = np.asarray(sigmas, dtype=np.float64)[s_idx].copy()
    if len(pairs):
        order = np.lexsort((pairs[:, 1], pairs[:, 0]))
        for i, j in pairs[order]:
            si, sj = sig[i], sig[j]
            if si == 0.0 and sj == 0.0:
                continue  # overlap of two dead blobs is 0
            # a dead blob (sigma 0) still "overlaps" a live one only if it lies inside it; the
            # reference then zeroes the smaller sigma, i.e. the dead one again: no effect
            if si == 0.0 or sj == 0.0:
                continue
            if si > sj:
                sig[j] = 0.0
            else:
                sig[i] = 0.0
    return sig


def blob_log(vol: Volume, min_sigma, max_sigma, num_sigma, threshold, overlap=0.5, cube=None, before_sync=None,
             peak_order=None):
    """skimage.feature.blob_log on the device.  Returns host float64 array (N, 4) rows
    [x, y, z, sigma] in peak_local_max order (descending response).  ``before_sync`` (optional
    callable) is run after the filters and the peak detection have been queued and before the host
    waits for the peak count -- callers queue independent device work there."""
    sigmas = T.sigma_list(min_sigma, max_sigma, num_sigma)
    ns = len(sigmas)
    nz, nx, ny = vol.data.shape
    dims = (ns, nx, ny, nz)
    if cube is None and ns == 1 and T.gaussian_radius(float(sigmas[0])) <= 16 and min(nz, nx) > 16:
        _, peaks = log_peaks_async(vol, float(sigmas[0]), threshold)  # fused LoG + NMS
    else:
        if cube is None:
            cube = log_cube(vol, sigmas)
        peaks = detect_peaks_async(cube, threshold)
    if before_sync is not None:
        before_sync()
    idx, val = peaks()
    if idx.numel() == 0:
        return np.empty((0, 3))
    if ns == 1:
        cutoff = T.prune_cutoff_sq(float(sigmas[0]), None, overlap)
        idx_r, _, alive = rank_prune_equal(idx, val, dims, max(cutoff, 0), peak_order)
        keep = alive.bool()
        idx_h = idx_r[keep].cpu().numpy()
        x, y, z, s = decode_raster(idx_h, dims)
        return np.stack([x, y, z, sigmas[s]], axis=1).astype(np.float64)
    # multi-scale: rank, candidate pairs on the device, sequential sigma rule on the host
    idx_r, _, _ = rank_prune_equal(idx, val, dims, 0, peak_order)
    cut = np.array([[T.prune_cutoff_sq(float(a), float(b), overlap) for b in sigmas] for a in sigmas], dtype=np.int64)
    pairs = overlap_pairs(idx_r, dims, cut)
    idx_h = idx_r.cpu().numpy()
    x, y, z, s = decode_raster(idx_h, dims)
    sig = prune_unequal_host(len(idx_h), s, sigmas, pairs)
    keep = sig > 0
    return np.stack([x[keep], y[keep], z[keep], sig[keep]], axis=1).astype(np.float64)


# ------------------------------------------------------------------ lookups
def _points_zxy(spots_xyz, device):
    p = np.asarray(spots_xyz)[:, :3].astype(np.int64)
    zxy = np.ascontiguousarray(np.stack([p[:, 2], p[:, 0], p[:, 1]], axis=1).astype(np.int32))
    return torch.from_numpy(zxy).to(device)


def gather(vol: torch.Tensor, spots_xyz) -> np.ndarray:
    """vol[x, y, z] at integer spot positions (rows [x, y, z, ...]); float32 or uint8 volume."""
    n = len(spots_xyz)
    if n == 0:
        return np.empty((0,), dtype=np.float32 if vol.dtype == torch.float32 else np.uint8)
    nz, nx, ny = vol.shape
    pts = _points_zxy(spots_xyz, vol.device)
    out = torch.empty(n, dtype=vol.dtype, device=vol.device)
    fn = L.lib().mfish_gather_f32 if vol.dtype == torch.float32 else L.lib().mfish_gather_u8
    L.check(fn(vol.data_ptr(), nz, nx, ny, pts.data_ptr(), n, out.data_ptr(), _stream()))
    return out.cpu().numpy()


def gather_points(vol: torch.Tensor, pts_zxy: torch.Tensor) -> torch.Tensor:
    """vol[z, x, y] at device points (int32 [n, 3] rows (z, x, y)); result stays on the device."""
    n = pts_zxy.shape[0]
    out = torch.empty(n, dtype=vol.dtype, device=vol.device)
    if n:
        nz, nx, ny = vol.shape
        fn = L.lib().mfish_gather_f32 if vol.dtype == torch.float32 else L.lib().mfish_gather_u8
        L.check(fn(vol.data_ptr(), nz, nx, ny, pts_zxy.data_ptr(), n, out.data_ptr(), _stream()))
    return out


def blur_at_points_dev(vol: Volume, pts_zxy: torch.Tensor, sigma_xyz=(3, 3, 0.3), mode="nearest") -> torch.Tensor:
    """``blur_at_points`` for device points (int32 [n, 3] rows (z, x, y)); float64 result on the device."""
    n = pts_zxy.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=vol.data.device)
    if n:
        nz, nx, ny = vol.data.shape
        sx, sy, sz = [float(s) for s in sigma_xyz]
        tz, rz = T.gaussian_taps(sz)
        tx, rx = T.gaussian_taps(sx)
        ty, ry = T.gaussian_taps(sy)
        b = {"nearest": L.BOUNDARY_NEAREST, "reflect": L.BOUNDARY_REFLECT}[mode]
        L.check(L.lib().mfish_blur_at_points(vol.data.data_ptr(), nz, nx, ny, pts_zxy.data_ptr(), n, L.darray(tz), rz,
                                             L.darray(tx), rx, L.darray(ty), ry, b,
                                             vol.minmax.data_ptr() if vol.minmax is not None else None,
                                             out.data_ptr(), _stream()))
    return out


def blur_at_points(vol: Volume, spots_xyz, sigma_xyz=(3, 3, 0.3), mode="nearest") -> np.ndarray:
    """skimage.filters.gaussian(normalize(img), sigma)[spot] in float64 at the given spots only."""
    n = len(spots_xyz)
    if n == 0:
        return np.empty((0,), dtype=np.float64)
    nz, nx, ny = vol.data.shape
    sx, sy, sz = [float(s) for s in sigma_xyz]
    tz, rz = T.gaussian_taps(sz)
    tx, rx = T.gaussian_taps(sx)
    ty, ry = T.gaussian_taps(sy)
    pts = _points_zxy(spots_xyz, vol.data.device)
    out = torch.empty(n, dtype=torch.float64, device=vol.data.device)
    b = {"nearest": L.BOUNDARY_NEAREST, "reflect": L.BOUNDARY_REFLECT}[mode]
    L.check(L.lib().mfish_blur_at_points(vol.data.data_ptr(), nz, nx, ny, pts.data_ptr(), n, L.darray(tz), rz,
                                         L.darray(tx), rx, L.darray(ty), ry, b,
                                         vol.minmax.data_ptr() if vol.minmax is not None else None,
                                         out.data_ptr(), _stream()))
    return out.cpu().numpy()


def masked_percentile_async(vol: Volume, mask: torch.Tensor, q_percent: float, minmax_out: torch.Tensor = None):
    """Launch the masked radix select; the returned ``finish()`` reads the two order statistics and
    completes np.percentile(normalize(img)[mask], q) in float64 on the host.  With ``minmax_out``
    (float32[2], = ``vol.minmax``) the select's one full pass also reduces the global min/max."""
    dev = vol.data.device
    ws = WS.get("qws", (L.QUANTILE_WS_BYTES,), torch.uint8, dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    info = torch.empty(3, dtype=torch.float64, device=dev)
    q = float(q_percent) / 100.0
    if minmax_out is not None:
        L.check(L.lib().mfish_masked_quantile_minmax_f32(vol.data.data_ptr(), mask.data_ptr(), vol.data.numel(), q,
                                                         ws.data_ptr(), out.data_ptr(), info.data_ptr(),
                                                         minmax_out.data_ptr(), _stream()))
    else:
        L.check(L.lib().mfish_masked_quantile_fast_f32(vol.data.data_ptr(), mask.data_ptr(), vol.data.numel(), q,
                                                       ws.data_ptr(), out.data_ptr(), info.data_ptr(), _stream()))

    def finish():
        gamma, n, status = info.tolist()
        if status != 0:  # the sample bracket missed (verified on the device): exact three-pass select
            L.check(L.lib().mfish_masked_quantile_f32(vol.data.data_ptr(), mask.data_ptr(), vol.data.numel(), q,
                                                      ws.data_ptr(), out.data_ptr(), info.data_ptr(), _stream()))
            gamma, n = info[:2].tolist()
        v0, v1 = out.tolist()
        if n == 0:
            raise IndexError("percentile of an empty selection (mask selects no voxel)")
        if vol.minmax is not None:
            lo, hi = vol.minmax_host()
            a, b = T.interp01(np.array([v0, v1], dtype=np.float64), lo, hi)
        else:
            a, b = float(v0), float(v1)
        return T.lerp(float(a), float(b), float(gamma))

    return finish


def masked_percentile_minmax_async(img_zxy: torch.Tensor, mask: torch.Tensor, q_percent: float):
    """``(Volume(img, minmax), finish)``: su.normalize_image's bounds and the masked percentile from a
    single read of the voxels (the percentile is selected on raw values and mapped afterwards)."""
    mm = torch.empty(2, dtype=torch.float32, device=img_zxy.device)
    vol = Volume(img_zxy, mm)
    return vol, masked_percentile_async(vol, mask, q_percent, minmax_out=mm)


def masked_percentile(vol: Volume, mask: torch.Tensor, q_percent: float) -> float:
    """np.percentile(normalize(img)[mask], q): exact order statistics selected on the device."""
    return masked_percentile_async(vol, mask, q_percent)()


def masked_quantile_distributed(vol: torch.Tensor, mask: torch.Tensor, q: float, reduce_hist, minmax_out=None):
    """Masked select of the two order statistics around quantile ``q`` over a volume that is split across
    ranks; ``reduce_hist(tensor)`` sums an integer tensor over the ranks in place.  ONE pass over the slab:
    a 1/64 sample (its 2x2048-bin histograms all-reduced) brackets the wanted ranks, the full pass counts the
    masked voxels below the bracket and collects those inside it, and the exact ranks are selected among the
    candidates of all ranks (their histograms all-reduced).  A missed bracket (checked identically on every
    rank) falls back to the three-pass radix select with 64-bit bins.  ``minmax_out`` (float32[2] on the device):
    the slab's own {min, max} over ALL voxels, reduced by the same pass.
    Returns (value[k], value[k+1], gamma, n_masked) of the RAW values."""
    dev = vol.device
    lib = L.lib()
    st = _stream()
    ws = WS.get("qws", (L.QUANTILE_WS_BYTES,), torch.uint8, dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    pair = torch.empty(2, dtype=torch.float32, device=dev)
    info = torch.zeros(3, dtype=torch.float64, device=dev)
    scratch = torch.empty(3, dtype=torch.float64, device=dev)
    counts = torch.empty(4, dtype=torch.int64, device=dev)
    hist = ws[:2 * 2048 * 4].view(torch.int32)
    n = vol.numel()
    L.check(lib.mfish_qselb_begin(ws.data_ptr(), float(q), st))
    for p in range(3):
        L.check(lib.mfish_qselb_sample_hist(vol.data_ptr(), mask.data_ptr(), n, p, ws.data_ptr(), st))
        reduce_hist(hist)
        L.check(lib.mfish_qsel_scan(p, ws.data_ptr(), pair.data_ptr(), scratch.data_ptr(), st))
    cap = (max(1 << 20, n // 64) + 63) // 64 * 64
    cand = WS.get("qcand", (cap,), torch.float32, dev)
    L.check(lib.mfish_qselb_collect(vol.data_ptr(), mask.data_ptr(), n, ws.data_ptr(), pair.data_ptr(),
                                    cand.data_ptr(), cap, counts.data_ptr(),
                                    minmax_out.data_ptr() if minmax_out is not None else None, st))
    reduce_hist(counts)
    L.check(lib.mfish_qselb_ranks(ws.data_ptr(), float(q), counts.data_ptr(), out.data_ptr(), info.data_ptr(), st))
    for p in range(3):
        L.check(lib.mfish_qselb_cand_hist(cand.data_ptr(), cap, p, ws.data_ptr(), st))
        reduce_hist(hist)
        L.check(lib.mfish_qsel_scan(p, ws.data_ptr(), out.data_ptr(), info.data_ptr(), st))
    gamma, nm, status = info.tolist()
    if status != 0.0:
        return masked_quantile_distributed_3pass(vol, mask, q, reduce_hist)
    v0, v1 = out.tolist()
    return v0, v1, gamma, nm


def masked_quantile_distributed_3pass(vol: torch.Tensor, mask: torch.Tensor, q: float, reduce_hist):
    """The three-pass radix select over a split volume: the 2x2048 32-bit bins of every rank are widened to 64
    bits before ``reduce_hist`` sums them (a mosaic of 2^32 voxels or more would wrap a 32-bit bin)."""
    dev = vol.device
    lib = L.lib()
    st = _stream()
    ws = WS.get("qws", (L.QUANTILE_WS_BYTES,), torch.uint8, dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    info = torch.empty(2, dtype=torch.float64, device=dev)
    hist = ws[:2 * 2048 * 4].view(torch.int32)
    L.check(lib.mfish_qsel_begin(ws.data_ptr(), float(q), st))
    for p in range(3):
        L.check(lib.mfish_qsel_hist(vol.data_ptr(), mask.data_ptr(), vol.numel(), p, ws.data_ptr(), st))
        h64 = hist.to(torch.int64) & 0xFFFFFFFF
        reduce_hist(h64)
        L.check(lib.mfish_qsel_scan_u64(p, ws.data_ptr(), h64.data_ptr(), out.data_ptr(), info.data_ptr(), st))
    v0, v1 = out.tolist()
    gamma, n = info.tolist()
    return v0, v1, gamma, n


def plane_means(vol: torch.Tensor, mask: torch.Tensor | None):
    """Per-z-plane (masked) sum and count, float64 on the host."""
    nz, nx, ny = vol.shape
    sums = torch.empty(nz, dtype=torch.float64, device=vol.device)
    cnts = torch.empty(nz, dtype=torch.float64, device=vol.device)
    L.check(L.lib().mfish_plane_sums_f32(vol.data_ptr(), mask.data_ptr() if mask is not None else None, nz, nx, ny,
                                         sums.data_ptr(), cnts.data_ptr(), _stream()))
    return sums.cpu().numpy(), cnts.cpu().numpy()


def scale_planes(vol: torch.Tensor, factors) -> torch.Tensor:
    nz, nx, ny = vol.shape
    f = torch.as_tensor(np.asarray(factors, dtype=np.float32), device=vol.device)
    out = torch.empty_like(vol)
    L.check(L.lib().mfish_scale_planes_f32(vol.data_ptr(), out.data_ptr(), nz, nx, ny, f.data_ptr(), _stream()))
    return out


# ---------------------------------------------------------------------- EDT
def edt(mask_zxy: torch.Tensor, sampling_xyz=None, want_sq=False):
    """Exact distance to the nearest ZERO voxel of ``mask_zxy`` (uint8 [nz,nx,ny]).
    Returns float32 distances [nz,nx,ny] (and int32 squared voxel distances when ``want_sq``)."""
    nz, nx, ny = mask_zxy.shape
    dev = mask_zxy.device
    if sampling_xyz is None:
        sampling_xyz = (1.0, 1.0, 1.0)
    sx, sy, sz = [float(s) for s in sampling_xyz]
    ws_bytes = int(L.lib().mfish_edt_workspace_bytes(nz, nx, ny))
    ws = WS.get("edt_ws", (ws_bytes,), torch.uint8, dev)
    dist = torch.empty((nz, nx, ny), dtype=torch.float32, device=dev)
    sq = torch.empty((nz, nx, ny), dtype=torch.int32, device=dev) if want_sq else None
    L.check(L.lib().mfish_edt3d_u8(mask_zxy.data_ptr(), nz, nx, ny, L.darray([sz, sx, sy]), dist.data_ptr(),
                                   sq.data_ptr() if sq is not None else None, ws.data_ptr(), _stream()))
    return (dist, sq) if want_sq else dist


def edt_inplane(mask_zxy: torch.Tensor, sampling_xyz=None, out: torch.Tensor | None = None):
    """y and x passes only: squared in-plane distance [nz,nx,ny] (int32 or float32) + its kind."""
    import ctypes
    nz, nx, ny = mask_zxy.shape
    dev = mask_zxy.device
    sx, sy, sz = [float(s) for s in (sampling_xyz or (1.0, 1.0, 1.0))]
    ws = WS.get("edt_ws", (int(L.lib().mfish_edt_workspace_bytes(nz, nx, ny)),), torch.uint8, dev)
    f2 = out if out is not None else torch.empty((nz, nx, ny), dtype=torch.int32, device=dev)
    kind = ctypes.c_int(-1)
    L.check(L.lib().mfish_edt_inplane_u8(mask_zxy.data_ptr(), nz, nx, ny, L.darray([sz, sx, sy]), f2.data_ptr(),
                                         ctypes.byref(kind), ws.data_ptr(), _stream()))
    if kind.value == L.EDT_F2_F32:
        f2 = f2.view(torch.float32)
    return f2, kind.value


def edt_rowscan(mask_zxy: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """pass 1 alone: uint16 distance along y to the nearest zero voxel of each row (0xFFFF = none)"""
    nz, nx, ny = mask_zxy.shape
    dy = out if out is not None else torch.empty((nz, nx, ny), dtype=torch.uint16, device=mask_zxy.device)
    L.check(L.lib().mfish_edt_rowscan_u8(mask_zxy.data_ptr(), nz * nx, ny, dy.data_ptr(), _stream()))
    return dy


def edt_xpass(dy: torch.Tensor, sampling_xyz=None, ny_full=None, out: torch.Tensor | None = None):
    """pass 2 alone on row distances [nz,nx,ny] (possibly a y-slab of rows ``ny_full`` long): squared in-plane
    distance (int32 or float32) + its kind"""
    import ctypes
    nz, nx, ny = dy.shape
    sx, sy, sz = [float(s) for s in (sampling_xyz or (1.0, 1.0, 1.0))]
    link = WS.get("edt_link", (nz * nx * ny,), torch.int32, dy.device)
    f2 = out if out is not None else torch.empty((nz, nx, ny), dtype=torch.int32, device=dy.device)
    kind = ctypes.c_int(-1)
    L.check(L.lib().mfish_edt_xpass_u16(dy.data_ptr(), nz, nx, ny, int(ny_full or ny), L.darray([sz, sx, sy]),
                                        f2.data_ptr(), ctypes.byref(kind), link.data_ptr(), _stream()))
    if kind.value == L.EDT_F2_F32:
        f2 = f2.view(torch.float32)
    return f2, kind.value


def edt_zpass(f2: torch.Tensor, kind: int, sampling_xyz=None, out: torch.Tensor | None = None, plane_xy=None):
    """z pass on (possibly re-partitioned) in-plane squared distances -> float32 distances.
    ``plane_xy=(nx, ny)``: extents of the plane the in-plane pass ran over when ``f2`` is an x-slab of
    a wider mosaic (bounds the integer payloads)."""
    nz, nx, ny = f2.shape
    dev = f2.device
    sx, sy, sz = [float(s) for s in (sampling_xyz or (1.0, 1.0, 1.0))]
    link = WS.get("edt_link", (nz * nx * ny,), torch.int32, dev)
    if out is None:
        out = torch.empty((nz, nx, ny), dtype=torch.float32, device=dev)
    f2 = f2.contiguous()
    bound = 0 if plane_xy is None else (int(plane_xy[0]) - 1) ** 2 + (int(plane_xy[1]) - 1) ** 2
    L.check(L.lib().mfish_edt_zpass(f2.data_ptr(), int(kind), nz, nx, ny, L.darray([sz, sx, sy]), bound,
                                    out.data_ptr(), None, link.data_ptr(), _stream()))
    return out


# ------------------------------------------------------ segmentation thresholds
def histogram(vol: torch.Tensor, nbins=256, value_range=None):
    """np.histogram(vol, bins=nbins, range=value_range or (min, max)) -> (hist int64, bin_edges float64), host"""
    if value_range is None:
        lo, hi = [float(v) for v in minmax(vol).tolist()]
    else:
        lo, hi = float(value_range[0]), float(value_range[1])
    if lo == hi:  # numpy widens a degenerate range by half a unit each way
        lo, hi = lo - 0.5, hi + 0.5
    edges = np.linspace(lo, hi, nbins + 1, endpoint=True)
    e_d = torch.from_numpy(edges).to(vol.device)
    h_d = torch.empty(nbins, dtype=torch.int64, device=vol.device)
    L.check(L.lib().mfish_histogram_f32(vol.data_ptr(), vol.numel(), lo, hi, int(nbins), e_d.data_ptr(),
                                        h_d.data_ptr(), _stream()))
    return h_d.cpu().numpy(), edges


def split_moments(vol: torch.Tensor, threshold: float):
    """(count, sum) of the voxels <= threshold and (count, sum) of those above, float64 on the host"""
    out = torch.empty(4, dtype=torch.float64, device=vol.device)
    L.check(L.lib().mfish_split_moments_f32(vol.data_ptr(), vol.numel(), float(threshold), out.data_ptr(), _stream()))
    c0, s0, c1, s1 = out.tolist()
    return (c0, s0), (c1, s1)


def binarize(vol: torch.Tensor, threshold: float, minmax_dev: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 (vol > threshold); with ``minmax_dev`` the compare is on the normalised voxels"""
    out = torch.empty(vol.shape, dtype=torch.uint8, device=vol.device)
    L.check(L.lib().mfish_binarize_f32(vol.data_ptr(), vol.numel(), float(threshold),
                                       minmax_dev.data_ptr() if minmax_dev is not None else None, out.data_ptr(),
                                       _stream()))
    return out


# ---------------------------------------------------------- compartment masks
def binary_morph(mask_zxy: torch.Tensor, op: str, n3d: int, n2d: int, out: torch.Tensor | None = None):
    """``n3d`` 3-D + ``n2d`` per-plane cross dilations ('dilate') or skimage-style erosions ('erode') of a
    uint8 mask [nz,nx,ny] in one capped city-block distance transform.  Returns uint8 0/1."""
    nz, nx, ny = mask_zxy.shape
    dev = mask_zxy.device
    ws = WS.get("morph_ws", (int(L.lib().mfish_morph_workspace_bytes(nz, nx, ny)),), torch.uint8, dev)
    if out is None:
        out = torch.empty((nz, nx, ny), dtype=torch.uint8, device=dev)
    code = {"dilate": L.MORPH_DILATE, "erode": L.MORPH_ERODE}[op]
    L.check(L.lib().mfish_binary_morph_u8(mask_zxy.data_ptr(), nz, nx, ny, code, int(n3d), int(n2d), out.data_ptr(),
                                          ws.data_ptr(), _stream()))
    return out


def compartments(nuc_zxy: torch.Tensor, fiber_zxy: torch.Tensor, erode=(1, 10), dilate=(4, 40)):
    """(region_nuc, region_peri, region_cyt) uint8 [nz,nx,ny] of general_pipeline/fish_analysis.py:169-186."""
    nz, nx, ny = nuc_zxy.shape
    dev = nuc_zxy.device
    if tuple(fiber_zxy.shape) != (nz, nx, ny):
        raise ValueError("nuclear and fibre masks differ in shape")
    ws = WS.get("morph_ws", (int(L.lib().mfish_morph_workspace_bytes(nz, nx, ny)),), torch.uint8, dev)
    r_nuc, r_peri, r_cyt = (torch.empty((nz, nx, ny), dtype=torch.uint8, device=dev) for _ in range(3))
    L.check(L.lib().mfish_compartments_u8(nuc_zxy.data_ptr(), fiber_zxy.data_ptr(), nz, nx, ny, int(erode[0]),
                                          int(erode[1]), int(dilate[0]), int(dilate[1]), r_nuc.data_ptr(),
                                          r_peri.data_ptr(), r_cyt.data_ptr(), ws.data_ptr(), _stream()))
    return r_nuc, r_peri, r_cyt


# ----------------------------------------------------- device-resident pipeline
class StageTimer:
    """CUDA-event timing of the individual C-ABI calls of a step, on the launching stream."""

    def __init__(self):
        self.events = []  # (name, start, end)

    def run(self, name, fn, *a, **k):
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        out = fn(*a, **k)
        e.record()
        self.events.append((name, s, e))
        return out

    def mark(self, name):
        """host time stamp (the host blocks on every read-back, so these trace the step's critical path)"""
        import time
        self.marks = getattr(self, "marks", [])
        self.marks.append((name, time.perf_counter()))

    def totals_ms(self):
        """{stage: [ms per recorded call]} -- call after a synchronize."""
        out = {}
        for name, s, e in self.events:
            out.setdefault(name, []).append(s.elapsed_time(e))
        return out


class _NoTimer:
    def run(self, name, fn, *a, **k):
        return fn(*a, **k)

    def mark(self, name):
        pass


def _spot_list_tail(vol, fiber_zxy, peaks, baseline_f, sigma, snr, dist_ready, t, fiber_ready=None):
    """Second half of the hot path, shared by the device-resident and the host-buffer entry.  After the
    peak count is known, everything per spot -- rank / prune, decode, blurred intensity, fibre-mask and
    distance look-ups -- runs on the high-priority stream for ALL pruned peaks and comes back in two small
    device-to-host copies: first what is ready, then the one look-up whose input arrives last (the distance
    map when the transform is still running beside the list work, the fibre mask when it is still crossing
    PCIe: ``fiber_ready`` given), so that the host's share of the work overlaps that wait.  The mask / SNR
    rules are applied on the host.  ``dist_ready()`` -> float32 [nz,nx,ny] distance map, valid on the calling
    stream."""
    nz, nx, ny = vol.data.shape
    dims = (1, nx, ny, nz)
    main = torch.cuda.current_stream()
    t.mark("tail_begin")
    idx, val = peaks()
    t.mark("peak_count")
    hp = _list_stream() if os.environ.get("MFISH_LIST_PRIO", "1") != "0" else main
    hp.wait_stream(main)
    fiber_last = fiber_ready is not None
    spots, inten, dists_all, inside = np.empty((0, 4)), np.empty(0), None, None
    baseline, keep = None, np.zeros(0, bool)

    def apply_rules():
        """mask filter, baseline, SNR rule: as soon as the mask look-up is on the host"""
        nonlocal spots, inten, baseline, snr, keep
        spots, inten = spots[inside], inten[inside]
        # no spot inside the mask: the reference fails on the empty row stack before it ever takes the
        # percentile (modules/muscle_fish.py:341), so an empty mask must not raise from here
        baseline = baseline_f() if len(spots) else None
        t.mark("baseline")
        if snr is None and len(spots):
            thresh = np.percentile(inten, 90) / (baseline * np.sqrt(10))
            snr = max(thresh, np.sqrt(10))
        keep = (inten / baseline > snr) if len(spots) else np.zeros(0, bool)

    with torch.cuda.stream(hp):
        n = 0
        if idx.numel():
            cutoff = max(T.prune_cutoff_sq(float(sigma), None, 0.5), 0)
            idx_r, _, alive = t.run("rank_prune", rank_prune_equal, idx, val, dims, cutoff)
            ras = idx_r[alive.bool()]  # survivors in peak order; raster index over (x, y, z)
            n = int(ras.numel())
            t.mark("pruned")
        if n:
            z = ras % nz
            xy = ras // nz
            pts = torch.stack([z, xy // ny, xy % ny], dim=1).to(torch.int32)
            inten_d = t.run("blur_pts", blur_at_points_dev, vol, pts, (3, 3, 0.3), "nearest")
            if fiber_last:
                dist = dist_ready()
                early_d = gather_points(dist, pts)
            else:
                early_d = gather_points(fiber_zxy, pts)
            early = torch.stack([ras.to(torch.float64), inten_d, early_d.to(torch.float64)]).cpu().numpy()
            t.mark("early_readback")
            x, y, zz, _ = decode_raster(early[0].astype(np.int64), dims)
            spots = np.stack([x, y, zz, np.full(n, float(sigma))], axis=1).astype(np.float64)
            inten = early[1]
            t.mark("decoded")
            if fiber_last:
                inside = gather_points(fiber_ready(), pts).cpu().numpy() != 0
                t.mark("late_readback")
                dists_all = early[2]
                apply_rules()
            else:
                inside = early[2] != 0
                apply_rules()  # everything but the distances, while the transform is still running
                dist = dist_ready()
                dists_all = gather_points(dist, pts).cpu().numpy().astype(np.float64)
                t.mark("late_readback")
            dists = dists_all[inside][keep]
        else:
            if fiber_last:
                fiber_ready()
            dist = dist_ready()
            dists = np.empty(0)
    main.wait_stream(hp)
    t.mark("tail_end")
    return {"spots": spots[keep], "spots_masked": spots, "intensity": inten, "baseline": baseline, "snr": snr,
            "spot_dist": dists, "dist": dist}


def analyze_stack(img_zxy: torch.Tensor, fiber_zxy: torch.Tensor, notnuc_zxy: torch.Tensor, sigma=2.0,
                  t_spot=0.025, sampling_xyz=(1.0, 1.0, 1.0), snr=None, timer=None):
    """The whole hot path on device-resident inputs (general_pipeline/fish_analysis.py:199 +
    nocodazole/fish_analysis_noco.py:218-221,297-298): normalise -> LoG -> NMS -> prune -> mask ->
    P25 baseline -> blur at spots -> SNR filter, then EDT of the NOT-nucleus mask and the per-spot
    distance gather.  ``img_zxy`` float32 [nz,nx,ny]; masks uint8.  Returns a dict of host arrays."""
    t = timer or _NoTimer()
    # everything that does not depend on the spot list is queued first, so the host round trips of
    # the list handling (count, D2H, decode) overlap the remaining device work
    # normalize_image's bounds come out of the baseline percentile's pass over the voxels
    vol, baseline_f = t.run("quantile_minmax", masked_percentile_minmax_async, img_zxy, fiber_zxy, 25)
    log = WS.get("log_out", img_zxy.shape, torch.float32, img_zxy.device)
    _, peaks = t.run("log3d_nms", log_peaks_async, vol, sigma, t_spot, log)
    # the distance transform runs on a side stream, after the streaming kernels above, so that the
    # small kernels and host round trips of the spot-list handling proceed while it runs
    main = torch.cuda.current_stream()
    side = _side_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        dist = t.run("edt", edt, notnuc_zxy, sampling_xyz)

    def dist_ready():
        cur = torch.cuda.current_stream()
        cur.wait_stream(side)
        dist.record_stream(cur)
        return dist

    return _spot_list_tail(vol, fiber_zxy, peaks, baseline_f, sigma, snr, dist_ready, t)


def analyze_host(img_fish, fiber_mask, region_nuc, sigma=2.0, t_spot=0.025, sampling_xyz=(1.0, 1.0, 1.0), snr=None,
                 z_first=False, timer=None):
    """The same path from HOST arrays in the reference's (x, y, z) order (any dtype: uint16 stacks go over
    PCIe as 2 bytes per voxel), arranged so that PCIe -- the slow link -- never waits for the GPU and the
    GPU's work after the last byte is as short as possible:

      copy stream :  nucleus mask | image ............................ | fibre mask |
      side stream :               | ingest + EDT (3 passes)                         | ingest, P25 pass
      main stream :                                  ingest(+min/max), LoG + NMS | count
      list stream :                                                               | rank, prune, blur, look-ups | D2H

    ``region_nuc`` is the nucleus mask itself (the NOT is taken on ingest, nocodazole/fish_analysis_noco.py:218).
    Returns the dict of ``analyze_stack``."""
    t = timer or _NoTimer()
    fiber_dev = MASK_HANDOFF.lookup(fiber_mask, z_first)  # a mask this package returned: already on the device
    (tn, en), (ti, ei), (tf, ef) = host_to_device_many([region_nuc, img_fish, None if fiber_dev is not None else fiber_mask],
                                                       kinds=("mask", "image", "mask"), pack_masks=True)
    main = torch.cuda.current_stream()
    side = _side_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        notnuc = ingest(tn, z_first=z_first, as_mask=True, negate=True, ready=en)
        dist = t.run("edt", edt, notnuc, sampling_xyz)
        edt_done = torch.cuda.Event()
        edt_done.record(side)
    vol = ingest(ti, z_first=z_first, want_minmax=True, ready=ei)
    vol_ready = torch.cuda.Event()
    vol_ready.record(main)
    log = WS.get("log_out", vol.data.shape, torch.float32, vol.data.device)
    _, peaks = t.run("log3d_nms", log_peaks_async, vol, sigma, t_spot, log)
    # the fibre mask is the last buffer to arrive; its ingest and the P25 pass over the fibre voxels follow the
    # distance transform on the side stream, so neither the peak count's read-back (main stream) nor the list
    # handling (high-priority stream) queues behind PCIe
    fiber_arrived = torch.cuda.Event()
    with torch.cuda.stream(side):
        fiber = fiber_dev if fiber_dev is not None else ingest(tf, z_first=z_first, as_mask=True, ready=ef)
        if tuple(fiber.shape) != tuple(vol.data.shape):
            raise IndexError("mask and image shapes differ")
        fiber_arrived.record(side)
        side.wait_event(vol_ready)
        pending = masked_percentile_async(vol, fiber, 25)

    def fiber_ready():
        torch.cuda.current_stream().wait_event(fiber_arrived)
        fiber.record_stream(torch.cuda.current_stream())
        return fiber

    def baseline_f():
        torch.cuda.current_stream().wait_stream(side)
        return pending()

    def dist_ready():
        cur = torch.cuda.current_stream()
        cur.wait_event(edt_done)  # not wait_stream: the side stream goes on to wait for the fibre mask
        dist.record_stream(cur)
        return dist

    return _spot_list_tail(vol, None, peaks, baseline_f, sigma, snr, dist_ready, t, fiber_ready=fiber_ready)
