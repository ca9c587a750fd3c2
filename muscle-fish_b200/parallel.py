"""Multi-GPU drivers: one process per GPU, ``torch.distributed`` for the plumbing.

Two partitionings:

* **batch** -- independent stacks, one per rank round-robin, no data-path collective.  This is the
  reference's own model: ``general_pipeline/batch_process.py:28-52`` writes one job per image.
* **z-slab** -- one large mosaic ``[nz][nx][ny]`` split into contiguous z-slabs.
    - global min/max:           all-reduce(MIN/MAX) of two floats
    - LoG / NMS / blur:         r+1 input planes from each z-neighbour (send/recv halo exchange);
                                every rank reports only peaks whose centre lies in its own planes
    - overlap prune:            all-gather of the (small) peak lists, same prune on every rank
    - P25 baseline:             three 2x2048-bin histogram all-reduces (masked radix select)
    - EDT:                      y and x passes slab-local; the z pass needs whole z lines, so the
                                in-plane squared distances are re-partitioned from z-slabs to
                                x-slabs (all-to-all of grouped send/recv), then the z pass runs
                                on ``[nz][nx/P][ny]``

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


class SlabLayout:
    """z-slab ownership (filters, NMS) and x-slab ownership (EDT z pass) of a [nz][nx][ny] mosaic."""

    def __init__(self, nz, nx, ny, rank=None, world=None):
        self.nz, self.nx, self.ny = int(nz), int(nx), int(ny)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.zb = split_bounds(self.nz, self.world)
        self.xb = split_bounds(self.nx, self.world)

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
    ext = torch.empty((lo + nzl + hi,) + tuple(slab.shape[1:]), dtype=slab.dtype, device=slab.device)
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


def repartition_z_to_x(t: torch.Tensor, lay: SlabLayout, group=None) -> torch.Tensor:
    """[nz_local][nx][ny] (z-slab) -> [nz][nx_local][ny] (x-slab): the all-to-all before the EDT z pass."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    ny = lay.ny
    out = torch.empty((lay.nz, lay.x1 - lay.x0, ny), dtype=t.dtype, device=t.device)
    sends = [t[:, b:e, :].contiguous() for (b, e) in lay.xb]
    ops = []
    for s in range(P):
        if s == r:
            continue
        ops.append(dist.P2POp(dist.isend, sends[s], s, group))
        ops.append(dist.P2POp(dist.irecv, out[lay.zb[s][0]:lay.zb[s][1]], s, group))
    out[lay.z0:lay.z1] = sends[r]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def repartition_x_to_z(t: torch.Tensor, lay: SlabLayout, group=None) -> torch.Tensor:
    """[nz][nx_local][ny] (x-slab) -> [nz_local][nx][ny] (z-slab)."""
    P, r = lay.world, lay.rank
    if P == 1:
        return t
    out = torch.empty((lay.z1 - lay.z0, lay.nx, lay.ny), dtype=t.dtype, device=t.device)
    recvs = [torch.empty((lay.z1 - lay.z0, e - b, lay.ny), dtype=t.dtype, device=t.device) for (b, e) in lay.xb]
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

    def minmax(self, slab):
        return self.D.minmax(slab)

    def log(self, ext, mm, sigma):
        return self.D.log_volume(self.D.Volume(ext, mm), sigma)

    def peaks(self, log_ext, thr):
        """-> int64 array [n, 3] of (z, x, y) in `log_ext` coordinates and float32 values"""
        idx, val = self.D.detect_peaks(log_ext, thr)
        nz, nx, ny = log_ext.shape
        x, y, z, _ = self.D.decode_raster(idx.cpu().numpy(), (1, nx, ny, nz))
        return np.stack([z, x, y], axis=1), val.cpu().numpy()

    def rank_prune(self, zxy, val, dims_zxy, cutoff_sq):
        """global peaks -> (order, alive): rank order (descending value, raster ties) and survivors"""
        nz, nx, ny = dims_zxy
        dev = torch.device("cuda", torch.cuda.current_device())
        raster = (zxy[:, 1].astype(np.int64) * ny + zxy[:, 2]) * nz + zxy[:, 0]
        idx = torch.from_numpy(raster).to(dev)
        v = torch.from_numpy(np.ascontiguousarray(val, dtype=np.float32)).to(dev)
        idx_r, _, alive = self.D.rank_prune_equal(idx, v, (1, nx, ny, nz), cutoff_sq)
        return idx_r.cpu().numpy(), alive.cpu().numpy().astype(bool)

    def gather_u8(self, vol, zxy):
        return self.D.gather(vol, zxy[:, [1, 2, 0]])

    def gather_f32(self, vol, zxy):
        return self.D.gather(vol, zxy[:, [1, 2, 0]])

    def blur_at_points(self, ext, mm, zxy, sigma_xyz):
        return self.D.blur_at_points(self.D.Volume(ext, mm), zxy[:, [1, 2, 0]].astype(np.float64), sigma_xyz, "nearest")

    # masked radix select in three passes; `reduce` all-reduces the histogram between hist and scan
    def masked_quantile(self, vol, mask, q, reduce):
        return self.D.masked_quantile_distributed(vol, mask, q, reduce)

    def edt_inplane(self, mask, sampling_xyz):
        return self.D.edt_inplane(mask, sampling_xyz)

    def edt_zpass(self, f2, kind, sampling_xyz):
        return self.D.edt_zpass(f2, kind, sampling_xyz)


# ------------------------------------------------------------------------ slab-mode pipeline
def slab_find_spots_snrfilter(img_slab: torch.Tensor, mask_slab: torch.Tensor, lay: SlabLayout, sigma=2.0,
                              t_spot=0.025, snr=None, ops=None, group=None):
    """``find_spots_snrfilter`` (modules/muscle_fish.py:281-402) on a z-slab decomposed mosaic.
    ``img_slab`` float32 / ``mask_slab`` uint8, both ``[nz_local][nx][ny]``.  Every rank returns the
    same global result: dict(spots (N,4) float64 [x,y,z,sigma], spots_masked, intensity, baseline, snr)."""
    ops = ops or DeviceOps()
    dev = img_slab.device
    nz, nx, ny = lay.nz, lay.nx, lay.ny
    # 1. global min / max
    with _phase("minmax"):
        mm = ops.minmax(img_slab).clone()
        lo_hi = mm.to(torch.float32)
        lo_t, hi_t = lo_hi[0:1].clone(), lo_hi[1:2].clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi_t, op=dist.ReduceOp.MAX, group=group)
        mm = torch.cat([lo_t, hi_t])
        lo_v, hi_v = float(lo_t.item()), float(hi_t.item())
    # 2. halo of raw input: r planes for the z filter pass + 1 plane of LoG for the NMS
    radius = T.gaussian_radius(sigma)
    with _phase("halo"):
        ext, lo, hi = exchange_halo(img_slab, radius + 1, lay, group)
    # 3. LoG on the extended slab (planes within `radius` of an interior cut are contaminated by the
    #    boundary rule and are not used); reflect applies only at the global faces
    with _phase("log"):
        log_ext = ops.log(ext, mm, sigma)
    a = radius if lo else 0
    b = log_ext.shape[0] - (radius if hi else 0)
    with _phase("peaks"):
        zxy, val = ops.peaks(log_ext[a:b], t_spot)
    zg = zxy[:, 0] + a - lo + lay.z0
    own = (zg >= lay.z0) & (zg < lay.z1)
    mine = np.stack([zg[own], zxy[own, 1], zxy[own, 2]], axis=1).astype(np.int64)
    vals = np.asarray(val)[own].astype(np.float32)
    # 4. all-gather the peak lists, prune identically everywhere
    with _phase("gather_peaks"):
        gathered = [None] * lay.world
        dist.all_gather_object(gathered, (mine, vals), group=group)
    all_zxy = np.concatenate([g[0] for g in gathered], axis=0) if gathered else mine
    all_val = np.concatenate([g[1] for g in gathered], axis=0)
    empty = {"spots": np.empty((0, 4)), "spots_masked": np.empty((0, 4)), "intensity": np.empty(0),
             "baseline": None, "snr": snr}
    if len(all_zxy) == 0:
        return empty
    cutoff = max(T.prune_cutoff_sq(float(sigma), None, 0.5), 0)
    with _phase("prune"):
        raster_ranked, alive = ops.rank_prune(all_zxy, all_val, (nz, nx, ny), cutoff)
    raster_ranked = raster_ranked[alive]
    z = raster_ranked % nz
    t = raster_ranked // nz
    y = t % ny
    x = t // ny
    # 5. mask filter: the owner of each spot looks it up
    ownz = (z >= lay.z0) & (z < lay.z1)
    local = np.stack([z[ownz] - lay.z0, x[ownz], y[ownz]], axis=1)
    inmask = np.zeros(len(z), dtype=np.int32)
    with _phase("mask"):
        if ownz.any():
            inmask[ownz] = (ops.gather_u8(mask_slab, local) != 0).astype(np.int32)
        inmask = _allreduce_np(inmask, dev, group=group) > 0
    x, y, z = x[inmask], y[inmask], z[inmask]
    spots = np.stack([x, y, z, np.full(len(x), float(sigma))], axis=1).astype(np.float64)
    if len(spots) == 0:
        return empty
    # 6. P25 baseline over the (global) mask
    def reduce_hist(h):
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
    with _phase("quantile"):
        v0, v1, gamma, n = ops.masked_quantile(img_slab, mask_slab, 0.25, reduce_hist)
    if n == 0:
        raise IndexError("percentile of an empty selection (mask selects no voxel)")
    a01 = T.interp01(np.array([v0, v1], dtype=np.float64), lo_v, hi_v)
    baseline = T.lerp(float(a01[0]), float(a01[1]), float(gamma))
    # 7. blurred intensity at the spots: owners evaluate on their extended raw slab
    ownz = (z >= lay.z0) & (z < lay.z1)
    inten = np.zeros(len(z), dtype=np.float64)
    with _phase("blur"):
        if ownz.any():
            loc = np.stack([z[ownz] - lay.z0 + lo, x[ownz], y[ownz]], axis=1)
            inten[ownz] = ops.blur_at_points(ext, mm, loc, (3, 3, 0.3))
        inten = _allreduce_np(inten, dev, group=group)
    if snr is None:
        thresh = np.percentile(inten, 90) / (baseline * np.sqrt(10))
        snr = max(thresh, np.sqrt(10))
    keep = inten / baseline > snr
    return {"spots": spots[keep], "spots_masked": spots, "intensity": inten, "baseline": baseline, "snr": snr}


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
        dx = ops.edt_zpass(f2x, kind, sampling_xyz)
    return repartition_x_to_z(dx, lay, group) if back_to_z else dx


def slab_spot_distances(dist_xslab: torch.Tensor, spots_xyz: np.ndarray, lay: SlabLayout, ops=None, group=None):
    """EDT value at each spot (the interpn gather of nocodazole/fish_analysis_noco.py:297-298) from the
    x-slab partitioned distance map; every rank gets the full vector."""
    ops = ops or DeviceOps()
    s = np.asarray(spots_xyz)
    out = np.zeros(len(s), dtype=np.float64)
    if len(s):
        x = s[:, 0].astype(np.int64)
        own = (x >= lay.x0) & (x < lay.x1)
        if own.any():
            loc = np.stack([s[own, 2].astype(np.int64), x[own] - lay.x0, s[own, 1].astype(np.int64)], axis=1)
            out[own] = ops.gather_f32(dist_xslab, loc).astype(np.float64)
    return _allreduce_np(out, dist_xslab.device, group=group)
