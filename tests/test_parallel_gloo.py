"""CPU, world_size 2 (gloo): the multi-GPU drivers' partition / halo-exchange / all-to-all logic.

The voxel work is injected as an oracle-backed ``ops`` object (scipy float64), so what is tested
here is exactly the host-side plumbing of ``muscle_fish_b200.parallel``: the slab-decomposed run
must reproduce the oracle on the un-split volume.  (The CUDA ``DeviceOps`` is exercised under
``-m gpu``.)"""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import ndimage as ndi

import oracle
import muscle_fish_b200  # noqa: F401
from muscle_fish_b200 import parallel as P, synth, taps as T
from conftest import match_fraction

ANISO = (0.0577, 0.0577, 0.5)


class OracleOps:
    """Same interface as parallel.DeviceOps, CPU tensors, scipy arithmetic (test infrastructure)."""

    def __init__(self, sigma=2.0):
        self.sigma = sigma

    def minmax(self, slab):
        return torch.stack([slab.min(), slab.max()]).to(torch.float32)

    def log(self, ext, mm, sigma, z_range=None):
        lo, hi = [float(v) for v in mm]
        a = (ext.numpy().astype(np.float64) - lo) * (1.0 / (hi - lo))
        return torch.from_numpy((-ndi.gaussian_laplace(a, sigma) * sigma ** 2))

    def peaks(self, log_ext, thr):
        cube = log_ext.numpy()
        nzs, nx, ny = cube.shape
        pk = oracle.peak_local_max(cube[..., None], threshold_abs=thr)[:, :3]
        raster = (pk[:, 1].astype(np.int64) * ny + pk[:, 2]) * nzs + pk[:, 0]
        return torch.from_numpy(raster), torch.from_numpy(cube[tuple(pk.T)].astype(np.float32))

    def rank_prune(self, raster, val, dims_zxy, cutoff_sq):
        nz, nx, ny = dims_zxy
        raster, val = raster.numpy(), val.numpy()
        order = np.lexsort((raster, -val.astype(np.float64)))
        r = raster[order]
        z, t = r % nz, r // nz
        y, x = t % ny, t // ny
        blobs = np.stack([x, y, z, np.full(len(r), self.sigma)], 1).astype(float)
        assert T.prune_cutoff_sq(self.sigma) == cutoff_sq
        kept = oracle.prune_blobs(blobs, 0.5, pair_order="sorted")
        keyset = {tuple(k[:3].astype(int)) for k in kept}
        alive = np.array([tuple(b[:3].astype(int)) in keyset for b in blobs], dtype=bool)
        return torch.from_numpy(r), torch.from_numpy(alive)

    def gather_u8(self, vol, zxy):
        return torch.from_numpy(vol.numpy()[tuple(zxy.numpy().T)])

    gather_f32 = gather_u8

    def blur_at_points(self, ext, mm, zxy, sigma_xyz):
        lo, hi = [float(v) for v in mm]
        a = (ext.numpy().astype(np.float64) - lo) * (1.0 / (hi - lo))
        sx, sy, sz = sigma_xyz
        return torch.from_numpy(oracle.gaussian(a, (sz, sx, sy))[tuple(zxy.numpy().T)])

    def masked_quantile(self, vol, mask, q, reduce):
        """three-pass 11/11/10-bit radix select on the monotone float key, histograms reduced"""
        v = vol.numpy()[mask.numpy() != 0].astype(np.float32)
        u = v.view(np.uint32)
        key = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)
        prefix, k, shifts, bits = [0, 0], [0, 0], (21, 10, 0), (11, 11, 10)
        n = gamma = 0
        for p in range(3):
            hist = torch.zeros(2 * 2048, dtype=torch.int32)
            for j in range(2):
                sel = key if p == 0 else key[(key >> (shifts[p] + bits[p])) == prefix[j]]
                d = (sel >> shifts[p]) & ((1 << bits[p]) - 1)
                hist[j * 2048:(j + 1) * 2048] += torch.from_numpy(np.bincount(d, minlength=2048).astype(np.int32))
            reduce(hist)
            h = hist.numpy().astype(np.int64).reshape(2, 2048)
            if p == 0:
                n = int(h[0].sum())
                if n == 0:
                    return 0.0, 0.0, 0.0, 0
                vi = (n - 1) * q
                k0 = min(max(int(np.floor(vi)), 0), n - 1)
                k = [k0, min(k0 + 1, n - 1)]
                gamma = vi - np.floor(vi)
            for j in range(2):
                cum = np.cumsum(h[j])
                b = int(np.searchsorted(cum, k[j], side="right"))
                k[j] -= int(cum[b - 1]) if b > 0 else 0
                prefix[j] = b if p == 0 else ((prefix[j] << bits[p]) | b)
        out = []
        for j in range(2):
            kk = np.uint32(prefix[j])
            uu = (kk & np.uint32(0x7FFFFFFF)) if (kk & np.uint32(0x80000000)) else ~kk
            out.append(float(np.array([uu], dtype=np.uint32).view(np.float32)[0]))
        return out[0], out[1], float(gamma), n

    def log_peaks(self, ext, mm, sigma, thr):
        log = self.log(ext, mm, sigma)
        cube = log.numpy()
        nz, nx, ny = cube.shape
        pk = oracle.peak_local_max(cube[..., None], threshold_abs=thr)[:, :3]
        raster = (pk[:, 1].astype(np.int64) * ny + pk[:, 2]) * nz + pk[:, 0]
        return torch.from_numpy(raster), torch.from_numpy(cube[tuple(pk.T)].astype(np.float32))

    def edt_rowscan(self, mask):
        """distance along y to the nearest zero voxel of the row; 0x7FFF = no zero in the row (int16 on the wire)"""
        m = mask.numpy() != 0
        ny = m.shape[2]
        yy = np.arange(ny)
        d = np.where(m[:, :, None, :], 10 ** 6, np.abs(yy[:, None] - yy[None, :])[None, None]).min(axis=3)
        return torch.from_numpy(np.where(d >= 10 ** 6, 0x7FFF, d).astype(np.int16))

    def edt_xz(self, dy, sampling_xyz, ny_full):
        sx, sy, sz = sampling_xyz
        d = dy.numpy().astype(np.float64)
        f = np.where(d >= 0x7FFF, np.inf, (d * sy) ** 2)  # [nz][nx][nyl]
        nz, nx, _ = f.shape
        dx = ((np.arange(nx)[:, None] - np.arange(nx)[None, :]) * sx) ** 2
        f2 = (f[:, None, :, :] + dx[None, :, :, None]).min(axis=2)  # [z][p][y] = min_u f[z][u][y] + dx[p][u]
        dz = ((np.arange(nz)[:, None] - np.arange(nz)[None, :]) * sz) ** 2
        out = (f2[None, :, :, :] + dz[:, :, None, None]).min(axis=1)
        return torch.from_numpy(np.sqrt(out).astype(np.float32))

    def edt_inplane(self, mask, sampling_xyz):
        sx, sy, _ = sampling_xyz
        m = mask.numpy()
        f2 = np.empty(m.shape, np.float32)
        for z in range(m.shape[0]):
            if m[z].all():
                f2[z] = np.inf
            else:
                f2[z] = ndi.distance_transform_edt(m[z], sampling=(sx, sy)) ** 2
        return torch.from_numpy(f2), 1

    def edt_zpass(self, f2, kind, sampling_xyz, plane_xy=None):
        sz = sampling_xyz[2]
        f = f2.numpy().astype(np.float64)
        nz = f.shape[0]
        dz = (np.arange(nz)[:, None] - np.arange(nz)[None, :]) * sz
        cand = f[None, :, :, :] + (dz ** 2)[:, :, None, None]  # [p, q, x, y]
        return torch.from_numpy(np.sqrt(cand.min(axis=1)).astype(np.float32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, outdir, args):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        res = fn(rank, world, *args)
        np.save(os.path.join(outdir, f"r{rank}.npy"), np.array(res, dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def run_ranks(world, fn, *args):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), fn, d, args), nprocs=world, join=True)
        return [np.load(os.path.join(d, f"r{r}.npy"), allow_pickle=True) for r in range(world)]


# ---------------------------------------------------------------------------------- unit level
def test_shard_stacks_and_bounds():
    assert P.shard_stacks(64, 3, 8) == list(range(3, 64, 8))
    assert sorted(sum((P.shard_stacks(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert P.split_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    lay = P.SlabLayout(25, 13, 7, rank=1, world=2)
    assert (lay.z0, lay.z1, lay.x0, lay.x1) == (13, 25, 7, 13) and lay.min_slab() == 12


def _comm_job(rank, world, nz, nx, ny, halo):
    full = torch.arange(nz * nx * ny, dtype=torch.float32).reshape(nz, nx, ny)
    lay = P.SlabLayout(nz, nx, ny)
    slab = full[lay.z0:lay.z1].clone()
    ext, lo, hi = P.exchange_halo(slab, halo, lay)
    ok_halo = torch.equal(ext, full[lay.z0 - lo:lay.z1 + hi])
    xs = P.repartition_z_to_x(slab, lay)
    ok_x = torch.equal(xs, full[:, lay.x0:lay.x1])
    back = P.repartition_x_to_z(xs, lay)
    return [bool(ok_halo), bool(ok_x), bool(torch.equal(back, slab)), lo, hi]


@pytest.mark.parametrize("shape,halo", [((10, 8, 5), 3), ((11, 7, 4), 5)])
def test_halo_exchange_and_all_to_all_world2(shape, halo):
    res = run_ranks(2, _comm_job, *shape, halo)
    for r, (a, b, c, lo, hi) in enumerate(res):
        assert a and b and c
        assert (lo, hi) == ((0, halo), (halo, 0))[r]


def test_halo_thicker_than_slab_is_refused():
    lay = P.SlabLayout(6, 4, 4, rank=0, world=2)
    with pytest.raises(ValueError):
        P.exchange_halo(torch.zeros(3, 4, 4), 5, lay)


# ------------------------------------------------------------------------------ pipeline level
def _spots_job(rank, world, seed):
    st = synth.make_stack(shape=(72, 64, 28), n_spots=60, n_nuclei=2, seed=seed, nuc_semi=(10.0, 7.0, 4.0))
    img = torch.from_numpy(np.ascontiguousarray(st["img"].transpose(2, 0, 1)))
    fib = torch.from_numpy(np.ascontiguousarray(st["fiber_mask"].transpose(2, 0, 1)))
    notnuc = torch.from_numpy(np.ascontiguousarray((st["nuc_mask"] == 0).astype(np.uint8).transpose(2, 0, 1)))
    lay = P.SlabLayout(*img.shape)
    ops = OracleOps(2.0)
    res = P.slab_find_spots_snrfilter(img[lay.z0:lay.z1].clone(), fib[lay.z0:lay.z1].clone(), lay, sigma=2.0,
                                      t_spot=0.025, ops=ops)
    dmap = P.slab_edt(notnuc[lay.z0:lay.z1].clone(), lay, ANISO, ops=ops)
    sd = P.slab_spot_distances(dmap, res["spots"], lay, ops=ops)
    dz = P.slab_edt(notnuc[lay.z0:lay.z1].clone(), lay, ANISO, ops=ops, back_to_z=True)
    # the combined driver (CPU tensors: no side stream) returns the same thing
    both = P.slab_analyze(img[lay.z0:lay.z1].clone(), fib[lay.z0:lay.z1].clone(), notnuc[lay.z0:lay.z1].clone(), lay,
                          sigma=2.0, t_spot=0.025, sampling_xyz=ANISO, ops=ops, edt_group=dist.new_group())
    assert np.array_equal(both["spots"], res["spots"])
    np.testing.assert_allclose(both["spot_dist"], sd, rtol=1e-6)  # (the oracle ops take two scipy routes)
    assert torch.equal(both["dist_slab"], P.slab_edt_rows(notnuc[lay.z0:lay.z1].clone(), lay, ANISO, ops=ops))
    sq = P.slab_analyze(img[lay.z0:lay.z1].clone(), fib[lay.z0:lay.z1].clone(), notnuc[lay.z0:lay.z1].clone(), lay,
                        sigma=2.0, t_spot=0.025, sampling_xyz=ANISO, ops=ops, edt_wire="squares")
    assert np.array_equal(sq["spot_dist"], sd) and torch.equal(sq["dist_xslab"], dmap)
    ymap = both["dist_slab"].numpy()
    return [res["spots"], res["spots_masked"], res["intensity"], res["baseline"], res["snr"], sd, dz.numpy(),
            (lay.z0, lay.z1), ymap, (lay.y0, lay.y1)]


def test_slab_pipeline_world2_equals_unsplit_oracle():
    seed = 31
    res = run_ranks(2, _spots_job, seed)
    st = synth.make_stack(shape=(72, 64, 28), n_spots=60, n_nuclei=2, seed=seed, nuc_semi=(10.0, 7.0, 4.0))
    ref_spots, ref_data = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"],
                                                      pair_order="sorted")
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO)
    for r in range(2):
        spots, masked, inten, baseline, snr, sd, dz, (z0, z1), ymap, (y0, y1) = res[r]
        assert len(ref_spots) > 10
        assert match_fraction(spots, ref_spots) >= 0.999
        assert len(masked) == len(ref_data) - 1
        ref_int = np.array([row[4] for row in ref_data[1:]])
        np.testing.assert_allclose(np.sort(inten), np.sort(ref_int), rtol=1e-6)
        np.testing.assert_allclose(sd, oracle.edt_gather(grass, spots, ANISO), rtol=1e-6)
        np.testing.assert_allclose(dz, grass.transpose(2, 0, 1)[z0:z1], rtol=1e-6)
        np.testing.assert_allclose(ymap, grass.transpose(2, 0, 1)[:, :, y0:y1], rtol=1e-6)
    np.testing.assert_array_equal(res[0][0], res[1][0])  # every rank holds the same global result


# ------------------------------------------------------------------------------------ x-slabs
def _xcomm_job(rank, world, nz, nx, ny, halo):
    full = torch.arange(nz * nx * ny, dtype=torch.float32).reshape(nz, nx, ny)
    lay = P.XSlabLayout(nz, nx, ny)
    slab = full[:, lay.x0:lay.x1].clone()
    ext, lo, hi = P.exchange_halo_x(slab, halo, lay)
    ok_halo = torch.equal(ext, full[:, lay.x0 - lo:lay.x1 + hi])
    ys = P.repartition_x_to_y(slab.to(torch.int16), lay)
    ok_y = torch.equal(ys, full[:, :, lay.y0:lay.y1].to(torch.int16))
    return [bool(ok_halo), bool(ok_y), lo, hi]


@pytest.mark.parametrize("world,shape,halo", [(2, (5, 10, 8), 3), (3, (4, 13, 7), 2)])
def test_x_halo_exchange_and_x_to_y_all_to_all(world, shape, halo):
    res = run_ranks(world, _xcomm_job, *shape, halo)
    for r, (a, b, lo, hi) in enumerate(res):
        assert a and b
        assert (lo, hi) == (halo if r > 0 else 0, halo if r < world - 1 else 0)


def _xspots_job(rank, world, seed, shape):
    st = synth.make_stack(shape=shape, n_spots=60, n_nuclei=2, seed=seed, nuc_semi=(10.0, 7.0, 4.0))
    img = torch.from_numpy(np.ascontiguousarray(st["img"].transpose(2, 0, 1)))
    fib = torch.from_numpy(np.ascontiguousarray(st["fiber_mask"].transpose(2, 0, 1)))
    notnuc = torch.from_numpy(np.ascontiguousarray((st["nuc_mask"] == 0).astype(np.uint8).transpose(2, 0, 1)))
    lay = P.XSlabLayout(*img.shape)
    ops = OracleOps(2.0)
    xs = slice(lay.x0, lay.x1)
    res = P.slab_analyze(img[:, xs].clone(), fib[:, xs].clone(), notnuc[:, xs].clone(), lay, sigma=2.0, t_spot=0.025,
                         sampling_xyz=ANISO, ops=ops)
    return [res["spots"], res["spots_masked"], res["intensity"], res["baseline"], res["spot_dist"],
            res["dist_slab"].numpy(), (lay.y0, lay.y1)]


@pytest.mark.parametrize("world", [2, 3])
def test_xslab_pipeline_equals_unsplit_oracle(world):
    """x-slabs for the filters / NMS / row scan, y-slabs for the EDT's envelope passes: the same global result
    as the un-split oracle on every rank"""
    seed, shape = 33, (96, 64, 20)
    res = run_ranks(world, _xspots_job, seed, shape)
    st = synth.make_stack(shape=shape, n_spots=60, n_nuclei=2, seed=seed, nuc_semi=(10.0, 7.0, 4.0))
    ref_spots, ref_data = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"],
                                                      pair_order="sorted")
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO)
    assert len(ref_spots) > 10
    for r in range(world):
        spots, masked, inten, baseline, sd, dmap, (y0, y1) = res[r]
        assert match_fraction(spots, ref_spots) >= 0.999
        assert len(masked) == len(ref_data) - 1
        np.testing.assert_allclose(np.sort(inten), np.sort([row[4] for row in ref_data[1:]]), rtol=1e-6)
        np.testing.assert_allclose(sd, oracle.edt_gather(grass, spots, ANISO), rtol=1e-6)
        np.testing.assert_allclose(dmap, grass.transpose(2, 0, 1)[:, :, y0:y1], rtol=1e-6)
        np.testing.assert_array_equal(spots, res[0][0])
