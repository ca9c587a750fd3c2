"""GPU: the slab-decomposed drivers with the CUDA ``DeviceOps`` (nccl).  World size 1 runs on any
box; world size 2 needs two GPUs (one process per GPU) and is skipped otherwise."""
import os
import socket
import tempfile

import numpy as np
import pytest

import oracle
from conftest import match_fraction

pytestmark = pytest.mark.gpu
ANISO = (0.0577, 0.0577, 0.5)
SHAPE = (160, 128, 40)  # x-slabs of 80 rows hold the 12-row halo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _job(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import parallel as P, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        st = synth.make_stack(shape=SHAPE, n_spots=250, n_nuclei=4, seed=77, nuc_semi=(14.0, 9.0, 5.0))
        lay = P.SlabLayout(SHAPE[2], SHAPE[0], SHAPE[1])
        z = slice(lay.z0, lay.z1)
        img = torch.from_numpy(np.ascontiguousarray(st["img"].transpose(2, 0, 1)[z])).cuda()
        fib = torch.from_numpy(np.ascontiguousarray(st["fiber_mask"].transpose(2, 0, 1)[z])).cuda()
        notnuc = torch.from_numpy(np.ascontiguousarray((st["nuc_mask"] == 0).astype(np.uint8).transpose(2, 0, 1)[z])).cuda()
        res = P.slab_find_spots_snrfilter(img, fib, lay, sigma=2.0, t_spot=0.025)
        dmap = P.slab_edt(notnuc, lay, ANISO)
        sd = P.slab_spot_distances(dmap, res["spots"], lay)
        dz = P.slab_edt(notnuc, lay, ANISO, back_to_z=True).cpu().numpy()
        # combined driver, EDT chain on a side stream with its own communicator: same results
        both = P.slab_analyze(img, fib, notnuc, lay, sigma=2.0, t_spot=0.025, sampling_xyz=ANISO,
                              edt_group=dist.new_group())
        torch.cuda.synchronize()
        assert np.array_equal(both["spots"], res["spots"]) and np.array_equal(both["spot_dist"], sd)
        zy = both["dist_slab"].cpu().numpy()  # z-slabs, uint16 rows on the wire: y-slab partitioned map
        sq = P.slab_analyze(img, fib, notnuc, lay, sigma=2.0, t_spot=0.025, sampling_xyz=ANISO, edt_wire="squares")
        assert np.array_equal(sq["spot_dist"], sd)
        # x-slab decomposition of the same stack: same global result
        xlay = P.XSlabLayout(SHAPE[2], SHAPE[0], SHAPE[1])
        xs = slice(xlay.x0, xlay.x1)
        full = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)[:, xs])).cuda()
        xres = P.slab_analyze(full(st["img"]), full(st["fiber_mask"]), full((st["nuc_mask"] == 0).astype(np.uint8)),
                              xlay, sigma=2.0, t_spot=0.025, sampling_xyz=ANISO, edt_group=dist.new_group())
        torch.cuda.synchronize()
        assert np.array_equal(xres["spots"], res["spots"]) and np.array_equal(xres["spot_dist"], sd)
        assert np.array_equal(xres["intensity"], res["intensity"]) and xres["baseline"] == res["baseline"]
        dy_ = xres["dist_slab"].cpu().numpy()
        np.save(os.path.join(outdir, f"r{rank}.npy"),
                np.array([res["spots"], res["intensity"], res["baseline"], sd, dz, (lay.z0, lay.z1), dy_,
                          (xlay.y0, xlay.y1), zy], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def _check(world):
    import torch.multiprocessing as mp
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_job, args=(world, _free_port(), d), nprocs=world, join=True)
        res = [np.load(os.path.join(d, f"r{r}.npy"), allow_pickle=True) for r in range(world)]
    st = synth.make_stack(shape=SHAPE, n_spots=250, n_nuclei=4, seed=77, nuc_semi=(14.0, 9.0, 5.0))
    ref_spots, ref_data = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"],
                                                      pair_order="sorted")
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO)
    assert len(ref_spots) > 50
    for spots, inten, baseline, sd, dz, (z0, z1), dy_, (y0, y1), zy in res:
        assert match_fraction(spots, ref_spots) >= 0.999
        np.testing.assert_allclose(np.sort(inten), np.sort([r[4] for r in ref_data[1:]]), rtol=1e-6)
        np.testing.assert_allclose(sd, oracle.edt_gather(grass, spots, ANISO), rtol=1e-6)
        np.testing.assert_allclose(dz, grass.transpose(2, 0, 1)[z0:z1], rtol=1e-6)
        np.testing.assert_allclose(dy_, grass.transpose(2, 0, 1)[:, :, y0:y1], rtol=1e-6)
        np.testing.assert_allclose(zy, grass.transpose(2, 0, 1)[:, :, y0:y1], rtol=1e-6)


def test_slab_drivers_world1():
    _check(1)


def test_slab_drivers_world2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    _check(2)


@pytest.mark.parametrize("dtype", ["uint16", "float32", "uint8"])
def test_block_copy_kernels_match_torch_slicing(dtype):
    """mfish_copy2d / mfish_copy3d (the pack-and-transfer kernel of the slab exchanges) on local buffers: every
    block shape the re-partitions and halo exchanges produce, plus an unaligned one (torch's copy takes it)."""
    import torch
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import parallel as P
    dt = getattr(torch, dtype)
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.randint(0, 200, (12, 40, 96), device="cuda", generator=g).to(dt)
    cases = [
        (slice(None), slice(None), slice(32, 64)),   # last-axis block of every row (z/x -> y re-partition)
        (slice(None), slice(8, 24), slice(None)),    # middle-axis rows (x halo)
        (slice(3, 9), slice(None), slice(None)),     # leading planes (z halo)
        (slice(2, 11), slice(8, 24), slice(16, 80)),  # all three at once
        (slice(None), slice(None), slice(1, 50)),    # unaligned start: falls back to torch
    ]
    for sl in cases:
        blk = src[sl]
        # into a contiguous destination, and into a sub-block of a larger one
        dst = torch.zeros(blk.shape, dtype=dt, device="cuda")
        P._block_copy(dst, blk)
        assert torch.equal(dst, blk)
        big = torch.zeros((16, 48, 128), dtype=dt, device="cuda")
        view = big[2:2 + blk.shape[0], 4:4 + blk.shape[1], 16:16 + blk.shape[2]]
        P._block_copy(view, blk)
        assert torch.equal(view, blk)
        big[2:2 + blk.shape[0], 4:4 + blk.shape[1], 16:16 + blk.shape[2]] = 0
        assert not big.view(torch.uint8).any(), "nothing outside the destination block was written"
