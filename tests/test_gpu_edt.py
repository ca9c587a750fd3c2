"""GPU parity: exact Euclidean distance transform + spot gather against scipy / brute force.

Bar (BASELINE.json north_star): distances exact in integer squared-voxel units for isotropic
spacing; within 1e-6 relative for anisotropic spacing.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

ANISO = (0.0577, 0.0577, 0.5)


def _edt_gpu(mask_xyz, sampling=None, want_sq=False):
    from muscle_fish_b200 import device as D
    m = D.ingest(mask_xyz, as_mask=True)
    res = D.edt(m, sampling, want_sq=want_sq)
    if want_sq:
        return (res[0].permute(1, 2, 0).cpu().numpy(), res[1].permute(1, 2, 0).cpu().numpy())
    return res.permute(1, 2, 0).cpu().numpy()


def _random_mask(shape, p_zero, seed):
    rng = np.random.default_rng(seed)
    return (rng.random(shape) >= p_zero).astype(np.uint8)


@pytest.mark.parametrize("shape", [(17, 19, 5), (32, 32, 32), (40, 70, 9), (3, 200, 2), (65, 33, 1), (1, 1, 7)])
@pytest.mark.parametrize("p_zero", [0.5, 0.05, 0.002])
def test_edt_isotropic_exact_integers(shape, p_zero):
    mask = _random_mask(shape, p_zero, seed=hash((shape, p_zero)) % 1000)
    mask.flat[0] = 0  # at least one site
    dist, sq = _edt_gpu(mask, None, want_sq=True)
    ref_sq = oracle.edt_bruteforce_sq(mask)
    np.testing.assert_array_equal(sq, ref_sq.astype(np.int64))
    ref = oracle.distance_transform_edt(mask)
    np.testing.assert_array_equal(np.rint(ref ** 2).astype(np.int64), sq)
    np.testing.assert_allclose(dist, ref, rtol=1e-7)  # float32 rounding of sqrt only
    assert np.all(dist[mask == 0] == 0)


@pytest.mark.parametrize("sampling", [ANISO, (1.0, 1.0, 8.666), (0.3, 0.7, 1.9), (2.0, 2.0, 2.0), (5.0, 0.1, 0.1)])
@pytest.mark.parametrize("seed", [1, 2])
def test_edt_anisotropic_within_1e6(sampling, seed):
    mask = _random_mask((61, 45, 23), 0.01, seed)
    mask[30, 20, 11] = 0
    dist = _edt_gpu(mask, sampling)
    ref = oracle.distance_transform_edt(mask, sampling=sampling)
    np.testing.assert_allclose(dist, ref, rtol=1e-6, atol=0)
    assert np.all(dist[mask == 0] == 0)


def test_edt_single_site_analytic():
    """distance to one zero voxel = sqrt(sum((delta*spacing)^2))"""
    shape = (50, 40, 12)
    mask = np.ones(shape, np.uint8)
    mask[7, 31, 4] = 0
    dist = _edt_gpu(mask, ANISO)
    gx, gy, gz = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
    ref = np.sqrt(((gx - 7) * ANISO[0]) ** 2 + ((gy - 31) * ANISO[1]) ** 2 + ((gz - 4) * ANISO[2]) ** 2)
    np.testing.assert_allclose(dist, ref, rtol=1e-6)


def test_edt_no_site_gives_inf_and_all_zero_gives_zero():
    """scipy returns garbage for an input without zeros (undefined upstream); we return +inf"""
    dist, sq = _edt_gpu(np.ones((9, 8, 7), np.uint8), None, want_sq=True)
    assert np.all(np.isinf(dist)) and np.all(sq == np.iinfo(np.int32).max)
    dist = _edt_gpu(np.zeros((9, 8, 7), np.uint8), ANISO)
    assert np.all(dist == 0)


def test_edt_rows_without_sites_and_long_lines():
    """sites confined to one corner: most rows/columns have no site of their own"""
    shape = (300, 150, 6)
    mask = np.ones(shape, np.uint8)
    mask[:3, :2, 0] = 0
    mask[299, 149, 5] = 0
    for sampling in (None, ANISO):
        dist = _edt_gpu(mask, sampling)
        ref = oracle.distance_transform_edt(mask, sampling=sampling)
        np.testing.assert_allclose(dist, ref, rtol=1e-6)


def test_edt_nuclei_geometry_vs_scipy():
    """the call the reference makes: distance_transform_edt(~region_nuc, sampling=dims)"""
    from muscle_fish_b200 import ndimage as nd, synth
    st = synth.make_config("small")
    not_nuc = np.logical_not(st["nuc_mask"])
    ref = oracle.distance_transform_edt(not_nuc, sampling=ANISO)
    got = nd.distance_transform_edt(not_nuc, sampling=ANISO)
    assert got.dtype == np.float64 and got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-6)
    assert np.all(got[st["nuc_mask"] != 0] == 0)
    # int64 0/1 input as produced by np.where(mask, 1, 0) upstream, isotropic
    ref1 = oracle.distance_transform_edt(not_nuc.astype(np.int64))
    got1 = nd.distance_transform_edt(not_nuc.astype(np.int64))
    np.testing.assert_array_equal(np.rint(got1 ** 2), np.rint(ref1 ** 2))


def test_spot_distance_gather_equals_interpn():
    """nocodazole/fish_analysis_noco.py:219-221,297-298"""
    from muscle_fish_b200 import ndimage as nd, synth
    st = synth.make_config("small")
    rng = np.random.default_rng(3)
    X, Y, Z = st["img"].shape
    spots = np.stack([rng.integers(0, X, 300), rng.integers(0, Y, 300), rng.integers(0, Z, 300),
                      np.full(300, 1.0)], axis=1).astype(np.float64)
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO)
    ref = oracle.edt_gather(grass, spots, ANISO)
    got = nd.spot_distances(st["nuc_mask"], spots, ANISO)
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, ref, rtol=1e-6)


def test_edt_split_inplane_then_zpass_equals_fused():
    """the two halves used by the z-slab mosaic driver"""
    import torch
    from muscle_fish_b200 import _lib as L, device as D
    import ctypes as C
    mask = _random_mask((40, 50, 21), 0.01, 9)
    m = D.ingest(mask, as_mask=True)
    nz, nx, ny = m.shape
    for sampling in (None, ANISO, (0.3, 0.7, 1.9)):
        ref = D.edt(m, sampling)
        sx, sy, sz = sampling or (1.0, 1.0, 1.0)
        sp = L.darray([sz, sx, sy])
        ws = torch.empty(int(L.lib().mfish_edt_workspace_bytes(nz, nx, ny)), dtype=torch.uint8, device="cuda")
        f2 = torch.empty((nz, nx, ny), dtype=torch.int32, device="cuda")
        kind = C.c_int(-1)
        L.check(L.lib().mfish_edt_inplane_u8(m.data_ptr(), nz, nx, ny, sp, f2.data_ptr(), C.byref(kind),
                                             ws.data_ptr(), D._stream()))
        out = torch.empty((nz, nx, ny), dtype=torch.float32, device="cuda")
        link = torch.empty(nz * nx * ny, dtype=torch.int32, device="cuda")
        L.check(L.lib().mfish_edt_zpass(f2.data_ptr(), kind.value, nz, nx, ny, sp, 0, out.data_ptr(), None,
                                        link.data_ptr(), D._stream()))
        assert torch.equal(out, ref)


def test_golden_edt(golden):
    nuc = golden["nuc_mask"]
    dist, sq = _edt_gpu(np.logical_not(nuc), None, want_sq=True)
    np.testing.assert_array_equal(sq, golden["edt_iso_sq"])
    dist = _edt_gpu(np.logical_not(nuc), ANISO)
    np.testing.assert_allclose(dist, golden["edt_aniso"], rtol=1e-6)


def _inplane_sq(mask_xyz, bands):
    """squared in-plane distances (passes 1 + 2) with the x pass on one / two threads per line (MFISH_EDT_BANDS)"""
    import ctypes as C
    import os
    import torch
    from muscle_fish_b200 import _lib as L, device as D
    m = D.ingest(mask_xyz, as_mask=True)
    nz, nx, ny = m.shape
    ws = torch.empty(int(L.lib().mfish_edt_workspace_bytes(nz, nx, ny)), dtype=torch.uint8, device="cuda")
    f2 = torch.empty((nz, nx, ny), dtype=torch.int32, device="cuda")
    kind = C.c_int(-1)
    old = os.environ.get("MFISH_EDT_BANDS")
    os.environ["MFISH_EDT_BANDS"] = str(bands)
    try:
        L.check(L.lib().mfish_edt_inplane_u8(m.data_ptr(), nz, nx, ny, None, f2.data_ptr(), C.byref(kind),
                                             ws.data_ptr(), D._stream()))
        torch.cuda.synchronize()
    finally:
        if old is None:
            del os.environ["MFISH_EDT_BANDS"]
        else:
            os.environ["MFISH_EDT_BANDS"] = old
    assert kind.value == 0  # MFISH_EDT_F2_I32
    return f2.permute(1, 2, 0).cpu().numpy()


def _inplane_bruteforce(mask_xyz):
    out = np.empty(mask_xyz.shape, np.int64)
    for z in range(mask_xyz.shape[2]):
        plane = mask_xyz[:, :, z]
        out[:, :, z] = oracle.edt_bruteforce_sq(plane[:, :, None])[:, :, 0] if (plane == 0).any() else 2 ** 31 - 1
    return out


@pytest.mark.parametrize("nx", [2, 3, 4, 5, 9, 16, 33, 64, 130])
@pytest.mark.parametrize("p_zero", [0.3, 0.02, 0.001])
def test_banded_x_pass_equals_single_thread_lines_and_bruteforce(nx, p_zero):
    """two threads per x line joined by the lower common tangent of their hulls (edt.cu, BANDS = 2)"""
    mask = _random_mask((nx, 70, 3), p_zero, seed=nx * 7 + int(p_zero * 1000))
    one = _inplane_sq(mask, 1)
    two = _inplane_sq(mask, 2)
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(two, _inplane_bruteforce(mask))


def test_banded_x_pass_empty_bands_blobs_and_the_default_policy():
    rng = np.random.default_rng(5)
    # sites only in the lower / only in the upper half of x, none at all in one plane, one full plane
    mask = np.ones((90, 40, 5), np.uint8)
    mask[:40, :, 0] = rng.random((40, 40)) > 0.05
    mask[50:, :, 1] = rng.random((40, 40)) > 0.05
    mask[44:46, 10:12, 2] = 0
    mask[:, :, 4] = 0
    one = _inplane_sq(mask, 1)
    two = _inplane_sq(mask, 2)
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(two, _inplane_bruteforce(mask))
    # smooth blobs (long pop runs, sparse hulls) on lines long enough for the default to pick the banded kernel
    from muscle_fish_b200 import synth
    st = synth.make_stack(shape=(1100, 96, 4), n_spots=5, n_nuclei=14, seed=11, nuc_semi=(60.0, 20.0, 3.0))
    not_nuc = (st["nuc_mask"] == 0).astype(np.uint8)
    one = _inplane_sq(not_nuc, 1)
    two = _inplane_sq(not_nuc, 2)
    dflt = _inplane_sq(not_nuc, 0)
    np.testing.assert_array_equal(two, one)
    np.testing.assert_array_equal(dflt, one)
    # and the full transform (default policy) against scipy
    ref = oracle.distance_transform_edt(not_nuc, sampling=ANISO)
    np.testing.assert_allclose(_edt_gpu(not_nuc, ANISO), ref, rtol=1e-6)
