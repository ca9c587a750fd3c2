"""GPU parity: ingest / normalise / separable filters / LoG against the oracle (scipy float64).

Tolerance (BASELINE.json north_star): filtered volumes within 1e-5 of the reference, relative
to the volume's max |value| (LoG has cancellation, per-voxel relative error is meaningless
near 0).  Integer/byte work (ingest, min/max, float64 normalise) is bit exact.
"""
import numpy as np
import pytest
from scipy import ndimage as ndi

import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


def _rel_err(got, ref):
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))


def _zxy_to_xyz(t):
    return t.permute(1, 2, 0).contiguous().cpu().numpy()


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int16, np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("shape", [(33, 47, 9), (64, 32, 70), (5, 130, 3)])
def test_ingest_layout_and_minmax(dtype, shape):
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(1)
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        a = rng.integers(max(info.min, -30000), min(info.max, 60000), size=shape).astype(dtype)
    else:
        a = (rng.standard_normal(shape) * 100).astype(dtype)
    vol = D.ingest(a)
    got = _zxy_to_xyz(vol.data)
    np.testing.assert_array_equal(got, a.astype(np.float32))
    lo, hi = vol.minmax_host()
    assert lo == np.float32(a.min()) and hi == np.float32(a.max())
    # z-first input: same device content
    vol2 = D.ingest(np.ascontiguousarray(a.transpose(2, 0, 1)), z_first=True)
    np.testing.assert_array_equal(vol2.data.cpu().numpy(), vol.data.cpu().numpy())
    # masks
    m = D.ingest(a, as_mask=True)
    np.testing.assert_array_equal(_zxy_to_xyz(m), (a != 0).astype(np.uint8))
    mneg = D.ingest(a, as_mask=True, negate=True)
    np.testing.assert_array_equal(_zxy_to_xyz(mneg), (a == 0).astype(np.uint8))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32, np.float64, np.int64])
def test_normalize_image_bit_exact(dtype):
    from muscle_fish_b200 import scope_utils3 as su
    rng = np.random.default_rng(2)
    a = (rng.random((40, 50, 12)) * 4000).astype(dtype)
    got = su.normalize_image(a)
    ref = oracle.normalize_image(a)
    assert got.dtype == np.float64 and got.shape == a.shape
    np.testing.assert_array_equal(got, ref)
    assert got.min() == 0.0 and got.max() == 1.0
    # other bounds: exact at the knots, within 1 ulp elsewhere (fma-free arithmetic => exact too)
    got2 = su.normalize_image(a, vmin=2, vmax=7)
    np.testing.assert_array_equal(got2, oracle.normalize_image(a, 2, 7))
    # constant image -> vmax everywhere, like np.interp
    c = np.full((4, 5, 6), 3, dtype=dtype)
    np.testing.assert_array_equal(su.normalize_image(c), oracle.normalize_image(c))


def _sepconv(vol_xyz, axis_xyz, sigma, order_op, mode, norm=False):
    """one pass through the C ABI; returns outputs in (x,y,z) order"""
    import torch
    from muscle_fish_b200 import _lib as L, device as D, taps as T
    v = D.ingest(vol_xyz.astype(np.float32), want_minmax=norm)
    nz, nx, ny = v.data.shape
    axis = {0: L.AXIS_X, 1: L.AXIS_Y, 2: L.AXIS_Z}[axis_xyz]
    g, r = T.gaussian_taps(sigma, 0)
    d, _ = T.gaussian_taps(sigma, 2)
    out0 = torch.empty_like(v.data)
    out1 = torch.empty_like(v.data)
    b = {"reflect": L.BOUNDARY_REFLECT, "nearest": L.BOUNDARY_NEAREST}[mode]
    L.check(L.lib().mfish_sepconv_f32(v.data.data_ptr(), None, out0.data_ptr(), out1.data_ptr(), nz, nx, ny, axis,
                                      L.darray(g), L.darray(d), r, b, order_op, 1.0,
                                      v.minmax.data_ptr() if norm else None, D._stream()))
    return _zxy_to_xyz(out0), _zxy_to_xyz(out1)


@pytest.mark.parametrize("sigma", [0.3, 0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.5])
@pytest.mark.parametrize("axis", [0, 1, 2])
@pytest.mark.parametrize("mode", ["reflect", "nearest"])
def test_sepconv_one_axis(sigma, axis, mode):
    from muscle_fish_b200 import _lib as L
    rng = np.random.default_rng(int(sigma * 100) + axis)
    a = rng.random((45, 70, 37)).astype(np.float32)  # ny = 70 is not a multiple of 4
    g, d = _sepconv(a, axis, sigma, L.CONV_GD, mode)
    ref_g = ndi.gaussian_filter1d(a.astype(np.float64), sigma, axis=axis, order=0, mode=mode)
    ref_d = ndi.gaussian_filter1d(a.astype(np.float64), sigma, axis=axis, order=2, mode=mode)
    assert _rel_err(g, ref_g) < REL_TOL
    assert _rel_err(d, ref_d) < REL_TOL
    g2, _ = _sepconv(a, axis, sigma, L.CONV_G, mode)
    assert _rel_err(g2, ref_g) < REL_TOL


@pytest.mark.parametrize("sigma", [0.5, 1.0, 1.5, 2.0])
@pytest.mark.parametrize("shape", [(48, 72, 37), (40, 256, 20), (33, 520, 9), (300, 8, 19)])
@pytest.mark.parametrize("mode", ["reflect", "nearest"])
def test_sepconv_tma_staged_walker(sigma, shape, mode):
    """rows that are a multiple of 4 floats take the TMA-staged z/x walker (radius <= 8): full and
    partial 256-float segments, lines shorter than one unrolled sweep, both boundary rules"""
    from muscle_fish_b200 import _lib as L
    rng = np.random.default_rng(int(sigma * 10) + shape[0])
    a = rng.random(shape).astype(np.float32)
    for axis in (0, 2):
        g, d = _sepconv(a, axis, sigma, L.CONV_GD, mode)
        ref_g = ndi.gaussian_filter1d(a.astype(np.float64), sigma, axis=axis, order=0, mode=mode)
        ref_d = ndi.gaussian_filter1d(a.astype(np.float64), sigma, axis=axis, order=2, mode=mode)
        assert _rel_err(g, ref_g) < REL_TOL, (axis,)
        assert _rel_err(d, ref_d) < REL_TOL, (axis,)
        g2, _ = _sepconv(a, axis, sigma, L.CONV_G, mode)
        assert _rel_err(g2, ref_g) < REL_TOL, (axis,)


@pytest.mark.parametrize("shape", [(48, 72, 37), (64, 256, 21), (20, 8, 150)])
@pytest.mark.parametrize("sigma", [1.0, 2.0])
def test_log_volume_tma_staged_shapes(shape, sigma):
    """LoG through the TMA-staged z and x walkers on small aligned volumes (partial segments, lines
    shorter than the stage ring)"""
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(11)
    a = rng.random(shape).astype(np.float32)
    vol = D.ingest(a)
    got = _zxy_to_xyz(D.log_volume(vol, sigma))
    ref = oracle.log_cube(oracle.normalize_image(a), [sigma])[..., 0]
    assert _rel_err(got, ref) < REL_TOL


@pytest.mark.parametrize("shape", [(3, 5, 2), (9, 4, 5), (130, 516, 20), (17, 1030, 3)])
def test_sepconv_short_and_ragged_lines(shape):
    """lines shorter than the kernel radius (reflect wraps with period 2N) and ragged rows"""
    from muscle_fish_b200 import _lib as L
    rng = np.random.default_rng(5)
    a = rng.random(shape).astype(np.float32)
    for axis in range(3):
        for mode in ("reflect", "nearest"):
            g, d = _sepconv(a, axis, 2.0, L.CONV_GD, mode)
            ref_g = ndi.gaussian_filter1d(a.astype(np.float64), 2.0, axis=axis, order=0, mode=mode)
            ref_d = ndi.gaussian_filter1d(a.astype(np.float64), 2.0, axis=axis, order=2, mode=mode)
            assert _rel_err(g, ref_g) < REL_TOL, (shape, axis, mode)
            assert _rel_err(d, ref_d) < REL_TOL, (shape, axis, mode)


def test_sepconv_normalise_on_load():
    from muscle_fish_b200 import _lib as L
    rng = np.random.default_rng(6)
    a = (rng.random((40, 64, 11)) * 3000 + 100).astype(np.float32)
    g, _ = _sepconv(a, 1, 1.0, L.CONV_G, "reflect", norm=True)
    ref = ndi.gaussian_filter1d(oracle.normalize_image(a), 1.0, axis=1, mode="reflect")
    assert _rel_err(g, ref) < REL_TOL


@pytest.mark.parametrize("sigma", [(3, 3, 0.3), (1, 2, 0.5), (10, 10, 1), (3, 3, 1), (0, 2, 0)])
def test_gaussian3d_nearest(sigma):
    """skimage.filters.gaussian call sites: modules/muscle_fish.py:166,239,357; noco:235"""
    from muscle_fish_b200 import ndimage as nd
    rng = np.random.default_rng(7)
    a = rng.random((90, 100, 21))
    got = nd.gaussian(a, sigma)
    ref = oracle.gaussian(a, sigma)
    assert got.dtype == np.float64 and got.shape == a.shape
    assert _rel_err(got, ref) < REL_TOL


@pytest.mark.parametrize("shape", [(90, 100, 21), (300, 530, 9), (47, 33, 5)])
@pytest.mark.parametrize("sigma", [(10, 10, 1), (20, 20, 2), (5, 30, 6)])
def test_segmentation_blurs_large_radius(shape, sigma):
    """threshold_fiber's (10,10,1) and threshold_nuclei's (20,20,2) (modules/muscle_fish.py:239,166): radii 40
    and 80, longer than the short axes of the small shapes (the 'nearest' rule folds everything beyond the
    ends onto the edge voxel); a radius-24 z pass exercises the strided large-radius kernel along z as well"""
    import os
    from muscle_fish_b200 import ndimage as nd
    rng = np.random.default_rng(8)
    a = rng.random(shape) + np.linspace(0, 3, shape[0])[:, None, None]
    ref = oracle.gaussian(a, sigma)
    got = nd.gaussian(a, sigma)
    assert _rel_err(got, ref) < REL_TOL
    # reflect mode through the same kernels
    assert _rel_err(nd.gaussian(a, sigma, mode="reflect"), oracle.gaussian(a, sigma, mode="reflect")) < REL_TOL


@pytest.mark.parametrize("sigma", [1.0, 1.5, 2.0, 2.5, 3.0])
def test_log_volume_matches_gaussian_laplace(sigma):
    from muscle_fish_b200 import device as D, synth
    st = synth.make_config("small")
    vol = D.ingest(st["img"])
    got = _zxy_to_xyz(D.log_volume(vol, sigma))
    ref = oracle.log_cube(oracle.normalize_image(st["img"]), [sigma])[..., 0]
    assert _rel_err(got, ref) < REL_TOL


def test_gaussian_laplace_dropin():
    from muscle_fish_b200 import ndimage as nd
    rng = np.random.default_rng(8)
    a = rng.random((50, 44, 18))
    got = nd.gaussian_laplace(a, 2.0)
    ref = ndi.gaussian_laplace(a, 2.0)
    assert _rel_err(got, ref) < REL_TOL


def test_log_delta_known_answer():
    """LoG of a unit impulse = sum over axes of (order-2 taps on that axis) x (order-0 taps on the others)"""
    from muscle_fish_b200 import device as D, taps as T
    a = np.zeros((41, 41, 41), dtype=np.float32)
    a[20, 20, 20] = 1.0
    vol = D.ingest(a, want_minmax=False)
    got = _zxy_to_xyz(D.log_volume(vol, 2.0))
    g, r = T.gaussian_taps(2.0, 0)
    d, _ = T.gaussian_taps(2.0, 2)
    G = np.concatenate([g[:0:-1], g])
    Dk = np.concatenate([d[:0:-1], d])
    k = (Dk[:, None, None] * G[None, :, None] * G[None, None, :] + G[:, None, None] * Dk[None, :, None] * G[None, None, :]
         + G[:, None, None] * G[None, :, None] * Dk[None, None, :])
    ref = np.zeros(a.shape)
    ref[20 - r:20 + r + 1, 20 - r:20 + r + 1, 20 - r:20 + r + 1] = -4.0 * k
    assert _rel_err(got, ref) < REL_TOL


def test_golden_filters(golden):
    from muscle_fish_b200 import device as D
    img = golden["img"]
    vol = D.ingest(img)
    for s, key in ((1.0, "log_s1"), (2.0, "log_s2")):
        got = _zxy_to_xyz(D.log_volume(vol, s))
        assert _rel_err(got, golden[key]) < REL_TOL
    got = _zxy_to_xyz(D.gaussian(vol, (3, 3, 0.3)))
    assert _rel_err(got, golden["blur_3_3_03"]) < REL_TOL


@pytest.mark.parametrize("sigma", [1.0, 2.0, 5.5])
def test_log_volume_plane_range_equals_full(sigma):
    """z-slab mosaics compute only the planes that have real neighbours on both sides"""
    import torch
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(9)
    a = rng.random((40, 36, 31)).astype(np.float32)
    vol = D.ingest(a)
    full = D.log_volume(vol, sigma).clone()
    for z0, z1 in ((0, 31), (8, 23), (0, 12), (20, 31), (15, 16)):
        part = D.log_volume(vol, sigma, z_range=(z0, z1))
        assert torch.equal(part[z0:z1], full[z0:z1]), (z0, z1)
