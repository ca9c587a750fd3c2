"""GPU parity: peak detection, prune, SNR filter and the drop-in spot API against the oracle.

Bar (BASELINE.json north_star): spot coordinate sets identical except for documented ties at
the threshold (>= 99.9 % exact match).  Index work (NMS on a given cube, ranking, prune,
percentile order statistics) is bit exact.
"""
import numpy as np
import pytest

import oracle
from conftest import match_fraction, spot_set

pytestmark = pytest.mark.gpu

MIN_MATCH = 0.999


def _cube_xyzs(cube_t):
    # device [ns, nz, nx, ny] -> reference (x, y, z, s)
    return cube_t.permute(2, 3, 1, 0).contiguous().cpu().numpy()


def _peaks_gpu(cube_t, thr):
    from muscle_fish_b200 import device as D
    ns, nz, nx, ny = cube_t.shape
    idx, val = D.detect_peaks(cube_t, thr)
    idx_r, val_r, alive = D.rank_prune_equal(idx, val, (ns, nx, ny, nz), 0)
    assert bool(alive.all())
    x, y, z, s = D.decode_raster(idx_r.cpu().numpy(), (ns, nx, ny, nz))
    return np.stack([x, y, z, s], axis=1), val_r.cpu().numpy()


@pytest.mark.parametrize("ns", [1, 3])
@pytest.mark.parametrize("shape", [(24, 31, 9), (64, 64, 12), (7, 132, 5)])
def test_nms_exact_on_random_cube(ns, shape):
    """exact peak set AND order on the same float32 cube, incl. borders (zero padding)"""
    import torch
    rng = np.random.default_rng(11 + ns)
    cube = rng.random(shape + (ns,)).astype(np.float32)
    ref = oracle.peak_local_max(cube, threshold_abs=0.5)
    cube_t = torch.from_numpy(np.ascontiguousarray(cube.transpose(3, 2, 0, 1))).cuda()
    got, vals = _peaks_gpu(cube_t, 0.5)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(vals, cube[tuple(ref.T)])


def test_nms_plateau_ties_are_all_kept():
    """quantised cube: plateaus produce ties; every tied voxel is a peak (documented ties)"""
    import torch
    rng = np.random.default_rng(12)
    cube = (rng.integers(0, 6, size=(40, 40, 10, 1)) / 5.0).astype(np.float32)
    ref = oracle.peak_local_max(cube, threshold_abs=0.1)
    cube_t = torch.from_numpy(np.ascontiguousarray(cube.transpose(3, 2, 0, 1))).cuda()
    got, _ = _peaks_gpu(cube_t, 0.1)
    assert len(ref) > 100
    np.testing.assert_array_equal(got, ref)


def test_nms_threshold_is_strict_and_capacity_grows():
    import torch
    from muscle_fish_b200 import device as D
    cube = np.zeros((16, 16, 4, 1), dtype=np.float32)
    cube[3, 3, 1, 0] = 0.25
    cube[9, 9, 2, 0] = 0.5
    cube_t = torch.from_numpy(np.ascontiguousarray(cube.transpose(3, 2, 0, 1))).cuda()
    got, _ = _peaks_gpu(cube_t, 0.25)  # strict '>': the 0.25 voxel is not a peak
    np.testing.assert_array_equal(got, [[9, 9, 2, 0]])
    # capacity smaller than the number of peaks: the wrapper re-runs with a larger buffer
    rng = np.random.default_rng(13)
    big = rng.random((1, 8, 64, 64)).astype(np.float32)
    idx, val = D.detect_peaks(torch.from_numpy(big).cuda(), 0.0, capacity=16)
    ref = oracle.peak_local_max(big.transpose(2, 3, 1, 0), threshold_abs=0.0)
    assert idx.numel() == len(ref) > 16
    with pytest.raises(ValueError):
        D.detect_peaks(torch.from_numpy(big).cuda(), -0.1)


@pytest.mark.parametrize("sigma,cutoff", [(1.0, 1), (2.0, 5), (3.0, 12)])
def test_prune_cutoff_known_answers(sigma, cutoff):
    from muscle_fish_b200 import taps as T
    assert T.prune_cutoff_sq(sigma) == oracle.prune_cutoff_sq(sigma)
    if sigma in (1.0, 2.0):
        assert T.prune_cutoff_sq(sigma) == cutoff


@pytest.mark.parametrize("sigma", [1.0, 2.0])
def test_equal_sigma_prune_matches_oracle(sigma):
    """dense random peaks so that many pairs (and chains) overlap"""
    import torch
    from muscle_fish_b200 import device as D, taps as T
    rng = np.random.default_rng(14)
    shape = (48, 40, 12)
    n = 2500
    flat = rng.choice(np.prod(shape), size=n, replace=False)
    vals = rng.random(n).astype(np.float32)
    vals[: n // 10] = vals[0]  # value ties -> order decided by raster index
    idx = torch.from_numpy(flat.astype(np.int64)).cuda()
    val = torch.from_numpy(vals).cuda()
    dims = (1,) + shape
    idx_r, val_r, alive = D.rank_prune_equal(idx, val, dims, T.prune_cutoff_sq(sigma))
    x, y, z, _ = D.decode_raster(idx_r.cpu().numpy(), dims)
    # oracle: same peaks in peak_local_max order
    order = np.lexsort((flat, -vals.astype(np.float64)))
    xr, yr, zr = np.unravel_index(flat[order], shape)
    np.testing.assert_array_equal(np.stack([x, y, z], 1), np.stack([xr, yr, zr], 1))
    blobs = np.stack([xr, yr, zr, np.full(n, sigma)], axis=1).astype(np.float64)
    ref_sorted = oracle.prune_blobs(blobs, 0.5, pair_order="sorted")
    keep = alive.cpu().numpy().astype(bool)
    got = np.stack([x[keep], y[keep], z[keep], np.full(keep.sum(), sigma)], axis=1)
    assert len(ref_sorted) < n  # something was pruned
    np.testing.assert_array_equal(got, ref_sorted)


@pytest.mark.parametrize("sigma", [1.0, 2.0])
def test_equal_sigma_prune_vs_reference_set_order(sigma):
    """The reference iterates the KD-tree pairs in Python-set order, which makes chains A-B-C
    order dependent; the device prune uses sorted-pair order.  At the dense-stress density of
    BASELINE config 5 (1 peak / 839 voxels) both give the same survivors."""
    import torch
    from muscle_fish_b200 import device as D, taps as T
    rng = np.random.default_rng(17)
    shape = (256, 256, 32)
    n = 2500
    flat = rng.choice(np.prod(shape), size=n, replace=False)
    vals = rng.random(n).astype(np.float32)
    dims = (1,) + shape
    idx_r, _, alive = D.rank_prune_equal(torch.from_numpy(flat.astype(np.int64)).cuda(),
                                         torch.from_numpy(vals).cuda(), dims, T.prune_cutoff_sq(sigma))
    x, y, z, _ = D.decode_raster(idx_r.cpu().numpy(), dims)
    keep = alive.cpu().numpy().astype(bool)
    got = np.stack([x[keep], y[keep], z[keep], np.full(keep.sum(), sigma)], axis=1)
    blobs = np.stack([x, y, z, np.full(n, sigma)], axis=1).astype(np.float64)
    ref_set = oracle.prune_blobs(blobs, 0.5, pair_order="set")
    assert len(ref_set) < n
    assert match_fraction(got, ref_set) >= MIN_MATCH


def test_blob_log_single_scale_vs_oracle():
    from muscle_fish_b200 import device as D, synth
    st = synth.make_config("small")
    img = oracle.normalize_image(st["img"])
    ref = oracle.blob_log(img, 2.0, 2.0, 1, 0.025, pair_order="sorted")
    got = D.blob_log(D.ingest(st["img"]), 2.0, 2.0, 1, 0.025)
    assert len(ref) > 100
    assert match_fraction(got, ref) >= MIN_MATCH
    # rows come in descending-response order; float32 vs float64 responses may swap near-equal
    # neighbours, so the order is compared through the oracle's own (float64) responses
    cube = oracle.log_cube(img, [2.0])
    resp = cube[tuple(got[:, :3].astype(int).T) + (0,)]
    assert np.all(np.diff(resp) <= 1e-5 * resp.max())


def test_blob_log_multi_scale_vs_oracle():
    """config-5 style sigma sweep (modules/scope_utils3.py:403): 80-neighbour NMS + unequal-sigma prune"""
    from muscle_fish_b200 import ndimage as nd, synth
    st = synth.make_stack(shape=(128, 96, 24), n_spots=1500, n_nuclei=2, seed=5)
    img = oracle.normalize_image(st["img"])
    ref = oracle.blob_log(img, 1.0, 3.0, 5, 0.02, pair_order="sorted")
    got = nd.blob_log(img, 1.0, 3.0, 5, 0.02)
    assert len(ref) > 300
    assert match_fraction(got, ref) >= MIN_MATCH
    ref_set = oracle.blob_log(img, 1.0, 3.0, 5, 0.02, pair_order="set")
    assert match_fraction(got, ref_set) >= MIN_MATCH


def test_blob_log_no_peaks_returns_empty_0x3():
    from muscle_fish_b200 import ndimage as nd
    got = nd.blob_log(np.zeros((20, 20, 8)), 1, 1, 1, 0.1)
    assert got.shape == (0, 3)


def test_find_spots_vs_oracle():
    """modules/muscle_fish.py:631-672 incl. the mask filter"""
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config("small")
    ref = oracle.find_spots(st["img"], sigma=1.0, t_spot=0.02, mask=st["fiber_mask"], pair_order="sorted")
    got = mf.find_spots(st["img"], sigma=1.0, t_spot=0.02, mask=st["fiber_mask"])
    assert got.dtype == np.float64 and got.shape[1] == 4
    assert match_fraction(got, ref) >= MIN_MATCH
    ref2 = oracle.find_spots(st["img"], sigma=1.0, t_spot=0.02, pair_order="sorted")
    got2 = mf.find_spots(st["img"], sigma=1.0, t_spot=0.02)
    assert match_fraction(got2, ref2) >= MIN_MATCH
    assert len(got2) >= len(got)
    # int64 0/1 masks, as the scripts build them with np.where(..., 1, 0)
    got3 = mf.find_spots(st["img"], sigma=1.0, t_spot=0.02, mask=st["fiber_mask"].astype(np.int64))
    np.testing.assert_array_equal(got3, got)


def test_find_spots_raises_valueerror_when_nothing_survives():
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config("tiny")
    empty = np.zeros(st["img"].shape, dtype=np.uint8)
    with pytest.raises(ValueError):
        mf.find_spots(st["img"], sigma=1.0, t_spot=0.02, mask=empty)
    with pytest.raises(ValueError):
        mf.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=empty)


@pytest.mark.parametrize("cfg", ["tiny", "small"])
def test_find_spots_snrfilter_vs_oracle(cfg, capsys):
    """modules/muscle_fish.py:281-402: spots, spot_data table, SNR print"""
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config(cfg)
    ref_spots, ref_data = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"],
                                                      pair_order="sorted")
    got_spots, got_data = mf.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"])
    assert "SNR threshold = " in capsys.readouterr().out
    assert got_spots.dtype == np.float64
    assert match_fraction(got_spots, ref_spots) >= MIN_MATCH
    assert got_data[0] == ref_data[0] == ['spot_id', 'x', 'y', 'z', 'intensity', 'SNR']
    if len(got_data) == len(ref_data):
        g = np.array([r[1:] for r in got_data[1:]], dtype=np.float64)
        r = np.array([r[1:] for r in ref_data[1:]], dtype=np.float64)
        # rows are in descending-response order; float32 responses may swap near-equal neighbours
        g = g[np.lexsort((g[:, 2], g[:, 1], g[:, 0]))]
        r = r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]
        np.testing.assert_array_equal(g[:, :3], r[:, :3])
        np.testing.assert_allclose(g[:, 3:], r[:, 3:], rtol=1e-6)  # float32 voxels vs float64
    # explicit snr and z_first
    zimg = np.ascontiguousarray(st["img"].transpose(2, 0, 1))
    zmask = np.ascontiguousarray(st["fiber_mask"].transpose(2, 0, 1))
    got_z, _ = mf.find_spots_snrfilter(zimg, snr=4.0, sigma=2.0, t_spot=0.025, mask=zmask, z_first=True)
    ref_z, _ = oracle.find_spots_snrfilter(zimg, snr=4.0, sigma=2.0, t_spot=0.025, mask=zmask, z_first=True,
                                           pair_order="sorted")
    assert match_fraction(got_z, ref_z) >= MIN_MATCH


def test_find_spots_snrfilter_mask_none_fails_like_reference():
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config("tiny")
    with pytest.raises((IndexError, ValueError)):
        mf.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=None)


@pytest.mark.parametrize("q", [0, 25, 50, 90, 99.9, 100])
def test_masked_percentile_exact(q):
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(15)
    a = (rng.standard_normal((61, 47, 13)) * 50 + 10).astype(np.float32)
    a[::7] = a[0, 0, 0]  # many duplicates
    mask = rng.random(a.shape) < 0.4
    vol = D.ingest(a)
    m = D.ingest(mask, as_mask=True)
    got = D.masked_percentile(vol, m, q)
    ref = np.percentile(oracle.normalize_image(a)[mask], q)
    assert got == ref
    with pytest.raises(IndexError):
        D.masked_percentile(vol, D.ingest(np.zeros(a.shape, np.uint8), as_mask=True), q)


@pytest.mark.parametrize("shape", [(200, 192, 40), (61, 47, 13), (64, 64, 9)])
@pytest.mark.parametrize("q", [25, 50, 99.9])
def test_masked_percentile_fused_minmax(shape, q):
    """one pass over the voxels: normalize_image's bounds (bit exact, over ALL voxels, masked or not)
    and the masked percentile"""
    import torch
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(19)
    a = (rng.standard_normal(shape) * 30 + 500).astype(np.float32)
    mask = rng.random(shape) < 0.3
    a[~mask] *= 3.0  # the extremes lie outside the mask
    vol = D.ingest(a)
    m = D.ingest(mask, as_mask=True)
    vol2, finish = D.masked_percentile_minmax_async(vol.data, m, q)
    got = finish()
    assert torch.equal(vol2.minmax, vol.minmax)
    assert vol2.minmax_host() == (float(a.min()), float(a.max()))
    assert got == np.percentile(oracle.normalize_image(a)[mask], q)


@pytest.mark.parametrize("kind", ["float", "uint8_ties", "two_values", "sparse_mask"])
@pytest.mark.parametrize("q", [25, 90, 0.01, 99.99])
def test_masked_percentile_bracketed_path_exact(kind, q):
    """large enough for the sampled bracket (one full pass) to engage; heavy ties and skewed data either
    stay inside the bracket or take the verified fallback -- the result is exact either way"""
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(18)
    shape = (200, 192, 40)  # 1.5 Mvoxel
    if kind == "float":
        a = (rng.standard_normal(shape) * 10 + 160).astype(np.float32)
        mask = rng.random(shape) < 0.5
    elif kind == "uint8_ties":
        a = rng.integers(0, 7, size=shape).astype(np.float32)
        mask = rng.random(shape) < 0.7
    elif kind == "two_values":
        a = np.where(rng.random(shape) < 0.25, 3.0, 900.0).astype(np.float32)
        mask = np.ones(shape, bool)
    else:
        a = (rng.random(shape) * 1e4).astype(np.float32)
        mask = rng.random(shape) < 0.004
    vol = D.ingest(a)
    got = D.masked_percentile(vol, D.ingest(mask, as_mask=True), q)
    assert got == np.percentile(oracle.normalize_image(a)[mask], q)


@pytest.mark.parametrize("kind", ["float", "uint8_ties", "two_values", "sparse_mask"])
@pytest.mark.parametrize("q", [0.25, 0.9])
@pytest.mark.parametrize("three_pass", [False, True])
def test_masked_quantile_distributed_steps(kind, q, three_pass):
    """the split-volume select (one pass per slab + all-reduced histograms / counters) on one rank: the
    reduce callback is the identity, the result must be the two order statistics numpy's 'linear' method
    interpolates between -- on the bracketed path, on its verified fallback (heavy ties overflow the
    candidate buffer), and through the 64-bit three-pass select directly"""
    from muscle_fish_b200 import device as D, taps as T
    rng = np.random.default_rng(21)
    shape = (200, 192, 40)
    if kind == "float":
        a = (rng.standard_normal(shape) * 10 + 160).astype(np.float32)
        mask = rng.random(shape) < 0.5
    elif kind == "uint8_ties":
        a = rng.integers(0, 7, size=shape).astype(np.float32)
        mask = rng.random(shape) < 0.7
    elif kind == "two_values":
        a = np.where(rng.random(shape) < 0.25, 3.0, 900.0).astype(np.float32)
        mask = np.ones(shape, bool)
    else:
        a = (rng.random(shape) * 1e4).astype(np.float32)
        mask = rng.random(shape) < 0.004
    vol = D.ingest(a, want_minmax=False)
    m = D.ingest(mask, as_mask=True)
    calls = []
    fn = D.masked_quantile_distributed_3pass if three_pass else D.masked_quantile_distributed
    v0, v1, gamma, n = fn(vol.data, m, q, lambda t: calls.append(tuple(t.shape)))
    srt = np.sort(a[mask])
    _, k, g = T.percentile_virtual_index(len(srt), q * 100)
    assert n == len(srt) and abs(gamma - g) < 1e-12
    assert v0 == srt[k] and v1 == srt[min(k + 1, len(srt) - 1)]
    assert len(calls) >= 3


@pytest.mark.parametrize("z_range", [(3, 25), (0, 20), (8, 28), (0, 28)])
def test_log_peaks_on_a_plane_range_equals_two_calls(z_range):
    """z-extended slabs: LoG + fused candidate test on the planes [a, b) only == ranged LoG, then the
    stand-alone NMS on that sub-cube (same peaks, same values; indices relative to the sub-cube)"""
    import torch
    from muscle_fish_b200 import device as D, synth
    st = synth.make_stack(shape=(96, 80, 28), n_spots=150, n_nuclei=2, seed=5)
    vol = D.ingest(st["img"])
    a, b = z_range
    log = D.log_volume(vol, 2.0, z_range=z_range)
    r0, v0 = D.detect_peaks(log[a:b].contiguous(), 0.025)
    log2, finish = D.log_peaks_async(vol, 2.0, 0.025, z_range=z_range)
    r1, v1 = finish()
    assert torch.equal(log2[a:b], log[a:b])
    o0, o1 = torch.argsort(r0), torch.argsort(r1)
    assert r0.numel() > 20
    assert torch.equal(r0[o0], r1[o1]) and torch.equal(v0[o0], v1[o1])


def test_blur_at_points_matches_dense_blur():
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(16)
    a = (rng.random((70, 66, 15)) * 1000).astype(np.float32)
    pts = np.stack([rng.integers(0, 70, 200), rng.integers(0, 66, 200), rng.integers(0, 15, 200)], 1)
    pts[:8] = [[0, 0, 0], [69, 65, 14], [0, 65, 0], [69, 0, 14], [1, 1, 1], [68, 64, 13], [0, 30, 7], [35, 0, 0]]
    ref = oracle.gaussian(oracle.normalize_image(a), (3, 3, 0.3))[tuple(pts.T)]
    got = D.blur_at_points(D.ingest(a), pts.astype(np.float64), (3, 3, 0.3))
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)


def test_fix_bleaching_vs_oracle():
    """modules/muscle_fish.py:405-512"""
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config("small")
    img = st["img"].astype(np.float64)
    decay = np.exp(-0.03 * np.arange(img.shape[2]))
    img = img * decay
    full = np.zeros(img.shape, np.uint8)
    full[:, 40:150, :] = 1
    for mask in (None, full):
        ref = oracle.fix_bleaching(img, mask=mask)
        got = mf.fix_bleaching(img, mask=mask)
        assert got.dtype == np.float64 and got.shape == img.shape
        np.testing.assert_allclose(got, ref, rtol=2e-6)
    zimg = np.ascontiguousarray(img.transpose(2, 0, 1))
    ref = oracle.fix_bleaching(zimg, second_half=False, z_first=True)
    got = mf.fix_bleaching(zimg, second_half=False, z_first=True)
    np.testing.assert_allclose(got, ref, rtol=2e-6)
    # a mask that is empty in some fitted plane gives a NaN mean: curve_fit raises in both
    with pytest.raises(ValueError):
        oracle.fix_bleaching(img, mask=st["fiber_mask"])
    with pytest.raises(ValueError):
        mf.fix_bleaching(img, mask=st["fiber_mask"])
    # brightening stack: k > 0 -> returned unchanged
    up = st["img"].astype(np.float64) * np.exp(0.02 * np.arange(img.shape[2]))
    np.testing.assert_array_equal(mf.fix_bleaching(up), up)


def test_optimize_threshold_blob_log_counts():
    """modules/scope_utils3.py:388-422: same spot count per threshold as the oracle"""
    from muscle_fish_b200 import device as D, synth, taps as T
    st = synth.make_stack(shape=(96, 80, 20), n_spots=300, n_nuclei=1, seed=21)
    img = oracle.normalize_image(st["img"])
    _, ref_counts = None, []
    vol = D.ingest(img, want_minmax=False)
    cube = D.log_cube(vol, T.sigma_list(1.0, 2.0, 2))
    for t in np.linspace(0.01, 0.08, 5):
        ref_counts.append(len(oracle.blob_log(img, 1.0, 2.0, 2, threshold=t, pair_order="sorted")))
        got = len(D.blob_log(vol, 1.0, 2.0, 2, t, cube=cube))
        assert abs(got - ref_counts[-1]) <= max(1, round(0.001 * ref_counts[-1]))


def test_optimize_threshold_blob_log_itself_vs_oracle(capsys):
    """modules/scope_utils3.py:388-422 called as the reference calls it: same chosen threshold, the
    'Optimal threshold value' print, and IndexError when the counts have no plateau inflection (which is
    what a monotonically falling count gives upstream)."""
    from muscle_fish_b200 import scope_utils3 as su, synth
    img = oracle.normalize_image(synth.make_threshold_stack())
    t_ref, counts = oracle.optimize_threshold_blob_log(img, 1.0, 5.0, 3, 0.0, 0.2, 21, pair_order="sorted")
    assert counts.tolist()[:13] == [8, 8, 6, 4, 4, 3, 3, 3, 3, 3, 3, 8, 13] and abs(t_ref - 0.02) < 1e-12
    t_got = su.optimize_threshold_blob_log(img, min_sigma=1., max_sigma=5., num_sigma=3, tmin=0., tmax=0.2, num=21)
    assert t_got == t_ref
    assert "Optimal threshold value for blob_log: " + str(t_ref) in capsys.readouterr().out
    # float32 input takes the same path
    assert su.optimize_threshold_blob_log(img.astype(np.float32), 1., 5., 3, 0., 0.2, 21) == t_ref
    # a natural stack: the count only falls with the threshold -> no inflection before the spike
    st = synth.make_stack(shape=(96, 80, 20), n_spots=300, n_nuclei=1, seed=21)
    nat = oracle.normalize_image(st["img"])
    with pytest.raises(IndexError):
        oracle.optimize_threshold_blob_log(nat, 1.0, 2.0, 2, 0.0, 0.1, 8, pair_order="sorted")
    with pytest.raises(IndexError):
        su.optimize_threshold_blob_log(nat, 1.0, 2.0, 2, 0.0, 0.1, 8)


def test_optimize_threshold_blob_log_default_sigma_ladder():
    """the signature defaults (max_sigma=50, num_sigma=10: radii up to 200 voxels, reflect folding many
    times over a short z axis) run through the large-radius filter path: counts equal the oracle's"""
    from muscle_fish_b200 import device as D, ndimage as nd, taps as T
    rng = np.random.default_rng(31)
    img = rng.random((40, 36, 12))
    img[20, 18, 6] += 6.0
    img[8, 30, 3] += 4.0
    img = oracle.normalize_image(img)
    ref = oracle.blob_log(img, 1.0, 50.0, 10, threshold=0.05, pair_order="sorted")
    got = nd.blob_log(img, 1.0, 50.0, 10, threshold=0.05)
    assert len(ref) >= 2
    assert match_fraction(got, ref) >= MIN_MATCH
    cube = D.log_cube(D.ingest(img, want_minmax=False), T.sigma_list(1.0, 50.0, 10))
    ref_cube = oracle.log_cube(img, oracle.sigma_list(1.0, 50.0, 10))
    got_cube = cube.permute(2, 3, 1, 0).cpu().numpy()
    assert np.max(np.abs(got_cube - ref_cube)) / np.max(np.abs(ref_cube)) < 1e-5


def test_blob_log_integer_images_are_scaled_like_img_as_float():
    """skimage.feature.blob_log starts with img_as_float: uint8 / 255, uint16 / 65535"""
    from muscle_fish_b200 import ndimage as nd, synth
    st = synth.make_stack(shape=(64, 56, 16), n_spots=60, n_nuclei=1, seed=33)
    u16 = np.clip(st["img"] * 20.0, 0, 65535).astype(np.uint16)
    ref = oracle.blob_log(u16.astype(np.float64) / 65535.0, 2.0, 2.0, 1, 0.005, pair_order="sorted")
    got = nd.blob_log(u16, 2.0, 2.0, 1, 0.005)
    assert len(ref) > 20 and match_fraction(got, ref) >= MIN_MATCH
    u8 = np.clip(st["img"] / 8.0, 0, 255).astype(np.uint8)
    ref8 = oracle.blob_log(u8.astype(np.float64) / 255.0, 2.0, 2.0, 1, 0.01, pair_order="sorted")
    got8 = nd.blob_log(u8, 2.0, 2.0, 1, 0.01)
    assert len(ref8) > 10 and match_fraction(got8, ref8) >= MIN_MATCH


@pytest.mark.parametrize("peak_order", ["response", "reverse_raster"])
def test_peak_order_decides_the_survivor_like_the_oracle(peak_order):
    """skimage >= 0.18 hands _prune_blobs the peaks in descending-response order, 0.14-0.17 (the
    reference's era) in reverse raster order; the lower INDEX of an overlapping equal-sigma pair is
    zeroed, so the order picks the survivor.  Both orders are implemented; each matches the oracle run
    with the same order, on a list dense enough that a quarter of the peaks is pruned."""
    import torch
    from muscle_fish_b200 import device as D, taps as T
    rng = np.random.default_rng(41)
    shape = (40, 36, 12)
    n = 1500
    flat = np.sort(rng.choice(np.prod(shape), size=n, replace=False))
    vals = rng.random(n).astype(np.float32)
    dims = (1,) + shape
    idx_r, val_r, alive = D.rank_prune_equal(torch.from_numpy(flat.astype(np.int64)).cuda(),
                                             torch.from_numpy(vals).cuda(), dims, T.prune_cutoff_sq(2.0),
                                             peak_order=peak_order)
    x, y, z, _ = D.decode_raster(idx_r.cpu().numpy(), dims)
    order = np.lexsort((flat, -vals.astype(np.float64))) if peak_order == "response" else np.arange(n)[::-1]
    xr, yr, zr = np.unravel_index(flat[order], shape)
    np.testing.assert_array_equal(np.stack([x, y, z], 1), np.stack([xr, yr, zr], 1))  # same ranking
    np.testing.assert_array_equal(val_r.cpu().numpy(), vals[order])
    blobs = np.stack([xr, yr, zr, np.full(n, 2.0)], axis=1).astype(np.float64)
    ref = oracle.prune_blobs(blobs, 0.5, pair_order="sorted")
    keep = alive.cpu().numpy().astype(bool)
    got = blobs[keep]
    assert len(ref) < 0.8 * n
    np.testing.assert_array_equal(got, ref)


def test_peak_orders_on_a_dense_stack_vs_oracle_and_each_other():
    """config-5 density: each order matches ITS oracle to 99.9 %; between the two orders 0.1-0.3 % of the
    survivors differ (none at the C1 / C2 densities, tools/measure_peak_order.py)"""
    from muscle_fish_b200 import device as D, synth
    full = np.prod((2048, 2048, 100))
    shape = (224, 224, 64)
    st = synth.make_stack(shape=shape, n_spots=int(500000 * np.prod(shape) / full), n_nuclei=1, seed=5)
    img = oracle.normalize_image(st["img"])
    vol = D.ingest(st["img"])
    res = {}
    for po in ("response", "reverse_raster"):
        ref = oracle.blob_log(img, 2.0, 2.0, 1, 0.025, pair_order="sorted", peak_order=po)
        got = D.blob_log(vol, 2.0, 2.0, 1, 0.025, peak_order=po)
        assert len(ref) > 1500
        assert match_fraction(got, ref) >= MIN_MATCH, po
        res[po] = got
    # reverse raster order of the rows themselves
    r = res["reverse_raster"]
    ras = (r[:, 0] * shape[1] + r[:, 1]) * shape[2] + r[:, 2]
    assert np.all(np.diff(ras) < 0)
    # the two orders pick different survivors of a few overlapping pairs: 5 of 1843 spots on this stack,
    # 3 of 4057 on the crop of tools/measure_peak_order.py -- the documented skimage-version divergence
    assert 0.995 <= match_fraction(res["response"], res["reverse_raster"]) < 1.0


def test_golden_spots(golden):
    from muscle_fish_b200 import muscle_fish as mf
    got = mf.find_spots(golden["img"], sigma=2.0, t_spot=0.025)
    assert match_fraction(got, golden["spots_s2"]) >= MIN_MATCH
    got_f, _ = mf.find_spots_snrfilter(golden["img"], sigma=2.0, t_spot=0.025, mask=golden["fiber_mask"])
    assert match_fraction(got_f, golden["spots_snr_s2"]) >= MIN_MATCH


@pytest.mark.parametrize("sigma,shape", [(2.0, (96, 80, 24)), (1.0, (61, 77, 19)), (3.0, (64, 64, 40))])
def test_fused_log_peaks_equals_two_call_form(sigma, shape):
    """mfish_log3d_peaks_f32 (candidates emitted by the last LoG pass) == LoG volume + stand-alone NMS"""
    import torch
    from muscle_fish_b200 import device as D, synth
    st = synth.make_stack(shape=shape, n_spots=80, n_nuclei=1, seed=23, nuc_semi=(8.0, 6.0, 3.0))
    vol = D.ingest(st["img"])
    log_a = D.log_volume(vol, sigma).clone()
    idx_a, val_a = D.detect_peaks(log_a, 0.02)
    log_b, fin = D.log_peaks_async(vol, sigma, 0.02)
    idx_b, val_b = fin()
    assert torch.equal(log_a, log_b)
    oa, ob = torch.argsort(idx_a), torch.argsort(idx_b)
    assert idx_a.numel() > 20
    assert torch.equal(idx_a[oa], idx_b[ob]) and torch.equal(val_a[oa], val_b[ob])
    # plateau / tie handling is identical too: a quantised image produces many ties
    q = np.round(st["img"] / 40.0) * 40.0
    volq = D.ingest(q.astype(np.float32))
    la = D.log_volume(volq, sigma).clone()
    ia, _ = D.detect_peaks(la, 0.0)
    _, finq = D.log_peaks_async(volq, sigma, 0.0)
    ib, _ = finq()
    assert torch.equal(torch.sort(ia).values, torch.sort(ib).values)
