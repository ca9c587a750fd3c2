"""CPU: pin the oracle (test infrastructure) on known answers, an independent brute force,
and the committed golden vectors.  The reference ships no tests or fixtures (parity unpinned
upstream), so these are the pins this repository adds itself."""
import math

import numpy as np
import pytest
from hypothesis import given, settings, strategies as hst
from scipy import ndimage as ndi

import oracle
import muscle_fish_b200  # noqa: F401
from muscle_fish_b200 import synth, taps as T
from conftest import match_fraction


# ----------------------------------------------------------------- normalise
def test_normalize_image_known_answers():
    a = np.array([[[2, 4], [6, 10]]], dtype=np.uint16)
    out = oracle.normalize_image(a)
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, (a - 2) / 8.0)
    n1 = oracle.normalize_image(np.random.default_rng(0).random((5, 6, 7)).astype(np.float32))
    # the second normalisation inside find_spots_snrfilter is an exact identity
    np.testing.assert_array_equal(oracle.normalize_image(n1), n1)
    np.testing.assert_array_equal(oracle.normalize_image(np.full((2, 2, 2), 7.0)), np.ones((2, 2, 2)))


# ----------------------------------------------------------------------- LoG
@pytest.mark.parametrize("sigma", [1.0, 2.0, 3.0])
def test_log_of_impulse_is_tap_outer_products(sigma):
    g, r = T.gaussian_taps(sigma, 0)
    d, _ = T.gaussian_taps(sigma, 2)
    n = 2 * r + 5
    c = n // 2
    a = np.zeros((n, n, n))
    a[c, c, c] = 1.0
    got = oracle.log_cube(a, [sigma])[..., 0]
    G = np.concatenate([g[:0:-1], g])
    D = np.concatenate([d[:0:-1], d])
    k = (D[:, None, None] * G[None, :, None] * G[None, None, :] + G[:, None, None] * D[None, :, None] * G[None, None, :]
         + G[:, None, None] * G[None, :, None] * D[None, None, :])
    ref = np.zeros_like(a)
    ref[c - r:c + r + 1, c - r:c + r + 1, c - r:c + r + 1] = -sigma ** 2 * k
    np.testing.assert_allclose(got, ref, atol=1e-15)


@pytest.mark.parametrize("sigma,radius", [(0.3, 1), (1.0, 4), (2.0, 8), (3.0, 12), (10.0, 40), (20.0, 80)])
def test_taps_equal_scipy_kernels(sigma, radius):
    from scipy.ndimage._filters import _gaussian_kernel1d
    for order in (0, 2):
        half, r = T.gaussian_taps(sigma, order)
        assert r == radius
        full = _gaussian_kernel1d(sigma, order, radius)[::-1]
        np.testing.assert_allclose(np.concatenate([half[:0:-1], half]), full, rtol=1e-13, atol=1e-18)
    g, _ = T.gaussian_taps(sigma, 0)
    assert abs(2 * g[1:].sum() + g[0] - 1.0) < 1e-15
    d, _ = T.gaussian_taps(2.0, 2)
    assert abs(2 * d[1:].sum() + d[0]) > 1e-6  # the DC leak of the order-2 taps is NOT corrected


def test_three_pass_factorisation_equals_gaussian_laplace():
    """y: (G, D) -> z: (G A, D A + G B) -> x: D C + G F, the schedule the CUDA path uses"""
    rng = np.random.default_rng(3)
    a = rng.random((30, 28, 26))

    def c1(v, axis, order):
        return ndi.gaussian_filter1d(v, 2.0, axis=axis, order=order, mode="reflect")

    A, B = c1(a, 1, 0), c1(a, 1, 2)
    C, F = c1(A, 2, 0), c1(A, 2, 2) + c1(B, 2, 0)
    log = c1(C, 0, 2) + c1(F, 0, 0)
    np.testing.assert_allclose(log, ndi.gaussian_laplace(a, 2.0), atol=1e-15)


# ----------------------------------------------------------------------- NMS
def test_peak_local_max_semantics():
    cube = np.zeros((9, 9, 9, 1))
    cube[0, 0, 0, 0] = 0.5       # corner voxel: out-of-volume neighbours are 0
    cube[4, 4, 4, 0] = 0.7
    cube[4, 4, 5, 0] = 0.7       # 2-voxel plateau: both reported
    cube[7, 7, 7, 0] = 0.3       # at the threshold: '>' is strict
    pk = oracle.peak_local_max(cube, threshold_abs=0.3)
    assert [tuple(p) for p in pk] == [(4, 4, 4, 0), (4, 4, 5, 0), (0, 0, 0, 0)]
    assert len(oracle.peak_local_max(np.full((4, 4, 4, 1), 0.9), 0.1)) == 0  # constant image


def test_single_blob_gives_one_spot_at_its_centre():
    a = np.zeros((41, 41, 21))
    gx = np.exp(-0.5 * ((np.arange(41) - 20) / 2.0) ** 2)
    gz = np.exp(-0.5 * ((np.arange(21) - 10) / 2.0) ** 2)
    a += gx[:, None, None] * gx[None, :, None] * gz[None, None, :]
    spots = oracle.find_spots(a, sigma=2.0, t_spot=0.025)
    np.testing.assert_array_equal(spots, [[20, 20, 10, 2.0]])


# --------------------------------------------------------------------- prune
def test_prune_threshold_distance():
    """equal sigma: pairs closer than 0.6946*sigma*sqrt(3) (2.41 voxels at sigma=2) lose one blob"""
    b = np.array([[10, 10, 10, 2.0], [12, 10, 10, 2.0]])
    out = oracle.prune_blobs(b)
    np.testing.assert_array_equal(out, [[12, 10, 10, 2.0]])  # the LOWER index is zeroed
    b = np.array([[10, 10, 10, 2.0], [13, 10, 10, 2.0]])
    assert len(oracle.prune_blobs(b)) == 2
    assert oracle.prune_cutoff_sq(1.0) == 1 and oracle.prune_cutoff_sq(2.0) == 5
    # unequal sigma: the larger sigma wins regardless of index
    b = np.array([[10, 10, 10, 3.0], [11, 10, 10, 1.0]])
    np.testing.assert_array_equal(oracle.prune_blobs(b), [[10, 10, 10, 3.0]])
    b = b[::-1].copy()
    np.testing.assert_array_equal(oracle.prune_blobs(b), [[10, 10, 10, 3.0]])


def test_prune_cutoff_tables_agree():
    for s1 in (1.0, 1.5, 2.0, 2.5, 3.0):
        assert T.prune_cutoff_sq(s1) == oracle.prune_cutoff_sq(s1)
        for s2 in (1.0, 2.0, 3.0):
            n = T.prune_cutoff_sq(s1, s2)
            if n >= 1:
                assert oracle.blob_overlap([0, 0, 0, s1], [math.sqrt(n), 0, 0, s2]) > 0.5
            assert oracle.blob_overlap([0, 0, 0, s1], [math.sqrt(n + 1), 0, 0, s2]) <= 0.5


def test_prune_order_independent_on_config_density():
    st = synth.make_config("small")
    img = oracle.normalize_image(st["img"])
    a, raw = oracle.blob_log(img, 2, 2, 1, 0.025, pair_order="set", return_raw=True)
    b = oracle.blob_log(img, 2, 2, 1, 0.025, pair_order="sorted")
    assert len(raw) >= len(a) > 100
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------- spots / SNR
def test_find_spots_snrfilter_shapes_and_errors():
    st = synth.make_config("tiny")
    spots, data = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"])
    assert spots.ndim == 2 and spots.shape[1] == 4 and spots.dtype == np.float64
    assert data[0] == ['spot_id', 'x', 'y', 'z', 'intensity', 'SNR'] and len(data) - 1 >= len(spots)
    truth = np.rint(st["centers"]).astype(int)
    hits = sum(np.min(np.abs(truth - s[:3]).max(axis=1)) <= 2 for s in spots)
    assert hits >= 0.9 * len(spots)
    with pytest.raises(ValueError):
        oracle.find_spots(st["img"], mask=np.zeros(st["img"].shape, np.uint8))
    zs, _ = oracle.find_spots_snrfilter(st["img"].transpose(2, 0, 1), sigma=2.0, t_spot=0.025,
                                        mask=st["fiber_mask"].transpose(2, 0, 1), z_first=True)
    np.testing.assert_array_equal(zs[:, [1, 2, 0, 3]], spots)


def test_fix_bleaching_recovers_decay():
    st = synth.make_config("small")
    img = st["img"].astype(np.float64) * np.exp(-0.03 * np.arange(st["img"].shape[2]))
    mask = np.zeros(img.shape, np.uint8)
    mask[:, 40:150, :] = 1  # a mask present in every plane (an empty plane gives a NaN mean)
    out, k, _ = oracle.fix_bleaching(img, mask=mask, return_fit=True)
    assert k is not None and -0.05 < k < -0.01
    up = st["img"].astype(np.float64) * np.exp(0.02 * np.arange(st["img"].shape[2]))
    out2, k2, _ = oracle.fix_bleaching(up, return_fit=True)
    assert k2 is None and out2 is up


# ----------------------------------------------------------------------- EDT
def test_edt_single_zero_voxel_analytic():
    a = np.ones((11, 12, 13), np.uint8)
    a[3, 4, 5] = 0
    s = (0.0577, 0.0577, 0.5)
    d = oracle.distance_transform_edt(a, sampling=s)
    gx, gy, gz = np.meshgrid(np.arange(11), np.arange(12), np.arange(13), indexing="ij")
    np.testing.assert_allclose(d, np.sqrt(((gx - 3) * s[0]) ** 2 + ((gy - 4) * s[1]) ** 2 + ((gz - 5) * s[2]) ** 2),
                               rtol=1e-14)


@settings(max_examples=25, deadline=None)
@given(hst.integers(1, 9), hst.integers(1, 9), hst.integers(1, 9), hst.floats(0.02, 0.6), hst.integers(0, 10 ** 6),
       hst.sampled_from([None, (0.0577, 0.0577, 0.5), (1.5, 0.4, 2.5)]))
def test_edt_scipy_equals_bruteforce(nx, ny, nz, p, seed, sampling):
    rng = np.random.default_rng(seed)
    a = (rng.random((nx, ny, nz)) >= p).astype(np.uint8)
    a.flat[rng.integers(a.size)] = 0
    ref = oracle.edt_bruteforce_sq(a, sampling)
    got = oracle.distance_transform_edt(a, sampling=sampling) ** 2
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    if sampling is None:
        np.testing.assert_array_equal(np.rint(got), ref)


def test_edt_gather_is_a_pure_gather():
    st = synth.make_config("tiny")
    dims = synth.SAMPLING_ANISO
    g = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=dims)
    spots = np.array([[1, 2, 3, 1.0], [50, 40, 10, 1.0], [95, 79, 23, 1.0]])
    np.testing.assert_array_equal(oracle.edt_gather(g, spots, dims), g[tuple(spots[:, :3].astype(int).T)])


# -------------------------------------------------------------------- golden
def test_oracle_reproduces_golden(golden):
    img = golden["img"]
    norm = oracle.normalize_image(img)
    np.testing.assert_array_equal(norm, golden["norm"])
    np.testing.assert_array_equal(oracle.log_cube(norm, [2.0])[..., 0], golden["log_s2"])
    np.testing.assert_array_equal(oracle.gaussian(norm, (3, 3, 0.3)), golden["blur_3_3_03"])
    np.testing.assert_array_equal(oracle.find_spots(img, sigma=2.0, t_spot=0.025), golden["spots_s2"])
    sp, data = oracle.find_spots_snrfilter(img, sigma=2.0, t_spot=0.025, mask=golden["fiber_mask"])
    np.testing.assert_array_equal(sp, golden["spots_snr_s2"])
    np.testing.assert_array_equal(np.array(data[1:], dtype=np.float64), golden["spot_data_snr_s2"])
    not_nuc = np.logical_not(golden["nuc_mask"])
    np.testing.assert_array_equal(np.rint(oracle.distance_transform_edt(not_nuc) ** 2), golden["edt_iso_sq"])
    np.testing.assert_array_equal(oracle.distance_transform_edt(not_nuc, sampling=(0.0577, 0.0577, 0.5)),
                                  golden["edt_aniso"])
    np.testing.assert_array_equal(golden["edt_iso_sq"], oracle.edt_bruteforce_sq(not_nuc))
    assert match_fraction(oracle.find_spots(img, sigma=2.0, t_spot=0.025, pair_order="sorted"),
                          golden["spots_s2"]) == 1.0


def test_compartment_oracle_matches_golden_and_the_closed_form():
    """general_pipeline/fish_analysis.py:165-187: the loop nest equals one L1-ball dilation per plane window
    (what the device computes), and reproduces the committed golden masks"""
    import os
    from scipy import ndimage as ndi
    from oracle import ref_morph as M
    gm = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_morph.npz"))
    nuc, fib = gm["nuc_mask"], gm["fiber_mask"]
    reg = M.compartments(nuc, fib, (1, 2), (2, 5))
    for k in reg:
        assert reg[k].dtype == np.int64
        np.testing.assert_array_equal(reg[k], gm[f"small_{k}"])

    def closed(mask, op, n3, n2):
        src = mask if op == "dilate" else ~mask
        nx, ny, nz = mask.shape
        d2 = np.stack([ndi.distance_transform_cdt(~src[:, :, z], metric="taxicab") if src[:, :, z].any()
                       else np.full((nx, ny), 10 ** 6) for z in range(nz)], axis=2)
        out = np.zeros(mask.shape, bool)
        for z in range(nz):
            for dz in range(-n3, n3 + 1):
                if 0 <= z + dz < nz:
                    out[:, :, z] |= d2[:, :, z + dz] <= n2 + n3 - abs(dz)
        return out if op == "dilate" else ~out

    m = nuc.astype(bool)
    for op, n3, n2 in [("dilate", 4, 7), ("erode", 1, 3), ("dilate", 0, 5), ("erode", 2, 0)]:
        np.testing.assert_array_equal(M.iterate(m, op, n3, n2), closed(m, op, n3, n2))


def test_oracle_label_and_fade_reproduce_golden():
    """tests/golden/golden_label.npz (real scipy label / gaussian_filter): the oracle's restatements of skimage's
    label / remove_small_objects / remove_small_holes (ref_segment) and of the per-nucleus loop (ref_fade)."""
    import os
    from oracle import ref_fade, ref_segment as S
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_label.npz"))
    m = g["mask"].astype(bool)
    lab, n = S.label(m)
    assert n == int(g["n_full"])
    np.testing.assert_array_equal(lab, g["label_full"])
    np.testing.assert_array_equal(S.remove_small_objects(m, 30), g["no_small_objects_30"].astype(bool))
    np.testing.assert_array_equal(S.remove_small_holes(m, 30), g["no_small_holes_30"].astype(bool))
    nuc_lab = g["nuc_labeled"]
    lab2, n2 = ref_fade.label_nuclei(nuc_lab >= 0)
    assert n2 == int(g["n_nuc"])
    np.testing.assert_array_equal(lab2, nuc_lab)
    for i in range(n2):
        np.testing.assert_allclose(ref_fade.nucleus_fade(nuc_lab, i), g[f"fade_{i}"], rtol=0, atol=1e-7)


def test_two_band_envelope_model_equals_bruteforce():
    """the specification of the banded x pass of the EDT (oracle.ref_edt.two_band_envelope; csrc/edt.cu BANDS = 2):
    mirrored bands joined by the lower common tangent give the plain lower envelope, on random, sparse, dense,
    one-sided and empty lines"""
    import random
    from oracle.ref_edt import two_band_envelope
    rnd = random.Random(3)
    for _ in range(3000):
        L = rnd.choice([1, 2, 3, 4, 5, 7, 8, 16, 33, 64, 100])
        dens = rnd.choice([0.0, 0.05, 0.3, 0.9, 1.0])
        mx = rnd.choice([1, 3, 10, 50, 2000])
        mode = rnd.randrange(4)
        F = []
        for u in range(L):
            if rnd.random() < dens and not (mode == 3 and u >= L // 2):
                d = rnd.randrange(mx) if mode in (0, 3) else (abs(u - L // 3) * 2 if mode == 1 else rnd.choice([0, mx]))
                F.append(d * d)
            else:
                F.append(None)
        sites = [(u, f) for u, f in enumerate(F) if f is not None]
        want = [min(f + (p - u) ** 2 for u, f in sites) for p in range(L)] if sites else [None] * L
        assert two_band_envelope(F) == want, F
