"""GPU parity: per-nucleus faded masks (nocodazole/fish_analysis_noco.py:227-235) against the
oracle's full-volume scipy Gaussian of each one-hot mask.  float32 filter sums: <= 1e-5 of max|ref| (= of 1.0)."""
import numpy as np
import pytest

import oracle
from oracle import ref_fade

pytestmark = pytest.mark.gpu


def _labels(shape, n, seed, touch_border=True):
    """n disjoint ellipsoids (some cut by the volume's faces), labels 0..n-1, background -1."""
    rng = np.random.default_rng(seed)
    lab = -np.ones(shape, np.int64)
    X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    k = 0
    tries = 0
    while k < n and tries < 200:
        tries += 1
        m_ = [0 if touch_border else min(8, s // 3) for s in shape]
        c = [rng.uniform(m, s - 1 - m) for m, s in zip(m_, shape)]
        r = (rng.uniform(4, 11), rng.uniform(3, 7), rng.uniform(1.5, 3.5))
        m = ((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2 <= 1.0
        if not m.any() or (lab[m] >= 0).any():
            continue
        lab[m] = k
        k += 1
    return lab, k


@pytest.mark.parametrize("shape,n,sigma", [((64, 48, 12), 5, (3, 3, 1)), ((40, 90, 9), 7, (3, 3, 1)),
                                           ((33, 31, 5), 3, (1.5, 2.0, 0.5))])
def test_nucleus_fades_match_full_volume_gaussian(shape, n, sigma):
    from muscle_fish_b200 import filters
    lab, n = _labels(shape, n, seed=sum(shape))
    fades = filters.nucleus_fades(lab, n, sigma)
    assert len(fades) == n
    for i in range(n):
        ref = ref_fade.nucleus_fade(lab, i, sigma)
        got = fades[i]
        assert got.shape == ref.shape and got.dtype == np.float64
        assert np.abs(got - ref).max() <= 1e-5
        sl, crop = fades.crop(i)
        outside = np.ones(shape, bool)
        outside[sl] = False
        assert np.all(ref[outside] == 0.0), "the reference is exactly zero outside the grown box"
        np.testing.assert_array_equal(got[sl], crop)


def test_nucleus_fades_reference_loop():
    """The loop of fish_analysis_noco.py:227-235 on a compartment mask: labels from the oracle's label step."""
    from muscle_fish_b200 import filters
    lab0, n0 = _labels((72, 60, 10), 6, seed=3, touch_border=False)
    region_nuc = (lab0 >= 0).astype(np.int64)
    lab, n = ref_fade.label_nuclei(region_nuc)
    fades = filters.nucleus_fades(lab, n)
    for i, got in enumerate(fades):
        nuc = np.where(lab == i, 1, 0)
        ref = oracle.gaussian(nuc.astype(float), (3, 3, 1))
        assert np.abs(got - ref).max() <= 1e-5
        assert (got > 0.5).sum() > 0  # the 0.5 iso-surface marching cubes extracts exists


def test_nucleus_fades_absent_label_and_empty():
    from muscle_fish_b200 import filters
    lab = -np.ones((20, 20, 6), np.int64)
    lab[3:8, 4:9, 1:4] = 1  # label 0 never occurs
    fades = filters.nucleus_fades(lab, 2)
    assert np.all(fades[0] == 0.0)
    assert fades.crop(0) == (None, None)
    ref = ref_fade.nucleus_fade(lab, 1)
    assert np.abs(fades[1] - ref).max() <= 1e-5
    assert len(filters.nucleus_fades(-np.ones((5, 5, 5), np.int64), 0)) == 0


def test_filters_gaussian_signature():
    """filters.gaussian(image, sigma) == skimage.filters.gaussian on a float volume (generic passes)."""
    from muscle_fish_b200 import filters
    rng = np.random.default_rng(0)
    img = rng.random((30, 25, 8))
    got = filters.gaussian(img, (3, 3, 1))
    ref = oracle.gaussian(img, (3, 3, 1))
    assert got.dtype == np.float64
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
