"""GPU parity: fibre / nucleus segmentation (modules/muscle_fish.py:114-278) -- large-radius blurs,
256-bin histogram, Otsu and Li thresholds, binarisation -- against the oracle's restatement on scipy / numpy.
The thresholds are compared to within one bin of the 256-bin histogram (the blurred volume is float32 here,
float64 there, and Li's iteration stops on a half-bin tolerance); histogram and split statistics on the SAME
float32 volume are bit exact."""
import numpy as np
import pytest

import oracle
from oracle import ref_segment as S

pytestmark = pytest.mark.gpu


def _dapi(st, seed=0):
    rng = np.random.default_rng(seed)
    return (st["nuc_mask"] * 900.0 + 100.0 + rng.normal(0, 20, st["img"].shape)).astype(np.float32)


def test_histogram_bit_exact_on_the_same_volume():
    import torch
    from muscle_fish_b200 import device as D
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((37, 45, 51)) * 3 + 10).astype(np.float32)
    a[::5] = a.max()  # values on the last edge
    a[1::7] = a.min()
    t = torch.from_numpy(a).cuda()
    hist, edges = D.histogram(t, 256)
    ref_h, ref_e = np.histogram(a.astype(np.float64).ravel(), 256)
    np.testing.assert_array_equal(edges, ref_e)
    np.testing.assert_array_equal(hist, ref_h)
    hist, edges = D.histogram(t, 100, (8.0, 12.5))
    ref_h, ref_e = np.histogram(a.astype(np.float64).ravel(), 100, range=(8.0, 12.5))
    np.testing.assert_array_equal(hist, ref_h)
    (c0, s0), (c1, s1) = D.split_moments(t, 10.0)
    assert c0 == (a <= 10.0).sum() and c1 == (a > 10.0).sum()
    np.testing.assert_allclose([s0, s1], [a[a <= 10.0].sum(dtype=np.float64), a[a > 10.0].sum(dtype=np.float64)],
                               rtol=1e-12)
    got = D.binarize(t, 10.0).cpu().numpy()
    np.testing.assert_array_equal(got, (a > 10.0).astype(np.uint8))


def test_otsu_and_li_thresholds_within_one_bin():
    from muscle_fish_b200 import device as D, segmentation as sg, synth
    st = synth.make_config("small")
    for img, sig in ((st["img"], (10, 10, 1)), (_dapi(st), (20, 20, 2))):
        norm = oracle.normalize_image(img)
        blur_ref = oracle.gaussian(norm, sig)
        blur = D.gaussian(D.ingest(img), sig)
        bin_w = (blur_ref.max() - blur_ref.min()) / 256
        assert abs(sg.threshold_otsu(blur) - S.threshold_otsu(blur_ref)) <= bin_w
        assert abs(sg.threshold_li(blur) - S.threshold_li(blur_ref)) <= bin_w
    # on identical float32 data both are exact / equal to rounding
    import torch
    b32 = blur.cpu().numpy()
    assert sg.threshold_otsu(blur) == S.threshold_otsu(b32.astype(np.float64))
    assert abs(sg.threshold_li(blur) - S.threshold_li(b32)) <= 1e-6 * bin_w + 1e-12


@pytest.mark.parametrize("z_first", [False, True])
def test_threshold_fiber_and_nuclei_vs_oracle(z_first, capsys):
    from muscle_fish_b200 import segmentation as sg, synth
    st = synth.make_config("small")
    img, dapi = st["img"], _dapi(st)
    if z_first:
        img, dapi = np.ascontiguousarray(img.transpose(2, 0, 1)), np.ascontiguousarray(dapi.transpose(2, 0, 1))
    ref_f, t_f, blur_f = S.threshold_fiber(img, z_first=z_first, return_threshold=True)
    got_f = sg.threshold_fiber(img, z_first=z_first, verbose=True)
    assert "Fiber threshold: " in capsys.readouterr().out
    assert got_f.shape == ref_f.shape and got_f.dtype == ref_f.dtype
    # voxels whose blurred value sits within float32 rounding of the threshold may flip
    assert (got_f != ref_f).mean() < 2e-3
    assert 0.3 < got_f.mean() < 0.9
    ref_n = S.threshold_nuclei(dapi, fiber_mask=ref_f, z_first=z_first)
    got_n = sg.threshold_nuclei(dapi, fiber_mask=ref_f, z_first=z_first, verbose=True)
    assert "DAPI threshold = " in capsys.readouterr().out
    assert got_n.max() == ref_n.max() >= 1
    assert ((got_n > 0) != (ref_n > 0)).mean() < 1e-4
    got_b = sg.threshold_nuclei(dapi, fiber_mask=ref_f, labeled=False, z_first=z_first)
    np.testing.assert_array_equal(got_b, np.where(got_n > 0, 1, 0))
    # explicit thresholds and the error paths of the reference
    np.testing.assert_array_equal(sg.threshold_fiber(img, t_fiber=float(t_f), z_first=z_first) > 0,
                                  S.threshold_fiber(img, t_fiber=float(t_f), z_first=z_first) > 0)
    np.testing.assert_array_equal(sg.threshold_nuclei(dapi, t_dapi=0.2, z_first=z_first) > 0,
                                  S.threshold_nuclei(dapi, t_dapi=0.2, z_first=z_first) > 0)
    with pytest.raises(TypeError):
        sg.threshold_fiber(img, t_fiber=3)
    with pytest.raises(TypeError):
        sg.threshold_nuclei(dapi, t_dapi="x")
    if not z_first:
        last_ref = S.threshold_fiber(img, t_fiber='last')
        last_got = sg.threshold_fiber(img, t_fiber='last')
        assert (last_got != last_ref).mean() < 2e-3


def _blobs(shape, p, seed):
    """random mask with structure at several scales: noise voxels, blobs, holes"""
    from scipy import ndimage as ndi
    rng = np.random.default_rng(seed)
    m = ndi.gaussian_filter(rng.random(shape), 1.5) > (0.5 + 0.02 * (1 - 2 * p))
    m |= rng.random(shape) < 0.02 * p
    m &= ~(rng.random(shape) < 0.01)
    return m


@pytest.mark.parametrize("shape", [(40, 33, 17), (9, 64, 5), (1, 50, 60), (70, 45), (3, 3, 3), (128, 96, 24)])
@pytest.mark.parametrize("full", [True, False])
@pytest.mark.parametrize("p", [0.2, 0.8])
def test_label_matches_scipy_numbering(shape, full, p):
    """csrc/ccl.cu: same components AND the same label numbers as scipy.ndimage.label / skimage.morphology.label
    (components ranked by their first voxel in raster order), both neighbourhoods, 2-D and 3-D"""
    import torch
    from scipy import ndimage as ndi
    from muscle_fish_b200 import segmentation as sg
    m = _blobs(shape, p, seed=len(shape) * 7 + shape[0])
    structure = np.ones((3,) * m.ndim, bool) if full else None
    ref, n_ref = ndi.label(m, structure=structure)
    lab, n = sg.label_device(torch.from_numpy(m).cuda(), full=full)
    assert n == n_ref and n >= 1
    np.testing.assert_array_equal(lab.cpu().numpy(), ref)
    got, n2 = sg.label(m, connectivity=None if full else 1, return_num=True)
    assert n2 == n_ref and got.dtype == np.int64
    np.testing.assert_array_equal(got, ref)


def test_label_edge_cases():
    import torch
    from muscle_fish_b200 import segmentation as sg, morphology
    z = torch.zeros((5, 6, 7), dtype=torch.uint8, device="cuda")
    lab, n = sg.label_device(z)
    assert n == 0 and not lab.any()
    o = torch.ones((5, 6, 7), dtype=torch.uint8, device="cuda")
    lab, n = sg.label_device(o, full=False)
    assert n == 1 and bool((lab == 1).all())
    # a long snake: many union steps along one chain
    s = np.zeros((64, 64), bool)
    s[::2, :] = True
    s[1::4, -1] = True
    s[3::4, 0] = True
    out, n = morphology.label(s, connectivity=1, return_num=True)
    assert n == 1 and np.array_equal(out != 0, s)


@pytest.mark.parametrize("min_size", [4, 40])
def test_remove_small_objects_and_holes_vs_oracle(min_size):
    """skimage.morphology.remove_small_objects / remove_small_holes on bool arrays (connectivity 1), as
    modules/muscle_fish.py:182-183 call them"""
    from muscle_fish_b200 import morphology
    m = _blobs((48, 40, 12), 0.5, seed=9)
    np.testing.assert_array_equal(morphology.remove_small_objects(m, min_size), S.remove_small_objects(m, min_size))
    np.testing.assert_array_equal(morphology.remove_small_holes(m, min_size), S.remove_small_holes(m, min_size))


def test_label_cleanup_and_fades_vs_golden():
    """the committed fixture of the real scipy routines (tests/golden/golden_label.npz)"""
    import os
    from muscle_fish_b200 import filters, morphology
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_label.npz"))
    m = g["mask"].astype(bool)
    lab, n = morphology.label(m, return_num=True)
    assert n == int(g["n_full"])
    np.testing.assert_array_equal(lab, g["label_full"])
    lab1, n1 = morphology.label(m, connectivity=1, return_num=True)
    assert n1 == int(g["n_cross"])
    np.testing.assert_array_equal(lab1, g["label_cross"])
    np.testing.assert_array_equal(morphology.remove_small_objects(m, 30), g["no_small_objects_30"].astype(bool))
    np.testing.assert_array_equal(morphology.remove_small_holes(m, 30), g["no_small_holes_30"].astype(bool))
    nuc_lab = g["nuc_labeled"]
    fades = filters.nucleus_fades(nuc_lab, int(g["n_nuc"]))
    for i in range(int(g["n_nuc"])):
        assert np.abs(fades[i] - g[f"fade_{i}"]).max() <= 1e-5
