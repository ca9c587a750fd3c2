"""GPU parity: compartment masks -- the erosion / dilation loops of
general_pipeline/fish_analysis.py:165-187 as one capped city-block distance transform, bit exact against
scipy.ndimage iterated exactly as the script iterates it (oracle.ref_morph) and against the committed
golden masks."""
import os

import numpy as np
import pytest

import oracle
from oracle import ref_morph as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gm():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_morph.npz"))


def _rand_mask(shape, p, seed):
    return (np.random.default_rng(seed).random(shape) < p)


@pytest.mark.parametrize("shape", [(40, 36, 12), (33, 70, 5), (64, 64, 1), (7, 9, 30), (130, 20, 3)])
@pytest.mark.parametrize("op,n3,n2", [("dilate", 4, 7), ("dilate", 0, 5), ("dilate", 3, 0), ("erode", 1, 3),
                                      ("erode", 2, 0), ("erode", 0, 4), ("dilate", 1, 40)])
def test_iterated_cross_morphology_bit_exact(shape, op, n3, n2):
    from muscle_fish_b200 import morphology as mo
    sparse = _rand_mask(shape, 0.01, seed=n3 * 10 + n2)
    # erosion needs bodies to eat into: a dilated version of the sparse mask (and its faces touch the border)
    mask = sparse if op == "dilate" else M.iterate(sparse, "dilate", 3, 3)
    ref = M.iterate(mask, op, n3, n2)
    got = mo.iterate(mask, op, n3, n2)
    assert got.dtype == bool and got.shape == mask.shape
    np.testing.assert_array_equal(got, ref)
    if op == "erode":
        assert ref.any() or n3 + n2 > 3


def test_single_steps_are_the_import_line_seam():
    """from muscle_fish_b200 import morphology  (was: from skimage import morphology)"""
    from muscle_fish_b200 import morphology as mo
    m3 = M.iterate(_rand_mask((30, 28, 9), 0.02, 5), "dilate", 2, 2)
    np.testing.assert_array_equal(mo.binary_erosion(m3), M.binary_erosion(m3))
    np.testing.assert_array_equal(mo.binary_dilation(m3), M.binary_dilation(m3))
    m2 = m3[:, :, 4]  # a non-contiguous 2-D plane view, as the scripts pass region[:, :, z]
    np.testing.assert_array_equal(mo.binary_erosion(m2), M.binary_erosion(m2))
    np.testing.assert_array_equal(mo.binary_dilation(m2), M.binary_dilation(m2))
    # int64 0/1 input (np.where(..., 1, 0)) and the literal loop of fish_analysis.py:171-175 run through the seam
    region = np.where(m3, 1, 0)
    ref = np.where(m3, 1, 0)
    region = mo.binary_erosion(region)
    ref = M.binary_erosion(ref)
    for n in range(2):
        for z in range(region.shape[2]):
            region[:, :, z] = mo.binary_erosion(region[:, :, z])
            ref[:, :, z] = M.binary_erosion(ref[:, :, z])
    np.testing.assert_array_equal(region, ref)
    with pytest.raises(NotImplementedError):
        mo.binary_dilation(m3, selem=np.ones((3, 3, 3)))


def test_compartments_vs_oracle_and_golden(gm):
    from muscle_fish_b200 import morphology as mo
    nuc, fib = gm["nuc_mask"], gm["fiber_mask"]
    for tag, er, di in (("ref", (1, 10), (4, 40)), ("small", (1, 2), (2, 5))):
        got = mo.compartments(nuc, fib, er, di)
        ref = M.compartments(nuc, fib, er, di)
        assert set(got) == {"nuc", "cyt", "peri"}
        for k in got:
            assert got[k].dtype == np.int64
            np.testing.assert_array_equal(got[k], ref[k])
            np.testing.assert_array_equal(got[k], gm[f"{tag}_{k}"])
        # the three regions are disjoint inside the fibre and leave nothing of it unassigned but the shell
        # between the eroded and the original nucleus, which the script assigns to 'peri'
        assert not (got["nuc"] & got["cyt"]).any() and not (got["peri"] & got["cyt"]).any()
    # labelled nuclei (values > 1) and int64 fibre masks, as the scripts hold them
    lab = nuc.astype(np.int64) * 7
    got = mo.compartments(lab, fib.astype(np.int64), (1, 2), (2, 5), dtype=np.uint8)
    np.testing.assert_array_equal(got["peri"], gm["small_peri"])
    np.testing.assert_array_equal(mo.binary_erosion(nuc), gm["erode_3d_1"])
    np.testing.assert_array_equal(mo.binary_dilation(nuc), gm["dilate_3d_1"])
    np.testing.assert_array_equal(mo.binary_dilation(nuc[:, :, 7]), gm["dilate_2d_plane7"])


def test_compartments_at_config2_size_properties():
    """2048x2048x100: too large for the oracle's Python loops, so (i) crops with a margin larger than the
    reach (44 voxels) are compared exactly, (ii) whole-volume identities: erode <= mask <= dilate,
    idempotence of the zero-iteration call, monotonicity in the iteration count"""
    import torch
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import device as D, synth
    st = synth.make_stack_device((2048, 2048, 100), 100, 120, 2)
    nuc, fib = st["nuc_mask"], st["fiber_mask"]
    r_nuc, r_peri, r_cyt = D.compartments(nuc, fib, (1, 10), (4, 40))
    dil = D.binary_morph(nuc, "dilate", 4, 40)
    assert bool((r_nuc <= nuc).all()) and bool((nuc <= dil).all())
    assert bool(torch.equal(D.binary_morph(nuc, "dilate", 0, 0), (nuc != 0).to(torch.uint8)))
    assert bool((D.binary_morph(nuc, "dilate", 4, 39) <= dil).all())
    assert bool(torch.equal(r_cyt, ((fib != 0) & (dil == 0)).to(torch.uint8)))
    assert bool(torch.equal(r_peri, ((dil > r_nuc) & ((fib != 0) | (nuc != 0))).to(torch.uint8)))
    H = 48
    for (x0, y0) in [(0, 500), (1000, 900), (2048 - 256, 1100)]:
        crop = lambda t: t[:, x0:x0 + 256, y0:y0 + 256].permute(1, 2, 0).contiguous().cpu().numpy()
        ref = M.compartments(crop(nuc), crop(fib), (1, 10), (4, 40))
        xa = 0 if x0 == 0 else H
        xb = 256 if x0 + 256 == 2048 else 256 - H
        for k, t in (("nuc", r_nuc), ("peri", r_peri), ("cyt", r_cyt)):
            np.testing.assert_array_equal(crop(t)[xa:xb, H:256 - H], ref[k][xa:xb, H:256 - H])
