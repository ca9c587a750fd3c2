"""Generate tests/golden/golden_small.npz (default), golden_morph.npz (`morph`), golden_label.npz (`label`).

The reference's own modules cannot be imported in this image (matplotlib / scikit-image /
czifile are absent), and it ships no test vectors.  The fixtures are therefore produced by the
REAL third-party routines the reference calls for the arithmetic -- scipy.ndimage
gaussian_laplace / gaussian_filter / maximum_filter / distance_transform_edt (scipy 1.18.1
here) -- driven by the oracle's restatement of skimage's orchestration.  Run from the repo
root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy
from scipy import ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
import muscle_fish_b200  # noqa: E402,F401
from muscle_fish_b200 import synth  # noqa: E402

ANISO = (0.0577, 0.0577, 0.5)


def main():
    st = synth.make_stack(shape=(56, 48, 16), n_spots=60, n_nuclei=2, seed=99, nuc_semi=(9.0, 6.0, 3.0))
    img = st["img"]
    norm = oracle.normalize_image(img)
    out = {
        "img": img,
        "fiber_mask": st["fiber_mask"],
        "nuc_mask": st["nuc_mask"],
        "norm": norm,
        "log_s1": -ndi.gaussian_laplace(norm, 1.0) * 1.0,
        "log_s2": -ndi.gaussian_laplace(norm, 2.0) * 4.0,
        "blur_3_3_03": ndi.gaussian_filter(norm, (3, 3, 0.3), mode="nearest", truncate=4.0),
        "spots_s2": oracle.find_spots(img, sigma=2.0, t_spot=0.025, pair_order="set"),
        "spots_s1_masked": oracle.find_spots(img, sigma=1.0, t_spot=0.02, mask=st["fiber_mask"], pair_order="set"),
        "edt_iso_sq": np.rint(ndi.distance_transform_edt(np.logical_not(st["nuc_mask"])) ** 2).astype(np.int32),
        "edt_aniso": ndi.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO),
        "scipy_version": np.array(scipy.__version__),
    }
    spots, data = oracle.find_spots_snrfilter(img, sigma=2.0, t_spot=0.025, mask=st["fiber_mask"], pair_order="set")
    out["spots_snr_s2"] = spots
    out["spot_data_snr_s2"] = np.array([r for r in data[1:]], dtype=np.float64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path), "bytes")


def main_morph():
    """tests/golden/golden_morph.npz: the compartment masks of general_pipeline/fish_analysis.py:165-187 by
    the real scipy.ndimage binary_erosion / binary_dilation, driven by the oracle's literal restatement of
    the script's loops (skimage's wrappers only add border_value=True to the erosion)."""
    st = synth.make_stack(shape=(72, 64, 20), n_spots=10, n_nuclei=3, seed=98, nuc_semi=(14.0, 9.0, 5.0))
    nuc, fib = st["nuc_mask"], st["fiber_mask"]
    out = {"nuc_mask": nuc, "fiber_mask": fib}
    for tag, er, di in (("ref", (1, 10), (4, 40)), ("small", (1, 2), (2, 5))):
        reg = oracle.ref_morph.compartments(nuc, fib, er, di)
        for k, v in reg.items():
            out[f"{tag}_{k}"] = v.astype(np.uint8)
    out["erode_3d_1"] = ndi.binary_erosion(nuc, border_value=True)
    out["dilate_3d_1"] = ndi.binary_dilation(nuc)
    out["dilate_2d_plane7"] = ndi.binary_dilation(nuc[:, :, 7])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_morph.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape, int(v.sum())) for k, v in out.items()}, os.path.getsize(path), "bytes")


def main_label():
    """tests/golden/golden_label.npz: connected components (both neighbourhoods), the small-object / small-hole
    clean-up and the per-nucleus faded masks of modules/muscle_fish.py:182-184 and
    nocodazole/fish_analysis_noco.py:227-235 by the real scipy.ndimage label / gaussian_filter."""
    rng = np.random.default_rng(97)
    shape = (48, 40, 12)
    m = ndi.gaussian_filter(rng.random(shape), 1.5) > 0.5
    m |= rng.random(shape) < 0.01
    m &= ~(rng.random(shape) < 0.01)
    lab26, n26 = ndi.label(m, structure=np.ones((3, 3, 3), bool))
    lab6, n6 = ndi.label(m)
    out = {"mask": m.astype(np.uint8), "label_full": lab26.astype(np.int32), "n_full": np.array(n26),
           "label_cross": lab6.astype(np.int32), "n_cross": np.array(n6)}
    sizes = np.bincount(lab6.ravel())
    small = sizes < 30
    small[0] = False
    out["no_small_objects_30"] = (m & ~small[lab6]).astype(np.uint8)
    inv = ~m
    labh, _ = ndi.label(inv)
    hs = np.bincount(labh.ravel())
    smallh = hs < 30
    smallh[0] = False
    out["no_small_holes_30"] = (~(inv & ~smallh[labh])).astype(np.uint8)
    st = synth.make_stack(shape=(72, 64, 20), n_spots=10, n_nuclei=3, seed=98, nuc_semi=(14.0, 9.0, 5.0))
    nuc_lab, n = ndi.label(st["nuc_mask"] > 0, structure=np.ones((3, 3, 3), bool))
    out["nuc_labeled"] = (nuc_lab - 1).astype(np.int32)  # bg = -1, nuclei 0..n-1 (fish_analysis_noco.py:229)
    for i in range(n):
        out[f"fade_{i}"] = ndi.gaussian_filter((nuc_lab - 1 == i).astype(float), (3, 3, 1), mode="nearest",
                                               truncate=4.0).astype(np.float32)
    out["n_nuc"] = np.array(n)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_label.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "label":
        main_label()
    elif len(sys.argv) > 1 and sys.argv[1] == "morph":
        main_morph()
    else:
        main()
