"""Generate tests/golden/golden_small.npz.

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


if __name__ == "__main__":
    main()
