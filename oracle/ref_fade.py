"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Per-nucleus faded masks of the nuclear-mesh step, ``nocodazole/fish_analysis_noco.py:227-235``::

    region_nuc_labeled, n_nuc = morphology.label(region_nuc, return_num=True)
    region_nuc_labeled = region_nuc_labeled - 1  # shift so bg = -1 and nuclei = [0:n_nuc]
    for i in range(n_nuc):
        nuc = np.where(region_nuc_labeled == i, 1, 0)
        nuc_fade = filters.gaussian(nuc.astype(float), (3, 3, 1))

``skimage.filters.gaussian`` = ``scipy.ndimage.gaussian_filter(mode='nearest', truncate=4)`` (``ref_spots.gaussian``).
``morphology.label`` (full connectivity by default) is scipy's ``label`` with the all-ones structure.
"""
import numpy as np
from scipy import ndimage as ndi

from .ref_spots import gaussian


def label_nuclei(region_nuc):
    """(region_nuc_labeled shifted so that bg = -1, n_nuc): fish_analysis_noco.py:227-229."""
    lab, n = ndi.label(np.asarray(region_nuc) > 0, structure=np.ones((3, 3, 3), bool))
    return lab.astype(np.int64) - 1, int(n)


def nucleus_fade(region_nuc_labeled, i, sigma=(3, 3, 1)):
    """fish_analysis_noco.py:233-235 for one nucleus: full-volume float64."""
    nuc = np.where(region_nuc_labeled == i, 1, 0)
    return gaussian(nuc.astype(float), sigma)
