"""Oracle (TEST INFRASTRUCTURE ONLY): compartment masks.

Follows general_pipeline/fish_analysis.py:165-187 (same code at nocodazole/fish_analysis_noco.py:189-208
and rna_decay/fish_analysis_actd.py:161-183), literally -- including the Python loop over z planes.
``skimage.morphology.binary_erosion / binary_dilation`` (not installable here) are thin wrappers:

    binary_erosion(image, selem=None)  = scipy.ndimage.binary_erosion(image, structure=selem, border_value=True)
    binary_dilation(image, selem=None) = scipy.ndimage.binary_dilation(image, structure=selem)

with ``structure=None`` meaning scipy's ``generate_binary_structure(ndim, 1)``, the cross -- the "cross-shaped
structuring element (connectivity=1)" of skimage's documentation.  The scipy calls are the real dependency.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


def binary_erosion(image, selem=None):
    return ndi.binary_erosion(np.asarray(image), structure=selem, border_value=True)


def binary_dilation(image, selem=None):
    return ndi.binary_dilation(np.asarray(image), structure=selem)


def iterate(image, op, n3d, n2d):
    fn = binary_erosion if op == "erode" else binary_dilation
    region = np.asarray(image)
    for _ in range(n3d):
        region = fn(region)
    region = np.array(region)
    for _ in range(n2d):
        for z in range(region.shape[2]):
            region[:, :, z] = fn(region[:, :, z])
    return region.astype(bool)


def compartments(nuclei_labeled, img_fiber_only, erode=(1, 10), dilate=(4, 40)):
    """general_pipeline/fish_analysis.py:166-186."""
    nuclei_labeled = np.asarray(nuclei_labeled)
    img_fiber_only = np.asarray(img_fiber_only)
    nuclei_binary = np.where(nuclei_labeled > 0, 1, 0)
    region_nuc = np.where(nuclei_labeled > 0, 1, 0)
    for n in range(erode[0]):
        region_nuc = binary_erosion(region_nuc)
    for n in range(erode[1]):
        for z in range(region_nuc.shape[2]):
            region_nuc[:, :, z] = binary_erosion(region_nuc[:, :, z])
    img_nuc_dilated = np.where(nuclei_labeled > 0, 1, 0)
    for n in range(dilate[0]):
        img_nuc_dilated = binary_dilation(img_nuc_dilated)
    for n in range(dilate[1]):
        for z in range(img_nuc_dilated.shape[2]):
            img_nuc_dilated[:, :, z] = binary_dilation(img_nuc_dilated[:, :, z])
    region_peri = np.where(np.logical_and(img_nuc_dilated > region_nuc,
                                          np.logical_or(img_fiber_only.astype(bool), nuclei_binary.astype(bool))), 1, 0)
    region_cyt = np.where(img_fiber_only > img_nuc_dilated, 1, 0)
    return {"nuc": np.asarray(region_nuc).astype(np.int64), "cyt": region_cyt, "peri": region_peri}
