"""Oracle (TEST INFRASTRUCTURE ONLY): fibre and nucleus segmentation.

Follows modules/muscle_fish.py:114-198 (threshold_nuclei) and :201-278 (threshold_fiber) line by line.  The
scikit-image pieces (not installable here) are restated from the published algorithms on top of the REAL
numpy / scipy routines they call:

* ``filters.gaussian``                = scipy.ndimage.gaussian_filter(mode='nearest', truncate=4)
* ``filters.threshold_otsu(image)``   = np.histogram(image.ravel(), 256) -> maximise the between-class variance
                                        over the bin centres (unchanged across releases)
* ``filters.threshold_li(image)``     = Li & Tam's iterative minimum cross entropy as shipped up to 0.15
                                        (tolerance half a 256th of the range, start at the mean); 0.16+ start
                                        the same iteration from the mean but stop at half the smallest gap
                                        between distinct values -- **parity unpinned**, like the other skimage
                                        restatements (oracle/__init__.py)
* ``morphology.label``                = scipy.ndimage.label with the full 3x3x3 structure (connectivity = ndim)
* ``morphology.remove_small_objects / remove_small_holes`` = components of connectivity 1 smaller than min_size
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

from .ref_spots import gaussian, normalize_image


def threshold_otsu(image, nbins=256):
    hist, bin_edges = np.histogram(np.asarray(image).ravel(), nbins)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.
    hist = hist.astype(float)
    weight1 = np.cumsum(hist)
    weight2 = np.cumsum(hist[::-1])[::-1]
    mean1 = np.cumsum(hist * bin_centers) / weight1
    mean2 = (np.cumsum((hist * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = np.argmax(variance12)
    return bin_centers[:-1][idx]


def threshold_li(image):
    image = np.array(image, dtype=np.float64, copy=True)
    immin = np.min(image)
    image -= immin
    imrange = np.max(image)
    tolerance = 0.5 * imrange / 256
    mean = np.mean(image)
    new_thresh = mean
    old_thresh = new_thresh + 2 * tolerance
    threshold = old_thresh
    while abs(new_thresh - old_thresh) > tolerance:
        old_thresh = new_thresh
        threshold = old_thresh + tolerance
        mean_back = image[image <= threshold].mean()
        mean_obj = image[image > threshold].mean()
        temp = (mean_back - mean_obj) / (np.log(mean_back) - np.log(mean_obj))
        if temp < 0:
            new_thresh = temp - tolerance
        else:
            new_thresh = temp + tolerance
    return threshold + immin


def label(mask):
    return ndi.label(mask, structure=np.ones((3,) * np.ndim(mask), dtype=bool))


def remove_small_objects(ar, min_size):
    lab, _ = ndi.label(ar)
    sizes = np.bincount(lab.ravel())
    too_small = sizes < min_size
    too_small[0] = False
    out = ar.copy()
    out[too_small[lab]] = False
    return out


def remove_small_holes(ar, area_threshold):
    return np.logical_not(remove_small_objects(np.logical_not(ar), area_threshold))


def threshold_fiber(img_fish, t_fiber=None, z_first=False, return_threshold=False):
    """modules/muscle_fish.py:201-278"""
    img_fish = np.asarray(img_fish)
    if z_first:
        img_fish = img_fish.transpose(1, 2, 0)
    img_avg = normalize_image(img_fish)
    img_blur = gaussian(img_avg, (10, 10, 1))
    if t_fiber is None:
        thresh = threshold_li(img_blur)
    elif type(t_fiber) == float:
        thresh = t_fiber
    elif t_fiber == 'last':
        thresh = threshold_li(img_blur[:, :, -1])
    else:
        raise TypeError('`t_fiber` argument not recognized.')
    bin_fiber = np.where(img_blur > thresh, 1, 0)
    if t_fiber == 'last':
        for z in range(bin_fiber.shape[2]):
            bin_fiber[:, :, z] = bin_fiber[:, :, -1]
    fiber_labeled, n_labels = label(bin_fiber)
    label_sizes = {i: np.count_nonzero(np.where(fiber_labeled == i, 1, 0)) for i in range(1, n_labels + 1)}
    fiber_idx = max(label_sizes, key=label_sizes.get)
    mask = np.where(fiber_labeled == fiber_idx, 1, 0)
    if z_first:
        mask = mask.transpose(2, 0, 1)
    return (mask, thresh, img_blur) if return_threshold else mask


def threshold_nuclei(img_dapi, t_dapi=None, fiber_mask=None, labeled=True, multiplier=0.75, z_first=False,
                     return_threshold=False):
    """modules/muscle_fish.py:114-198"""
    sig = (2, 20, 20) if z_first else (20, 20, 2)
    img_dapi = normalize_image(img_dapi)
    img_blur = gaussian(img_dapi, sig)
    if t_dapi is None:
        thresh = threshold_otsu(img_blur) * multiplier
    elif type(t_dapi) == float:
        thresh = t_dapi
    else:
        raise TypeError('`t_dapi` argument not recognized.')
    bin_dapi = np.where(img_dapi > thresh, 1, 0)
    bin_dapi = remove_small_objects(bin_dapi.astype(bool), 2048)
    bin_dapi = remove_small_holes(bin_dapi.astype(bool), 2048)
    nuclei_labeled, n_nuc = label(bin_dapi)
    if fiber_mask is not None:
        for i in range(1, n_nuc + 1):
            overlap = np.logical_and(np.where(nuclei_labeled == i, True, False), np.asarray(fiber_mask).astype(bool))
            if np.count_nonzero(overlap) == 0:
                nuclei_labeled[nuclei_labeled == i] = 0
    out = nuclei_labeled if labeled else np.where(nuclei_labeled > 0, 1, 0)
    return (out, thresh, img_blur) if return_threshold else out
