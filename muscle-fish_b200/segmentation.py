"""Drop-in host API for the segmentation functions of ``modules/muscle_fish.py`` (reference paths relative to
the upstream repository root):

* ``threshold_fiber``   modules/muscle_fish.py:201-278   normalise -> gaussian (10,10,1) -> Li -> largest object
* ``threshold_nuclei``  modules/muscle_fish.py:114-198   normalise -> gaussian (20,20,2) -> Otsu x 0.75 -> clean-up,
                                                         labelling, nuclei outside the fibre dropped

The voxel work -- normalisation, the two large-radius blurs (81 / 161 taps per axis), the 256-bin histogram,
the iterated means of Li's method, the binarisation -- runs on the GPU.  Connected-component labelling and
the small-object / small-hole clean-up stay on the host (scipy.ndimage.label), as SURVEY 8f-2 plans.

skimage version notes (the reference pins none; API usage implies 0.14-0.17): ``threshold_li`` is the
iterative form those releases ship up to 0.15 (tolerance = half a 256th of the range, start at the mean);
``threshold_otsu`` is unchanged across releases.
"""
from __future__ import annotations

import numpy as np

from . import device as D


def threshold_otsu_from_hist(hist, bin_edges):
    """skimage.filters.threshold_otsu on the (hist, bin_edges) of np.histogram(image, 256)."""
    hist = hist.astype(float)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.
    weight1 = np.cumsum(hist)
    weight2 = np.cumsum(hist[::-1])[::-1]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean1 = np.cumsum(hist * bin_centers) / weight1
        mean2 = (np.cumsum((hist * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    idx = np.argmax(variance12)
    return bin_centers[:-1][idx]


def threshold_otsu(vol_zxy, nbins=256):
    hist, edges = D.histogram(vol_zxy, nbins)
    return threshold_otsu_from_hist(hist, edges)


def threshold_li(vol_zxy):
    """skimage.filters.threshold_li (iterative minimum cross entropy, releases <= 0.15) on a device volume."""
    lo, hi = [float(v) for v in D.minmax(vol_zxy).tolist()]
    if lo == hi:
        return lo
    immin = lo
    imrange = hi - lo
    tolerance = 0.5 * imrange / 256
    (c0, s0), (c1, s1) = D.split_moments(vol_zxy, hi)  # everything: the mean of the shifted image
    mean = s0 / c0 - immin
    new_thresh = mean
    old_thresh = new_thresh + 2 * tolerance
    threshold = old_thresh
    while abs(new_thresh - old_thresh) > tolerance:
        old_thresh = new_thresh
        threshold = old_thresh + tolerance
        (c0, s0), (c1, s1) = D.split_moments(vol_zxy, threshold + immin)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean_back = s0 / c0 - immin
            mean_obj = s1 / c1 - immin
            temp = (mean_back - mean_obj) / (np.log(mean_back) - np.log(mean_obj))
        if temp < 0:
            new_thresh = temp - tolerance
        else:
            new_thresh = temp + tolerance
    return threshold + immin


def _label_full(mask):
    from scipy import ndimage as ndi
    return ndi.label(mask, structure=np.ones((3,) * mask.ndim, dtype=bool))  # skimage label: connectivity = ndim


def _remove_small_objects(ar, min_size):
    """skimage.morphology.remove_small_objects(bool array, min_size, connectivity=1)"""
    from scipy import ndimage as ndi
    lab, _ = ndi.label(ar)  # cross: connectivity 1
    sizes = np.bincount(lab.ravel())
    too_small = sizes < min_size
    too_small[0] = False
    out = ar.copy()
    out[too_small[lab]] = False
    return out


def _remove_small_holes(ar, area_threshold):
    return np.logical_not(_remove_small_objects(np.logical_not(ar), area_threshold))


def _host_mask(m_zxy, z_first):
    t = m_zxy if z_first else m_zxy.permute(1, 2, 0).contiguous()
    return t.cpu().numpy()


def threshold_fiber(img_fish, t_fiber=None, z_first=False, verbose=False):
    """Segment the muscle fibre from a 3-D FISH image: 0/1 int64 mask of its largest bright object."""
    if verbose:
        print('Segmenting fiber...')
    if t_fiber is not None and type(t_fiber) != float and t_fiber != 'last':
        raise TypeError('`t_fiber` argument not recognized. \\\n            Must be either float or "last".')
    vol = D.HANDOFF.lookup(img_fish, z_first) or D.ingest(np.asarray(img_fish), z_first=z_first)
    blur = D.gaussian(vol, (10, 10, 1), mode="nearest")  # of normalize_image(img): mapped on load
    if t_fiber is None:
        thresh = threshold_li(blur)
    elif type(t_fiber) == float:
        thresh = t_fiber
    else:
        thresh = threshold_li(blur[-1].contiguous())
    if verbose:
        print('Fiber threshold: ' + str(thresh))
    bin_d = D.binarize(blur, thresh)
    if t_fiber == 'last':
        bin_d = bin_d[-1:].expand_as(bin_d).contiguous()
        if verbose:
            print('OVERRIDE: using final frame mask for entire z-stack.')
    bin_fiber = _host_mask(bin_d, False)  # (x, y, z), as the reference labels it
    fiber_labeled, n_labels = _label_full(bin_fiber)
    sizes = np.bincount(fiber_labeled.ravel())
    sizes[0] = 0
    if n_labels == 0:
        raise ValueError("max() arg is an empty sequence")  # what max(label_sizes, ...) raises upstream
    mask = np.where(fiber_labeled == int(np.argmax(sizes)), 1, 0)
    if z_first:
        mask = mask.transpose(2, 0, 1)
    return mask


def threshold_nuclei(img_dapi, t_dapi=None, fiber_mask=None, labeled=True, multiplier=0.75, z_first=False,
                     verbose=False):
    """Segment nuclei in a 3-D DAPI image: labelled int mask (or 0/1 with ``labeled=False``)."""
    if verbose:
        print('\nSegmenting nuclei...')
    if t_dapi is not None and type(t_dapi) != float:
        raise TypeError('`t_dapi` argument not recognized. \\\n            Must be either float or None.')
    vol = D.HANDOFF.lookup(img_dapi, z_first) or D.ingest(np.asarray(img_dapi), z_first=z_first)
    if t_dapi is None:
        blur = D.gaussian(vol, (20, 20, 2), mode="nearest")
        thresh = threshold_otsu(blur) * multiplier
    else:
        thresh = t_dapi
    if verbose:
        print('DAPI threshold = ' + str(thresh))
    # the NORMALISED, unblurred image is what gets binarised (modules/muscle_fish.py:181)
    bin_dapi = _host_mask(D.binarize(vol.data, thresh, vol.minmax), z_first).astype(bool)
    bin_dapi = _remove_small_objects(bin_dapi, 2048)
    bin_dapi = _remove_small_holes(bin_dapi, 2048)
    nuclei_labeled, n_nuc = _label_full(bin_dapi)
    if fiber_mask is not None:
        if verbose:
            print('Removing nuclei that are not connected to main fiber segment...')
        inside = np.bincount(nuclei_labeled[np.asarray(fiber_mask).astype(bool)].ravel(), minlength=n_nuc + 1)
        drop = inside == 0
        drop[0] = False
        nuclei_labeled[drop[nuclei_labeled]] = 0
    if labeled:
        return nuclei_labeled
    return np.where(nuclei_labeled > 0, 1, 0)
