"""Drop-in host API for the segmentation functions of ``modules/muscle_fish.py`` (reference paths relative to
the upstream repository root):

* ``threshold_fiber``   modules/muscle_fish.py:201-278   normalise -> gaussian (10,10,1) -> Li -> largest object
* ``threshold_nuclei``  modules/muscle_fish.py:114-198   normalise -> gaussian (20,20,2) -> Otsu x 0.75 -> clean-up,
                                                         labelling, nuclei outside the fibre dropped

The voxel work -- normalisation, the two large-radius blurs (81 / 161 taps per axis), the 256-bin histogram,
the iterated means of Li's method, the binarisation, and (round 2) the connected-component labelling with the
small-object / small-hole clean-up built on it -- runs on the GPU: ``label`` = ``skimage.morphology.label`` /
``scipy.ndimage.label`` with the same numbering (components in raster order of their first voxel), union-find
kernels in ``csrc/ccl.cu``.

skimage version notes (the reference pins none; API usage implies 0.14-0.17): ``threshold_li`` is the
iterative form those releases ship up to 0.15 (tolerance = half a 256th of the range, start at the mean);
``threshold_otsu`` is unchanged across releases.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
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


def label_device(mask: torch.Tensor, full: bool = True):
    """Connected components of a C-ordered 2-D / 3-D uint8 / bool device mask, in the order the mask is laid out:
    ``(labels int32, n)`` with the numbering of ``scipy.ndimage.label`` / ``skimage.morphology.label`` (components
    ranked by their first voxel in raster order).  ``full``: 8- / 26-neighbourhood (skimage's default for
    ``label``), else 4- / 6-neighbourhood (``connectivity=1``: what remove_small_objects / _holes use on bools)."""
    m = mask.contiguous()
    if m.dtype == torch.bool:
        m = m.view(torch.uint8)
    if m.dim() not in (2, 3) or m.dtype != torch.uint8:
        raise ValueError("label_device: 2-D or 3-D uint8 / bool mask expected")
    dims = (1,) + tuple(m.shape) if m.dim() == 2 else tuple(m.shape)
    n = m.numel()
    parent = torch.empty(n, dtype=torch.int32, device=m.device)
    if n == 0:
        return parent.view(m.shape), 0
    L.check(L.lib().mfish_ccl_roots_u8(m.data_ptr(), dims[0], dims[1], dims[2], 1 if full else 0, parent.data_ptr(),
                                       D._stream()))
    # every foreground voxel now holds its component's smallest raster index; rank the roots, gather the rank
    roots = torch.nonzero(parent == torch.arange(n, dtype=torch.int32, device=m.device)).view(-1)
    ncomp = int(roots.numel())
    rank = torch.zeros(n + 1, dtype=torch.int32, device=m.device)  # slot n: the background's "root"
    rank[roots] = torch.arange(1, ncomp + 1, dtype=torch.int32, device=m.device)
    labels = rank[torch.where(parent >= 0, parent, n).long()]
    return labels.view(m.shape), ncomp


def label_sizes(labels: torch.Tensor, ncomp: int) -> torch.Tensor:
    """voxel count per label 0 .. ncomp (np.bincount(labels.ravel())), int64 on the device"""
    lab = labels.contiguous()
    sizes = torch.empty(ncomp + 1, dtype=torch.int64, device=lab.device)
    L.check(L.lib().mfish_label_sizes_i32(lab.data_ptr(), lab.numel(), ncomp + 1, sizes.data_ptr(), D._stream()))
    return sizes


def remove_small_objects_device(mask: torch.Tensor, min_size: int) -> torch.Tensor:
    """skimage.morphology.remove_small_objects(bool array, min_size) (connectivity 1) -> bool device mask"""
    lab, n = label_device(mask, full=False)
    keep = label_sizes(lab, n) >= int(min_size)
    keep[0] = False
    return keep[lab.long()]


def remove_small_holes_device(mask: torch.Tensor, area_threshold: int) -> torch.Tensor:
    """skimage.morphology.remove_small_holes(bool array, area_threshold) (connectivity 1)"""
    m = mask if mask.dtype == torch.bool else mask != 0
    return ~remove_small_objects_device(~m, area_threshold)


def remove_small_objects(ar, min_size=64, connectivity=1):
    """``skimage.morphology.remove_small_objects`` for a host BOOL array (modules/muscle_fish.py:182)."""
    a = np.asarray(ar)
    if a.dtype != np.bool_ or connectivity != 1:
        raise NotImplementedError("remove_small_objects: bool arrays with connectivity 1 (what the reference passes)")
    return remove_small_objects_device(D.host_to_device(np.ascontiguousarray(a)), min_size).cpu().numpy()


def remove_small_holes(ar, area_threshold=64, connectivity=1):
    """``skimage.morphology.remove_small_holes`` for a host BOOL array (modules/muscle_fish.py:183)."""
    a = np.asarray(ar)
    if a.dtype != np.bool_ or connectivity != 1:
        raise NotImplementedError("remove_small_holes: bool arrays with connectivity 1 (what the reference passes)")
    return remove_small_holes_device(D.host_to_device(np.ascontiguousarray(a)), area_threshold).cpu().numpy()


def label(image, connectivity=None, return_num=False):
    """``skimage.morphology.label`` for a host array (the import-line seam of
    nocodazole/fish_analysis_noco.py:227: ``morphology.label(region_nuc, return_num=True)``): labels in the
    array's own raster order, int64 like skimage's."""
    a = np.asarray(image)
    if connectivity not in (None, 1, a.ndim):
        raise NotImplementedError("label: connectivity 1 or the full neighbourhood")
    m = D.host_to_device(np.ascontiguousarray(a != 0))
    lab, n = label_device(m, full=connectivity != 1 or a.ndim == 1)
    out = lab.cpu().numpy().astype(np.int64)
    return (out, n) if return_num else out


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
    bin_xyz = bin_d.permute(1, 2, 0).contiguous()  # (x, y, z), as the reference labels it
    fiber_labeled, n_labels = label_device(bin_xyz, full=True)
    if n_labels == 0:
        raise ValueError("max() arg is an empty sequence")  # what max(label_sizes, ...) raises upstream
    sizes = label_sizes(fiber_labeled, n_labels)
    sizes[0] = 0
    biggest = int(torch.argmax(sizes).item())  # the first of equally large ones, as np.argmax / max() upstream
    mask_d = (fiber_labeled == biggest).to(torch.uint8)  # (x, y, z)
    mask_zxy = mask_d.permute(2, 0, 1).contiguous()
    out = (mask_zxy if z_first else mask_d).cpu().numpy().astype(np.int64)  # np.where(..., 1, 0)
    # the fibre mask is the `mask=` of every later call (general_pipeline/fish_analysis.py:152-199): keep its
    # device twin, so those calls find it resident instead of uploading 8 bytes per voxel again
    return D.MASK_HANDOFF.register(out, mask_zxy, z_first)


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
    bin_d = D.binarize(vol.data, thresh, vol.minmax)  # [z][x][y]
    bin_dapi = (bin_d if z_first else bin_d.permute(1, 2, 0).contiguous()) != 0  # the caller's own axis order
    bin_dapi = remove_small_objects_device(bin_dapi, 2048)
    bin_dapi = remove_small_holes_device(bin_dapi, 2048)
    nuclei_labeled, n_nuc = label_device(bin_dapi, full=True)
    if fiber_mask is not None:
        if verbose:
            print('Removing nuclei that are not connected to main fiber segment...')
        fm = D.MASK_HANDOFF.lookup(fiber_mask, z_first)
        if fm is not None:
            fm = fm if z_first else fm.permute(1, 2, 0)  # [z][x][y] -> the caller's axis order
        else:
            (fm, ev), = D.host_to_device_many([fiber_mask], kinds=("mask",))
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)
        if tuple(fm.shape) != tuple(nuclei_labeled.shape):
            raise ValueError("operands could not be broadcast together: fiber_mask and image shapes differ")
        inside = torch.bincount(nuclei_labeled[fm != 0].view(-1), minlength=n_nuc + 1)
        drop = inside == 0
        drop[0] = False
        nuclei_labeled = torch.where(drop[nuclei_labeled.long()], torch.zeros_like(nuclei_labeled), nuclei_labeled)
    if labeled:
        return nuclei_labeled.cpu().numpy().astype(np.int64)
    return (nuclei_labeled > 0).cpu().numpy().astype(np.int64)
