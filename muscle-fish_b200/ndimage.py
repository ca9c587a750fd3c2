"""Replacements for the third-party calls the reference's scripts make directly on the hot
path (reference paths relative to the upstream repository root):

* ``distance_transform_edt(input, sampling)``  scipy.ndimage, called at
  nocodazole/fish_analysis_noco.py:218 and simulation/analyze_simulation.py:101
* ``spot_distances``   the interpn gather of nocodazole/fish_analysis_noco.py:219-221,297-298
* ``gaussian``         skimage.filters.gaussian (modules/muscle_fish.py:166,239,357)
* ``gaussian_laplace`` scipy.ndimage.gaussian_laplace (inside skimage.feature.blob_log)
* ``blob_log``         skimage.feature.blob_log (modules/muscle_fish.py:333,661;
                       modules/scope_utils3.py:403)

A script switches over by changing its import line, e.g.
``from muscle_fish_b200.ndimage import distance_transform_edt``.
Inputs/outputs are host numpy arrays in the reference's (x, y, z) order, float64 out.
"""
from __future__ import annotations

import numpy as np

from . import device as D


def _sampling3(sampling):
    if sampling is None:
        return (1.0, 1.0, 1.0)
    if np.isscalar(sampling):
        return (float(sampling),) * 3
    s = [float(v) for v in sampling]
    if len(s) != 3:
        raise RuntimeError("sampling must have one entry per axis")
    return tuple(s)


def distance_transform_edt(input, sampling=None, return_distances=True, return_indices=False):
    """Exact Euclidean distance (float64, input's (x,y,z) shape) from every voxel to the
    nearest zero voxel.  3-D inputs only; ``return_indices`` is not on the hot path."""
    if return_indices or not return_distances:
        raise NotImplementedError("only return_distances=True is implemented on the device")
    a = np.asarray(input)
    if a.ndim != 3:
        raise NotImplementedError("the device EDT is 3-D (the reference's 2-D uses are off the hot path)")
    mask = D.ingest(a, as_mask=True)
    dist = D.edt(mask, _sampling3(sampling))
    return D.export_f64(dist).cpu().numpy()


def spot_distances(region_nuc, spots, dims_xyz):
    """Distance (float64, physical units) of each spot to the nearest nucleus voxel: the
    ``distance_transform_edt(np.logical_not(region_nuc), sampling=dims)`` + ``interpn`` pair of
    nocodazole/fish_analysis_noco.py:218-221,297-298 without the dense map leaving the GPU."""
    a = np.asarray(region_nuc)
    (tn, en), = D.host_to_device_many([a], kinds=("mask",), pack_masks=True)  # a large mask crosses PCIe as bits
    not_nuc = D.ingest(tn, as_mask=True, negate=True, ready=en)
    dist = D.edt(not_nuc, _sampling3(dims_xyz))
    return D.gather(dist, np.asarray(spots)).astype(np.float64)


def gaussian(image, sigma, mode="nearest"):
    """skimage.filters.gaussian for a float image: gaussian_filter(mode='nearest', truncate=4)."""
    a = np.asarray(image)
    vol = D.ingest(a, want_minmax=False)
    return D.export_f64(D.gaussian(vol, sigma, mode=mode)).cpu().numpy()


def gaussian_laplace(image, sigma):
    """scipy.ndimage.gaussian_laplace(image, sigma) (mode='reflect', truncate=4)."""
    a = np.asarray(image)
    vol = D.ingest(a, want_minmax=False)
    out = D.log_volume(vol, float(sigma))  # = -sigma^2 * LoG
    res = D.export_f64(out).cpu().numpy()
    return res / (-(float(sigma) ** 2))


def blob_log(image, min_sigma=1, max_sigma=50, num_sigma=10, threshold=.2, overlap=.5):
    """skimage.feature.blob_log for a 3-D float image (linear sigma ladder, scalar sigmas)."""
    vol = D.ingest_as_float(np.asarray(image))
    return D.blob_log(vol, min_sigma, max_sigma, num_sigma, threshold, overlap)
