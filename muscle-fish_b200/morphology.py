"""Replacements for the ``skimage.morphology`` calls that build the reference's compartment masks
(general_pipeline/fish_analysis.py:169-187, nocodazole/fish_analysis_noco.py:193-208,
rna_decay/fish_analysis_actd.py:165-183; reference paths relative to the upstream repository root):

* ``binary_erosion(image)`` / ``binary_dilation(image)``  one step with the default cross structuring
  element, 2-D or 3-D -- the import-line seam (``from muscle_fish_b200 import morphology``): the scripts'
  loops run unchanged, one GPU call per iteration.
* ``iterate(image, op, n3d, n2d)``  the whole loop nest (n3d 3-D steps, then n2d steps on every z plane)
  in one call: n steps with the cross are one step with the L1 ball of radius n, so the 50 sweeps of the
  reference are one capped city-block distance transform on the device.
* ``label`` / ``remove_small_objects`` / ``remove_small_holes``  connected components on the device, with
  skimage's numbering (components in raster order of their first voxel).
* ``compartments(nuclei_binary, fiber_mask)``  region_nuc / region_peri / region_cyt exactly as
  fish_analysis.py:169-186 defines them, from one upload of the two masks.

Inputs are host arrays in the reference's (x, y, z) order; outputs are host arrays of the reference's
dtypes (bool for the skimage functions, int64 0/1 for the ``np.where(..., 1, 0)`` regions).
"""
from __future__ import annotations

import numpy as np

from . import device as D


def _mask_to_device(image):
    """host 2-D [nx][ny] or 3-D (x,y,z) array -> uint8 device mask [nz][nx][ny] (nonzero = set)"""
    a = np.asarray(image)
    if a.ndim == 3:
        return D.ingest_mask(image if isinstance(image, np.ndarray) else a), a.shape
    if a.ndim == 2:
        import torch
        t = D.host_to_device(a)
        return (t != 0).to(torch.uint8).reshape((1,) + tuple(a.shape)), a.shape
    raise NotImplementedError("binary morphology on the device is 2-D / 3-D")


def _mask_to_host(m_zxy, shape, dtype):
    if len(shape) == 2:
        return m_zxy.reshape(shape).cpu().numpy().astype(dtype, copy=False)
    return m_zxy.permute(1, 2, 0).contiguous().cpu().numpy().astype(dtype, copy=False)


def _step(image, op, selem, out):
    if selem is not None:
        raise NotImplementedError("only the default structuring element (the cross, connectivity 1) is implemented")
    m, shape = _mask_to_device(image)
    res = D.binary_morph(m, op, 1 if len(shape) == 3 else 0, 0 if len(shape) == 3 else 1)
    res = _mask_to_host(res, shape, bool)
    if out is not None:
        out[...] = res
        return out
    return res


def binary_erosion(image, selem=None, out=None):
    """skimage.morphology.binary_erosion(image): cross structuring element, border_value=True."""
    return _step(image, "erode", selem, out)


def binary_dilation(image, selem=None, out=None):
    """skimage.morphology.binary_dilation(image): cross structuring element."""
    return _step(image, "dilate", selem, out)


def iterate(image, op, n3d, n2d):
    """``n3d`` 3-D steps followed by ``n2d`` steps on every z plane (op 'erode' | 'dilate') -> bool (x,y,z)."""
    m, shape = _mask_to_device(image)
    if len(shape) != 3:
        raise ValueError("iterate() takes a 3-D (x, y, z) mask")
    return _mask_to_host(D.binary_morph(m, op, n3d, n2d), shape, bool)


def compartments(nuclei_binary, fiber_mask, erode=(1, 10), dilate=(4, 40), dtype=np.int64):
    """The three compartments of general_pipeline/fish_analysis.py:169-186 as
    ``{'nuc': region_nuc, 'cyt': region_cyt, 'peri': region_peri}`` (0/1 arrays of ``dtype``; the reference's
    ``np.where(..., 1, 0)`` gives int64)."""
    nuc = D.ingest(np.asarray(nuclei_binary), as_mask=True)
    fib = D.ingest(np.asarray(fiber_mask), as_mask=True)
    r_nuc, r_peri, r_cyt = D.compartments(nuc, fib, erode, dilate)
    shape = np.asarray(nuclei_binary).shape
    return {"nuc": _mask_to_host(r_nuc, shape, dtype), "cyt": _mask_to_host(r_cyt, shape, dtype),
            "peri": _mask_to_host(r_peri, shape, dtype)}


# ``morphology.label`` / ``remove_small_objects`` / ``remove_small_holes`` (modules/muscle_fish.py:182-184, 266-272,
# nocodazole/fish_analysis_noco.py:227): connected components on the device (csrc/ccl.cu), same numbering as skimage
def label(image, connectivity=None, return_num=False):
    from . import segmentation
    return segmentation.label(image, connectivity=connectivity, return_num=return_num)


def remove_small_objects(ar, min_size=64, connectivity=1):
    from . import segmentation
    return segmentation.remove_small_objects(ar, min_size, connectivity)


def remove_small_holes(ar, area_threshold=64, connectivity=1):
    from . import segmentation
    return segmentation.remove_small_holes(ar, area_threshold, connectivity)
