"""``skimage.filters.gaussian`` as the nuclear-mesh step uses it, on the device.

``nocodazole/fish_analysis_noco.py:233-235`` blurs the one-hot mask of every nucleus over the WHOLE volume::

    nuc = np.where(region_nuc_labeled == i, 1, 0)
    nuc_fade = filters.gaussian(nuc.astype(float), (3, 3, 1))

* ``gaussian(image, sigma)`` -- the same signature; any volume (the generic separable passes).
* ``nucleus_fades(region_nuc_labeled, n_nuc, sigma)`` -- all nuclei in one batch: the label volume goes to the
  device once, the bounding boxes come from one sweep, and every label is filtered inside its own box grown by the
  filter radius (outside of it the reference's result is exactly 0.0).  ``NucleusFades[i]`` is the full-volume
  float64 array the reference's loop body computes; ``.crop(i)`` is the non-zero part and its position, which is
  what ``measure.marching_cubes`` needs.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import device as D
from .taps import gaussian_radius


def gaussian(image, sigma, mode="nearest"):
    """skimage.filters.gaussian(image, sigma) (mode 'nearest', truncate 4) -> float64, same shape."""
    from . import ndimage
    return ndimage.gaussian(image, sigma, mode=mode)


class NucleusFades:
    """Result of :func:`nucleus_fades`: ``len()`` nuclei; ``[i]`` full volume, ``crop(i)`` the non-zero box."""

    def __init__(self, shape_xyz, boxes_xyz, offsets, data):
        self.shape = tuple(shape_xyz)
        self.boxes = boxes_xyz      # [n][6] x0, x1, y0, y1, z0, z1 of the grown boxes; x1 <= x0: label absent
        self._offsets = offsets
        self._data = data           # concatenated crops, device layout [z][x][y], float32 (host)

    def __len__(self):
        return len(self.boxes)

    def crop(self, i):
        """(slices, float64 array in (x, y, z) order) of nucleus i; (None, None) if the label does not occur."""
        x0, x1, y0, y1, z0, z1 = (int(v) for v in self.boxes[i])
        if x1 <= x0:
            return None, None
        flat = self._data[self._offsets[i]:self._offsets[i + 1]]
        c = flat.reshape(z1 - z0, x1 - x0, y1 - y0).transpose(1, 2, 0).astype(np.float64)
        return (slice(x0, x1), slice(y0, y1), slice(z0, z1)), c

    def __getitem__(self, i):
        out = np.zeros(self.shape, np.float64)
        sl, c = self.crop(i)
        if sl is not None:
            out[sl] = c
        return out

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def nucleus_fades(region_nuc_labeled, n_nuc=None, sigma=(3, 3, 1), first_label=0) -> NucleusFades:
    """filters.gaussian((region_nuc_labeled == i).astype(float), sigma) for i = first_label .. first_label+n_nuc-1.

    ``region_nuc_labeled``: integer (x, y, z) array, background below ``first_label`` (the reference shifts
    ``morphology.label``'s output by -1, so nuclei are 0 .. n_nuc-1 and the background is -1)."""
    D._require_cuda()
    lab = np.asarray(region_nuc_labeled)
    if lab.ndim != 3:
        raise ValueError("nucleus_fades: 3-D label volume expected")
    if n_nuc is None:
        n_nuc = int(lab.max()) - first_label + 1 if lab.size else 0
    n_nuc = max(int(n_nuc), 0)
    if np.isscalar(sigma):
        sigma = (sigma,) * 3
    sx, sy, sz = (float(s) for s in sigma)
    nx, ny, nz = lab.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    if n_nuc == 0:
        return NucleusFades(lab.shape, np.zeros((0, 6), np.int64), np.zeros(1, np.int64), np.zeros(0, np.float32))
    lab_dev = D.host_to_device(np.ascontiguousarray(lab.astype(np.int32, copy=False)))
    lab_zxy = lab_dev.permute(2, 0, 1).contiguous()
    boxes = torch.empty((n_nuc, 6), dtype=torch.int32, device=dev)
    lib = L.lib()
    st = D._stream()
    L.check(lib.mfish_label_bbox_i32(lab_zxy.data_ptr(), nz, nx, ny, int(first_label), n_nuc, boxes.data_ptr(), st))
    b = boxes.cpu().numpy().astype(np.int64)          # z0, z1, x0, x1, y0, y1 (tight)
    present = b[:, 1] > b[:, 0]
    rz, rx, ry = gaussian_radius(sz), gaussian_radius(sx), gaussian_radius(sy)
    g = b.copy()
    g[:, 0] = np.maximum(b[:, 0] - rz, 0)
    g[:, 1] = np.minimum(b[:, 1] + rz, nz)
    g[:, 2] = np.maximum(b[:, 2] - rx, 0)
    g[:, 3] = np.minimum(b[:, 3] + rx, nx)
    g[:, 4] = np.maximum(b[:, 4] - ry, 0)
    g[:, 5] = np.minimum(b[:, 5] + ry, ny)
    g[~present] = 0
    sizes = (g[:, 1] - g[:, 0]) * (g[:, 3] - g[:, 2]) * (g[:, 5] - g[:, 4])
    offsets = np.zeros(n_nuc + 1, np.int64)
    np.cumsum(sizes, out=offsets[1:])
    total = int(offsets[-1])
    idx = np.nonzero(present)[0]
    crops = np.zeros((len(idx), 8), np.int32)
    crops[:, :6] = g[idx]
    crops[:, 6] = idx + int(first_label)
    out = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    if total:
        tmp = torch.empty(total, dtype=torch.float32, device=dev)
        crops_dev = torch.as_tensor(crops).to(dev)
        offs_dev = torch.as_tensor(np.ascontiguousarray(offsets[np.append(idx, n_nuc)])).to(dev)
        sig = L.darray([sz, sx, sy])
        L.check(lib.mfish_onehot_gaussian_f32(lab_zxy.data_ptr(), nz, nx, ny, crops_dev.data_ptr(),
                                              offs_dev.data_ptr(), len(idx), total, sig, out.data_ptr(),
                                              tmp.data_ptr(), st))
    data = out[:total].cpu().numpy()
    boxes_xyz = np.stack([g[:, 2], g[:, 3], g[:, 4], g[:, 5], g[:, 0], g[:, 1]], axis=1)
    return NucleusFades(lab.shape, boxes_xyz, offsets, data)
