"""Drop-in host API for the spot path of ``modules/muscle_fish.py`` (reference paths are
relative to the upstream repository root).

Same names, argument meaning, return types and error behaviour as the reference:

* ``find_spots``            modules/muscle_fish.py:631-672
* ``find_spots_snrfilter``  modules/muscle_fish.py:281-402
* ``fix_bleaching``         modules/muscle_fish.py:405-512

Inputs are host numpy arrays in the reference's (x, y, z) order; results are host float64
arrays.  All voxel work runs on the GPU through ``device`` (no CPU fallback).  ``draw=True``
is accepted for signature compatibility and ignored (plotting is outside the hot path).
"""
from __future__ import annotations

import numpy as np

from . import device as D
from . import scope_utils3 as su  # noqa: F401  (reference modules import it the same way)


class _Lazy:
    """A mask whose host->device copy is already queued; converted on first use, so that the
    filters on the image overlap the mask's trip over PCIe."""

    def __init__(self, tensor, event, z_first, shape, ready=None):
        self.tensor, self.event, self.z_first, self.shape, self._m = tensor, event, z_first, shape, ready
        if ready is not None and tuple(ready.shape) != tuple(shape):
            raise IndexError("mask and image shapes differ: %r vs %r" % (tuple(ready.shape), tuple(shape)))

    def get(self):
        if self._m is None:
            self._m = D.ingest(self.tensor, z_first=self.z_first, as_mask=True, ready=self.event)
            if tuple(self._m.shape) != tuple(self.shape):
                raise IndexError("mask and image shapes differ: %r vs %r" % (tuple(self._m.shape), tuple(self.shape)))
            self.tensor = None
        return self._m


def _prep(img_fish, mask, z_first):
    vol = D.HANDOFF.lookup(img_fish, z_first)  # an array this package returned: already on the device
    mdev = D.MASK_HANDOFF.lookup(mask, z_first)  # the same for a mask (threshold_fiber's result)
    up_mask = None if mdev is not None else mask
    if vol is not None:
        (tm, em), = D.host_to_device_many([up_mask], kinds=("mask",), pack_masks=True)
    else:
        (ti, ei), (tm, em) = D.host_to_device_many([img_fish, up_mask], kinds=("image", "mask"), pack_masks=True)
        vol = D.ingest(ti, z_first=z_first, ready=ei)
    m = _Lazy(tm, em, z_first, vol.data.shape, ready=mdev) if mask is not None else None
    return vol, m


def _row_stack(rows):
    # np.row_stack([]) raises "need at least one array to concatenate" (ValueError), which is what
    # callers of the reference see when no spot survives (modules/muscle_fish.py:341, 402)
    if len(rows) == 0:
        raise ValueError("need at least one array to concatenate")
    return np.vstack(rows)


def _nonempty(rows):
    if len(rows) == 0:
        return _row_stack([])
    return rows


def find_spots(img_fish, sigma=1., t_spot=0.025, mask=None):
    """Find FISH spots with blob_log; optionally keep only spots inside ``mask``.
    Returns an (N, 4) float64 array of rows [x, y, z, sigma]."""
    vol, m = _prep(img_fish, mask, False)
    spots = D.blob_log(vol, sigma, sigma, 1, t_spot)
    if mask is not None:
        if len(spots) == 0:
            return _row_stack([])
        return _nonempty(spots[D.gather(m.get(), spots) != 0])
    return spots


def find_spots_snrfilter(img_fish, snr=None, sigma=1., t_spot=0.025, mask=None, z_first=False, draw=False,
                         imgprefix='fiber'):
    """blob_log spots, kept when inside ``mask`` and when their blurred intensity over the
    fibre baseline exceeds an (automatic) signal-to-noise threshold.

    Returns ``(spots, spot_data)``: (N, 4) float64 rows [x, y, z, sigma] ([z, x, y, sigma] when
    ``z_first``) and the per-spot table ``[['spot_id','x','y','z','intensity','SNR'], ...]``.
    """
    vol, m = _prep(img_fish, mask, z_first)
    pending = {}

    def queue_baseline():  # P25 of the fibre voxels does not depend on the spots: queue it early
        if m is not None:
            pending["baseline"] = D.masked_percentile_async(vol, m.get(), 25)

    spots = D.blob_log(vol, sigma, sigma, 1, t_spot, before_sync=queue_baseline)

    if mask is not None:
        if len(spots) == 0:
            return _row_stack([]), None
        spots_masked = _nonempty(spots[D.gather(m.get(), spots) != 0])
    else:
        spots_masked = spots

    if m is None:
        # np.logical_not(None) masks every voxel in the reference, so the percentile below is
        # taken over an empty selection and fails (modules/muscle_fish.py:355-356)
        raise IndexError("find_spots_snrfilter needs a mask: percentile of an empty selection")
    spot_int = D.blur_at_points(vol, spots_masked, (3, 3, 0.3), mode="nearest")
    baseline = pending["baseline"]()

    spots_filtered, spot_data, _ = _snr_outputs(spots_masked, spot_int, baseline, snr, z_first)
    return _nonempty(spots_filtered), spot_data


def _snr_outputs(spots_masked, spot_int, baseline, snr, z_first):
    """modules/muscle_fish.py:367-402: automatic SNR threshold, its print, the filtered spots, the
    per-spot table of ALL mask-passing spots -- and which of them were kept."""
    if snr is None:
        thresh = np.percentile(spot_int, 90) / (baseline * np.sqrt(10))
        snr_min = np.sqrt(10)
        snr = thresh if thresh > snr_min else snr_min

    print('SNR threshold = ' + str(snr))

    # the reference's per-spot loop (modules/muscle_fish.py:382-400), vectorised
    ratio = spot_int / baseline
    keep = ratio > snr
    spots_filtered = spots_masked[keep]
    if z_first:
        spots_filtered = spots_filtered[:, [2, 0, 1, 3]]
    spot_data = np.column_stack([np.zeros(len(spot_int)), spots_masked[:, 0:3], spot_int, ratio]).tolist()
    for c, r in enumerate(spot_data):
        r[0] = c  # integer spot id, as the reference's enumerate()
    spot_data.insert(0, ['spot_id', 'x', 'y', 'z', 'intensity', 'SNR'])
    return spots_filtered, spot_data, keep


def find_spots_snrfilter_with_distances(img_fish, region_nuc, dims_xyz, snr=None, sigma=1., t_spot=0.025, mask=None,
                                        z_first=False, draw=False, imgprefix='fiber'):
    """``find_spots_snrfilter(img_fish, snr, sigma, t_spot, mask, z_first)`` AND the distance of every
    kept spot to the nearest nucleus voxel -- the ``distance_transform_edt(np.logical_not(region_nuc),
    sampling=dims_xyz)`` + ``interpn`` pair of nocodazole/fish_analysis_noco.py:218-221,297-298 -- from one
    call, so that the three uploads are queued back to back and the distance transform runs while the
    image is still crossing PCIe (``device.analyze_host``).  Returns ``(spots, spot_data, distances)``:
    the first two exactly as ``find_spots_snrfilter`` returns them, ``distances`` float64 per kept spot."""
    if mask is None:
        raise IndexError("find_spots_snrfilter needs a mask: percentile of an empty selection")
    from .ndimage import _sampling3
    # every mask-passing spot comes back with its distance (snr = -inf keeps all); the SNR rule is applied
    # here, with the reference's own arithmetic
    res = D.analyze_host(img_fish, mask, region_nuc, sigma=sigma, t_spot=t_spot, sampling_xyz=_sampling3(dims_xyz),
                         snr=-np.inf, z_first=z_first)
    if len(res["spots_masked"]) == 0:
        return _row_stack([]), None, np.empty(0)
    spots_filtered, spot_data, keep = _snr_outputs(res["spots_masked"], res["intensity"], res["baseline"], snr,
                                                   z_first)
    return _nonempty(spots_filtered), spot_data, res["spot_dist"][keep]


def fix_bleaching(img, second_half=True, mask=None, z_first=False, draw=False, imgprefix='fiber'):
    """Exponential photobleaching correction along z (per-plane masked means on the GPU,
    the two-parameter fit on the host, the per-plane rescale on the GPU)."""
    from scipy import optimize

    def expo_fit(x, A, k):
        return A * np.exp(k * x)

    img = np.asarray(img)
    vol = D.HANDOFF.lookup(img, z_first) or D.ingest(img, z_first=z_first, want_minmax=False)
    m = D.ingest_mask(mask, z_first=z_first) if mask is not None else None
    nz = vol.data.shape[0]
    start_frame = nz // 2 if second_half else 0
    sums, cnts = D.plane_means(vol.data, m)
    with np.errstate(invalid="ignore", divide="ignore"):
        means = sums / cnts
    intensities = list(means[start_frame:])
    total_mean = float(np.sum(D.plane_means(vol.data, None)[0]) / vol.data.numel()) if m is not None \
        else float(np.sum(sums) / vol.data.numel())
    try:
        e_params, _ = optimize.curve_fit(expo_fit, list(range(start_frame, nz)), intensities,
                                         p0=[total_mean, -0.1])
    except RuntimeError:
        print('WARNING: Bleaching correction optimization failed. Continuing without correction...')
        return img
    if e_params[1] > 0:
        print('WARNING: Bleaching correction optimization failed. Continuing without correction...')
        return img
    correction = np.array([expo_fit(z, 1, -1. * e_params[1]) for z in range(nz)])
    out = D.scale_planes(vol.data, correction)
    res = D.export_f64(out, z_first=z_first).cpu().numpy()
    return D.HANDOFF.register(res, D.Volume(out, D.minmax(out)), z_first)
