"""Oracle (TEST INFRASTRUCTURE ONLY): float64 restatement of the spot path.

Every function cites the reference lines it follows (paths relative to the
upstream repository root).  scipy calls are the real dependency; the
scikit-image pieces are restated from the published algorithm because
scikit-image is not installable here (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage as ndi
from scipy import optimize
from scipy.spatial import cKDTree


# --------------------------------------------------------------------------- #
# modules/scope_utils3.py:381-386
# --------------------------------------------------------------------------- #
def normalize_image(image, vmin=0, vmax=1):
    """``np.interp(image, (min, max), (vmin, vmax))`` -- always float64."""
    image = np.asarray(image)
    return np.interp(image, (np.min(image), np.max(image)), (vmin, vmax))


# --------------------------------------------------------------------------- #
# skimage.filters.gaussian as called at modules/muscle_fish.py:166,239,357 and
# nocodazole/fish_analysis_noco.py:235: img_as_float (identity for float64)
# followed by scipy.ndimage.gaussian_filter(mode='nearest', truncate=4.0).
# --------------------------------------------------------------------------- #
def gaussian(image, sigma, mode="nearest", truncate=4.0):
    image = np.asarray(image, dtype=np.float64)
    return ndi.gaussian_filter(image, sigma, mode=mode, truncate=truncate)


# --------------------------------------------------------------------------- #
# skimage.feature.blob_log, step 2 (call sites modules/muscle_fish.py:333,661;
# modules/scope_utils3.py:403): scale-normalised negative LoG, one volume per
# sigma, stacked on a trailing scale axis.
# --------------------------------------------------------------------------- #
def sigma_list(min_sigma, max_sigma, num_sigma):
    scale = np.linspace(0, 1, num_sigma)
    return scale * (float(max_sigma) - float(min_sigma)) + float(min_sigma)


def log_cube(image, sigmas):
    image = np.asarray(image, dtype=np.float64)
    vols = [-ndi.gaussian_laplace(image, float(s)) * float(s) ** 2 for s in sigmas]
    return np.stack(vols, axis=-1)


# --------------------------------------------------------------------------- #
# skimage.feature.peak_local_max(cube, threshold_abs=t, threshold_rel=0.0,
# footprint=ones(3,...), exclude_border=False) as blob_log calls it.
#   mask = (cube == maximum_filter(cube, footprint, mode='constant'))
#          & (cube > max(t, threshold_rel * cube.max()))
# Ordering rule frozen here (documented divergence between skimage releases):
# stable descending intensity over the raster-ordered nonzero() coordinates
# (skimage >= 0.18).  skimage 0.14-0.17 returned reverse raster order; on
# tie-free data both give the same *set*.  ensure_spacing(min_distance=1,
# p=inf) in new releases rejects only points at distance < 1, i.e. nothing on
# integer coordinates, so plateau ties are all kept.
# --------------------------------------------------------------------------- #
def peak_local_max(cube, threshold_abs, threshold_rel=0.0, peak_order="response"):
    """``peak_order='response'``: skimage >= 0.18 (stable descending intensity over raster order).
    ``peak_order='reverse_raster'``: skimage 0.14-0.17, the reference's era -- with ``num_peaks=inf``
    ``_get_high_intensity_peaks`` returns ``np.column_stack(np.nonzero(mask))[::-1]``.  Same SET of peaks;
    the order decides which blob of an overlapping equal-sigma pair ``_prune_blobs`` zeroes (the lower
    index: the brighter one under 'response', the raster-later one under 'reverse_raster')."""
    cube = np.asarray(cube)
    fp = np.ones((3,) * cube.ndim, dtype=bool)
    mx = ndi.maximum_filter(cube, footprint=fp, mode="constant")
    mask = cube == mx
    if np.all(mask):  # trivial (constant) image -> no peaks
        mask[...] = False
    thr = threshold_abs
    if threshold_rel is not None:
        thr = max(thr, threshold_rel * cube.max())
    mask &= cube > thr
    coord = np.nonzero(mask)
    if peak_order == "reverse_raster":
        return np.transpose(coord)[::-1]
    if peak_order != "response":
        raise ValueError(peak_order)
    intens = cube[coord]
    order = np.argsort(-intens, kind="stable")
    return np.transpose(coord)[order]


# --------------------------------------------------------------------------- #
# skimage.feature.blob._blob_overlap / _compute_sphere_overlap /
# _compute_disk_overlap and _prune_blobs(blobs, overlap, sigma_dim=1).
# --------------------------------------------------------------------------- #
def _sphere_overlap(d, r1, r2):
    vol = (math.pi / (12 * d) * (r1 + r2 - d) ** 2
           * (d ** 2 + 2 * d * (r1 + r2) - 3 * (r1 - r2) ** 2))
    return vol / (4.0 / 3 * math.pi * min(r1, r2) ** 3)


def _disk_overlap(d, r1, r2):
    ratio1 = (d ** 2 + r1 ** 2 - r2 ** 2) / (2 * d * r1)
    ratio1 = min(max(ratio1, -1), 1)
    acos1 = math.acos(ratio1)
    ratio2 = (d ** 2 + r2 ** 2 - r1 ** 2) / (2 * d * r2)
    ratio2 = min(max(ratio2, -1), 1)
    acos2 = math.acos(ratio2)
    a, b, c, dd = -d + r2 + r1, d - r2 + r1, d + r2 - r1, d + r2 + r1
    area = r1 ** 2 * acos1 + r2 ** 2 * acos2 - 0.5 * math.sqrt(abs(a * b * c * dd))
    return area / (math.pi * min(r1, r2) ** 2)


def blob_overlap(blob1, blob2):
    """Overlap fraction of two blobs [coords..., sigma] (scalar sigma)."""
    ndim = len(blob1) - 1
    if ndim > 3:
        return 0.0
    root_ndim = math.sqrt(ndim)
    s1, s2 = blob1[-1], blob2[-1]
    if s1 == s2 == 0:
        return 0.0
    elif s1 > s2:
        max_sigma = s1
        r1, r2 = 1.0, s2 / s1
    else:
        max_sigma = s2
        r2, r1 = 1.0, s1 / s2
    pos1 = np.asarray(blob1[:ndim], dtype=float) / (max_sigma * root_ndim)
    pos2 = np.asarray(blob2[:ndim], dtype=float) / (max_sigma * root_ndim)
    d = math.sqrt(float(np.sum((pos2 - pos1) ** 2)))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    if ndim == 2:
        return _disk_overlap(d, r1, r2)
    return _sphere_overlap(d, r1, r2)


def prune_blobs(blobs, overlap=0.5, pair_order="set"):
    """skimage ``_prune_blobs``.

    ``pair_order='set'`` iterates the KD-tree pairs in Python ``set`` order
    exactly like the reference dependency does; ``'sorted'`` is the
    deterministic lexicographic order the CUDA path implements.  Survivor
    sets were identical on every synthetic case tried; the
    tests measure it.
    """
    blobs = np.array(blobs, dtype=np.float64, copy=True)
    if len(blobs) == 0:
        return blobs
    sigma = blobs[:, -1].max()
    distance = 2 * sigma * math.sqrt(blobs.shape[1] - 1)
    tree = cKDTree(blobs[:, :-1])
    pairs = tree.query_pairs(distance)
    if pair_order == "sorted":
        pairs = sorted(pairs)
    elif pair_order == "set":
        pairs = list(pairs)
    else:
        raise ValueError(pair_order)
    for (i, j) in pairs:
        b1, b2 = blobs[i], blobs[j]
        if blob_overlap(b1, b2) > overlap:
            if b1[-1] > b2[-1]:
                b2[-1] = 0
            else:
                b1[-1] = 0
    return blobs[blobs[:, -1] > 0]


def prune_cutoff_sq(sigma, ndim=3, overlap=0.5):
    """Largest integer squared voxel distance n for which two equal-sigma
    blobs on integer coordinates overlap by more than ``overlap``.
    (Known answers: sigma=1 -> 1, sigma=2 -> 5.)"""
    n = 0
    best = -1
    lim = int((2 * sigma * math.sqrt(ndim)) ** 2) + 2
    for n in range(1, lim + 1):
        b1 = [0.0] * ndim + [float(sigma)]
        b2 = [math.sqrt(n)] + [0.0] * (ndim - 1) + [float(sigma)]
        if blob_overlap(b1, b2) > overlap:
            best = n
    return best


# --------------------------------------------------------------------------- #
# skimage.feature.blob_log(image, min_sigma, max_sigma, num_sigma, threshold)
# --------------------------------------------------------------------------- #
def blob_log(image, min_sigma=1, max_sigma=50, num_sigma=10, threshold=0.2,
             overlap=0.5, pair_order="set", return_raw=False, peak_order="response"):
    image = np.asarray(image, dtype=np.float64)
    sigmas = sigma_list(min_sigma, max_sigma, num_sigma)
    cube = log_cube(image, sigmas)
    peaks = peak_local_max(cube, threshold_abs=threshold, threshold_rel=0.0, peak_order=peak_order)
    if peaks.size == 0:
        out = np.empty((0, 3))
        return (out, out) if return_raw else out
    lm = peaks.astype(np.float64)
    lm[:, -1] = sigmas[peaks[:, -1]]
    pruned = prune_blobs(lm, overlap, pair_order=pair_order)
    return (pruned, lm) if return_raw else pruned


# --------------------------------------------------------------------------- #
# modules/muscle_fish.py:631-672
# --------------------------------------------------------------------------- #
def find_spots(img_fish, sigma=1.0, t_spot=0.025, mask=None, pair_order="set", peak_order="response"):
    spots = blob_log(normalize_image(img_fish), sigma, sigma, num_sigma=1,
                     threshold=t_spot, pair_order=pair_order, peak_order=peak_order)
    if mask is not None:
        keep = [s for s in spots if mask[tuple(s[0:3].astype(int))]]
        return np.vstack(keep)  # np.row_stack([]) -> ValueError, as upstream
    return spots


# --------------------------------------------------------------------------- #
# modules/muscle_fish.py:281-402 (draw=False path)
# --------------------------------------------------------------------------- #
def find_spots_snrfilter(img_fish, snr=None, sigma=1.0, t_spot=0.025, mask=None,
                         z_first=False, pair_order="set", verbose=False, peak_order="response"):
    if z_first:
        img_fish = normalize_image(np.asarray(img_fish).transpose(1, 2, 0))
        if mask is not None:
            mask = np.asarray(mask).transpose(1, 2, 0)
    else:
        img_fish = normalize_image(img_fish)

    spots = blob_log(normalize_image(img_fish), sigma, sigma, num_sigma=1,
                     threshold=t_spot, pair_order=pair_order, peak_order=peak_order)
    if mask is not None:
        spots_masked = [s for s in spots if mask[tuple(s[0:3].astype(int))]]
        spots_masked = np.vstack(spots_masked)
    else:
        spots_masked = spots

    fiber_intensities = np.ma.masked_where(np.logical_not(mask), img_fish).compressed()
    baseline = np.percentile(fiber_intensities, 25)
    img_blur = gaussian(img_fish, (3, 3, 0.3))

    if snr is None:
        spot_int = [img_blur[tuple(s[0:3].astype(int))] for s in spots_masked]
        thresh = np.percentile(spot_int, 90) / (baseline * np.sqrt(10))
        snr_min = np.sqrt(10)
        snr = thresh if thresh > snr_min else snr_min
    if verbose:
        print("SNR threshold = " + str(snr))

    spots_filtered = []
    spot_data = [["spot_id", "x", "y", "z", "intensity", "SNR"]]
    for c, spot in enumerate(spots_masked):
        x_int, y_int, z_int = tuple(spot[0:3].astype(int))
        intensity = img_blur[x_int, y_int, z_int]
        if intensity / baseline > snr:
            spots_filtered.append(spot[[2, 0, 1, 3]] if z_first else spot)
        spot_data.append([c] + list(spot[0:3]) + [intensity, intensity / baseline])
    return np.vstack(spots_filtered), spot_data


# --------------------------------------------------------------------------- #
# modules/muscle_fish.py:405-512 (draw=False path)
# --------------------------------------------------------------------------- #
def fix_bleaching(img, second_half=True, mask=None, z_first=False, return_fit=False):
    def expo_fit(x, A, k):
        return A * np.exp(k * x)

    img = np.asarray(img)
    if z_first:
        img = img.transpose(1, 2, 0)
        if mask is not None:
            mask = np.asarray(mask).transpose(1, 2, 0)
    start = img.shape[2] // 2 if second_half else 0
    if mask is not None:
        m = np.asarray(mask).astype(bool)
        intens = [np.mean(img[:, :, z][m[:, :, z]]) for z in range(start, img.shape[2])]
    else:
        intens = [np.mean(img[:, :, z]) for z in range(start, img.shape[2])]

    def _ret(a, k):
        a = a.transpose(2, 0, 1) if z_first else a
        return (a, k, intens) if return_fit else a

    try:
        e_params, _ = optimize.curve_fit(expo_fit, list(range(start, img.shape[2])),
                                         intens, p0=[np.mean(img), -0.1])
    except RuntimeError:
        return _ret(img, None)
    if e_params[1] > 0:
        return _ret(img, None)
    corr = np.array([expo_fit(z, 1, -1.0 * e_params[1]) for z in range(img.shape[2])])
    return _ret(img * corr, e_params[1])


# --------------------------------------------------------------------------- #
# modules/scope_utils3.py:388-422: spot count per threshold (the image is used
# as given -- this function does not normalise).  Returns the counts too.
# --------------------------------------------------------------------------- #
def optimize_threshold_blob_log(image, min_sigma=1.0, max_sigma=50.0, num_sigma=10,
                                tmin=0.0, tmax=0.1, num=20, pair_order="set"):
    from scipy.signal import argrelextrema
    t_range = np.linspace(tmin, tmax, num=num)
    num_spots = [len(blob_log(image, min_sigma, max_sigma, num_sigma, threshold=t,
                              pair_order=pair_order)) for t in t_range]
    dn_dt = np.gradient(num_spots, t_range[1] - t_range[0])
    infl = argrelextrema(dn_dt, np.less)[0]
    spike = int(np.argmax(num_spots))
    cands = [pt for pt in infl if pt < spike]
    t_opt = t_range[cands[-1]]  # IndexError if no inflection, as upstream
    return t_opt, np.asarray(num_spots)
