"""Drop-in host API for the spot-path functions of ``modules/scope_utils3.py``
(reference paths relative to the upstream repository root):

* ``normalize_image``              modules/scope_utils3.py:381-386
* ``optimize_threshold_blob_log``  modules/scope_utils3.py:388-422
* ``gauss_1d`` / ``gauss_2d`` / ``gauss_3d``  modules/scope_utils3.py:427-440 (pure formulas)
"""
from __future__ import annotations

import numpy as np

from . import device as D


def normalize_image(image, vmin=0, vmax=1):
    """np.interp(image, (min, max), (vmin, vmax)): float64, same shape; the min/max reduction
    and the affine map run on the GPU with np.interp's exact arithmetic."""
    a = np.asarray(image)
    t = D.host_to_device(a)
    out = D.normalize_f64(t, vmin, vmax)
    res = out.cpu().numpy().reshape(a.shape)
    if a.ndim == 3 and D.HANDOFF.enabled:
        # keep what ingest(res) would upload again in the next call of the script (see device.Handoff)
        D.HANDOFF.register(res, D.ingest(out.reshape(a.shape)), z_first=False)
    return res


def optimize_threshold_blob_log(image, min_sigma=1., max_sigma=50., num_sigma=10, tmin=0., tmax=0.1, num=20,
                                plot_opt=False):
    """Threshold at the plateau inflection of (#spots vs threshold).  The reference recomputes
    the full multi-sigma LoG cube for each of the ``num`` thresholds; here the cube is computed
    once and only peak detection + prune are repeated per threshold."""
    from scipy.signal import argrelextrema
    from . import taps as T

    vol = D.ingest_as_float(np.asarray(image))  # blob_log applies img_as_float, nothing else
    sigmas = T.sigma_list(min_sigma, max_sigma, num_sigma)
    # the whole scale cube is resident (num_sigma float32 volumes + 3 scratch volumes): C2 with the default
    # 10 scales is 22 GB of the 180 GB; refuse loudly rather than thrash when it cannot fit
    import torch
    need = (len(sigmas) + 4) * vol.data.numel() * 4
    free, _ = torch.cuda.mem_get_info(vol.data.device)
    if need > free:
        raise MemoryError(f"optimize_threshold_blob_log: the {len(sigmas)}-scale cube needs {need / 2**30:.1f} GiB "
                          f"of device memory, {free / 2**30:.1f} GiB are free; lower num_sigma or crop the image")
    cube = D.log_cube(vol, sigmas)
    t_range = np.linspace(tmin, tmax, num=num)
    num_spots = [len(D.blob_log(vol, min_sigma, max_sigma, num_sigma, t, cube=cube)) for t in t_range]
    dn_dt = np.gradient(num_spots, t_range[1] - t_range[0])
    inflection_pts = argrelextrema(dn_dt, np.less)[0]
    spike = np.argmax(num_spots)
    t_opt = t_range[[pt for pt in inflection_pts if pt < spike][-1]]
    print('Optimal threshold value for blob_log: ' + str(t_opt))
    return t_opt


def gauss_1d(x, A, mu, sig):
    return A * np.exp(-1. * ((x - mu) ** 2.) / (2. * (sig ** 2.)))


def gauss_2d(x, y, x_off, y_off, mu, sigma):
    d = np.sqrt((x - x_off) ** 2. + (y - y_off) ** 2.)
    return np.exp(-1. * (d ** 2.) / (2. * (sigma ** 2.)))


def gauss_3d(pos, A, mu_x, mu_y, mu_z, sig_xy, sig_z, m_x, m_y, m_z, b):
    x, y, z = pos[:, 0], pos[:, 1], pos[:, 2]
    return A * np.exp(-1. * ((((x - mu_x) ** 2.) / (2. * (sig_xy ** 2.))) + (((y - mu_y) ** 2.) / (2. * (sig_xy ** 2.)))
                             + (((z - mu_z) ** 2.) / (2. * (sig_z ** 2.))))) + m_x * x + m_y * y + m_z * z + b
