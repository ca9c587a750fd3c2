"""Seeded synthetic HCR-FISH stacks (C1..C5 geometries).

The reference ships no sample images, so every parity test and benchmark
runs on stacks made here.  ``make_stack`` is the numpy (PCG64) generator used
for tests, golden fixtures and the CPU-baseline sample; ``make_stack_device``
draws the same distribution with torch on the GPU for the full-size bench
configurations (a 2048x2048x100 stack takes too long to draw on the host).

Layout follows the reference: logical (x, y, z), C-contiguous, z fastest
(``modules/muscle_fish.py:60-71`` -- ``open_image`` returns (c, x, y, z)).
"""
from __future__ import annotations

import numpy as np

# name -> (shape_xyz, n_spots, n_nuclei, seed)
CONFIGS = {
    "C1": ((1024, 1024, 60), 5000, 30, 1234),
    "C2": ((2048, 2048, 100), 20000, 120, 2),
    "C5": ((2048, 2048, 100), 500000, 120, 5),
    "tiny": ((96, 80, 24), 40, 3, 7),
    "small": ((256, 192, 40), 300, 6, 11),
}

SAMPLING_ANISO = (0.0577, 0.0577, 0.5)  # um per voxel (x, y, z), SURVEY 8a-a7
SPOT_SIGMA_XY = 1.6
SPOT_SIGMA_Z = 0.9
NUC_SEMI = (45.0, 20.0, 7.0)


def _fiber_bounds(shape):
    X, Y, Z = shape
    y0, y1 = int(round(Y * 200 / 1024)), int(round(Y * 824 / 1024))
    z0, z1 = max(1, int(round(Z * 5 / 60))), min(Z - 1, int(round(Z * 55 / 60)))
    return y0, y1, z0, z1


def make_stack(shape=(1024, 1024, 60), n_spots=5000, n_nuclei=30, seed=1234,
               nuc_semi=NUC_SEMI, dtype=np.float32):
    """Return dict(img, fiber_mask, nuc_mask, centers) for one channel.

    background N(100, 10); fibre slab (+60) over all x, y in [0.195Y, 0.805Y),
    z in [Z/12, 11Z/12); spots: anisotropic Gaussians (sigma_xy 1.6, sigma_z
    0.9 vox, amplitude U(300, 1500)) at sub-voxel centres inside the fibre;
    nuclei: ellipsoids with semi-axes ``nuc_semi`` centred inside the fibre.
    """
    X, Y, Z = shape
    rng = np.random.default_rng(seed)
    img = rng.standard_normal(size=shape, dtype=np.float32) * np.float32(10.0) + np.float32(100.0)
    y0, y1, z0, z1 = _fiber_bounds(shape)
    fiber = np.zeros(shape, dtype=np.uint8)
    fiber[:, y0:y1, z0:z1] = 1
    img[:, y0:y1, z0:z1] += np.float32(60.0)

    mxy = min(10, max(2, (y1 - y0) // 8))
    mz = min(3, max(1, (z1 - z0) // 8))
    cx = rng.uniform(mxy, X - mxy, n_spots)
    cy = rng.uniform(y0 + mxy, y1 - mxy, n_spots)
    cz = rng.uniform(z0 + mz, z1 - mz, n_spots)
    amp = rng.uniform(300.0, 1500.0, n_spots)
    hx, hz = 6, 3
    ox = np.arange(-hx, hx + 1)
    oz = np.arange(-hz, hz + 1)
    for i in range(n_spots):
        ix, iy, iz = int(round(cx[i])), int(round(cy[i])), int(round(cz[i]))
        xs, ys, zs = ix + ox, iy + ox, iz + oz
        gx = np.exp(-0.5 * ((xs - cx[i]) / SPOT_SIGMA_XY) ** 2)
        gy = np.exp(-0.5 * ((ys - cy[i]) / SPOT_SIGMA_XY) ** 2)
        gz = np.exp(-0.5 * ((zs - cz[i]) / SPOT_SIGMA_Z) ** 2)
        vx = (xs >= 0) & (xs < X)
        vy = (ys >= 0) & (ys < Y)
        vz = (zs >= 0) & (zs < Z)
        patch = (amp[i] * gx[vx, None, None] * gy[None, vy, None] * gz[None, None, vz])
        img[xs[vx][0]:xs[vx][-1] + 1, ys[vy][0]:ys[vy][-1] + 1, zs[vz][0]:zs[vz][-1] + 1] += \
            patch.astype(np.float32)

    nuc = np.zeros(shape, dtype=np.uint8)
    a, b, c = nuc_semi
    a = min(a, X / 4.0)
    b = min(b, (y1 - y0) / 4.0)
    c = min(c, (z1 - z0) / 3.0)
    ncx = rng.uniform(0, X, n_nuclei)
    ncy = rng.uniform(y0, y1, n_nuclei)
    ncz = rng.uniform(z0, z1, n_nuclei)
    for i in range(n_nuclei):
        xa, xb = max(0, int(ncx[i] - a) - 1), min(X, int(ncx[i] + a) + 2)
        ya, yb = max(0, int(ncy[i] - b) - 1), min(Y, int(ncy[i] + b) + 2)
        za, zb = max(0, int(ncz[i] - c) - 1), min(Z, int(ncz[i] + c) + 2)
        gx = ((np.arange(xa, xb) - ncx[i]) / a) ** 2
        gy = ((np.arange(ya, yb) - ncy[i]) / b) ** 2
        gz = ((np.arange(za, zb) - ncz[i]) / c) ** 2
        inside = (gx[:, None, None] + gy[None, :, None] + gz[None, None, :]) <= 1.0
        nuc[xa:xb, ya:yb, za:zb] |= inside.astype(np.uint8)

    return {
        "img": img.astype(dtype, copy=False),
        "fiber_mask": fiber,
        "nuc_mask": nuc,
        "centers": np.stack([cx, cy, cz], axis=1),
        "shape": shape,
    }


def make_config(name, **over):
    shape, n_spots, n_nuc, seed = CONFIGS[name]
    kw = dict(shape=shape, n_spots=n_spots, n_nuclei=n_nuc, seed=seed)
    kw.update(over)
    return make_stack(**kw)


def make_threshold_stack(shape=(72, 72, 24)):
    """Noise-free float64 stack whose blob count first falls, then rises, then falls with the blob_log
    threshold -- the shape ``su.optimize_threshold_blob_log`` (modules/scope_utils3.py:388-422) looks
    for.  Three broad dim blobs (sigma 5), each with six bright small spots 4-6.5 voxels from its centre
    (inside its blob sphere: pruned for as long as the broad blob is detected), plus five isolated weak
    spots that drop out one after the other.  With sigmas (1, 3, 5) and thresholds 0, 0.01 ... 0.2 the
    counts are 8 8 6 4 4 3 3 3 3 3 3 8 13 18 18 ... and the chosen threshold is 0.02."""
    X, Y, Z = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    img = np.zeros(shape)

    def g(cx, cy, cz, a, s):
        return a * np.exp(-((X - cx) ** 2 + (Y - cy) ** 2 + (Z - cz) ** 2) / (2 * s * s))

    for (cx, cy, amp) in [(18, 18, 9.0), (52, 20, 8.0), (20, 52, 7.0)]:
        img += g(cx, cy, 12, amp, 5.0)
        for k, (dx, dy, dz) in enumerate([(5, 0, 0), (-5, 1, 0), (0, 5, 1), (1, -5, -1), (3, 3, 2), (-3, -4, -2)]):
            img += g(cx + dx, cy + dy, 12 + dz, 40.0 + 3 * k, 1.0)
    for k, a in enumerate([1.2, 1.8, 2.3, 2.8, 5.5]):
        img += g(44 + 6 * (k % 3), 44 + 8 * (k // 3), 6 + 3 * k, a, 1.0)
    return img


def make_stack_device(shape, n_spots, n_nuclei, seed, device="cuda", z_first=True, z_range=None, x_range=None):
    """Same distribution drawn on the GPU with torch (bench inputs only).

    ``z_range=(z0, z1)`` / ``x_range=(x0, x1)`` draw only those planes / x-rows of the stack (one slab of a
    mosaic): spots,
    fibre and nuclei are the global ones and the noise is seeded per global plane, so the slabs of any
    decomposition are exactly the planes of the whole stack.

    Returns float32 image and uint8 masks as torch tensors on ``device`` in
    (z, x, y) order when ``z_first`` else (x, y, z).  Spot/nucleus centres
    come from the same numpy PCG64 streams as ``make_stack``; only the
    background noise differs (torch Philox instead of PCG64).
    """
    import torch

    X, Y, Z = shape
    za, zb = (0, Z) if z_range is None else (int(z_range[0]), int(z_range[1]))
    xa, xb_ = (0, X) if x_range is None else (int(x_range[0]), int(x_range[1]))
    ZL, XL = zb - za, xb_ - xa
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=device)
    img = torch.empty((ZL, XL, Y), dtype=torch.float32, device=device)
    plane = torch.empty((X, Y), dtype=torch.float32, device=device) if XL != X else None
    for z in range(za, zb):  # one noise stream per GLOBAL plane: a slab of any decomposition draws the same voxels
        g.manual_seed((int(seed) * 1000003 + z * 7919) & 0x7FFFFFFFFFFFFFFF)
        if plane is None:
            img[z - za].normal_(100.0, 10.0, generator=g)
        else:
            plane.normal_(100.0, 10.0, generator=g)
            img[z - za] = plane[xa:xb_]
    del plane
    y0, y1, z0, z1 = _fiber_bounds(shape)
    fiber = torch.zeros((ZL, XL, Y), dtype=torch.uint8, device=device)
    fz0, fz1 = max(z0, za) - za, min(z1, zb) - za
    if fz1 > fz0:
        fiber[fz0:fz1, :, y0:y1] = 1
        img[fz0:fz1, :, y0:y1] += 60.0

    mxy = min(10, max(2, (y1 - y0) // 8))
    mz = min(3, max(1, (z1 - z0) // 8))
    cx = rng.uniform(mxy, X - mxy, n_spots)
    cy = rng.uniform(y0 + mxy, y1 - mxy, n_spots)
    cz = rng.uniform(z0 + mz, z1 - mz, n_spots)
    amp = rng.uniform(300.0, 1500.0, n_spots)
    hx, hz = 6, 3
    near = ((cz >= za - hz - 1) & (cz <= zb + hz + 1)  # spots that can touch this slab
            & (cx >= xa - hx - 1) & (cx <= xb_ + hx + 1))
    cx, cy, cz, amp = cx[near], cy[near], cz[near], amp[near]
    n_spots = int(near.sum())
    # spot patches are accumulated as 2^-16 fixed-point integers: integer atomics commute, so the voxels do
    # not depend on the order the patches land in, nor on which slab of a decomposition draws them
    acc = torch.zeros(ZL * XL * Y, dtype=torch.int64, device=device)
    ox = torch.arange(-hx, hx + 1, device=device)
    oz = torch.arange(-hz, hz + 1, device=device)
    chunk = 16384
    for s in range(0, n_spots, chunk):
        tcx = torch.as_tensor(cx[s:s + chunk], device=device)
        tcy = torch.as_tensor(cy[s:s + chunk], device=device)
        tcz = torch.as_tensor(cz[s:s + chunk], device=device)
        tam = torch.as_tensor(amp[s:s + chunk], device=device)
        ix, iy, iz = tcx.round().long(), tcy.round().long(), tcz.round().long()
        xs = ix[:, None] + ox[None, :]
        ys = iy[:, None] + ox[None, :]
        zs = iz[:, None] + oz[None, :]
        gx = torch.exp(-0.5 * ((xs - tcx[:, None]) / SPOT_SIGMA_XY) ** 2)
        gy = torch.exp(-0.5 * ((ys - tcy[:, None]) / SPOT_SIGMA_XY) ** 2)
        gz = torch.exp(-0.5 * ((zs - tcz[:, None]) / SPOT_SIGMA_Z) ** 2)
        gx = gx * ((xs >= xa) & (xs < xb_))
        gy = gy * ((ys >= 0) & (ys < Y))
        gz = gz * ((zs >= za) & (zs < zb))
        val = torch.round(tam[:, None, None, None] * gz[:, :, None, None] * gx[:, None, :, None]
                          * gy[:, None, None, :] * 65536.0).to(torch.int64)
        lin = (((zs - za).clamp(0, ZL - 1)[:, :, None, None] * XL + (xs - xa).clamp(0, XL - 1)[:, None, :, None]) * Y
               + ys.clamp(0, Y - 1)[:, None, None, :])
        acc.index_add_(0, lin.reshape(-1), val.reshape(-1))
    img += (acc.to(torch.float64) * (1.0 / 65536.0)).to(torch.float32).view(ZL, XL, Y)
    del acc

    nuc = torch.zeros((ZL, XL, Y), dtype=torch.uint8, device=device)
    a, b, c = NUC_SEMI
    a = min(a, X / 4.0)
    b = min(b, (y1 - y0) / 4.0)
    c = min(c, (z1 - z0) / 3.0)
    ncx = rng.uniform(0, X, n_nuclei)
    ncy = rng.uniform(y0, y1, n_nuclei)
    ncz = rng.uniform(z0, z1, n_nuclei)
    for i in range(n_nuclei):
        xlo, xhi = max(xa, int(ncx[i] - a) - 1), min(xb_, int(ncx[i] + a) + 2)
        ya, yb = max(0, int(ncy[i] - b) - 1), min(Y, int(ncy[i] + b) + 2)
        na, nb = max(za, int(ncz[i] - c) - 1), min(zb, int(ncz[i] + c) + 2)
        if nb <= na or xhi <= xlo:
            continue
        gx = ((torch.arange(xlo, xhi, device=device) - ncx[i]) / a) ** 2
        gy = ((torch.arange(ya, yb, device=device) - ncy[i]) / b) ** 2
        gz = ((torch.arange(na, nb, device=device) - ncz[i]) / c) ** 2
        inside = (gz[:, None, None] + gx[None, :, None] + gy[None, None, :]) <= 1.0
        nuc[na - za:nb - za, xlo - xa:xhi - xa, ya:yb] |= inside.to(torch.uint8)

    if not z_first:
        img = img.permute(1, 2, 0).contiguous()
        fiber = fiber.permute(1, 2, 0).contiguous()
        nuc = nuc.permute(1, 2, 0).contiguous()
    return {"img": img, "fiber_mask": fiber, "nuc_mask": nuc, "shape": shape}
