import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_small.npz")
    return np.load(path, allow_pickle=False)


def spot_set(spots):
    """set of integer (x, y, z[, sigma*1000]) tuples of an (N, >=3) spot array."""
    s = np.asarray(spots)
    if s.size == 0:
        return set()
    cols = [s[:, 0].astype(np.int64), s[:, 1].astype(np.int64), s[:, 2].astype(np.int64)]
    if s.shape[1] > 3:
        cols.append(np.round(s[:, 3] * 1000).astype(np.int64))
    return set(zip(*[c.tolist() for c in cols]))


def match_fraction(a, b):
    """|A & B| / max(|A|, |B|) for two spot arrays."""
    sa, sb = spot_set(a), spot_set(b)
    if not sa and not sb:
        return 1.0
    return len(sa & sb) / max(len(sa), len(sb))


def boundary_safe_peaks(raw_blobs, crop_xy, edge_unsafe, overlap=0.5):
    """Which raw peaks of a CROP have a blob_log outcome that provably does not depend on the crop's
    artificial x/y faces: a peak is unsafe if it lies within ``edge_unsafe`` voxels of such a face
    (filter radius + NMS: its response or its very existence may differ from the whole volume), and so
    is every peak connected to an unsafe one through a chain of blob pairs overlapping by more than
    ``overlap`` -- only such pairs make ``_prune_blobs`` zero anything, so only they carry a change
    along.  ``raw_blobs`` rows [x, y, z, sigma]; ``crop_xy`` = ((nx, ny), ((x_lo_cut, x_hi_cut),
    (y_lo_cut, y_hi_cut))): is that face a cut (True) or a real volume face (False).
    Returns a set of integer (x, y, z) tuples."""
    import math
    import oracle
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    b = np.asarray(raw_blobs, dtype=np.float64)
    n = len(b)
    if n == 0:
        return set()
    (nx, ny), ((xl, xh), (yl, yh)) = crop_xy[0], crop_xy[1]
    unsafe = np.zeros(n, bool)
    if xl:
        unsafe |= b[:, 0] < edge_unsafe
    if xh:
        unsafe |= b[:, 0] >= nx - edge_unsafe
    if yl:
        unsafe |= b[:, 1] < edge_unsafe
    if yh:
        unsafe |= b[:, 1] >= ny - edge_unsafe
    cand = cKDTree(b[:, :3]).query_pairs(2 * b[:, 3].max() * math.sqrt(3.0), output_type="ndarray")
    edges = [(i, j) for i, j in cand if oracle.blob_overlap(b[i], b[j]) > overlap]
    if edges:
        e = np.asarray(edges)
        g = coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n))
        _, lab = connected_components(g, directed=False)
        unsafe = np.isin(lab, np.unique(lab[unsafe]))
    pts = b[~unsafe, :3].astype(np.int64)
    return set(map(tuple, pts.tolist()))
