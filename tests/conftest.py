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
