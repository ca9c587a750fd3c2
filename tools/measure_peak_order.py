"""CPU measurement (oracle against oracle): how much does the peak ORDER that skimage's peak_local_max
returns change blob_log's pruned survivors?  'response' = skimage >= 0.18 (descending intensity),
'reverse_raster' = skimage 0.14-0.17 (the reference's era).  Crops at the spot densities of the
BASELINE configs; also 'set' (upstream Python set iteration) against 'sorted' pair order.

    python tools/measure_peak_order.py            # prints one line per case, JSON at the end
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
import muscle_fish_b200  # noqa: F401,E402
from muscle_fish_b200 import synth  # noqa: E402


def sset(a):
    return {tuple(r) for r in np.round(np.asarray(a) * 1000).astype(np.int64).tolist()}


def case(name, density_cfg, shape, sig, thr):
    full_shape, n_spots, n_nuc, seed = synth.CONFIGS[density_cfg]
    frac = (shape[0] * shape[1] * shape[2]) / float(np.prod(full_shape))
    st = synth.make_stack(shape=shape, n_spots=max(8, int(round(n_spots * frac))), n_nuclei=1, seed=seed)
    img = oracle.normalize_image(st["img"])
    res = {}
    for po in ("response", "reverse_raster"):
        for pair in ("sorted", "set"):
            pruned, raw = oracle.blob_log(img, sig[0], sig[1], sig[2], thr, pair_order=pair, peak_order=po,
                                          return_raw=True)
            res[(po, pair)] = sset(pruned)
            nraw = len(raw)
    a, b = res[("response", "sorted")], res[("reverse_raster", "sorted")]
    out = {
        "case": name, "crop": list(shape), "raw_peaks": nraw, "survivors_response": len(a),
        "survivors_reverse_raster": len(b), "pruned": nraw - len(a),
        "match_response_vs_reverse_raster": len(a & b) / max(len(a), len(b), 1),
        "match_set_vs_sorted_response": len(res[("response", "set")] & a) / max(len(a), 1),
        "match_set_vs_sorted_reverse_raster": len(res[("reverse_raster", "set")] & b) / max(len(b), 1),
    }
    print(out, flush=True)
    return out


if __name__ == "__main__":
    rows = [
        case("C1 density, sigma 2", "C1", (384, 384, 60), (2.0, 2.0, 1), 0.025),
        case("C2 density, sigma 2", "C2", (384, 384, 100), (2.0, 2.0, 1), 0.025),
        case("C5 density, sigma 2", "C5", (256, 256, 100), (2.0, 2.0, 1), 0.025),
        case("C5 density, sigma 1..3 x5", "C5", (192, 192, 100), (1.0, 3.0, 5), 0.02),
    ]
    print(json.dumps(rows))
