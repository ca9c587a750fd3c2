"""ctypes binding of ``libmfish_sm100.so`` (the C ABI declared in ``include/mfish.h``).

There is NO fallback: if the shared library is missing the first use raises.
Build it with ``python __graft_entry__.py build`` (or ``make -C muscle-fish_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (MFISH_LIB: an alternative build of the library, for A/B timing of kernel variants -- tools/)
LIB_PATH = os.environ.get("MFISH_LIB") or os.path.join(_HERE, "libmfish_sm100.so")

# constants mirrored from include/mfish.h
ABI_VERSION = 6
OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
U8, U16, I16, I32, I64, F32, F64 = range(7)
LAYOUT_XYZ, LAYOUT_ZXY = 0, 1
DST_F32, DST_NONZERO, DST_ISZERO = 0, 1, 2
BOUNDARY_REFLECT, BOUNDARY_NEAREST = 0, 1
AXIS_Z, AXIS_X, AXIS_Y = 0, 1, 2
CONV_G, CONV_GD, CONV_MID, CONV_LAST = 0, 1, 2, 3
QUANTILE_WS_BYTES = 64 * 1024
EDT_F2_I32, EDT_F2_F32 = 0, 1
ORDER_RESPONSE, ORDER_REVERSE_RASTER = 0, 1
MORPH_DILATE, MORPH_ERODE = 0, 1


class MfishError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmfish_sm100 error {code}: {msg}")
        self.code = code


_p, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); every entry point of include/mfish.h
SIGNATURES = {
    "mfish_abi_version": (_i, []),
    "mfish_last_error": (C.c_char_p, []),
    "mfish_launch_count": (_i64, []),
    "mfish_profile_enable": (None, [_i]),
    "mfish_profile_collect": (_i64, [C.c_char_p, _i64]),
    "mfish_ingest": (_i, [_p, _i, _i, _i64, _i64, _i64, _p, _i, _p, _p]),
    "mfish_mask_bits_bytes": (_i64, [_i64]),
    "mfish_host_pack_mask": (_i, [_p, _i, _i64, _p, _i]),
    "mfish_ingest_mask_bits": (_i, [_p, _i, _i64, _i64, _i64, _p, _i, _p]),
    "mfish_minmax_f32": (_i, [_p, _i64, _p, _p]),
    "mfish_normalize_f64": (_i, [_p, _i, _i64, _d, _d, _p, _p, _p]),
    "mfish_sepconv_f32": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i, _p, _p, _i, _i, _i, _f, _p, _p]),
    "mfish_gaussian3d_f32": (_i, [_p, _p, _p, _i64, _i64, _i64, _p, _i, _p, _p]),
    "mfish_log3d_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _d, _p, _p]),
    "mfish_log3d_range_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _d, _p, _i64, _i64, _p]),
    "mfish_log3d_peaks_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _d, _p, _f, _p, _i64, _p, _p, _p, _i64, _p]),
    "mfish_log3d_peaks_range_f32": (_i, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _d, _p, _i64, _i64, _f, _p, _i64, _p,
                                         _p, _p, _i64, _p]),
    "mfish_nms_compact_f32": (_i, [_p, _i64, _i64, _i64, _i64, _f, _p, _p, _p, _i64, _p]),
    "mfish_rank_prune": (_i, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i, _p, _p, _p, _p]),
    "mfish_overlap_pairs": (_i, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _p]),
    "mfish_masked_quantile_f32": (_i, [_p, _p, _i64, _d, _p, _p, _p, _p]),
    "mfish_masked_quantile_fast_f32": (_i, [_p, _p, _i64, _d, _p, _p, _p, _p]),
    "mfish_masked_quantile_minmax_f32": (_i, [_p, _p, _i64, _d, _p, _p, _p, _p, _p]),
    "mfish_qsel_begin": (_i, [_p, _d, _p]),
    "mfish_qsel_hist": (_i, [_p, _p, _i64, _i, _p, _p]),
    "mfish_qsel_scan": (_i, [_i, _p, _p, _p, _p]),
    "mfish_qsel_scan_u64": (_i, [_i, _p, _p, _p, _p, _p]),
    "mfish_qselb_begin": (_i, [_p, _d, _p]),
    "mfish_qselb_sample_hist": (_i, [_p, _p, _i64, _i, _p, _p]),
    "mfish_qselb_collect": (_i, [_p, _p, _i64, _p, _p, _p, _i64, _p, _p, _p]),
    "mfish_qselb_ranks": (_i, [_p, _d, _p, _p, _p, _p]),
    "mfish_qselb_cand_hist": (_i, [_p, _i64, _i, _p, _p]),
    "mfish_blur_at_points": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _i, _p, _i, _p, _i, _i, _p, _p, _p]),
    "mfish_gather_f32": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "mfish_gather_u8": (_i, [_p, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "mfish_plane_sums_f32": (_i, [_p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "mfish_scale_planes_f32": (_i, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "mfish_edt_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "mfish_edt3d_u8": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "mfish_edt_rowscan_u8": (_i, [_p, _i64, _i64, _p, _p]),
    "mfish_edt_xpass_u16": (_i, [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "mfish_edt_inplane_u8": (_i, [_p, _i64, _i64, _i64, _p, _p, _p, _p, _p]),
    "mfish_edt_zpass": (_i, [_p, _i, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _p]),
    "mfish_export_f64": (_i, [_p, _i64, _i64, _i64, _i, _p, _p]),
    "mfish_morph_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "mfish_binary_morph_u8": (_i, [_p, _i64, _i64, _i64, _i, _i, _i, _p, _p, _p]),
    "mfish_histogram_f32": (_i, [_p, _i64, _d, _d, _i, _p, _p, _p]),
    "mfish_split_moments_f32": (_i, [_p, _i64, _f, _p, _p]),
    "mfish_binarize_f32": (_i, [_p, _i64, _f, _p, _p, _p]),
    "mfish_compartments_u8": (_i, [_p, _p, _i64, _i64, _i64, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "mfish_ccl_roots_u8": (_i, [_p, _i64, _i64, _i64, _i, _p, _p]),
    "mfish_label_sizes_i32": (_i, [_p, _i64, C.c_int32, _p, _p]),
    "mfish_copy2d": (_i, [_p, _i64, _p, _i64, _i64, _i64, _p]),
    "mfish_copy3d": (_i, [_p, _i64, _i64, _p, _i64, _i64, _i64, _i64, _i64, _p]),
    "mfish_label_bbox_i32": (_i, [_p, _i64, _i64, _i64, C.c_int32, C.c_int32, _p, _p]),
    "mfish_onehot_gaussian_f32": (_i, [_p, _i64, _i64, _i64, _p, _p, C.c_int32, _i64, _p, _p, _p, _p]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built and this package has no CPU "
            "fallback.  Run `python __graft_entry__.py build` (nvcc, sm_100a)."
        )
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the .so is stale: loud by design
        fn.restype = res
        fn.argtypes = args
    ver = handle.mfish_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError(f"libmfish_sm100 ABI {ver} != expected {ABI_VERSION}: rebuild the extension")
    _lib = handle
    return _lib


def check(code):
    if code != OK:
        raise MfishError(code, lib().mfish_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().mfish_launch_count())


def profile_enable(on: bool):
    lib().mfish_profile_enable(1 if on else 0)


def profile_collect():
    """{kernel name: (calls, total_ms)} for the launches recorded since the last collect."""
    buf = C.create_string_buffer(1 << 16)
    lib().mfish_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms = line.split()
        out[name] = (int(calls), float(ms))
    return out


def darray(values):
    """float64 host array argument (kept alive by the caller for the call's duration)."""
    arr = (C.c_double * len(values))(*[float(v) for v in values])
    return arr


def i64array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])
