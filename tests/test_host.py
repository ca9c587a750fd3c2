"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
import muscle_fish_b200  # noqa: F401
from muscle_fish_b200 import taps as T, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mfish.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfish_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from muscle_fish_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python __graft_entry__.py build"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/mfish.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    lib = _lib.lib()
    assert lib.mfish_abi_version() == _lib.ABI_VERSION
    hdr = open(os.path.join(ROOT, "include", "mfish.h")).read()
    assert f"#define MFISH_ABI_VERSION {_lib.ABI_VERSION}" in hdr
    assert lib.mfish_launch_count() == 0  # nothing was launched by loading


def test_header_constants_match_python_mirror():
    from muscle_fish_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mfish.h")).read()
    consts = dict(re.findall(r"#define (MFISH_[A-Z0-9_]+) +(\(?-?\d+\)?)", hdr))
    val = lambda k: int(consts[k].strip("()"))
    assert (val("MFISH_U8"), val("MFISH_U16"), val("MFISH_I16"), val("MFISH_I32"), val("MFISH_I64"), val("MFISH_F32"),
            val("MFISH_F64")) == (_lib.U8, _lib.U16, _lib.I16, _lib.I32, _lib.I64, _lib.F32, _lib.F64)
    assert (val("MFISH_LAYOUT_XYZ"), val("MFISH_LAYOUT_ZXY")) == (_lib.LAYOUT_XYZ, _lib.LAYOUT_ZXY)
    assert (val("MFISH_DST_F32"), val("MFISH_DST_NONZERO"), val("MFISH_DST_ISZERO")) == \
        (_lib.DST_F32, _lib.DST_NONZERO, _lib.DST_ISZERO)
    assert (val("MFISH_BOUNDARY_REFLECT"), val("MFISH_BOUNDARY_NEAREST")) == (_lib.BOUNDARY_REFLECT, _lib.BOUNDARY_NEAREST)
    assert (val("MFISH_AXIS_Z"), val("MFISH_AXIS_X"), val("MFISH_AXIS_Y")) == (_lib.AXIS_Z, _lib.AXIS_X, _lib.AXIS_Y)
    assert (val("MFISH_CONV_G"), val("MFISH_CONV_GD"), val("MFISH_CONV_MID"), val("MFISH_CONV_LAST")) == \
        (_lib.CONV_G, _lib.CONV_GD, _lib.CONV_MID, _lib.CONV_LAST)
    assert (val("MFISH_EDT_F2_I32"), val("MFISH_EDT_F2_F32")) == (_lib.EDT_F2_I32, _lib.EDT_F2_F32)
    assert (val("MFISH_OK"), val("MFISH_ERR_INVALID"), val("MFISH_ERR_CUDA")) == (0, -1, -2)


def test_argument_validation_without_a_gpu():
    """error paths that return before any CUDA call"""
    from muscle_fish_b200 import _lib
    lib = _lib.lib()
    assert lib.mfish_edt_workspace_bytes(0, 4, 4) == 0
    n = 10 * 20 * 30
    assert lib.mfish_edt_workspace_bytes(10, 20, 30) >= n * 10
    rc = lib.mfish_nms_compact_f32(None, 1, 4, 4, 4, 0.1, None, None, None, 16, None)
    assert rc == _lib.ERR_INVALID and b"null" in lib.mfish_last_error()
    with pytest.raises(_lib.MfishError):
        _lib.check(lib.mfish_minmax_f32(None, 0, None, None))
    rc = lib.mfish_log3d_f32(None, None, None, None, None, 4, 4, 4, 2.0, None, None)
    assert rc == _lib.ERR_INVALID


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "muscle-fish_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "scipy.ndimage" not in text or f == "ndimage.py" or f.endswith((".cu", ".cuh", ".py")), f
    # and the device module refuses to run without CUDA instead of falling back
    from muscle_fish_b200 import device as D
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            D.host_to_device(np.zeros((2, 2, 2), np.float32))


def test_missing_extension_fails_loudly(monkeypatch):
    from muscle_fish_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmfish_sm100.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.lib()


# ------------------------------------------------------------ host arithmetic
def test_interp01_and_percentile_helpers_match_numpy():
    rng = np.random.default_rng(0)
    a = (rng.random(1000) * 4000).astype(np.float32)
    lo, hi = float(a.min()), float(a.max())
    np.testing.assert_array_equal(T.interp01(a, lo, hi), np.interp(a, (lo, hi), (0, 1)))
    np.testing.assert_array_equal(T.interp01(np.full(3, 5.0), 5.0, 5.0), np.interp(np.full(3, 5.0), (5, 5), (0, 1)))
    srt = np.sort(np.interp(a, (lo, hi), (0, 1)))
    for q in (0, 1, 25, 50, 90, 99.5, 100):
        vi, k, g = T.percentile_virtual_index(len(srt), q)
        k1 = min(k + 1, len(srt) - 1)
        assert T.lerp(srt[k], srt[k1], g) == np.percentile(srt, q)
    for n in (1, 2, 3, 7):
        s = np.sort(rng.random(n))
        for q in (0, 25, 33.3, 90, 100):
            vi, k, g = T.percentile_virtual_index(n, q)
            assert T.lerp(s[k], s[min(k + 1, n - 1)], g) == np.percentile(s, q)


def test_decode_raster_and_sigma_list():
    from muscle_fish_b200.device import decode_raster
    dims = (3, 7, 5, 4)  # ns, nx, ny, nz
    cube = np.arange(np.prod(dims)).reshape(7, 5, 4, 3)
    idx = np.array([0, 17, 100, 419])
    x, y, z, s = decode_raster(idx, dims)
    np.testing.assert_array_equal(cube[x, y, z, s], idx)
    np.testing.assert_array_equal(T.sigma_list(1, 3, 5), oracle.sigma_list(1, 3, 5))
    np.testing.assert_array_equal(T.sigma_list(2, 2, 1), [2.0])


def test_unequal_sigma_prune_host_equals_oracle_sorted_order():
    from muscle_fish_b200.device import prune_unequal_host
    rng = np.random.default_rng(4)
    sigmas = T.sigma_list(1, 3, 5)
    n = 600
    pts = rng.integers(0, 40, size=(n, 3))
    pts = np.unique(pts, axis=0)
    n = len(pts)
    s_idx = rng.integers(0, 5, size=n)
    blobs = np.concatenate([pts, sigmas[s_idx][:, None]], axis=1).astype(np.float64)
    ref = oracle.prune_blobs(blobs, 0.5, pair_order="sorted")
    cut = np.array([[T.prune_cutoff_sq(a, b) for b in sigmas] for a in sigmas])
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    ii, jj = np.nonzero(np.triu(d2 <= cut[s_idx][:, s_idx], k=1))
    pairs = np.stack([ii, jj], axis=1)
    sig = prune_unequal_host(n, s_idx, sigmas, pairs)
    got = blobs[sig > 0]
    assert 0 < len(ref) < n
    np.testing.assert_array_equal(got, ref)


def test_synthetic_generator_is_seeded_and_sane():
    a = synth.make_config("tiny")
    b = synth.make_config("tiny")
    np.testing.assert_array_equal(a["img"], b["img"])
    assert a["img"].dtype == np.float32 and a["img"].shape == (96, 80, 24)
    assert a["fiber_mask"].dtype == np.uint8 and 0 < a["nuc_mask"].mean() < 0.2
    c = synth.make_stack(shape=(96, 80, 24), n_spots=40, n_nuclei=3, seed=8)
    assert not np.array_equal(a["img"], c["img"])


def test_dropin_modules_expose_reference_signatures():
    import inspect
    from muscle_fish_b200 import muscle_fish as mf, scope_utils3 as su, ndimage as nd
    assert str(inspect.signature(mf.find_spots)) == "(img_fish, sigma=1.0, t_spot=0.025, mask=None)"
    assert str(inspect.signature(mf.find_spots_snrfilter)) == \
        "(img_fish, snr=None, sigma=1.0, t_spot=0.025, mask=None, z_first=False, draw=False, imgprefix='fiber')"
    assert str(inspect.signature(mf.fix_bleaching)) == \
        "(img, second_half=True, mask=None, z_first=False, draw=False, imgprefix='fiber')"
    assert str(inspect.signature(su.normalize_image)) == "(image, vmin=0, vmax=1)"
    assert list(inspect.signature(su.optimize_threshold_blob_log).parameters)[:7] == \
        ["image", "min_sigma", "max_sigma", "num_sigma", "tmin", "tmax", "num"]
    assert list(inspect.signature(nd.distance_transform_edt).parameters)[:2] == ["input", "sampling"]
    # PYTHONPATH drop-in: `import muscle_fish as mf; import scope_utils3 as su`
    import importlib
    import sys
    d = os.path.join(ROOT, "muscle-fish_b200", "dropin")
    sys.path.insert(0, d)
    try:
        mf2 = importlib.import_module("muscle_fish")
        su2 = importlib.import_module("scope_utils3")
        assert mf2.find_spots_snrfilter is mf.find_spots_snrfilter and su2.normalize_image is su.normalize_image
    finally:
        sys.path.remove(d)
        sys.modules.pop("muscle_fish", None)
        sys.modules.pop("scope_utils3", None)


def test_slab_generators_draw_the_voxels_of_the_whole_stack():
    """the bench's mosaic is generated slab by slab on each rank: any z- or x-slab must be bit-identical to the
    same voxels of the whole stack (noise seeded per global plane, spots accumulated in fixed point)"""
    import torch
    shape, ns, nn, seed = (96, 64, 20), 400, 5, 4
    full = synth.make_stack_device(shape, ns, nn, seed, device="cpu")
    for z0, z1 in ((0, 7), (7, 20)):
        part = synth.make_stack_device(shape, ns, nn, seed, device="cpu", z_range=(z0, z1))
        for k in ("img", "fiber_mask", "nuc_mask"):
            assert torch.equal(part[k], full[k][z0:z1]), (k, z0)
    for x0, x1 in ((0, 31), (31, 64), (64, 96)):
        part = synth.make_stack_device(shape, ns, nn, seed, device="cpu", x_range=(x0, x1))
        for k in ("img", "fiber_mask", "nuc_mask"):
            assert torch.equal(part[k], full[k][:, x0:x1]), (k, x0)
    again = synth.make_stack_device(shape, ns, nn, seed, device="cpu")
    assert torch.equal(again["img"], full["img"])


def test_host_mask_packer_equals_numpy_packbits():
    """mfish_host_pack_mask is plain host code (no CUDA call): every dtype, ragged lengths, both polarities, and
    the split into chunks the worker threads use"""
    from muscle_fish_b200 import _lib
    lib = _lib.lib()
    rng = np.random.default_rng(0)
    for dt, code in ((np.uint8, _lib.U8), (np.uint16, _lib.U16), (np.int16, _lib.I16), (np.int32, _lib.I32),
                     (np.int64, _lib.I64), (np.float32, _lib.F32), (np.float64, _lib.F64)):
        for n in (0, 1, 7, 8, 63, 64, 65, 1000, 4099):
            a = (rng.random(n) < 0.3).astype(dt) * 3
            nb = int(lib.mfish_mask_bits_bytes(n))
            assert nb == (0 if n == 0 else ((n + 31) // 32 + 1) * 4)
            out = np.full(max(nb, 1), 0xAA, np.uint8)
            for neg in (0, 1):
                _lib.check(lib.mfish_host_pack_mask(a.ctypes.data, code, n, out.ctypes.data, neg))
                ref = np.packbits((a != 0) != bool(neg), bitorder="little")
                np.testing.assert_array_equal(out[:len(ref)], ref)
    a = (rng.random(100_000) < 0.5).astype(np.uint8)
    out = np.zeros(int(lib.mfish_mask_bits_bytes(a.size)), np.uint8)
    for i0 in range(0, a.size, 4096):  # chunk starts are multiples of 8 elements
        n = min(4096, a.size - i0)
        _lib.check(lib.mfish_host_pack_mask(a.ctypes.data + i0, _lib.U8, n, out.ctypes.data + i0 // 8, 0))
    np.testing.assert_array_equal(out[:a.size // 8], np.packbits(a, bitorder="little"))
    assert lib.mfish_host_pack_mask(None, _lib.U8, 8, out.ctypes.data, 0) == _lib.ERR_INVALID
    assert lib.mfish_host_pack_mask(a.ctypes.data, 99, 8, out.ctypes.data, 0) == _lib.ERR_UNSUPPORTED


def test_which_masks_travel_as_bits():
    """device._packable: C-contiguous 3-D arrays of the ingest dtypes above the size threshold (bool as uint8);
    everything else takes the whole-byte path.  No CUDA call involved."""
    import torch
    from muscle_fish_b200 import device as D
    big = np.zeros((64, 64, 8), np.uint8)
    assert D._packable(big) is None  # below the threshold
    assert D._packable(big, min_elems=0) is big
    assert D._packable_any_size(big.astype(np.bool_)).dtype == np.uint8
    assert D._packable_any_size(big.astype(np.int64)).dtype == np.int64
    assert D._packable_any_size(np.asfortranarray(big)) is None  # not C-contiguous
    assert D._packable_any_size(big[:, :, ::2]) is None
    assert D._packable_any_size(big[0]) is None  # not 3-D
    assert D._packable_any_size(big.astype(np.complex64)) is None  # not an ingest dtype
    t = torch.zeros((4, 4, 4), dtype=torch.bool)
    a = D._packable_any_size(t)
    assert a is not None and a.dtype == np.uint8 and a.shape == (4, 4, 4)
    assert D._packable_any_size(t.permute(2, 0, 1)) is None
    # the default follows the number of ranks per host
    assert D.PACK_MASKS == (int(os.environ.get("LOCAL_WORLD_SIZE", "1")) <= 2
                            if os.environ.get("MFISH_PACK_MASKS") is None else os.environ["MFISH_PACK_MASKS"] != "0")
