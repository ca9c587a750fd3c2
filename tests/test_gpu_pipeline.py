"""GPU: the host-buffer entries around the hot path -- overlapped uploads (device.analyze_host and the
fused drop-in call), the device-resident hand-off between the reference's consecutive calls
(general_pipeline/fish_analysis.py:132-142,196-199) and the stack-per-GPU batch driver
(general_pipeline/batch_process.py:28-52).  Each is compared with the plain call sequence / the oracle."""
import numpy as np
import pytest

import oracle
from conftest import match_fraction

pytestmark = pytest.mark.gpu

ANISO = (0.0577, 0.0577, 0.5)


def _two_calls(st, sigma=2.0, t_spot=0.025, snr=None):
    from muscle_fish_b200 import muscle_fish as mf, ndimage as nd
    spots, data = mf.find_spots_snrfilter(st["img"], snr=snr, sigma=sigma, t_spot=t_spot, mask=st["fiber_mask"])
    return spots, data, nd.spot_distances(st["nuc_mask"], spots, ANISO)


@pytest.mark.parametrize("cfg", ["tiny", "small"])
def test_fused_call_equals_the_two_reference_calls_and_the_oracle(cfg, capsys):
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config(cfg)
    spots, data, dist = _two_calls(st)
    f_spots, f_data, f_dist = mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"], ANISO, sigma=2.0,
                                                                     t_spot=0.025, mask=st["fiber_mask"])
    assert capsys.readouterr().out.count("SNR threshold = ") == 2
    np.testing.assert_array_equal(f_spots, spots)
    assert f_data == data
    np.testing.assert_array_equal(f_dist, dist)
    ref_spots, _ = oracle.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=st["fiber_mask"],
                                               pair_order="sorted")
    grass = oracle.distance_transform_edt(np.logical_not(st["nuc_mask"]), sampling=ANISO)
    assert match_fraction(f_spots, ref_spots) >= 0.999
    np.testing.assert_allclose(f_dist, oracle.edt_gather(grass, f_spots, ANISO), rtol=1e-6)
    # explicit snr, int64 masks (np.where(..., 1, 0) upstream), z-first stacks
    a = mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"].astype(np.int64), ANISO, snr=4.0, sigma=2.0,
                                               t_spot=0.025, mask=st["fiber_mask"].astype(np.int64))
    b = _two_calls(st, snr=4.0)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[2], b[2])
    zf = lambda v: np.ascontiguousarray(v.transpose(2, 0, 1))
    c = mf.find_spots_snrfilter_with_distances(zf(st["img"]), zf(st["nuc_mask"]), ANISO, snr=4.0, sigma=2.0,
                                               t_spot=0.025, mask=zf(st["fiber_mask"]), z_first=True)
    np.testing.assert_array_equal(c[0], b[0][:, [2, 0, 1, 3]])
    np.testing.assert_array_equal(c[2], b[2])


def test_fused_call_error_behaviour():
    from muscle_fish_b200 import muscle_fish as mf, synth
    st = synth.make_config("tiny")
    with pytest.raises(ValueError):  # nothing inside the mask: np.row_stack([]) upstream
        mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"], ANISO, sigma=2.0, t_spot=0.025,
                                               mask=np.zeros(st["img"].shape, np.uint8))
    with pytest.raises(IndexError):
        mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"], ANISO, sigma=2.0, mask=None)


def test_uint16_stack_goes_up_as_it_is():
    """raw detector counts: 2 bytes per voxel over PCIe, same spots as the widened copy"""
    from muscle_fish_b200 import device as D, synth
    st = synth.make_config("small")
    u16 = np.clip(np.rint(st["img"] * 16.0), 0, 65535).astype(np.uint16)
    a = D.analyze_host(u16, st["fiber_mask"], st["nuc_mask"], 2.0, 0.025, ANISO)
    b = D.analyze_host(u16.astype(np.float64), st["fiber_mask"], st["nuc_mask"], 2.0, 0.025, ANISO)
    np.testing.assert_array_equal(a["spots"], b["spots"])
    np.testing.assert_array_equal(a["spot_dist"], b["spot_dist"])
    ref, _ = oracle.find_spots_snrfilter(u16, sigma=2.0, t_spot=0.025, mask=st["fiber_mask"], pair_order="sorted")
    assert match_fraction(a["spots"], ref) >= 0.999


def test_device_handoff_between_consecutive_calls():
    """normalize_image -> fix_bleaching -> find_spots_snrfilter as general_pipeline/fish_analysis.py:132-199
    chains them: with the hand-off the image is uploaded once; the results are bit-identical to the chain
    that goes through host memory at every step"""
    from muscle_fish_b200 import device as D, muscle_fish as mf, scope_utils3 as su, synth
    st = synth.make_config("small")
    raw = np.clip(np.rint(st["img"] * 16.0 * np.exp(-0.01 * np.arange(st["img"].shape[2]))), 0, 65535)
    raw = raw.astype(np.uint16)
    fiber = np.zeros(raw.shape, np.int64)
    fiber[:, 40:150, :] = 1

    def chain():
        img = su.normalize_image(raw)
        corr = mf.fix_bleaching(img, mask=fiber, draw=False)
        spots, data = mf.find_spots_snrfilter(corr, sigma=2, snr=None, t_spot=0.025, mask=fiber)
        return img, corr, spots, data

    assert D.HANDOFF.enabled
    h0 = D.HANDOFF.hits
    img, corr, spots, data = chain()
    assert D.HANDOFF.hits == h0 + 2  # fix_bleaching and find_spots_snrfilter found their argument on the device
    assert not img.flags.writeable and not corr.flags.writeable
    assert img.dtype == np.float64 and corr.dtype == np.float64
    np.testing.assert_array_equal(img, oracle.normalize_image(raw))
    D.HANDOFF.enabled = False
    try:
        img2, corr2, spots2, data2 = chain()
    finally:
        D.HANDOFF.enabled = True
    assert img2.flags.writeable
    np.testing.assert_array_equal(img, img2)
    np.testing.assert_array_equal(corr, corr2)
    np.testing.assert_array_equal(spots, spots2)
    assert data == data2
    # an edited copy is an ordinary array again
    mine = img.copy()
    mine[:8] = 0.0
    h1 = D.HANDOFF.hits
    s3 = mf.find_spots(mine, sigma=2.0, t_spot=0.025)
    assert D.HANDOFF.hits == h1
    np.testing.assert_array_equal(s3, mf.find_spots(np.array(mine), sigma=2.0, t_spot=0.025))
    with pytest.raises(ValueError):
        img[0, 0, 0] = 1.0  # read-only: the pairing with the device copy cannot be broken silently
    # the entry dies with the array
    n = len(D.HANDOFF._entries)
    del img, corr
    import gc
    gc.collect()
    assert len(D.HANDOFF._entries) <= n - 2


def test_run_batch_equals_stack_by_stack():
    """config-3 driver: round-robin shards, double-buffered staging; every stack's result equals the
    one-at-a-time call on the same arrays"""
    import torch
    from muscle_fish_b200 import device as D, parallel as P, synth
    shape = (96, 80, 24)
    n = 7
    stacks = [synth.make_stack(shape=shape, n_spots=40, n_nuclei=3, seed=100 + i) for i in range(n)]
    u16 = [np.clip(np.rint(s["img"] * 16.0), 0, 65535).astype(np.uint16) for s in stacks]
    calls = []

    def loader(i, img, fiber, nuc):
        calls.append(i)
        img[...] = u16[i]
        fiber[...] = stacks[i]["fiber_mask"]
        nuc[...] = stacks[i]["nuc_mask"]

    seen = []
    out = []
    for rank in range(2):
        out += P.run_batch(n, loader, shape, torch.uint16, 2.0, 0.025, ANISO, rank=rank, world=2,
                           on_result=lambda i, r: seen.append(i))
    assert sorted(i for i, _ in out) == list(range(n)) and sorted(calls) == list(range(n))
    assert seen == [0, 2, 4, 6, 1, 3, 5]
    for i, res in out:
        ref = D.analyze_host(u16[i], stacks[i]["fiber_mask"], stacks[i]["nuc_mask"], 2.0, 0.025, ANISO)
        np.testing.assert_array_equal(res["spots"], ref["spots"])
        np.testing.assert_array_equal(res["spot_dist"], ref["spot_dist"])
        assert res["baseline"] == ref["baseline"] and "dist" not in res
    assert P.run_batch(1, loader, shape, torch.uint16, rank=1, world=2) == []
    with pytest.raises(RuntimeError, match="boom"):
        def bad(i, *a):
            raise RuntimeError("boom")
        P.run_batch(3, bad, shape, torch.uint16, rank=0, world=1)


def test_staged_uploads_of_pageable_arrays(monkeypatch):
    """Large pageable arrays are narrowed on the host (float64 / int64 image -> float32, any mask -> 0/1 uint8:
    what ingest makes of them on the device) while they are staged through a ring of pinned buffers.  Forced on
    for small arrays and small chunks here, so that the ring wraps several times: same device data as the
    conversion done directly, same spots / table / distances as the plain path."""
    import torch
    from muscle_fish_b200 import device as D, muscle_fish as mf, synth
    monkeypatch.setattr(D, "_STAGE_MIN_BYTES", 0)
    monkeypatch.setattr(D, "_STAGE_CHUNK_ELEMS", 1 << 16)
    D._STAGE.clear()
    try:
        rng = np.random.default_rng(4)
        img = rng.random((90, 70, 31)) * 4000.0
        m64 = (rng.random((90, 70, 31)) < 0.4).astype(np.int64) * 5
        u16 = (img * 10).astype(np.uint16)
        fimg = np.asfortranarray(img)  # not C-contiguous
        (t_img, e1), (t_m, e2), (t_u, e3), (t_f, e4) = D.host_to_device_many(
            [img, m64, u16, fimg], kinds=("image", "mask", "image", "image"))
        for e in (e1, e2, e3, e4):
            assert e is not None
            torch.cuda.current_stream().wait_event(e)
        assert t_img.dtype == torch.float32 and torch.equal(t_img.cpu(), torch.from_numpy(img.astype(np.float32)))
        assert t_m.dtype == torch.uint8 and torch.equal(t_m.cpu(), torch.from_numpy((m64 != 0).astype(np.uint8)))
        assert t_u.dtype == torch.uint16 and np.array_equal(t_u.cpu().numpy(), u16)
        assert torch.equal(t_f.cpu(), torch.from_numpy(img.astype(np.float32)))
        st = synth.make_config("small")
        args = (st["img"].astype(np.float64), st["nuc_mask"].astype(np.int64), ANISO)
        kw = dict(sigma=2.0, t_spot=0.025, mask=st["fiber_mask"].astype(np.int64))
        a = mf.find_spots_snrfilter_with_distances(*args, **kw)
        monkeypatch.setattr(D, "STAGED_UPLOADS", False)
        b = mf.find_spots_snrfilter_with_distances(*args, **kw)
        np.testing.assert_array_equal(a[0], b[0])
        assert a[1] == b[1]
        np.testing.assert_array_equal(a[2], b[2])
        assert len(a[0]) > 20
    finally:
        D._STAGE.clear()


def test_mask_handoff_from_threshold_fiber():
    """threshold_fiber's result is the `mask=` of every later call of general_pipeline/fish_analysis.py:152-199;
    its device twin is found again (no 8-byte-per-voxel upload), and the results equal those with a plain copy
    of the mask (which is uploaded like any other array)."""
    from muscle_fish_b200 import device as D, muscle_fish as mf, segmentation as sg, morphology, synth
    st = synth.make_config("small")
    D.MASK_HANDOFF.clear()
    fiber = sg.threshold_fiber(st["img"])
    assert fiber.dtype == np.int64 and not fiber.flags.writeable
    plain = fiber.copy()
    h0 = D.MASK_HANDOFF.hits
    a = mf.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=fiber)
    assert D.MASK_HANDOFF.hits == h0 + 1
    b = mf.find_spots_snrfilter(st["img"], sigma=2.0, t_spot=0.025, mask=plain)
    assert D.MASK_HANDOFF.hits == h0 + 1
    np.testing.assert_array_equal(a[0], b[0])
    assert a[1] == b[1]
    c = mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"], ANISO, sigma=2.0, t_spot=0.025, mask=fiber)
    d = mf.find_spots_snrfilter_with_distances(st["img"], st["nuc_mask"], ANISO, sigma=2.0, t_spot=0.025, mask=plain)
    assert D.MASK_HANDOFF.hits == h0 + 2
    np.testing.assert_array_equal(c[0], d[0])
    np.testing.assert_array_equal(c[2], d[2])
    # (fix_bleaching takes the same mask through device.ingest_mask; the synthetic fibre leaves the outer planes
    # empty, whose NaN means make curve_fit raise here as upstream, so it is exercised in test_boundary instead)
    np.testing.assert_array_equal(morphology.binary_erosion(fiber), morphology.binary_erosion(plain))
    n1 = sg.threshold_nuclei(st["img"], fiber_mask=fiber)
    n2 = sg.threshold_nuclei(st["img"], fiber_mask=plain)
    np.testing.assert_array_equal(n1, n2)
    assert D.MASK_HANDOFF.hits >= h0 + 4


def test_masks_as_bits_equal_whole_byte_masks(monkeypatch):
    """mfish_host_pack_mask + mfish_ingest_mask_bits against mfish_ingest on the whole-byte mask (every dtype the
    reference's callers pass, both layouts, both polarities, ragged sizes), and the host-buffer entry with packed
    masks against the same call with MFISH_PACK_MASKS off."""
    import torch
    from muscle_fish_b200 import device as D, muscle_fish as mf, synth
    monkeypatch.setattr(D, "_PACK_MIN_ELEMS", 0)
    monkeypatch.setattr(D, "_PACK_CHUNK_ELEMS", 1 << 12)
    monkeypatch.setattr(D, "_SPLIT_MIN_BYTES", 0)
    rng = np.random.default_rng(8)
    for shape in [(40, 36, 21), (33, 17, 100), (8, 64, 32), (5, 3, 1), (64, 8, 77)]:
        for dt in (np.uint8, np.bool_, np.int64, np.float64, np.uint16):
            m = (rng.random(shape) < 0.3).astype(dt)
            if dt not in (np.bool_,):
                m = m * 3
            for z_first in (False, True):
                for negate in (False, True):
                    (t, e), = D.host_to_device_many([m], kinds=("mask",), pack_masks=True)
                    assert isinstance(t, D.PackedMask) and e is not None
                    got = D.ingest(t, z_first=z_first, as_mask=True, negate=negate, ready=e)
                    ref = D.ingest(m, z_first=z_first, as_mask=True, negate=negate)
                    assert torch.equal(got, ref), (shape, dt, z_first, negate)
    # two masks and a pinned image in one call: the first mask is sent between the halves of the image
    img = torch.from_numpy(rng.random((64, 48, 20)).astype(np.float32)).pin_memory()
    m1 = (rng.random((64, 48, 20)) < 0.5).astype(np.uint8)
    m2 = (rng.random((64, 48, 20)) < 0.5).astype(np.uint8)
    (a, ea), (b, eb), (c, ec) = D.host_to_device_many([m1, img.numpy(), m2], kinds=("mask", "image", "mask"),
                                                      pack_masks=True)
    for e in (ea, eb, ec):
        torch.cuda.current_stream().wait_event(e)
    assert torch.equal(b.cpu(), img)
    assert torch.equal(D.ingest(a, as_mask=True), D.ingest(m1, as_mask=True))
    assert torch.equal(D.ingest(c, as_mask=True), D.ingest(m2, as_mask=True))
    st = synth.make_config("small")
    args = (st["img"], st["nuc_mask"], ANISO)
    kw = dict(sigma=2.0, t_spot=0.025, mask=st["fiber_mask"])
    before = D.H2D_BYTES
    a = mf.find_spots_snrfilter_with_distances(*args, **kw)
    packed_bytes = D.H2D_BYTES - before
    monkeypatch.setattr(D, "PACK_MASKS", False)
    before = D.H2D_BYTES
    b = mf.find_spots_snrfilter_with_distances(*args, **kw)
    assert packed_bytes < D.H2D_BYTES - before
    np.testing.assert_array_equal(a[0], b[0])
    assert a[1] == b[1]
    np.testing.assert_array_equal(a[2], b[2])
    assert len(a[0]) > 20
