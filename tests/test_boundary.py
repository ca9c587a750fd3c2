"""The drop-in boundary: a reference script reaches the hot path through
``import scope_utils3 as su; import muscle_fish as mf`` on PYTHONPATH (general_pipeline/fish_analysis.py:21-22).

The reference's own modules cannot be imported in this image (matplotlib / skimage / czifile are
missing), so the "original" modules here are stubs with the same names the scripts use
(``open_image``, ``threshold_fiber``, ``threshold_nuclei``, ``animate_zstacks`` ...).  The tests run the
literal call sequence of general_pipeline/fish_analysis.py:111-199 in a fresh interpreter with nothing but
PYTHONPATH set -- on the CPU up to the first voxel call (which must fail loudly: no fallback), on the GPU
end to end against the oracle.
"""
import ast
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "muscle-fish_b200", "dropin")

STUB_MF = '''
"""stand-in for the reference's modules/muscle_fish.py (non-hot-path functions only matter here)"""
import numpy as np
import scope_utils3 as su          # the original imports it the same way (modules/muscle_fish.py:18)
ORIGINAL = "muscle_fish"

def open_image(path):              # modules/muscle_fish.py:21-111 -> (img (c,x,y,z), name, metadata)
    d = np.load(path)
    return d["img"], "stack", None

def threshold_fiber(img, t_fiber=None, verbose=False):      # modules/muscle_fish.py:201-278
    return np.where(img > np.percentile(img, 45), 1, 0)      # int64 0/1, as np.where gives upstream

def threshold_nuclei(img, t_dapi=None, fiber_mask=None, verbose=False):   # modules/muscle_fish.py:114-198
    return np.where((img > 0.5) & (fiber_mask > 0), 1, 0)

def find_spots_snrfilter(*a, **k):
    raise AssertionError("the ORIGINAL find_spots_snrfilter must be shadowed by the drop-in")

def fix_bleaching(*a, **k):
    raise AssertionError("the ORIGINAL fix_bleaching must be shadowed by the drop-in")
'''

STUB_SU = '''
"""stand-in for the reference's modules/scope_utils3.py"""
ORIGINAL = "scope_utils3"
CALLS = []

def animate_zstacks(imgs, **kw):   # modules/scope_utils3.py:58-93
    CALLS.append(("animate_zstacks", len(imgs)))

def normalize_image(*a, **k):
    raise AssertionError("the ORIGINAL normalize_image must be shadowed by the drop-in")
'''

# general_pipeline/fish_analysis.py:21-22, 111, 132-142, 146, 152, 165, 196, 199 -- the same calls, same
# keyword arguments, in the same order
SCRIPT = '''
import json, sys
import numpy as np
import scope_utils3 as su
import muscle_fish as mf

out = {"mf_file": mf.__file__, "su_file": su.__file__}
out["open_image_from"] = mf.open_image.__module__
out["animate_from"] = su.animate_zstacks.__module__
out["hot_from"] = [mf.find_spots_snrfilter.__module__, mf.fix_bleaching.__module__, su.normalize_image.__module__]
try:
    mf.no_such_function
except AttributeError as e:
    out["missing"] = str(e)
img, img_name, mtree = mf.open_image(sys.argv[1])
stage = "normalize_image"
try:
    img_dapi = su.normalize_image(img[-1,:,:,:])
    img_rna = su.normalize_image(img[0,:,:,:])
    su.animate_zstacks([img_dapi, np.clip(img_rna, a_min=0, a_max=np.amax(img_rna)/5.)], titles=['DAPI', 'FISH, gain=5'], gif_name='x.gif')
    stage = "threshold_fiber"
    img_fiber_only = mf.threshold_fiber(img_rna, t_fiber=None, verbose=True)
    nuclei_labeled = mf.threshold_nuclei(img_dapi, t_dapi=None, fiber_mask=img_fiber_only, verbose=True)
    stage = "fix_bleaching"
    img_rna_corr = mf.fix_bleaching(img_rna, mask=img_fiber_only, draw=False, imgprefix=img_name)
    stage = "find_spots_snrfilter"
    spots_masked, spot_data = mf.find_spots_snrfilter(img_rna_corr, sigma=2, snr=None, t_spot=0.025, mask=img_fiber_only, imgprefix=img_name)
    stage = "done"
    np.savez(sys.argv[2], spots=spots_masked, img_rna=img_rna, img_rna_corr=img_rna_corr, fiber=img_fiber_only,
             table=np.array(spot_data[1:], dtype=np.float64))
    out["header"] = spot_data[0]
except RuntimeError as e:
    out["error"] = str(e)
out["stage"] = stage
print("RESULT " + json.dumps(out))
'''


def _run_script(tmp_path, stack, via_env=False, segmentation=False):
    orig = tmp_path / "modules"
    orig.mkdir()
    (orig / "muscle_fish.py").write_text(textwrap.dedent(STUB_MF))
    (orig / "scope_utils3.py").write_text(textwrap.dedent(STUB_SU))
    script = tmp_path / "fish_analysis_calls.py"
    script.write_text(textwrap.dedent(SCRIPT))
    img = np.stack([stack["img"], (stack["nuc_mask"] * 900.0 + 100.0).astype(np.float32)], axis=0)  # (c,x,y,z)
    np.savez(tmp_path / "stack.npz", img=img)
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "MUSCLE_FISH_MODULES")}
    # threshold_fiber / threshold_nuclei: the drop-in's own (GPU blurs + thresholds) or the original's
    env["MFISH_DROPIN_SEGMENTATION"] = "1" if segmentation else "0"
    if via_env:
        env["PYTHONPATH"] = DROPIN
        env["MUSCLE_FISH_MODULES"] = str(orig)
    else:
        env["PYTHONPATH"] = DROPIN + os.pathsep + str(orig)
    res = subprocess.run([sys.executable, str(script), str(tmp_path / "stack.npz"), str(tmp_path / "out.npz")],
                         env=env, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):]), res.stdout


@pytest.mark.parametrize("via_env", [False, True])
def test_pythonpath_seam_resolves_hot_path_here_and_the_rest_in_the_original(tmp_path, via_env):
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    out, _ = _run_script(tmp_path, synth.make_config("tiny"), via_env)
    assert os.path.dirname(out["mf_file"]) == DROPIN and os.path.dirname(out["su_file"]) == DROPIN
    assert out["open_image_from"] == "_muscle_fish_original"
    assert out["animate_from"] == "_scope_utils3_original"
    assert out["hot_from"] == ["muscle_fish_b200.muscle_fish", "muscle_fish_b200.muscle_fish",
                               "muscle_fish_b200.scope_utils3"]
    assert "no attribute 'no_such_function'" in out["missing"]
    import torch
    if not torch.cuda.is_available():
        # the first voxel call refuses to run without a GPU instead of falling back to the original
        assert out["stage"] == "normalize_image" and "no CPU fallback" in out["error"]


def test_missing_original_gives_a_clear_attribute_error(tmp_path):
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH", "MUSCLE_FISH_MODULES")}
    env["PYTHONPATH"] = DROPIN
    code = "import muscle_fish as mf\ntry:\n    mf.open_image\nexcept AttributeError as e:\n    print('AE', e)\n"
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stderr
    assert "AE" in res.stdout and "MUSCLE_FISH_MODULES" in res.stdout


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout exists only in the build container")
def test_every_name_the_reference_scripts_use_is_served():
    """AST scan of the five caller scripts: each ``mf.X`` / ``su.X`` they touch is either defined by the
    drop-in or by the original module of that name (so the fall-through finds it)."""
    scripts = ["general_pipeline/fish_analysis.py", "nocodazole/fish_analysis_noco.py",
               "rna_decay/fish_analysis_actd.py", "simulation/get_fiber_dimensions.py",
               "granule_intensity/fish_analysis_granule_intensity.py"]

    def defs(path):
        tree = ast.parse(open(path).read())
        names = set()
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                names.add(node.name)
            elif isinstance(node, ast.Assign):
                names.update(t.id for t in node.targets if isinstance(t, ast.Name))
        return names

    orig = {"mf": defs(os.path.join(REF, "modules/muscle_fish.py")),
            "su": defs(os.path.join(REF, "modules/scope_utils3.py"))}
    sys.path.insert(0, DROPIN)
    try:
        import importlib
        ours = {"mf": set(importlib.import_module("muscle_fish").HOT_PATH),
                "su": set(importlib.import_module("scope_utils3").HOT_PATH)}
    finally:
        sys.path.remove(DROPIN)
        sys.modules.pop("muscle_fish", None)
        sys.modules.pop("scope_utils3", None)
    used = {"mf": set(), "su": set()}
    for s in scripts:
        for node in ast.walk(ast.parse(open(os.path.join(REF, s)).read())):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in used:
                used[node.value.id].add(node.attr)
    assert {"open_image", "threshold_fiber", "threshold_nuclei", "find_spots_snrfilter"} <= used["mf"]
    for alias in ("mf", "su"):
        unserved = used[alias] - ours[alias] - orig[alias]
        assert not unserved, (alias, unserved)
        # and every hot-path name the scripts call is ours, not the original's
    assert {"find_spots", "find_spots_snrfilter", "fix_bleaching"} & used["mf"] <= ours["mf"]
    assert "normalize_image" in ours["su"]


@pytest.mark.gpu
def test_fish_analysis_call_sequence_end_to_end_vs_oracle(tmp_path):
    """general_pipeline/fish_analysis.py:111-199 through PYTHONPATH only, compared with the oracle run
    on the same arrays."""
    import oracle
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    from conftest import match_fraction
    st = synth.make_config("small")
    out, stdout = _run_script(tmp_path, st)
    assert out["stage"] == "done", out
    assert "SNR threshold = " in stdout
    assert out["header"] == ['spot_id', 'x', 'y', 'z', 'intensity', 'SNR']
    got = np.load(tmp_path / "out.npz")
    img_rna = oracle.normalize_image(st["img"])
    np.testing.assert_array_equal(got["img_rna"], img_rna)  # float64, bit exact
    fiber = np.where(img_rna > np.percentile(img_rna, 45), 1, 0)
    np.testing.assert_array_equal(got["fiber"], fiber)
    corr = oracle.fix_bleaching(img_rna, mask=fiber)
    np.testing.assert_allclose(got["img_rna_corr"], corr, rtol=2e-6)
    ref_spots, ref_data = oracle.find_spots_snrfilter(corr, sigma=2, t_spot=0.025, mask=fiber, pair_order="sorted")
    assert got["spots"].dtype == np.float64 and got["spots"].shape[1] == 4
    assert len(ref_spots) > 50
    assert match_fraction(got["spots"], ref_spots) >= 0.999
    assert abs(len(got["table"]) - (len(ref_data) - 1)) <= 1


@pytest.mark.gpu
def test_fish_analysis_call_sequence_with_the_dropin_segmentation(tmp_path):
    """the same script with threshold_fiber / threshold_nuclei served by the drop-in as well: the fibre mask
    equals the oracle's restatement of modules/muscle_fish.py:201-278 up to voxels within float32 rounding of
    the threshold, and the spot calls run on it"""
    import oracle
    from oracle import ref_segment as S
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import synth
    st = synth.make_config("small")
    # the synthetic fibre stops short of the first / last planes; fix_bleaching's fit needs fibre voxels in
    # every fitted plane (an empty plane mean is NaN and curve_fit raises, upstream too): extend it over z
    y0, y1, z0, z1 = synth._fiber_bounds(st["img"].shape)
    img = st["img"].copy()
    img[:, y0:y1, :z0] += 60.0
    img[:, y0:y1, z1:] += 60.0
    st = dict(st, img=img)
    out, stdout = _run_script(tmp_path, st, segmentation=True)
    assert out["stage"] == "done", out
    assert "Fiber threshold: " in stdout and "DAPI threshold = " in stdout
    got = np.load(tmp_path / "out.npz")
    ref_fiber = S.threshold_fiber(oracle.normalize_image(st["img"]))
    assert got["fiber"].dtype == ref_fiber.dtype and (got["fiber"] != ref_fiber).mean() < 2e-3
    assert len(got["spots"]) > 50


@pytest.mark.gpu
def test_nocodazole_call_sequence_vs_oracle():
    """nocodazole/fish_analysis_noco.py:189-298 with the import-line seams (morphology, distance_transform_edt,
    filters.gaussian) and the drop-in find_spots: nuclear region by erosions, grassfire, per-nucleus faded
    masks, spots, distance of every spot to the nearest nucleus -- each stage against the oracle's restatement."""
    import oracle
    from oracle import ref_fade, ref_morph
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import filters, morphology, ndimage, synth
    from muscle_fish_b200 import muscle_fish as mf
    from conftest import match_fraction
    st = synth.make_config("small")
    dims_xyz = np.array([0.0577, 0.0577, 0.5])
    nuclei_labeled = st["nuc_mask"].astype(np.int64)
    img_fiber_only = st["fiber_mask"].astype(np.int64)
    # :193-198  region_nuc: 1 erosion in 3-D, then 10 per plane
    region_nuc = np.where(nuclei_labeled > 0, 1, 0)
    for n in range(1):
        region_nuc = morphology.binary_erosion(region_nuc)
    for n in range(2):  # (10 upstream; the small test nuclei would vanish)
        for z in range(region_nuc.shape[2]):
            region_nuc[:, :, z] = morphology.binary_erosion(region_nuc[:, :, z])
    ref_region = ref_morph.iterate(np.where(nuclei_labeled > 0, 1, 0), "erode", 1, 2)
    np.testing.assert_array_equal(region_nuc != 0, ref_region != 0)
    assert region_nuc.any()
    # :218-221  grassfire
    grassfire = ndimage.distance_transform_edt(np.logical_not(region_nuc), sampling=dims_xyz)
    ref_grass = oracle.distance_transform_edt(np.logical_not(ref_region), sampling=dims_xyz)
    assert grassfire.dtype == np.float64
    np.testing.assert_allclose(grassfire, ref_grass, rtol=1e-6)
    # :227-235  labels (morphology.label on the device, skimage's numbering) and the faded mask of every nucleus
    region_nuc_labeled, n_nuc = morphology.label(region_nuc, return_num=True)
    lab = region_nuc_labeled - 1  # shift so bg = -1 and nuclei = [0:n_nuc]
    ref_lab, ref_n = ref_fade.label_nuclei(region_nuc)
    assert n_nuc == ref_n >= 2
    np.testing.assert_array_equal(lab, ref_lab)
    fades = filters.nucleus_fades(lab, n_nuc)
    for i in range(n_nuc):
        nuc = np.where(lab == i, 1, 0)
        assert np.abs(fades[i] - oracle.gaussian(nuc.astype(float), (3, 3, 1))).max() <= 1e-5
    # :269  spots; :297-298 their distance to the nearest nucleus (interpn on grid nodes = gather)
    spots = mf.find_spots(st["img"], t_spot=0.02, mask=img_fiber_only)
    ref_spots = oracle.find_spots(st["img"], t_spot=0.02, mask=img_fiber_only, pair_order="sorted")
    assert len(ref_spots) > 50 and match_fraction(spots, ref_spots) >= 0.999
    got_d = ndimage.spot_distances(region_nuc, spots, dims_xyz)
    np.testing.assert_allclose(got_d, oracle.edt_gather(ref_grass, spots, dims_xyz), rtol=1e-6)
