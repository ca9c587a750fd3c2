"""muscle-fish_b200: B200-native (sm_100a) implementation of muscle-FISH's volumetric
spot-analysis hot path -- LoG spot detection, SNR filter, nucleus distance transform.

The directory name carries a hyphen, so import it through the loader at the repository
root (``import muscle_fish_b200``).  Sub-modules:

* ``muscle_fish`` / ``scope_utils3``  drop-in host API with the reference's signatures
  (``modules/muscle_fish.py``, ``modules/scope_utils3.py``)
* ``ndimage``      ``distance_transform_edt`` / ``gaussian`` / ``gaussian_laplace`` replacements
* ``segmentation`` ``threshold_fiber`` / ``threshold_nuclei`` (blurs, Li / Otsu statistics on the GPU)
* ``morphology``   ``binary_erosion`` / ``binary_dilation`` / ``compartments`` (the compartment masks)
* ``filters``      ``gaussian`` and the batched per-nucleus faded masks of the nuclear-mesh step
* ``device``       device-resident pipeline on torch CUDA tensors
* ``parallel``     multi-GPU drivers (stack-per-GPU batches, z-slab mosaics)
* ``taps`` / ``synth``  host-side filter taps and the synthetic stack generator

Every voxel operation runs in ``libmfish_sm100.so`` (hand-written CUDA, C ABI in
``include/mfish.h``).  There is no CPU fallback: without the built extension or without a
CUDA device the calls raise.
"""
__version__ = "0.1.0"

from . import taps, synth  # noqa: F401  (pure numpy, importable anywhere)


def __getattr__(name):
    import importlib
    if name in ("device", "muscle_fish", "scope_utils3", "ndimage", "morphology", "segmentation", "filters", "parallel",
                "_lib"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
