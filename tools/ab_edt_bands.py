"""EDT on the C2 nucleus mask (or EDT_SHAPE=x,y,z) with the x pass on one thread per line (MFISH_EDT_BANDS=1) and
on two (the default where the kernel's policy picks it; =2 forces it): per-kernel CUDA-event times, results
compared bit for bit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import muscle_fish_b200  # noqa: E402,F401
from muscle_fish_b200 import _lib, device as D, synth  # noqa: E402


def timed(notnuc, samp, bands, reps=10):
    os.environ["MFISH_EDT_BANDS"] = str(bands)
    for _ in range(3):
        d = D.edt(notnuc, samp)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d = D.edt(notnuc, samp)
    e1.record()
    torch.cuda.synchronize()
    k = _lib.profile_collect()
    _lib.profile_enable(False)
    line = " ".join(f"{n}={ms / c:.3f}" for n, (c, ms) in sorted(k.items()))
    return d, e0.elapsed_time(e1) / reps, line


def main():
    shapes = [synth.CONFIGS["C2"][0]]
    if os.environ.get("EDT_SHAPE"):
        shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ["EDT_SHAPE"].split(";")]
    nn, seed = synth.CONFIGS["C2"][2], synth.CONFIGS["C2"][3]
    for shape in shapes:
        scale = max(1, round(shape[0] * shape[1] * shape[2] / (2048 * 2048 * 100)))
        st = synth.make_stack_device(shape, 100, nn * scale, seed)
        notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
        del st
        for name, samp in (("aniso", (0.0577, 0.0577, 0.5)), ("iso", (1.0, 1.0, 1.0))):
            ref = None
            for bands in (1, 0, 2, 1, 0):
                d, ms, line = timed(notnuc, samp, bands)
                if ref is None:
                    ref = d.clone()
                same = bool(torch.equal(d, ref))
                print(f"{shape} {name} bands={bands}: total {ms:.3f} ms | {line} | identical={same}", flush=True)
                assert same
            _, sq1 = D.edt(notnuc, samp, want_sq=True) if name == "iso" else (None, None)
        del notnuc
        torch.cuda.empty_cache()
    os.environ.pop("MFISH_EDT_BANDS", None)


if __name__ == "__main__":
    main()
