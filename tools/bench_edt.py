"""EDT only, on the C2 nucleus mask: per-kernel CUDA-event times (mfish_profile_*), for the
anisotropic and the isotropic transform.  The environment switch MFISH_EDT_V2=0
selects the linked-hull kernels of round 1; the result is compared against theirs bit for bit
(isotropic squares) / to 1e-6 (anisotropic)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(tag, save=None):
    import torch
    import muscle_fish_b200  # noqa: F401
    from muscle_fish_b200 import _lib, device as D, synth
    wl = os.environ.get("EDT_WL", "C2")
    shape, ns, nn, seed = synth.CONFIGS[wl]
    if os.environ.get("EDT_NZ"):
        shape = (shape[0], shape[1], int(os.environ["EDT_NZ"]))
    if os.environ.get("EDT_SHAPE"):  # "x,y,z"
        shape = tuple(int(v) for v in os.environ["EDT_SHAPE"].split(","))
    st = synth.make_stack_device(shape, 100, nn, seed)
    notnuc = (st["nuc_mask"] == 0).to(torch.uint8)
    del st
    nvox = notnuc.numel()
    out = {}
    for name, samp in (("aniso", (0.0577, 0.0577, 0.5)), ("iso", (1.0, 1.0, 1.0))):
        for _ in range(3):
            d = D.edt(notnuc, samp)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        _lib.profile_collect()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            d = D.edt(notnuc, samp)
        e1.record()
        torch.cuda.synchronize()
        k = _lib.profile_collect()
        _lib.profile_enable(False)
        line = " ".join(f"{n}={ms / c:.3f}" for n, (c, ms) in sorted(k.items()))
        print(f"[{tag}] {wl} {shape} {name}: total {e0.elapsed_time(e1) / 10:.3f} ms | {line}", flush=True)
        out[name] = d.cpu()
        if name == "iso":
            _, sq = D.edt(notnuc, samp, want_sq=True)
            out["iso_sq"] = sq.cpu()
    if save:
        torch.save(out, save)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        run(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
        sys.exit(0)
    import torch
    os.makedirs("gpurun_out", exist_ok=True)
    base_env = dict(os.environ, MFISH_EDT_V2="0")
    subprocess.run([sys.executable, __file__, "--child", "round1", "/tmp/edt_old.pt"], env=base_env, check=True)
    new = run("new")
    old = torch.load("/tmp/edt_old.pt")
    ok_sq = bool(torch.equal(old["iso_sq"], new["iso_sq"]))
    rel = ((old["aniso"] - new["aniso"]).abs() / old["aniso"].clamp(min=1e-30)).max().item()
    rel_iso = ((old["iso"] - new["iso"]).abs() / old["iso"].clamp(min=1e-30)).max().item()
    print(f"iso squares identical: {ok_sq}; max rel diff aniso {rel:.3e}, iso {rel_iso:.3e}")
    assert ok_sq and rel < 1e-6 and rel_iso < 1e-6
