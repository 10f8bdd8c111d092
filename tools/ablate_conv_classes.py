"""Which convolution classes sit on the critical path of the graph-replayed guided step?  Re-capture the step with the
convs of ONE class (by spatial size / role) stubbed out and report the time they cost in-graph (not cold, not
serialised like the ncu launch list).   python tools/ablate_conv_classes.py [--batch B]
With --batch B the step advances B independent edits (the throughput mode); the GroupNorm and attention families are
ablated the same way."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper


def main():
    B = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 1
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), "bf16", dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    if B > 1:
        geo = [geo] * B
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 96, 128, 128, generator=g).to(dev)
    origin = torch.randn(*((B,) if B > 1 else ()), 3, 64, 64, 170, generator=g).to(dev)
    ops = model._get_ops()
    orig = ops.conv
    stats = {}

    def make(pred):
        def conv(a, w, bias, ksize, out, **kw):
            H = a.shape[1]
            key = (H, ksize)
            if pred(H, ksize):
                e = stats.setdefault(key, [0, 0.0])
                e[0] += 1
                e[1] += 2.0 * a.shape[0] * a.shape[1] * a.shape[2] * w.shape[0] * w.shape[1]
                return out
            return orig(a, w, bias, ksize, out, **kw)
        return conv

    def run(pred):
        stats.clear()
        ops.conv = make(pred)
        try:
            st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=True)
            st.img.copy_(x)
            st.step(49, origin)
            snap = {k: tuple(v) for k, v in stats.items()}
            for k in range(1, 5):
                st.step(49 - k, origin)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(30):
                st.step(49 - k, origin)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 30, snap
        finally:
            ops.conv = orig

    run(lambda H, k: False)                      # throw-away: a fresh box ramps its clocks during the first seconds
    run(lambda H, k: True)
    full, _ = run(lambda H, k: False)
    print(f"full step {full:.3f} ms")
    for H in (() if "--only-full" in sys.argv else (8, 16, 32, 64, 128)):
        for ks in (3, 1):
            t, snap = run(lambda h, k, H=H, ks=ks: h == H and k == ks)
            n = sum(v[0] for v in snap.values())
            gf = sum(v[1] for v in snap.values()) / 1e9
            if n:
                d = full - t
                print(f"H={H:3d} k{ks}: {n:3d} launches {gf:7.1f} GFLOP  cost {d:6.3f} ms  ({1e3 * d / n:5.1f} us/launch, "
                      f"{gf / max(d, 1e-6):7.1f} TF/s in-graph)")
    if "--only-full" not in sys.argv:
        t, snap = run(lambda h, k: True)
        print(f"all convs: cost {full - t:.3f} ms")
    noop = lambda *a, **kw: None     # noqa: E731
    for fam, names in (("GroupNorm", ("gn_forward", "gn_backward")),
                       ("attention", ("attention_flash_forward", "attention_flash_backward")),
                       ("drag loss + gradient", ("drag_loss_grad",))):
        saved = {n: getattr(ops, n) for n in names}
        for n in names:
            setattr(ops, n, noop)
        try:
            t, _ = run(lambda H, k: False)
        finally:
            for n, f in saved.items():
                setattr(ops, n, f)
        print(f"{fam}: cost {full - t:.3f} ms")
    full2, _ = run(lambda H, k: False)
    print(f"full step again {full2:.3f} ms")
    os._exit(0)


main()
