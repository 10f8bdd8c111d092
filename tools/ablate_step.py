"""Where does the graph-replayed guided step spend its time?  Re-capture the step with one operator
family stubbed out (no kernels launched for it; results are garbage, timing is what matters) and
report the difference to the full step.  python tools/ablate_step.py [--mode bf16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper

FAMILIES = ["conv", "gn_forward", "gn_backward", "attention_forward", "attention_backward",
            "attention_flash_forward", "attention_flash_backward", "drag_loss_grad",
            "time_embed", "ddpm_step"]


def time_stepper(st, origin, reps=20):
    for k in range(3):
        st.step(49 - k, origin)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        st.step(49 - (k % 50), origin)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--batch", type=int, default=1, help="edits advanced together (throughput mode)")
    args = ap.parse_args()
    B = args.batch
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), args.mode, dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    if B > 1:
        geo = [DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170) for _ in range(B)]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 96, 128, 128, generator=g).to(dev)
    origin = torch.randn(3, 64, 64, 170, generator=g).to(dev)
    if B > 1:
        origin = torch.stack([origin] * B)
    ops = model._get_ops()
    orig = {f: getattr(ops, f) for f in FAMILIES}

    def run(stub):
        for f in FAMILIES:
            setattr(ops, f, orig[f])
        for f in stub:
            setattr(ops, f, lambda *a, **k: None)
        st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=True)
        st.img.copy_(x)
        return time_stepper(st, origin)

    full = run([])
    print(f"full step                {full:7.3f} ms")
    for f in FAMILIES:
        t = run([f])
        print(f"without {f:20s} {t:7.3f} ms   -> {f} costs {full - t:6.3f} ms ({100 * (full - t) / full:4.1f}%)")
    t = run(FAMILIES)
    print(f"all stubbed (torch copies/noise only) {t:7.3f} ms")


if __name__ == "__main__":
    main()
