"""A/B the graph-replayed guided step under environment switches, interleaved to cancel clock drift.
python tools/ab_step.py "ISB_SIDE_BWD=0" "ISB_FILM_CACHE=0" ...   (the empty variant = defaults is always run)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper


def main():
    variants = [""] + sys.argv[1:]
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), "bf16", dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 128, 128, generator=g).to(dev)
    origin = torch.randn(3, 64, 64, 170, generator=g).to(dev)
    steppers = []
    for v in variants:
        saved = dict(os.environ)
        for kv in v.split():
            k, val = kv.split("=")
            os.environ[k] = val
        if "ISB_TC_" in v:
            model.invalidate()      # conv tiling policy changed: the plan's statistics slots depend on it
        st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=True)
        st.img.copy_(x)
        for k in range(50):
            st.step(49 - k, origin)
        os.environ.clear()
        os.environ.update(saved)
        steppers.append(st)
    torch.cuda.synchronize()
    res = {v: [] for v in variants}
    for rnd in range(5):
        for v, st in zip(variants, steppers):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(50):
                st.step(49 - k, origin)
            e1.record()
            torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / 50)
    for v in variants:
        r = sorted(res[v])
        print(f"{v or '(defaults)':40s} median {r[len(r)//2]:.3f} ms   min {r[0]:.3f}   max {r[-1]:.3f}")
    os._exit(0)


main()
