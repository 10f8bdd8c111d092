"""In-graph (L2-warm) time of every distinct GroupNorm / conv / attention call of one guided step: the calls of an
eager step are recorded with their real tensors, each distinct signature is then replayed 10x inside a CUDA graph.
    python tools/time_ops.py [family ...]      (default: gn_forward gn_backward)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper


def sig(name, args, kwargs):
    parts = [name]
    for a in list(args) + [kwargs[k] for k in sorted(kwargs)]:
        if torch.is_tensor(a):
            parts.append(f"{tuple(a.shape)}:{str(a.dtype)[6:]}")
        elif hasattr(a, "cout"):
            parts.append(f"W{a.cout}x{a.k}")
        elif a is None:
            parts.append("-")
        else:
            parts.append(str(a))
    return " ".join(parts)


def nbytes(args, kwargs):
    return sum(a.numel() * a.element_size() for a in list(args) + list(kwargs.values()) if torch.is_tensor(a))


def main():
    fams = sys.argv[1:] or ["gn_forward", "gn_backward"]
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), "bf16", dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=False)
    g = torch.Generator().manual_seed(1)
    st.img.copy_(torch.randn(1, 96, 128, 128, generator=g).to(dev))
    origin = torch.randn(3, 64, 64, 170, generator=g).to(dev)
    st.step(49, origin)
    ops = st.ops
    calls = {}
    orig = {f: getattr(ops, f) for f in fams}
    for f in fams:
        def rec(*a, _f=f, **k):
            s = sig(_f, a, k)
            e = calls.setdefault(s, [0, _f, a, k])
            e[0] += 1
            return orig[_f](*a, **k)
        setattr(ops, f, rec)
    st.step(48, origin)
    torch.cuda.synchronize()
    for f in fams:
        setattr(ops, f, orig[f])
    rows = []
    for s, (cnt, f, a, k) in calls.items():
        orig[f](*a, **k)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                orig[f](*a, **k)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        rows.append((cnt * us, cnt, us, nbytes(a, k) / us / 1e3, s))
    rows.sort(reverse=True)
    print(f"total {sum(r[0] for r in rows):.0f} us over {sum(r[1] for r in rows)} calls")
    for tot, cnt, us, gbs, s in rows:
        print(f"{tot:7.0f} us = {cnt:2d} x {us:6.1f} us  ({gbs:5.0f} GB/s of arg bytes)  {s[:200]}")
    os._exit(0)


main()
