"""One eager (no CUDA graph) drag-guided step of the full NFD configuration between
cudaProfilerStart/Stop, for `ncu --profile-from-start off`.  Synthetic origin feature (no 200-step
cache build) so the profiled process stays short.  Prints per-kernel-family CUDA-event timings when
run without a profiler:   python tools/profile_step.py [--mode bf16] [--steps 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--time-ops", action="store_true", help="CUDA-event time per op family (eager)")
    args = ap.parse_args()
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), args.mode, dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=False)
    g = torch.Generator().manual_seed(1)
    st.img.copy_(torch.randn(1, 96, 128, 128, generator=g).to(dev))
    origin = torch.randn(3, 64, 64, 170, generator=g).to(dev)
    st.step(49, origin)          # warm-up (sizes workspaces)
    torch.cuda.synchronize()
    if args.time_ops:
        ops = st.ops
        rec = {}

        def wrap(name):
            fn = getattr(ops, name)

            def timed(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                r = fn(*a, **k)
                e.record()
                key = name
                if name == "conv":
                    A, Wt = a[0], a[1]
                    key = f"conv k{a[3]} {A.shape[1]}x{A.shape[2]} {A.shape[3]}->{Wt.shape[0]} K{Wt.shape[1]}"
                rec.setdefault(key, []).append((s, e))
                return r

            setattr(ops, name, timed)

        for n in ("conv", "gn_forward", "gn_backward", "attention_forward", "attention_backward", "time_embed",
                  "ddpm_step", "drag_loss_grad", "to_nhwc", "to_nchw", "cast_lo"):
            wrap(n)
        st.step(48, origin)
        torch.cuda.synchronize()
        rows = [(k, len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in rec.items()]
        tot = sum(r[2] for r in rows)
        fam = {}
        for k, n, t in rows:
            f = k.split(" ")[0]
            fam[f] = fam.get(f, 0) + t
        print(f"total {tot:.3f} ms over {sum(r[1] for r in rows)} op calls")
        for f, t in sorted(fam.items(), key=lambda x: -x[1]):
            print(f"  {f:20s} {t:8.3f} ms  {100 * t / tot:5.1f}%")
        print("conv shapes (calls, ms, TFLOP/s):")
        for k, n, t in sorted(rows, key=lambda x: -x[2]):
            if not k.startswith("conv"):
                continue
            p = k.split(" ")
            hw = p[2].split("x")
            cout = int(p[3].split("->")[1])
            K = int(p[4][1:])
            fl = 2.0 * int(hw[0]) * int(hw[1]) * cout * K * n
            print(f"  {k:44s} x{n:3d} {t:8.3f} ms {fl / t / 1e9:8.1f}")
        return
    torch.cuda.profiler.start()
    for s in range(args.steps):
        st.step(48 - s, origin)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", args.steps, "step(s)")


if __name__ == "__main__":
    main()
