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
    ap.add_argument("--algbytes", default=None, help="write per-family algorithmic bytes of one step to this JSON")
    ap.add_argument("--batch", type=int, default=1, help="advance this many independent edits as one batch")
    ap.add_argument("--decode", type=int, default=0, help="also profile one dense decode at this resolution")
    args = ap.parse_args()
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), args.mode, dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    B = args.batch
    st = GuidedStepper(model, diff, [geo] * B if B > 1 else geo, 8, 0.2, "l2", 600.0, use_graph=False)
    g = torch.Generator().manual_seed(1)
    st.img.copy_(torch.randn(B, 96, 128, 128, generator=g).to(dev))
    origin = torch.randn(*((B,) if B > 1 else ()), 3, 64, 64, 170, generator=g).to(dev)
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
    if args.algbytes:
        import json
        ops = st.ops
        bpe = lambda t: 0 if t is None else t.numel() * t.element_size()  # noqa: E731
        acc = {"conv": {"launches": 0, "bytes": 0}, "groupnorm": {"launches": 0, "bytes": 0}}
        conv0, gnf0, gnb0 = ops.conv, ops.gn_forward, ops.gn_backward

        def conv(a, w, bias, ksize, out, a2=None, residual=None, **kw):
            wd = w.data if hasattr(w, "data") and not torch.is_tensor(w) else w
            acc["conv"]["launches"] += 1
            acc["conv"]["bytes"] += bpe(a) + bpe(a2) + bpe(wd) + bpe(out) + bpe(residual) + bpe(bias)
            return conv0(a, w, bias, ksize, out, a2=a2, residual=residual, **kw)

        def gnf(x1, x2, *a, **kw):
            y = a[7]
            acc["groupnorm"]["launches"] += 1
            acc["groupnorm"]["bytes"] += bpe(x1) + bpe(x2) + bpe(y) + bpe(kw.get("raw")) + bpe(kw.get("xres"))
            return gnf0(x1, x2, *a, **kw)

        def gnb(x1, x2, gamma, beta, film, film_off, silu, resample, stats, dy, gres, gres_at_input, gx1, acc1, gx1_lo,
                gx2, acc2, gx2_lo, **kw):
            acc["groupnorm"]["launches"] += 1
            acc["groupnorm"]["bytes"] += (bpe(x1) + bpe(x2) + bpe(dy) + bpe(gres) + bpe(gx1) + bpe(gx1_lo) + bpe(gx2)
                                          + bpe(gx2_lo) + (bpe(gx1) if acc1 else 0) + (bpe(gx2) if acc2 else 0))
            return gnb0(x1, x2, gamma, beta, film, film_off, silu, resample, stats, dy, gres, gres_at_input, gx1, acc1,
                        gx1_lo, gx2, acc2, gx2_lo, **kw)

        ops.conv, ops.gn_forward, ops.gn_backward = conv, gnf, gnb
        st.step(48, origin)
        torch.cuda.synchronize()
        ops.conv, ops.gn_forward, ops.gn_backward = conv0, gnf0, gnb0
        if args.decode:
            acc["decode"] = {"launches": 1, "bytes": 4 * args.decode ** 3 + 3 * 32 * 128 * 128 * 4}
        with open(args.algbytes, "w") as f:
            json.dump(acc, f)
    dec = None
    if args.decode:
        from tests.helpers import build_decoder
        from ishapediting_b200.triplane_decoder.visualize import query_volume
        dec, _, planes = build_decoder(128, dev)
        for p in range(3):
            dec.embeddings[p] = planes[[p]].to(dev)
        query_volume(dec, 0, res=args.decode)
        torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for s in range(args.steps):
        st.step(48 - s, origin)
    if dec is not None:
        query_volume(dec, 0, res=args.decode)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", args.steps, "step(s)")


if __name__ == "__main__":
    main()
