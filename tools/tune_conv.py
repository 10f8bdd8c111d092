"""Sweep conv_tc tile / split-K / pipeline settings per layer shape with COLD weights (a ring of weight
copies larger than L2 is cycled, as in the real step where 1.57 GB of panels stream through a 126 MB L2).
    python tools/tune_conv.py [--quick]"""
import argparse
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps, PackedWeight
from ishapediting_b200._lib import IsbError

SHAPES_BIG = [(32, 3, 512, 512), (32, 3, 768, 768), (64, 3, 256, 256), (64, 3, 512, 512), (64, 3, 512, 256),
              (128, 3, 256, 256), (128, 3, 512, 256), (128, 3, 128, 256)]
SHAPES = [  # H(=W), ksize, Cin, Cout
    (8, 3, 1024, 1024), (8, 1, 1024, 1024), (8, 1, 3072, 1024), (8, 1, 1024, 3072),
    (16, 3, 768, 768), (16, 3, 1024, 1024), (16, 1, 768, 768), (16, 1, 768, 2304),
    (32, 3, 512, 512), (32, 3, 768, 768), (32, 1, 512, 512), (32, 1, 512, 1536),
    (64, 3, 256, 256), (64, 3, 512, 512),
    (128, 3, 256, 256), (128, 3, 512, 256),
]


def bench(ops, a, weights, bias, k, out, tune, tiled, reps=4):
    n = len(weights)
    ws = [PackedWeight(w, out.shape[3], w.numel() // out.shape[3], tiled) for w in weights]
    try:
        ops.conv(a, ws[0], bias, k, out, tune=tune)     # validates the config, sizes workspace
    except (IsbError, AssertionError) as e:
        return None, str(e)[:60]
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for w in ws:
            ops.conv(a, w, bias, k, out, tune=tune)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * n), ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--warm", action="store_true", help="one weight copy: weights stay L2-resident")
    ap.add_argument("--auto-only", action="store_true")
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--batch", type=int, default=1, help="images per launch (throughput mode: 8)")
    args = ap.parse_args()
    ops = CudaOps(torch.device("cuda", 0), "bf16")
    dev = ops.device
    for (H, k, Cin, Cout) in (SHAPES_BIG if args.big else SHAPES):
        K = k * k * Cin
        wbytes = Cout * K * 2
        ncopies = max(4, min(64, int(400e6 // wbytes)))           # ring > L2 (126 MB) when weights are big
        if args.warm:
            ncopies = 1
        a = torch.randn(args.batch, H, H, Cin, device=dev).to(torch.bfloat16)
        out = torch.empty(args.batch, H, H, Cout, device=dev)
        bias = torch.randn(Cout, device=dev)
        base = [torch.randn(Cout, K, device=dev).to(torch.bfloat16) for _ in range(ncopies)]
        if args.warm:
            base = base * 16
        flops = 2.0 * args.batch * H * H * Cout * K
        rows = []
        bns = [64, 128, 256] if Cout % 256 == 0 else [64, 128]
        splits = [1, 2, 4, 8]
        stages = [3, 6] if args.quick else [3, 4, 6]
        t_auto, _ = bench(ops, a, base, bias, k, out, None, False)
        if args.auto_only:
            print(f"H={H:3d} k{k} {Cin:4d}->{Cout:4d}  auto {t_auto:7.1f} us  ({flops / t_auto / 1e6:6.1f} TF/s, weights {wbytes / t_auto / 1e3:6.0f} GB/s)", flush=True)
            continue
        for bn, sp, stg in itertools.product(bns, splits, stages):
            t, err = bench(ops, a, base, bias, k, out, {"block_n": bn, "split_k": sp, "stages": stg}, False)
            if t is not None:
                rows.append((t, bn, sp, stg))
        if H * H * args.batch >= 1024 and Cout % 128 == 0:      # CTA-pair kernel (cta_group::2)
            for bn in ([128, 256] if Cout % 256 == 0 else [128]):
                for stg in (4, 5, 6):
                    t, err = bench(ops, a, base, bias, k, out, {"block_n": bn, "split_k": 1, "stages": stg, "two_cta": 1}, False)
                    if t is not None:
                        rows.append((t, f"pair{bn}", 1, stg))
        rows.sort()
        best = rows[0]
        print(f"H={H:3d} k{k} {Cin:4d}->{Cout:4d}  auto {t_auto:7.1f} us | best {best[0]:7.1f} us bn={best[1]} split={best[2]} "
              f"stages={best[3]}  ({flops / best[0] / 1e6:6.1f} TF/s, weights {wbytes / best[0] / 1e3:6.0f} GB/s) | "
              + "  ".join(f"{r[0]:.1f}@{r[1]}/{r[2]}/{r[3]}" for r in rows[1:5]), flush=True)


if __name__ == "__main__":
    main()
