"""Per-CTA phase timeline of conv_tc_kernel (debug bit 2: %globaltimer stamps, see tc_stamp in conv_tc.cu).
A graph of `n` back-to-back launches with cold weights is replayed; for every launch but the first the phases are
reported relative to the previous launch's last CTA exit.   python tools/trace_conv.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps, PackedWeight

PHASES = ["entry", "prologue done", "producer past pdl_wait", "first stage full", "last MMA committed",
          "accumulator ready", "tile parked / stored", "cluster sync 1", "fold stored", "exit", "first fold batch summed",
          "park: chunk 0 in registers", "park: chunk 0 staged", "park: chunk 0 stored"]

CASES_BIG = [
    (128, 3, 256, 256, {"two_cta": 2, "block_n": 128, "split_k": 1}),
    (128, 3, 256, 256, {"two_cta": 2, "block_n": 256, "split_k": 1}),
    (128, 3, 256, 256, {"two_cta": 1, "block_n": 128, "split_k": 1, "stages": 4}),
    (128, 3, 256, 256, {"two_cta": 1, "block_n": 256, "split_k": 1, "stages": 4}),
    (128, 3, 256, 256, {"two_cta": 1, "block_n": 256, "split_k": 1, "stages": 6}),
    (64, 3, 512, 512, {"two_cta": 1, "block_n": 256, "split_k": 1, "stages": 6}),
    (64, 3, 512, 512, {}),
]
CASES = [  # H, k, Cin, Cout, tune
    (8, 3, 1024, 1024, {}),
    (16, 3, 768, 768, {}),
    (32, 3, 512, 512, {}),
    (64, 3, 256, 256, {}),
    (128, 3, 256, 256, {"two_cta": 2}),
    (8, 1, 1024, 1024, {}),
]


def main():
    ops = CudaOps(torch.device("cuda", 0), "bf16")
    dev = ops.device
    for (H, k, Cin, Cout, tune) in (CASES_BIG if '--big' in sys.argv else CASES):
        K = k * k * Cin
        n = max(6, min(24, int(300e6 // (Cout * K * 2))))
        a = torch.randn(1, H, H, Cin, device=dev).to(torch.bfloat16)
        out = torch.empty(1, H, H, Cout, device=dev)
        bias = torch.randn(Cout, device=dev)
        ws = [PackedWeight(torch.randn(Cout, K, device=dev).to(torch.bfloat16), Cout, K, False) for _ in range(n)]
        traces = [torch.zeros(4096 * 16, dtype=torch.int64, device=dev) for _ in range(n)]
        for dbg in ((4,) if '--big' in sys.argv else (4, 7)):
            t = dict(tune)
            t["debug"] = dbg
            ops.conv(a, ws[0], bias, k, out, tune={**t, "trace": traces[0]})
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for w, tr in zip(ws, traces):
                    ops.conv(a, w, bias, k, out, tune={**t, "trace": tr})
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            raw = torch.stack(traces).cpu()
            it = raw[n // 2, 3000 * 16:3000 * 16 + 640].view(160, 4).double()
            tr = raw.view(n, 4096, 16)
            ncta = int((tr[0, :, 0] != 0).sum())
            tr = tr[:, :ncta, :].double()
            period = (tr[2:, :, 9].max(dim=1).values - tr[1:-1, :, 9].max(dim=1).values).mean().item() / 1e3
            print(f"\nH={H} k{k} {Cin}->{Cout} {tune} debug={dbg}: {ncta} CTAs on {len(set(tr[1,:,15].tolist()))} SMs, "
                  f"launch period {period:.2f} us")
            print(f"  {'phase':26s} {'min':>8s} {'median':>8s} {'max':>8s}   (us after the previous launch's last exit)")
            for ph in (0, 1, 2, 3, 4, 5, 11, 12, 13, 6, 7, 10, 8, 9):
                rel = []
                for li in range(1, n):
                    base = tr[li - 1, :, 9].max()
                    v = tr[li, :, ph]
                    v = v[v != 0]
                    if len(v):
                        rel.append((v - base) / 1e3)
                if not rel:
                    continue
                mins = sum(r.min().item() for r in rel) / len(rel)
                meds = sum(r.median().item() for r in rel) / len(rel)
                maxs = sum(r.max().item() for r in rel) / len(rel)
                print(f"  {PHASES[ph]:26s} {mins:8.2f} {meds:8.2f} {maxs:8.2f}")
            if "--iters" in sys.argv:
                t0 = it[it > 0].min()
                print("  per-iteration SM-clock stamps of CTA 0 (cycles after its first stamp): producer empty-wait passed / TMA issued | MMA full-wait passed / committed")
                for i in range(160):
                    if (it[i] > 0).any():
                        print("   it %3d  " % i + "  ".join("%8d" % (v - t0) if v > 0 else "       -" for v in it[i].tolist()))
    os._exit(0)


main()
