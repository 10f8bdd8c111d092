"""GPU probe (not a pytest file): run ONE conv case through CudaOps and print error statistics
against the torch mirror.  Used from gpurun so that a faulting kernel cannot hide the other cases:
    python tests/probe_conv.py <mode> <case-name|all>"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.conv_cases import CASES
from tests.ref_ops import RefOps


def run_case(ops, case, mode, packed=True):
    name, N, H, W, Cin, Cout, k, Cin2, has_bias, has_res, acc, out_bf16, tune = case
    lo = torch.bfloat16 if mode == "bf16" else torch.float32
    g = torch.Generator().manual_seed(hash(name) % 1000)
    a = torch.randn(N, H, W, Cin, generator=g).to(lo)
    a2 = torch.randn(N, H, W, Cin2, generator=g).to(lo) if Cin2 else None
    w = (torch.randn(Cout, k * k * Cin + Cin2, generator=g) / (k * k * Cin + Cin2) ** 0.5).to(lo)
    bias = torch.randn(Cout, generator=g) if has_bias else None
    res = torch.randn(N, H, W, Cout, generator=g) if has_res else None
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out0 = torch.randn(N, H, W, Cout, generator=g).to(odt)
    ref = RefOps(mode)
    out_ref = out0.clone()
    ref.conv(a, w, bias, k, out_ref, a2=a2, residual=res, accumulate=acc)
    d = ops.device
    out = out0.clone().to(d)
    w_dev = ops.pack_weight(w.to(d)) if packed else w.to(d)     # panel-tiled layout when the shape allows
    ops.conv(a.to(d), w_dev, bias.to(d) if has_bias else None, k, out, a2=a2.to(d) if Cin2 else None,
             residual=res.to(d) if has_res else None, accumulate=acc, tune=tune)
    torch.cuda.synchronize()
    o = out.float().cpu()
    r = out_ref.float()
    err = (o - r).norm() / r.norm()
    return float(err), float((o - r).abs().max()), o, r


if __name__ == "__main__":
    from ishapediting_b200.ops import CudaOps

    mode, which = sys.argv[1], sys.argv[2]
    ops = CudaOps(torch.device("cuda", 0), mode)
    for case in CASES:
        if which != "all" and case[0] != which:
            continue
        if mode == "fp32" and case[11]:
            continue
        err, mx, o, r = run_case(ops, case, mode)
        err_u, _, _, _ = run_case(ops, case, mode, packed=False)
        err = max(err, err_u)
        tol = 1e-2 if case[11] else (2e-3 if mode == "bf16" else 1e-5)
        print(f"[probe_conv {mode}] {case[0]:34s} rel_l2={err:.3e} max_abs={mx:.3e} {'OK' if err < tol else 'FAIL'}", flush=True)
        if err >= tol:
            bad = (o - r).abs()
            idx = torch.nonzero(bad > 10 * tol * r.abs().mean())
            print("   first bad idx:", idx[:8].tolist(), " n_bad:", idx.shape[0], " of", o.numel(),
                  " nan:", int(torch.isnan(o).sum()), flush=True)
            # per-row / per-col error pattern helps to spot descriptor / swizzle mistakes
            rowerr = bad.reshape(-1, o.shape[-1]).mean(1)
            colerr = bad.reshape(-1, o.shape[-1]).mean(0)
            print("   row err (first 16):", [f"{v:.2e}" for v in rowerr[:16].tolist()])
            print("   col err (first 16):", [f"{v:.2e}" for v in colerr[:16].tolist()])
    sys.stdout.flush()
    os._exit(0)     # skip interpreter/CUDA teardown
