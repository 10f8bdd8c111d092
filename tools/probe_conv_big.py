"""One large conv launch in isolation (for ncu --set full): N x 128 x 128, Cin -> Cout 3x3, bf16 in, fp32 out,
optional residual.   python tools/probe_conv_big.py [N] [Cin] [Cout] [residual 0/1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    Cin = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    Cout = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    res = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    ops = CudaOps(torch.device("cuda", 0), "bf16")
    dev = ops.device
    g = torch.Generator().manual_seed(0)
    a = torch.randn(N, 128, 128, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, 9 * Cin, generator=g) / (9 * Cin) ** 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g).to(dev)
    residual = torch.randn(N, 128, 128, Cout, generator=g).to(dev) if res else None
    out = torch.empty(N, 128, 128, Cout, device=dev)
    for _ in range(3):
        ops.conv(a, w, bias, 3, out, residual=residual)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv(a, w, bias, 3, out, residual=residual)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * N * 128 * 128 * Cout * 9 * Cin
    print(f"N={N} {Cin}->{Cout} residual={res}: {1e3 * ms:.1f} us per launch, {fl / ms / 1e9:.0f} TF/s")


main()
