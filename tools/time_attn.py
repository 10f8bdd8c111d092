"""Time the fused attention forward at the UNet's three shapes (CUDA events, graph-free loop of 200 launches).
ISB_FA_TC=0 selects the mma.sync kernel for every shape, the default the tcgen05 kernel where T % 128 == 0.
    python tools/time_attn.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps


def main():
    dev = "cuda:0"
    ops = CudaOps(dev, "bf16")
    g = torch.Generator().manual_seed(0)
    for S, heads in ((32, 8), (16, 12), (8, 16)):
        T, C = S * S, heads * 64
        qkv = torch.randn(1, S, S, 3 * C, generator=g).to(torch.bfloat16).to(dev)
        out = torch.zeros(1, S, S, C, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(1, heads, T, device=dev)
        for _ in range(20):
            ops.attention_flash_forward(qkv, heads, out, lse)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(50):
                ops.attention_flash_forward(qkv, heads, out, lse)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 500
        fl = 4.0 * T * T * 64 * heads
        print(f"T={T:5d} heads={heads:2d}: {us:7.2f} us per launch  {fl / us / 1e6:7.1f} TFLOP/s  (ISB_FA_TC={os.environ.get('ISB_FA_TC', '1')})")


main()
