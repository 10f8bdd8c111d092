"""One fused-attention forward + backward at the 32x32 level (T = 1024, 8 heads), for `ncu -k regex:fa_`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps

ops = CudaOps(torch.device("cuda", 0), "bf16")
dev = ops.device
bf = torch.bfloat16
qkv = torch.randn(1, 32, 32, 1536, device=dev).to(bf)
out = torch.empty(1, 32, 32, 512, device=dev, dtype=bf)
lse = torch.empty(1, 8, 1024, device=dev)
d_out = torch.randn(1, 32, 32, 512, device=dev).to(bf)
delta = torch.empty(1, 8, 1024, device=dev)
d_qkv = torch.empty(1, 32, 32, 1536, device=dev, dtype=bf)
for _ in range(3):
    ops.attention_flash_forward(qkv, 8, out, lse)
    ops.attention_flash_backward(qkv, out, d_out, lse, 8, delta, d_qkv)
torch.cuda.synchronize()
os._exit(0)
