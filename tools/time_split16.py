"""A/B of split-K factors for the weight-streaming layers (cold weights, graph of back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.tune_conv import bench
from ishapediting_b200.ops import CudaOps

ops = CudaOps(torch.device("cuda", 0), "bf16")
dev = ops.device
for (H, k, Cin, Cout) in [(8, 3, 1024, 1024), (8, 3, 2048, 1024), (8, 1, 1024, 1024), (8, 1, 1024, 3072), (16, 3, 768, 768), (16, 3, 1536, 768)]:
    K = k * k * Cin
    ncopies = max(4, min(64, int(400e6 // (Cout * K * 2))))
    a = torch.randn(1, H, H, Cin, device=dev).to(torch.bfloat16)
    out = torch.empty(1, H, H, Cout, device=dev)
    bias = torch.randn(Cout, device=dev)
    base = [torch.randn(Cout, K, device=dev).to(torch.bfloat16) for _ in range(ncopies)]
    res = []
    for tune in (None, {"block_n": 64, "split_k": 8, "stages": 4}, {"block_n": 64, "split_k": 16, "stages": 4},
                 {"block_n": 128, "split_k": 16, "stages": 3}, {"block_n": 128, "split_k": 8, "stages": 3}):
        t, err = bench(ops, a, base, bias, k, out, tune, False)
        res.append(f"{tune}: {t if t is None else round(t, 2)} {err}")
    print(f"H={H} k{k} {Cin}->{Cout}:\n   " + "\n   ".join(res), flush=True)
    # correctness of split 16 against split 8
    o8, o16 = torch.empty_like(out), torch.empty_like(out)
    from ishapediting_b200.ops import PackedWeight
    w = PackedWeight(base[0], Cout, K, False)
    ops.conv(a, w, bias, k, o8, tune={"block_n": 64, "split_k": 8, "stages": 4})
    ops.conv(a, w, bias, k, o16, tune={"block_n": 64, "split_k": 16, "stages": 4})
    torch.cuda.synchronize()
    print("   split16 vs split8 rel diff", float((o16 - o8).norm() / o8.norm()))
os._exit(0)
