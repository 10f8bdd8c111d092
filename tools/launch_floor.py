"""Floor of one dependent kernel launch inside a CUDA graph on this GPU: chains of empty kernels, with and without
programmatic dependent launch.   python tools/launch_floor.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishapediting_b200.ops import CudaOps

ops = CudaOps(torch.device("cuda", 0), "bf16")
lib = ops.lib
N = 200
for (ctas, threads) in ((1, 32), (148, 256), (592, 256), (2048, 256)):
    for pdl in (0, 1):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            lib.isb_debug_launch_chain(N, ctas, threads, pdl, s.cuda_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                lib.isb_debug_launch_chain(N, ctas, threads, pdl, torch.cuda.current_stream().cuda_stream)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"{ctas:5d} CTAs x {threads:3d} threads  pdl={pdl}:  {e0.elapsed_time(e1) * 1e3 / (5 * N):6.2f} us per dependent launch")
os._exit(0)
