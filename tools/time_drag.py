"""Time isb_drag_loss_grad (sample + gather + loss fold) at the NFD step's size: 50 back-to-back calls captured into a
CUDA graph and replayed (warm L2, like inside the step).   python tools/time_drag.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ishapediting_b200.drag_utils import DragGeometry, align_maps
from ishapediting_b200.ops import CudaOps


def main():
    dev = torch.device("cuda", 0)
    ops = CudaOps(dev, "bf16")
    S, Cf = 64, 512
    chan_map, inv_map, Ca = align_maps(Cf)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, S, Ca).to(dev)
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(1, S, S, Cf, generator=g).to(dev)
    origin = torch.randn(3, S, S, Ca, generator=g).to(dev)
    d_feat = torch.empty_like(feat)
    gbuf, pt = ops.empty((3, geo.npts, Ca)), ops.empty((3, geo.npts, 4))
    partial = ops.zeros((ops.drag_partial_len(S, Cf, geo.npts),), torch.float64)
    loss = ops.zeros((1,))
    cm, im = chan_map.to(dev), inv_map.to(dev)

    def call():
        ops.drag_loss_grad(feat, origin, cm, im, geo.patch_xy, geo.shift_xy, geo.weight, geo.group_size, geo.bbox,
                           geo.mask, geo.mask_count, geo.inv_count, 0.2, 0, gbuf, pt, partial, loss, d_feat)

    call()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(50):
            call()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"isb_drag_loss_grad: {1e3 * e0.elapsed_time(e1) / 500:.2f} us per call (3 launches), loss {float(loss):.6f}, "
          f"|d_feat| {float(d_feat.abs().sum()):.4f}")


main()
