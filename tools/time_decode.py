"""Time the dense triplane decode (visualize.py:76-98 replacement) and check it against the oracle on a sub-grid.
    python tools/time_decode.py [--res 256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import nfd_oracle as O
from ishapediting_b200.triplane_decoder.axisnetworks import MultiTriplane
from ishapediting_b200.triplane_decoder.visualize import query_volume


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, nargs="+", default=[128, 256, 512])
    args = ap.parse_args()
    dev = "cuda:0"
    w, planes = O.synth_decoder()
    dec = MultiTriplane(1).to(dev)
    dec.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        dec.net[idx].weight.data.copy_(w["w" + k])
        dec.net[idx].bias.data.copy_(w["b" + k])
    for p in range(3):
        dec.embeddings[p] = planes[[p]].to(dev)
    # parity on a 64^3 grid (oracle on CPU takes a few seconds)
    vol = query_volume(dec, 0, res=64).cpu().reshape(-1)
    ref = O.decode_grid(w, planes, 64)
    occ, occ_ref = vol > 0, ref > 0
    iou = float((occ & occ_ref).sum()) / float((occ | occ_ref).sum())
    print(f"64^3 parity: max|d logit| {float((vol - ref).abs().max()):.3e}  IoU {iou:.6f}  occupancy {float(occ_ref.float().mean()):.3f}")
    for res in args.res:
        out = torch.empty(res ** 3, device=dev)
        query_volume(dec, 0, res=res, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            query_volume(dec, 0, res=res, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        pts = res ** 3
        print(f"res {res:4d}^3: {ms:8.3f} ms  {pts / ms / 1e6:8.2f} Gpts/s  {pts * 69888 / ms / 1e9:7.1f} TFLOP/s (algorithmic)  "
              f"{pts * 4 / ms / 1e6:7.1f} GB/s written")


if __name__ == "__main__":
    main()
