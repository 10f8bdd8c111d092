"""Time one iteration of the reconstruction guidance (train_triplane loop body, drag_utils.py:445-463) at the NFD size:
UNet forward + full input-gradient backward + triplane decoder forward / backward at 40 000 sample points.
    python tools/time_recon.py [--mode bf16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_decoder, build_model
from ishapediting_b200.drag_utils import recon_guided_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="bf16")
    args = ap.parse_args()
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), args.mode, dev)
    dec, _, _ = build_decoder(128, dev)
    g = torch.Generator().manual_seed(2)
    img = torch.randn(1, 96, 128, 128, generator=g).to(dev)
    pts = (torch.rand(40000, 3, generator=g) * 2 - 1).to(dev)
    occ = (torch.rand(40000, 1, generator=g) < 0.3).float().to(dev)
    noise = torch.randn(1, 96, 128, 128, generator=g).to(dev)
    for k in range(3):
        img, loss = recon_guided_step(model, diff, dec, img, 199 - k, pts, occ, noise=noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for k in range(reps):
        img, loss = recon_guided_step(model, diff, dec, img, 196 - k, pts, occ, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    print(f"recon-guided step ({args.mode}, eager, 40000 points): {e0.elapsed_time(e1) / reps:.2f} ms  loss {float(loss):.4f}")
    from ishapediting_b200.drag_utils import ReconStepper

    st = ReconStepper(model, diff, dec, 40000, scale=600.0, use_graph=True)
    st.img.copy_(img)
    for k in range(3):
        st.step(185 - k, pts, occ, noise=noise)
    torch.cuda.synchronize()
    e0.record()
    for k in range(20):
        st.step(180 - k, pts, occ, noise=noise)
    e1.record()
    torch.cuda.synchronize()
    print(f"recon-guided step ({args.mode}, ReconStepper, CUDA graph): {e0.elapsed_time(e1) / 20:.2f} ms  loss {float(st.loss):.4f}")
    # decoder kernels alone
    e0.record()
    for _ in range(20):
        planes = img.reshape(3, 32, 128, 128).detach().requires_grad_(True)
        for j in range(3):
            dec.embeddings[j] = planes[[j]]
        out = dec(0, pts.unsqueeze(0))
        out.sum().backward()
    e1.record()
    torch.cuda.synchronize()
    print(f"decoder forward + backward at 40000 points (incl. layout kernels): {e0.elapsed_time(e1) / 20:.3f} ms")
    os._exit(0)


main()
