"""Time examples/host_guided_step.c (the plain-C host of the handle-level ABI) at the full NFD size: builds the blob
(421 M parameters, synthetic seeded weights), compiles the host with gcc and runs it eagerly and from its own CUDA
graph.   python tools/run_c_host_nfd.py [steps]"""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import nfd_oracle as O
from tests.helpers import build_model
from tests.test_gpu_c_host import build_host, read_blob, write_blob
from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper, align_maps


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    dev = "cuda:0"
    cfg = O.NFD_CFG
    model, diff = build_model(cfg, O.synth_state_dict(cfg), "bf16", dev)
    rng = np.random.RandomState(4)
    src = rng.uniform(-0.5, 0.5, size=(4, 3)).astype(np.float32)
    tgt = (src + rng.uniform(-0.2, 0.2, size=(4, 3))).astype(np.float32)
    geo = DragGeometry(src, tgt, 12, 2.0 / 256, 64, 170)
    st = GuidedStepper(model, diff, geo, 8, 0.2, "l2", 600.0, use_graph=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 128, 128, generator=g)
    origin = torch.randn(3, 64, 64, 170, generator=g)
    noise = torch.randn(1, 96, 128, 128, generator=g)
    idx = [49 - (k % 50) for k in range(steps)]
    # the Python host on the same inputs (graph replay), for the comparison of results and time
    st.img.copy_(x.to(dev))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for k, i in enumerate(idx):
        if k == 2:
            ev[0].record()
        st.step(i, origin.to(dev), noise.to(dev))
    ev[1].record()
    torch.cuda.synchronize()
    print(f"python host (GuidedStepper, CUDA graph): {ev[0].elapsed_time(ev[1]) / (steps - 2):.3f} ms/step "
          f"(incl. the per-step H2D of origin and noise)")
    chan_map, inv_map, _ = align_maps(st.Cf)
    mult = [int(m) for m in cfg["channel_mult"]]
    ds = sorted(int(a) for a in model.attention_resolutions)
    R = cfg["image_size"]
    ci = ([model.in_channels, model.model_channels, model.out_channels, model.num_res_blocks, len(mult)]
          + mult + [0] * (8 - len(mult)) + [len(ds)] + ds + [0] * (8 - len(ds))
          + [model.num_heads, model.num_head_channels, model.num_heads_upsample, 1, R, R, 1,
             8, steps, geo.group_size, geo.mask_count, 0, 1])
    entries = [("cfg", np.array(ci, dtype=np.int32)),
               ("scalars", np.array([geo.inv_count, 0.2], dtype=np.float32)), ("dyn", st.dyn.cpu().numpy()),
               ("w:time_embed.freqs", st.plan.freqs.cpu().numpy()),
               ("x", x.numpy()), ("t", st.t_table.cpu().numpy()[idx].astype(np.float32)),
               ("coef", st.coef_table.cpu().numpy()[idx].astype(np.float32)),
               ("noise", noise.numpy().reshape(1, -1).repeat(steps, 0)),
               ("origin", origin.numpy().reshape(1, 3 * 64 * 64, 170).repeat(steps, 0)),
               ("chan_map", chan_map.numpy()), ("inv_map", inv_map.numpy()),
               ("patch_xy", geo.patch_xy.numpy()), ("shift_xy", geo.shift_xy.numpy()), ("weight", geo.weight.numpy()),
               ("bbox", geo.bbox.numpy().reshape(3, -1, 4)), ("mask", geo.mask.numpy())]
    for name, t in model.state_dict().items():
        entries.append(("w:" + name, t.detach().float().cpu().numpy()))
    ref = st.img.cpu().numpy().ravel()
    del st, model
    torch.cuda.empty_cache()
    with tempfile.TemporaryDirectory() as tmp:
        exe = build_host(tmp)
        blob, out = os.path.join(tmp, "in.blob"), os.path.join(tmp, "out.blob")
        write_blob(blob, entries)
        for variant in ([], ["graph"]):
            r = subprocess.run([exe, blob, out] + variant, capture_output=True, text=True, timeout=600)
            print(r.stdout.strip(), r.stderr.strip())
            got = read_blob(out)
            print("  final latent bit-identical to the Python host:", bool(np.array_equal(got["img"], ref)))


main()
