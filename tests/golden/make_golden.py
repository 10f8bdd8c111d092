"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference, imported in
place through oracle/ref_import.py) on seeded synthetic inputs.  Run in the build container:

    python tests/golden/make_golden.py

Inputs are NOT stored: they are regenerated from the seeds below (torch CPU generators are
deterministic for a given torch version; the fixtures record torch.__version__).  Outputs:
  small_unet_step.npz  reference UNetModel forward (out, inter_feat), p_sample_guidance dict, and
                       two iterations of the reference's own DragStuff.training loop on the shrunken
                       5-level config (oracle.small_cfg)
  nfd_step.npz         the same for the real NFD architecture at 96x128x128 — strided sub-samples
                       (every 61st element) + L2 norms, one guided step
  decoder.npz          reference MultiTriplane.forward logits at 4096 seeded points + a 24^3 grid
  recon_step.npz       one iteration of the reference's reconstruction guidance (train_triplane loop body,
                       drag_utils.py:445-463) on the small config with 96 latent channels
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nfd_oracle as O  # noqa: E402
from oracle import ref_import as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STRIDE = 61


def sub(t):
    return t.detach().reshape(-1)[::STRIDE].numpy().astype(np.float32)


def run_reference_training(ns, cfg, sd, x, origin_list, sources, targets, r1, voxel, w_time, seed, scale=600, cof=0.2):
    """Drive the reference's DragStuff.training generator (drag_utils.py:302-399) for w_time steps."""
    du = ns.drag_utils
    a = du.DragStuff.args
    a.use_fp16 = False
    a.image_size, a.num_channels = cfg["image_size"], cfg["num_channels"]
    a.attention_resolutions, a.in_out_channels = cfg["attention_resolutions"], cfg["in_out_channels"]
    a.channel_mult = "" if cfg["image_size"] == 128 else ",".join(str(m) for m in cfg["channel_mult"])
    a.num_steps, a.timestep_respacing, a.w_time = 200, "200", w_time
    a.feat_layer, a.loss_type = cfg["feat_layer"], "l2"
    ds = du.DragStuff()
    ds.model.load_state_dict(sd, strict=True)
    ds.model.eval()
    ds.w = x.clone()
    ds.feature_guidance = [o.clone() for o in origin_list]
    ds.r1 = r1
    ds.offset1 = du.make_offsets(r1, ds.device)
    ds.voxel_size = voxel
    cap = {}

    def fake_get_mesh(tri_feat=None, img=None, t=0):
        cap["img"] = img.detach().clone()
        return None

    ds.get_mesh = fake_get_mesh
    torch.manual_seed(seed)
    progress = [p for p in ds.training(sources, targets, scale=scale, cof=cof)]
    return cap["img"], progress


def make_case(cfg, name, w_time, r1, voxel, full):
    ns, model, diffusion = R.reference_model_and_diffusion(cfg)
    sd = O.synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1)
    C, Rz = cfg["in_out_channels"], cfg["image_size"]
    x = torch.randn(1, C, Rz, Rz, generator=g)
    x2 = torch.randn(1, C, Rz, Rz, generator=g)
    noise = torch.randn(1, C, Rz, Rz, generator=g)
    i = 49
    fl = cfg["feat_layer"]
    keep = (lambda t: t.detach().numpy().astype(np.float32)) if full else sub
    out = {"torch_version": np.array(torch.__version__), "stride": np.array(1 if full else STRIDE)}
    with torch.no_grad():
        t_orig = torch.tensor([diffusion.timestep_map[i]])
        mo, feat = model(x, t_orig, feat_layer=fl)
        out["unet_out"], out["unet_feat"] = keep(mo), keep(feat)
        out["unet_out_norm"], out["unet_feat_norm"] = np.array(float(mo.norm())), np.array(float(feat.norm()))
        ps = diffusion.p_sample_guidance(model, x, torch.tensor([i]), noise=noise, feat_layer=fl)
        for k in ("sample", "pred_xstart", "model_output", "variance", "mean"):
            out["ps_" + k] = keep(ps[k])
        # origin features for the guided steps come from a second latent (reference modules)
        origin_list = []
        for s in range(w_time):
            o2 = diffusion.p_sample_guidance(model, x2, torch.tensor([w_time - 1 - s]), noise=noise, feat_layer=fl)
            origin_list.append(ns.drag_utils.resize_feat_align(o2["inter_feat"]))
        out["align_feat"] = keep(origin_list[0])
    gs = torch.Generator().manual_seed(4)
    sources = (torch.rand(4, 3, generator=gs) - 0.5).numpy()
    targets = sources + (torch.rand(4, 3, generator=gs).numpy() - 0.5) * 0.4
    out["sources"], out["targets"] = sources.astype(np.float32), targets.astype(np.float32)
    img, progress = run_reference_training(ns, cfg, sd, x, origin_list, sources, targets, r1, voxel, w_time, seed=77)
    out["train_img"], out["train_img_norm"] = keep(img), np.array(float(img.norm()))
    out["progress"] = np.array(progress, dtype=np.float64)
    out["meta"] = np.array([w_time, r1, voxel, i], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("unet", "train"))})


def make_decoder():
    ns = R.import_reference()
    w, planes = O.synth_decoder(R=128)
    dec = ns.axisnetworks.MultiTriplane(1, input_dim=3, output_dim=1, device="cpu")
    dec.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        dec.net[idx].weight.data.copy_(w["w" + k])
        dec.net[idx].bias.data.copy_(w["b" + k])
    for p in range(3):
        dec.embeddings[p] = planes[[p]]
    dec.eval()
    g = torch.Generator().manual_seed(8)
    pts = torch.rand(4096, 3, generator=g) * 2 - 1
    with torch.no_grad():
        logits = dec(0, pts[None]).reshape(-1)
        grid = dec(0, O.dense_grid_coords(24)[None]).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "decoder.npz"), logits=logits.numpy(), grid24=grid.numpy(),
                        occupancy=np.array(float((grid > 0).float().mean())), torch_version=np.array(torch.__version__))
    print("decoder occupancy", float((grid > 0).float().mean()), float((logits > 0).float().mean()))


def recon_cfg():
    """small_cfg with the 96 latent channels the triplane decoder needs (3 planes x 32 features)."""
    c = O.small_cfg()
    c.update(in_out_channels=96)
    return c


def make_recon():
    """One iteration of the reference's reconstruction guidance (drag_utils.py:445-463) with the reference's own
    diffusion, UNet and MultiTriplane modules: classifier guidance on pred_xstart through the decoder."""
    cfg = recon_cfg()
    ns, model, diffusion = R.reference_model_and_diffusion(cfg)
    sd = O.synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    Rz = cfg["image_size"]
    w, _ = O.synth_decoder(R=Rz)
    dec = ns.axisnetworks.MultiTriplane(1, input_dim=3, output_dim=1, device="cpu")
    dec.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        dec.net[idx].weight.data.copy_(w["w" + k])
        dec.net[idx].bias.data.copy_(w["b" + k])
    for prm in dec.parameters():
        prm.requires_grad_(False)                     # drag_utils.py:248-249
    dec.eval()
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 96, Rz, Rz, generator=g)
    noise = torch.randn(1, 96, Rz, Rz, generator=g)
    coords = torch.rand(3000, 3, generator=g) * 2 - 1
    gt = (torch.rand(3000, 1, generator=g) < 0.3).float()
    i, scale = 120, 600
    img = x.clone().requires_grad_(True)
    outs = diffusion.p_sample_guidance(model, img, torch.tensor([i]), noise=noise)
    predict_x0 = (outs["pred_xstart"] * 1.0 + 0.0).reshape(3, 32, Rz, Rz)
    for j in range(3):
        dec.embeddings[j] = predict_x0[[j]]
    prediction = dec(0, coords.unsqueeze(0)).squeeze(0)
    assert gt.shape == prediction.shape
    loss = -torch.nn.BCEWithLogitsLoss()(prediction, gt)
    loss.backward()
    grads = scale * img.grad.clone().detach()
    with torch.no_grad():
        nxt = (outs["sample"] + outs["variance"] * grads).clone().detach()
    np.savez_compressed(os.path.join(HERE, "recon_step.npz"), img=nxt.numpy(), grad=img.grad.numpy(),
                        loss=np.array(float(loss)), logits=prediction.detach().reshape(-1).numpy(),
                        meta=np.array([i, scale], dtype=np.float64), torch_version=np.array(torch.__version__))
    print("recon_step.npz loss", float(loss), "|grad|", float(img.grad.norm()), "|img|", float(nxt.norm()))


if __name__ == "__main__":
    assert R.available(), "needs /root/reference"
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        make_case(O.small_cfg(), "small_unet_step.npz", w_time=2, r1=3, voxel=2.0 / 64, full=True)
    if which in ("all", "decoder"):
        make_decoder()
    if which in ("all", "recon"):
        make_recon()
    if which in ("all", "nfd"):
        make_case(O.NFD_CFG, "nfd_step.npz", w_time=2, r1=12, voxel=2.0 / 256, full=False)
