"""Shared builders for the parity tests (test code only)."""
import torch

from oracle import nfd_oracle as O


def build_model(cfg, sd, mode, device, ops=None):
    """The product UNetModel + SpacedDiffusion for an oracle config, weights = sd."""
    from ishapediting_b200.guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults

    kw = model_and_diffusion_defaults()
    kw.update(image_size=cfg["image_size"], num_channels=cfg["num_channels"], num_res_blocks=cfg["num_res_blocks"],
              num_head_channels=cfg["num_head_channels"], attention_resolutions=cfg["attention_resolutions"],
              channel_mult=",".join(str(m) for m in cfg["channel_mult"]), dropout=0.1, use_scale_shift_norm=True,
              resblock_updown=True, use_fp16=(mode == "bf16"), in_out_channels=cfg["in_out_channels"],
              learn_sigma=cfg["learn_sigma"], diffusion_steps=cfg["diffusion_steps"],
              timestep_respacing=cfg["timestep_respacing"])
    model, diffusion = create_model_and_diffusion(**kw)
    model.load_state_dict(sd, strict=True)
    model.eval()
    model.to(device)
    if ops is not None:
        model.set_ops(ops)
    return model, diffusion


def seeded_inputs(cfg, seed=1):
    g = torch.Generator().manual_seed(seed)
    C, R = cfg["in_out_channels"], cfg["image_size"]
    x = torch.randn(1, C, R, R, generator=g)
    x2 = torch.randn(1, C, R, R, generator=g)
    noise = torch.randn(1, C, R, R, generator=g)
    return g, x, x2, noise


def drag_problem(cfg, sd, sched, x2, noise, i, g, n_handles=4, r1=None, voxel=None):
    """Origin features from a second latent + random handles."""
    with torch.no_grad():
        o0 = O.p_sample_guidance(sd, cfg, sched, x2, i, noise, feat_layer=cfg["feat_layer"])
    origin = O.resize_feat_align(o0["inter_feat"])
    S = origin.shape[-1]
    src = (torch.rand(n_handles, 3, generator=g) - 0.5).numpy()
    tgt = src + (torch.rand(n_handles, 3, generator=g).numpy() - 0.5) * 0.4
    r1 = 12 if r1 is None else r1
    voxel = 2.0 / 256 if voxel is None else voxel
    pg, sg, masks = O.drag_setup(src, tgt, r1, voxel, S)
    return origin, src, tgt, r1, voxel, pg, sg, masks


def recon_cfg():
    """small_cfg with the 96 latent channels the triplane decoder needs (tests/golden/make_golden.py:recon_cfg)."""
    c = O.small_cfg()
    c.update(in_out_channels=96)
    return c


def recon_inputs(R, n_pts=3000, seed=21):
    """The seeded inputs of tests/golden/recon_step.npz: latent, noise, sample points, occupancies."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 96, R, R, generator=g)
    noise = torch.randn(1, 96, R, R, generator=g)
    coords = torch.rand(n_pts, 3, generator=g) * 2 - 1
    gt = (torch.rand(n_pts, 1, generator=g) < 0.3).float()
    return x, noise, coords, gt


def build_decoder(R, device, ops=None):
    """The product MultiTriplane carrying oracle.synth_decoder's MLP."""
    from ishapediting_b200.triplane_decoder.axisnetworks import MultiTriplane

    w, planes = O.synth_decoder(R=R)
    dec = MultiTriplane(1, device=str(device))
    dec.net[0]._B.data.copy_(w["B"])
    for idx, k in ((1, "1"), (3, "2"), (5, "3")):
        dec.net[idx].weight.data.copy_(w["w" + k])
        dec.net[idx].bias.data.copy_(w["b" + k])
    dec.to(device)
    for prm in dec.parameters():
        prm.requires_grad_(False)
    if ops is not None:
        dec.set_ops(ops)
    return dec, w, planes
