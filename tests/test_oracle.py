"""Pin the CPU oracle: against the committed golden fixtures (generated from the UNMODIFIED reference by
tests/golden/make_golden.py) everywhere, and live against /root/reference where it exists."""
import os

import numpy as np
import pytest
import torch

from oracle import nfd_oracle as O
from oracle import ref_import as R
from tests.conftest import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _inputs(cfg):
    g = torch.Generator().manual_seed(1)
    C, Rz = cfg["in_out_channels"], cfg["image_size"]
    return (torch.randn(1, C, Rz, Rz, generator=g), torch.randn(1, C, Rz, Rz, generator=g),
            torch.randn(1, C, Rz, Rz, generator=g))


def _check_case(cfg, fixture, tol):
    z = np.load(os.path.join(GOLD, fixture))
    stride = int(z["stride"])
    pick = (lambda t: t.detach().numpy()) if stride == 1 else (lambda t: t.detach().reshape(-1)[::stride].numpy())
    w_time, r1, voxel, i = int(z["meta"][0]), int(z["meta"][1]), float(z["meta"][2]), int(z["meta"][3])
    sd = O.synth_state_dict(cfg)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    x, x2, noise = _inputs(cfg)
    fl = cfg["feat_layer"]
    with torch.no_grad():
        mo, feat = O.unet_forward(sd, cfg, x, torch.tensor([sched.timestep_map[i]]), fl)
        assert rel_l2(torch.from_numpy(pick(mo)), torch.from_numpy(z["unet_out"])) < tol
        assert rel_l2(torch.from_numpy(pick(feat)), torch.from_numpy(z["unet_feat"])) < tol
        assert abs(float(mo.norm()) - float(z["unet_out_norm"])) / float(z["unet_out_norm"]) < tol
        ps = O.p_sample_guidance(sd, cfg, sched, x, i, noise, feat_layer=fl)
        for k in ("sample", "pred_xstart", "model_output", "variance", "mean"):
            assert rel_l2(torch.from_numpy(pick(ps[k])), torch.from_numpy(z["ps_" + k])) < tol, k
        origins = []
        for s in range(w_time):
            o2 = O.p_sample_guidance(sd, cfg, sched, x2, w_time - 1 - s, noise, feat_layer=fl)
            origins.append(O.resize_feat_align(o2["inter_feat"]))
        assert rel_l2(torch.from_numpy(pick(origins[0])), torch.from_numpy(z["align_feat"])) < tol
    # the reference's own training loop: noise is drawn by th.randn_like after manual_seed(77), one per step
    S = origins[0].shape[-1]
    pg, sg, masks = O.drag_setup(z["sources"], z["targets"], r1, voxel, S)
    torch.manual_seed(77)
    img = x
    for s in range(w_time):
        nz = torch.randn_like(img)
        out = O.guided_step(sd, cfg, sched, img, w_time - 1 - s, origins[s], nz, pg, sg, masks, scale=600, cof=0.2)
        img = out["img"]
    assert rel_l2(torch.from_numpy(pick(img)), torch.from_numpy(z["train_img"])) < tol
    assert abs(float(img.norm()) - float(z["train_img_norm"])) / float(z["train_img_norm"]) < tol
    progress = [1 - k / (w_time - 1.0) for k in range(w_time - 1, -1, -1)]
    assert np.allclose(progress, z["progress"])


def test_oracle_matches_reference_golden_small():
    """UNet forward, p_sample_guidance, resize_feat_align and two iterations of the reference's
    DragStuff.training loop on the shrunken NFD-shaped config."""
    _check_case(O.small_cfg(), "small_unet_step.npz", 2e-5)


@pytest.mark.slow
def test_oracle_matches_reference_golden_nfd():
    """Same at the real NFD size (96x128x128, 421M parameters; strided sub-samples + norms)."""
    _check_case(O.NFD_CFG, "nfd_step.npz", 5e-5)


def test_oracle_decoder_matches_reference_golden():
    z = np.load(os.path.join(GOLD, "decoder.npz"))
    w, planes = O.synth_decoder(R=128)
    g = torch.Generator().manual_seed(8)
    pts = torch.rand(4096, 3, generator=g) * 2 - 1
    with torch.no_grad():
        logits = O.triplane_forward(w, planes, pts)
        grid = O.decode_grid(w, planes, 24)
    assert float((logits - torch.from_numpy(z["logits"])).abs().max()) < 1e-5
    assert float((grid - torch.from_numpy(z["grid24"])).abs().max()) < 1e-5
    assert 0.1 < float((grid > 0).float().mean()) < 0.5      # non-trivial surface


def test_oracle_recon_guidance_matches_reference_golden():
    """One iteration of the reference's reconstruction guidance (train_triplane loop body, drag_utils.py:445-463,
    run with the reference's own UNet / diffusion / MultiTriplane): next latent, input gradient, loss, logits."""
    from tests.helpers import recon_cfg, recon_inputs

    z = np.load(os.path.join(GOLD, "recon_step.npz"))
    cfg = recon_cfg()
    sd = O.synth_state_dict(cfg)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    x, noise, coords, gt = recon_inputs(cfg["image_size"])
    w, _ = O.synth_decoder(R=cfg["image_size"])
    i, scale = int(z["meta"][0]), float(z["meta"][1])
    r = O.recon_guided_step(sd, cfg, sched, x, i, noise, w, coords, gt, scale=scale)
    assert rel_l2(r["img"], torch.from_numpy(z["img"])) < 2e-5
    assert rel_l2(r["grad"], torch.from_numpy(z["grad"])) < 2e-4
    assert abs(float(r["loss"]) - float(z["loss"])) < 1e-6
    assert float((r["logits"] - torch.from_numpy(z["logits"])).abs().max()) < 2e-5


def test_schedule_matches_published_constants():
    """200-of-1000 respacing map and table shapes."""
    s = O.Schedule(1000, "200")
    assert s.timestep_map[:3] == [0, 5, 10] and s.timestep_map[-3:] == [989, 994, 999]
    assert len(s.timestep_map) == 200 and s.num_timesteps == 200
    assert np.all(s.betas > 0) and np.all(s.betas < 1)
    assert np.isclose(s.posterior_variance[0], 0.0)


@pytest.mark.skipif(not R.available(), reason="/root/reference only exists in the build container")
def test_oracle_matches_live_reference_unet_and_schedule():
    cfg = O.small_cfg()
    ns, model, diffusion = R.reference_model_and_diffusion(cfg)
    sd = O.synth_state_dict(cfg)
    assert list(sd.keys()) == list(model.state_dict().keys())
    model.load_state_dict(sd, strict=True)
    x, _, noise = _inputs(cfg)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    assert sched.timestep_map == diffusion.timestep_map
    assert np.array_equal(sched.betas, diffusion.betas)
    with torch.no_grad():
        for i in (49, 1, 0):
            ref = diffusion.p_sample_guidance(model, x, torch.tensor([i]), noise=noise, feat_layer=8)
            mine = O.p_sample_guidance(sd, cfg, sched, x, i, noise, feat_layer=8)
            for k in ("sample", "pred_xstart", "model_output", "variance", "mean", "inter_feat"):
                assert torch.equal(ref[k], mine[k]), (i, k)
        assert torch.equal(ns.drag_utils.resize_feat_align(ref["inter_feat"]), O.resize_feat_align(mine["inter_feat"]))


@pytest.mark.skipif(not R.available(), reason="needs /root/reference or the oracle/_ref snapshot")
def test_oracle_ddim_and_cond_fn_match_live_reference(monkeypatch):
    """DDIM variants (gaussian_diffusion.py:654-761) and cond_fn conditioning (:364-398, respace.py:97-101) of the
    oracle against the reference's own SpacedDiffusion + UNet; th.randn_like is pinned so that eta > 0 is comparable."""
    cfg = O.small_cfg()
    ns, model, diffusion = R.reference_model_and_diffusion(cfg)
    sd = O.synth_state_dict(cfg)
    model.load_state_dict(sd, strict=True)
    x, x2, noise = _inputs(cfg)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: noise.to(t.device))
    seen = []

    def cond_fn(xx, t, **kw):
        seen.append(int(t[0]))
        return 0.05 * torch.sin(3.0 * xx) * (1.0 + 0.001 * t.float().view(-1, 1, 1, 1))

    with torch.no_grad():
        for i in (150, 49, 0):
            t = torch.tensor([i])
            for eta, cf in ((0.0, None), (0.5, None), (0.0, cond_fn)):
                ref = diffusion.ddim_sample(model, x, t, eta=eta, cond_fn=cf, model_kwargs={})
                mine = O.ddim_sample(sd, cfg, sched, x, i, noise, eta=eta, cond_fn=cf)
                assert rel_l2(mine["sample"], ref["sample"]) < 1e-6, (i, eta)
                assert rel_l2(mine["pred_xstart"], ref["pred_xstart"]) < 1e-6, (i, eta)
            ref = diffusion.ddim_reverse_sample(model, x, t)
            assert rel_l2(O.ddim_reverse_sample(sd, cfg, sched, x, i)["sample"], ref["sample"]) < 1e-6
            ref = diffusion.ddim_guidance_sample(x2.clone(), 0.1 * noise, x, t)
            assert rel_l2(O.ddim_guidance_sample(sched, x2.clone(), 0.1 * noise, x, i), ref) < 1e-6
            ref = diffusion.p_sample(model, x, t, cond_fn=cond_fn, model_kwargs={})
            mine = O.p_sample_cond(sd, cfg, sched, x, i, noise, cond_fn)
            assert rel_l2(mine["sample"], ref["sample"]) < 1e-6, i
    assert set(seen) == {sched.timestep_map[i] for i in (150, 49, 0)}       # cond_fn sees ORIGINAL timesteps
