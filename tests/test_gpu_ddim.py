"""GPU parity of the DDIM variants and cond_fn conditioning (reference
gaussian_diffusion.py:364-398, 654-761, respace.py:97-101): the product's SpacedDiffusion driving the kernel plan
(UNet) and the fused posterior kernel on the B200, against the CPU oracle (pinned to the live reference in
tests/test_oracle.py::test_oracle_ddim_and_cond_fn_match_live_reference)."""
import pytest
import torch

from oracle import nfd_oracle as O
from tests.conftest import rel_l2
from tests.helpers import build_model, seeded_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"bf16": 2e-2, "fp32": 1e-4}


def _cond_fn(xx, t, **kw):
    return 0.05 * torch.sin(3.0 * xx) * (1.0 + 0.001 * t.float().view(-1, 1, 1, 1))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_ddim_variants_and_cond_fn(mode, monkeypatch):
    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, diff = build_model(cfg, sd, mode, DEV)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g, x, x2, noise = seeded_inputs(cfg)
    xd, nd = x.to(DEV), noise.to(DEV)
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: noise.to(t.device))     # eta > 0: same noise on both sides
    tol = TOL[mode]
    seen = []

    def cond_dev(xx, t, **kw):
        seen.append(int(t[0]))
        return _cond_fn(xx, t)

    with torch.no_grad():
        for i in (150, 49, 0):
            t = torch.tensor([i], device=DEV)
            for eta, cf in ((0.0, None), (0.5, None), (0.0, _cond_fn)):
                ref = O.ddim_sample(sd, cfg, sched, x, i, noise, eta=eta, cond_fn=cf, feat_layer=cfg["feat_layer"])
                got = diff.ddim_sample(model, xd, t, eta=eta, cond_fn=cond_dev if cf else None, model_kwargs={},
                                       feat_layer=cfg["feat_layer"])
                errs = {k: rel_l2(got[k], ref[k]) for k in ("sample", "pred_xstart", "inter_feat", "model_output")}
                print("ddim_sample", mode, i, eta, cf is not None, errs)
                assert max(errs.values()) < tol, (i, eta, errs)
            ref = O.ddim_reverse_sample(sd, cfg, sched, x, i)
            got = diff.ddim_reverse_sample(model, xd, t)
            assert rel_l2(got["sample"], ref["sample"]) < tol, i
            ref = O.ddim_guidance_sample(sched, x2.clone(), 0.1 * noise, x, i)
            got = diff.ddim_guidance_sample(x2.clone().to(DEV), 0.1 * nd, xd, t)
            assert rel_l2(got, ref) < 1e-5, i
            ref = O.p_sample_cond(sd, cfg, sched, x, i, noise, _cond_fn)
            got = diff.p_sample(model, xd, t, cond_fn=cond_dev, model_kwargs={})
            assert rel_l2(got["sample"], ref["sample"]) < tol, i
    assert set(seen) == {sched.timestep_map[i] for i in (150, 49, 0)}       # cond_fn sees ORIGINAL timesteps


def test_ddim_loop_round_trip_fp32():
    """A size-independent property: DDIM inversion (ddim_reverse_sample, t -> t+1) followed by deterministic DDIM
    sampling (eta = 0) returns to the start up to the discretisation error of the ODE — and the product's loop must
    make exactly the oracle's trajectory."""
    cfg = O.mid_cfg()
    cfg.update(timestep_respacing="20")
    sd = O.synth_state_dict(cfg)
    model, diff = build_model(cfg, sd, "fp32", DEV)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g, x, _, _ = seeded_inputs(cfg)
    x = x * 0.3
    xd, xo = x.to(DEV), x
    with torch.no_grad():
        for i in range(0, 6):                      # invert 6 steps
            xd = diff.ddim_reverse_sample(model, xd, torch.tensor([i], device=DEV))["sample"]
            xo = O.ddim_reverse_sample(sd, cfg, sched, xo, i)["sample"]
        assert rel_l2(xd, xo) < 1e-4
        for i in range(6, 0, -1):                  # and come back: sample from step i lands on step i-1
            xd = diff.ddim_sample(model, xd, torch.tensor([i], device=DEV), eta=0.0)["sample"]
            xo = O.ddim_sample(sd, cfg, sched, xo, i, torch.zeros_like(xo), eta=0.0)["sample"]
        assert rel_l2(xd, xo) < 1e-4
