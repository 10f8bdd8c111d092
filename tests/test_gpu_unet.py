"""GPU parity of the UNet plan (forward + input-gradient backward) and of one full drag-guided step
against the CPU oracle.  Tolerances are BASELINE.json's: 2e-2 relative L2 in bf16 mode, 1e-4 in fp32
mode, for the predicted eps / model output, the intermediate feature and the guidance gradient."""
import pytest
import torch

from oracle import nfd_oracle as O
from tests.conftest import rel_l2
from tests.helpers import build_model, drag_problem, seeded_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"bf16": 2e-2, "fp32": 1e-4}


def _unet_case(cfg, mode, t_val=246):
    sd = O.synth_state_dict(cfg)
    model, _ = build_model(cfg, sd, mode, DEV)
    g, x, _, _ = seeded_inputs(cfg)
    t = torch.tensor([t_val])
    fl = cfg["feat_layer"]
    xr = x.clone().requires_grad_(True)
    o_ref, f_ref = O.unet_forward(sd, cfg, xr, t, fl)
    proj = torch.randn(f_ref.shape, generator=g)
    (f_ref * proj).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    o, f = model(xd, t.to(DEV), feat_layer=fl)
    (f * proj.to(DEV)).sum().backward()
    return dict(out=rel_l2(o.detach(), o_ref.detach()), feat=rel_l2(f.detach(), f_ref.detach()),
                grad=rel_l2(xd.grad, xr.grad))


@pytest.mark.parametrize("native", ["1", "0"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_unet_mid_forward_backward(mode, native, monkeypatch):
    """The public `UNetModel.forward` + autograd: through the handle-level plan (default) and through the per-operator
    Python plan (ISB_NATIVE_EAGER=0, what the steppers capture)."""
    monkeypatch.setenv("ISB_NATIVE_EAGER", native)
    errs = _unet_case(O.mid_cfg(), mode)
    print("mid", mode, "native" if native == "1" else "python plan", errs)
    assert max(errs.values()) < TOL[mode], errs


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_unet_nfd_forward_backward(mode):
    """The real NFD architecture at 96x128x128 (BASELINE config 1)."""
    errs = _unet_case(O.NFD_CFG, mode)
    print("nfd", mode, errs)
    assert max(errs.values()) < TOL[mode], errs


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_guided_step_mid(mode):
    """One DragStuff.training loop body (drag_utils.py:340-392): gradient and next latent."""
    from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper

    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, diff = build_model(cfg, sd, mode, DEV)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g, x, x2, noise = seeded_inputs(cfg)
    i = 49
    origin, src, tgt, r1, voxel, pg, sg, masks = drag_problem(cfg, sd, sched, x2, noise, i, g, r1=4, voxel=2.0 / 64)
    ref = O.guided_step(sd, cfg, sched, x, i, origin, noise, pg, sg, masks, scale=600.0, cof=0.2)
    geo = DragGeometry(src, tgt, r1, voxel, origin.shape[-1], origin.shape[1])
    for use_graph in (False, True):
        st = GuidedStepper(model, diff, geo, cfg["feat_layer"], 0.2, "l2", 600.0, use_graph=use_graph)
        reps = 3 if use_graph else 1      # warm-up, capture, replay must all give the same step
        for _ in range(reps):
            st.img.copy_(x.to(DEV))
            st.step(i, origin.permute(0, 2, 3, 1).contiguous().to(DEV), noise.to(DEV))
        errs = dict(grad=rel_l2(st.grad, ref["grad"]), img=rel_l2(st.img, ref["img"]),
                    sample=rel_l2(st.sample, ref["sample"]), loss=abs(float(st.loss) - float(ref["loss"])) / abs(float(ref["loss"])))
        print("guided", mode, use_graph, errs)
        assert max(errs.values()) < TOL[mode], (use_graph, errs)


def test_host_step_pipeline_matches_sequential():
    """HostStepPipeline (inputs prefetched / results read back on copy streams, one step late) must produce exactly
    the latents and losses of the plain sequential loop over the same host inputs."""
    from ishapediting_b200.drag_utils import DragGeometry, GuidedStepper, HostStepPipeline

    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, diff = build_model(cfg, sd, "bf16", DEV)
    sched = O.Schedule(cfg["diffusion_steps"], cfg["timestep_respacing"])
    g, x, x2, noise = seeded_inputs(cfg)
    origin, src, tgt, r1, voxel, pg, sg, masks = drag_problem(cfg, sd, sched, x2, noise, 49, g, r1=4, voxel=2.0 / 64)
    geo = DragGeometry(src, tgt, r1, voxel, origin.shape[-1], origin.shape[1])
    n = 6
    origins = [(origin.permute(0, 2, 3, 1).contiguous() * (1.0 + 0.05 * k)).pin_memory() for k in range(n)]
    noises = [torch.randn(noise.shape, generator=g).pin_memory() for _ in range(n)]
    st = GuidedStepper(model, diff, geo, cfg["feat_layer"], 0.2, "l2", 600.0, use_graph=True)
    for _ in range(2):                                   # warm-up + capture
        st.img.copy_(x.to(DEV))
        st.step(49, origins[0].to(DEV), noises[0].to(DEV))
    st.img.copy_(x.to(DEV))
    seq = []
    for k in range(n):
        st.step(49 - k, origins[k].to(DEV), noises[k].to(DEV))
        seq.append((st.img.cpu().clone(), st.loss.cpu().clone()))
    st.img.copy_(x.to(DEV))
    pipe = HostStepPipeline(st)
    got = []
    pipe.prefetch(0, origins[0], noises[0])
    for k in range(n):
        if k + 1 < n:
            pipe.prefetch(k + 1, origins[k + 1], noises[k + 1])
        pipe.run(k, 49 - k)
        if k:
            r = pipe.result(k - 1)
            got.append((r[0].clone(), r[1].clone()))
    r = pipe.result(n - 1)
    got.append((r[0].clone(), r[1].clone()))
    for k in range(n):
        assert torch.equal(got[k][0], seq[k][0]), k
        assert torch.equal(got[k][1], seq[k][1]), k
