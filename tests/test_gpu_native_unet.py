"""The UNet through the handle-level C ABI (isb_unet_*): the native plan inside libishape_b200.so
against (a) the Python host's plan — same kernels, same order, so BIT-identical — and (b) the CPU oracle, at the
tolerances of BASELINE.json (2e-2 rel-L2 bf16 mode, 1e-4 fp32 mode)."""
import ctypes as C

import pytest
import torch

from oracle import nfd_oracle as O
from tests.conftest import rel_l2
from tests.helpers import build_model, seeded_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {"bf16": 2e-2, "fp32": 1e-4}


def _python_plan_pass(model, x, t, fl, proj_nhwc, d_out=None):
    """forward + backward through guided_diffusion/unet.py::_Plan (the Python host), one stream."""
    N, _, H, W = x.shape
    plan = model.plan(N, H, W, want_backward=True)
    inter = plan.forward(x, t, fl)
    out = plan.ops.to_nchw(plan.out_nhwc, torch.empty(N, plan.out_nhwc.shape[3], H, W, device=x.device))
    feat = inter.val.clone()
    plan.begin_backward()
    if d_out is not None:
        plan.backward_out_layer(d_out)
    plan.seed_grad(inter, proj_nhwc)
    dx = plan.backward(torch.empty_like(x))
    return out, feat, dx


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("cfg_name", ["mid", "nfd"])
def test_native_unet_matches_python_plan_and_oracle(cfg_name, mode):
    from ishapediting_b200.native_unet import NativeUNet

    cfg = O.mid_cfg() if cfg_name == "mid" else O.NFD_CFG
    sd = O.synth_state_dict(cfg)
    model, _ = build_model(cfg, sd, mode, DEV)
    g, x, _, _ = seeded_inputs(cfg)
    fl = cfg["feat_layer"]
    t = torch.tensor([246.0])
    N, Cc, R = 1, cfg["in_out_channels"], cfg["image_size"]
    nat = NativeUNet(model, N, R, R, side_stream=False)
    xd, td = x.to(DEV), t.to(DEV)
    out = nat.forward(xd, td, feat_layer=fl)
    val, grad = nat.feat(fl)
    proj = torch.randn(val.shape, generator=g).to(DEV)          # NHWC
    dx = nat.backward_input(fl, d_feat=proj)
    out_p, feat_p, dx_p = _python_plan_pass(model, xd, td, fl, proj)
    torch.cuda.synchronize()
    assert torch.equal(out, out_p), rel_l2(out, out_p)
    assert torch.equal(val, feat_p), rel_l2(val, feat_p)
    assert torch.equal(dx, dx_p), rel_l2(dx, dx_p)
    # the same pass with the ResBlock skip backward on the handle's side stream, the gradient written in place
    nat2 = NativeUNet(model, N, R, R, side_stream=True)
    nat2.forward(xd, td, feat_layer=fl, stop_at_feat=True)
    val2, grad2 = nat2.feat(fl)
    grad2.copy_(proj)
    dx2 = nat2.backward_input(fl, in_place=True)
    out2 = nat2.forward_tail()
    torch.cuda.synchronize()
    assert torch.equal(out2, out) and torch.equal(dx2, dx)
    # oracle
    if cfg_name == "nfd" and mode == "fp32":
        return      # the fp32 NFD oracle pass is covered by test_gpu_unet.py; the equalities above tie this path to it
    xr = x.clone().requires_grad_(True)
    o_ref, f_ref = O.unet_forward(sd, cfg, xr, t.long(), fl)
    (f_ref * proj.cpu().permute(0, 3, 1, 2)).sum().backward()
    errs = dict(out=rel_l2(out, o_ref.detach()), feat=rel_l2(val.permute(0, 3, 1, 2), f_ref.detach()),
                grad=rel_l2(dx, xr.grad))
    print("native", cfg_name, mode, errs)
    assert max(errs.values()) < TOL[mode], errs


def test_native_unet_output_gradient_root_and_batch():
    """d_out root (the reconstruction guidance differentiates through the UNet OUTPUT, drag_utils.py:443-463) together
    with the feature root, batch 2, NHWC output written straight into the caller's buffer."""
    from ishapediting_b200.native_unet import NativeUNet

    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, _ = build_model(cfg, sd, "bf16", DEV)
    g = torch.Generator().manual_seed(5)
    Cc, R, fl = cfg["in_out_channels"], cfg["image_size"], cfg["feat_layer"]
    x = torch.randn(2, Cc, R, R, generator=g).to(DEV)
    t = torch.tensor([246.0, 31.0]).to(DEV)
    nat = NativeUNet(model, 2, R, R)
    out_nhwc = nat.forward(x, t, feat_layer=fl, out_nhwc=True)
    val, _ = nat.feat(fl)
    proj = torch.randn(val.shape, generator=g).to(DEV)
    d_out = torch.randn(2, out_nhwc.shape[3], R, R, generator=g).to(DEV)
    dx = nat.backward_input(fl, d_feat=proj, d_out=d_out)
    out_p, feat_p, dx_p = _python_plan_pass(model, x, t, fl, proj, d_out=d_out)
    torch.cuda.synchronize()
    assert torch.equal(out_nhwc.permute(0, 3, 1, 2), out_p)
    assert torch.equal(val, feat_p)
    # the Python plan runs the skip backward inline here, the handle on its side stream: same kernels, same operands
    assert torch.equal(dx, dx_p), rel_l2(dx, dx_p)


def test_native_unet_graph_capture_and_errors():
    """forward(stop at the feature) -> backward on the capturing stream with the tail forked to a second stream: the
    whole pass replays from ONE CUDA graph with identical results; misuse fails loudly with a message."""
    from ishapediting_b200 import _lib
    from ishapediting_b200.native_unet import NativeUNet

    cfg = O.mid_cfg()
    sd = O.synth_state_dict(cfg)
    model, _ = build_model(cfg, sd, "bf16", DEV)
    g, x, _, _ = seeded_inputs(cfg)
    fl, R = cfg["feat_layer"], cfg["image_size"]
    xd, td = x.to(DEV), torch.tensor([246.0], device=DEV)
    nat = NativeUNet(model, 1, R, R)
    val, grad = nat.feat(fl)
    proj = torch.randn(val.shape, generator=g).to(DEV)
    out_e = nat.forward(xd, td, feat_layer=fl)
    dx_e = nat.backward_input(fl, d_feat=proj)
    torch.cuda.synchronize()
    out_g, dx_g = torch.empty_like(out_e), torch.empty_like(dx_e)
    side, cap = torch.cuda.Stream(), torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cap):
        with torch.cuda.graph(graph, stream=cap):
            nat.forward(xd, td, feat_layer=fl, stop_at_feat=True)
            fork = torch.cuda.Event()
            fork.record(cap)
            with torch.cuda.stream(side):
                side.wait_event(fork)
                nat.forward_tail(out=out_g)
                join = torch.cuda.Event()
                join.record(side)
            nat.backward_input(fl, d_feat=proj, dx=dx_g)
            cap.wait_event(join)
    for _ in range(2):
        out_g.zero_(); dx_g.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out_g, out_e) and torch.equal(dx_g, dx_e)
    # errors: wrong feat_layer, too small a workspace, a missing parameter
    with pytest.raises(_lib.IsbError, match="feat_layer"):
        nat.forward(xd, td, feat_layer=99)
    rc = nat.lib.isb_unet_forward(nat._h, C.c_void_p(xd.data_ptr()), C.c_void_p(td.data_ptr()), -1, 0, None, 0,
                                  nat._ws_ptr, 1024, None)
    assert rc == -3 and b"workspace" in nat.lib.isb_last_error()
    sd_missing = {k: v for k, v in model.state_dict().items() if k != "middle_block.1.qkv.weight"}

    class _Partial:
        def __getattr__(self, k):
            return getattr(model, k)

        def state_dict(self):
            return sd_missing

        def parameters(self):
            return model.parameters()

    with pytest.raises(_lib.IsbError, match="middle_block.1.qkv.weight"):
        NativeUNet(_Partial(), 1, R, R)
