"""GPU parity of every non-GEMM kernel against the pure-torch mirror (tests/ref_ops.py, CPU fp32).
Call-site citations are in include/ishape_b200.h."""
import math

import pytest
import torch

from tests.conftest import rel_l2
from tests.ref_ops import RefOps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from ishapediting_b200.ops import CudaOps

    return CudaOps(torch.device(DEV), "bf16")


@pytest.fixture(scope="module")
def ref():
    return RefOps("bf16")


def G(seed):
    return torch.Generator().manual_seed(seed)


def test_layout_roundtrip(ops):
    x = torch.randn(2, 12, 16, 24, generator=G(0))
    for dt, tol in ((torch.float32, 0.0), (torch.bfloat16, 4e-3)):
        y = ops.to_nhwc(x.to(DEV), ops.empty((2, 16, 24, 64), dt))
        yc = y.float().cpu()
        assert rel_l2(yc[..., :12], x.permute(0, 2, 3, 1)) <= tol
        assert float(yc[..., 12:].abs().max()) == 0.0
        back = ops.to_nchw(y, ops.empty((2, 12, 16, 24)))
        assert rel_l2(back, x) <= tol


GN_CASES = [
    # N,H,W,C1,C2,film,silu,resample,raw
    (1, 16, 16, 256, 0, False, True, 0, False),
    (2, 8, 8, 512, 0, True, True, 0, False),
    (1, 8, 8, 1024, 768, False, True, 0, True),      # concat: 56-channel groups straddle the two sources
    (1, 16, 16, 256, 0, False, True, 1, False),      # down
    (1, 8, 8, 256, 0, False, True, 2, False),        # up
    (1, 32, 32, 512, 0, False, False, 0, False),     # attention norm (no silu)
    (1, 64, 64, 256, 0, True, True, 0, False),       # largest slice still on the single-launch path
    (1, 128, 128, 256, 0, True, True, 0, False),     # two-kernel path, multi-split statistics
    (1, 128, 128, 256, 0, False, True, 1, False),    # two-kernel path + avg-pool
]


@pytest.mark.parametrize("rows_mb", [None, "0"], ids=["default", "row-kernel"])
@pytest.mark.parametrize("case", GN_CASES)
def test_groupnorm_forward_backward(ops, ref, case, rows_mb, monkeypatch):
    """rows_mb = "0": the whole-row apply kernel that large (beyond-L2) tensors take, forced at these small sizes."""
    if rows_mb is not None:
        if case[7] != 0:
            pytest.skip("the row kernel serves the un-resampled apply only")
        monkeypatch.setenv("ISB_GN_ROWS_MIN_MB", rows_mb)
    N, H, W, C1, C2, film_on, silu, rs, want_raw = case
    g = G(3)
    C = C1 + C2
    x1 = torch.randn(N, H, W, C1, generator=g) * 1.5 + 0.3
    x2 = torch.randn(N, H, W, C2, generator=g) if C2 else None
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    film = 0.3 * torch.randn(N, 2 * C + 64, generator=g) if film_on else None
    foff = 32 if film_on else 0
    Ho, Wo = (H // 2, W // 2) if rs == 1 else ((H * 2, W * 2) if rs == 2 else (H, W))
    d = lambda t: None if t is None else t.to(DEV)
    for lo in (torch.float32, torch.bfloat16):
        outs = {}
        for name, o in (("ref", ref), ("cuda", ops)):
            dev = "cpu" if name == "ref" else DEV
            mv = (lambda t: None if t is None else t.to(dev))
            stats = torch.zeros(N, 32, 2, device=dev)
            y = torch.zeros(N, Ho, Wo, C, dtype=lo, device=dev)
            raw = torch.zeros(N, H, W, C, dtype=lo, device=dev) if want_raw else None
            xres = torch.zeros(N, Ho, Wo, C, device=dev) if rs else None
            o.gn_forward(mv(x1), mv(x2), mv(gamma), mv(beta), mv(film), foff, silu, rs, stats, y, raw=raw, xres=xres)
            outs[name] = (stats, y, raw, xres)
        tol = 1e-5 if lo == torch.float32 else 6e-3
        assert rel_l2(outs["cuda"][0], outs["ref"][0]) < 1e-5
        assert rel_l2(outs["cuda"][1].float(), outs["ref"][1].float()) < tol
        if want_raw:
            assert rel_l2(outs["cuda"][2].float(), outs["ref"][2].float()) < tol
        if rs:
            assert rel_l2(outs["cuda"][3], outs["ref"][3]) < 1e-6
    # backward
    dy = torch.randn(N, Ho, Wo, C, generator=g)
    for gres_mode in (None, "out", "in"):
        gres = None if gres_mode is None else (torch.randn(N, Ho, Wo, C, generator=g) if gres_mode == "out"
                                               else torch.randn(N, H, W, C, generator=g))
        res = {}
        for name, o in (("ref", ref), ("cuda", ops)):
            dev = "cpu" if name == "ref" else DEV
            mv = (lambda t: None if t is None else t.to(dev))
            stats = torch.zeros(N, 32, 2, device=dev)
            y = torch.zeros(N, Ho, Wo, C, device=dev)
            o.gn_forward(mv(x1), mv(x2), mv(gamma), mv(beta), mv(film), foff, silu, rs, stats, y)
            gx1 = torch.full((N, H, W, C1), 0.5, device=dev)      # accumulate into a non-zero buffer
            gx1_lo = torch.zeros(N, H, W, C1, dtype=torch.bfloat16, device=dev)
            gx2 = torch.zeros(N, H, W, C2, device=dev) if C2 else None
            gx2_lo = torch.zeros(N, H, W, C2, dtype=torch.bfloat16, device=dev) if C2 else None
            o.gn_backward(mv(x1), mv(x2), mv(gamma), mv(beta), mv(film), foff, silu, rs, stats, mv(dy), mv(gres),
                          gres_mode == "in", gx1, True, gx1_lo, gx2, False, gx2_lo)
            res[name] = (gx1, gx1_lo, gx2, gx2_lo)
        assert rel_l2(res["cuda"][0], res["ref"][0]) < 2e-5, (case, gres_mode)
        assert rel_l2(res["cuda"][1].float(), res["ref"][1].float()) < 6e-3
        if C2:
            assert rel_l2(res["cuda"][2], res["ref"][2]) < 2e-5
            assert rel_l2(res["cuda"][3].float(), res["ref"][3].float()) < 6e-3


@pytest.mark.parametrize("N,H,Cin,C,k,film,silu,gres,acc,tune", [
    (1, 64, 128, 256, 3, True, True, False, False, None),                      # ResBlock GN2 backward, direct epilogue
    (1, 32, 256, 512, 3, True, True, False, False, {"block_n": 128, "split_k": 4}),   # cluster fold
    (2, 8, 512, 1024, 3, False, True, True, True, None),                       # GN1 with skip gradient, accumulate
    (1, 32, 1536, 512, 1, False, False, True, False, None),                    # attention norm (qkv dgrad), no SiLU
    (3, 8, 256, 256, 1, True, True, True, False, {"block_n": 64, "split_k": 2}),
])
def test_gn_backward_fused_reduction(ops, N, H, Cin, C, k, film, silu, gres, acc, tune):
    """dgrad conv with gn_bwd=... + gn_backward(partials=...) against the ordinary reduce + apply on the same dy."""
    g = G(9)
    dev = ops.device
    a = torch.randn(N, H, H, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(C, k * k * Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev).to(torch.bfloat16)
    x = torch.randn(N, H, H, C, generator=g).to(dev)
    gamma, beta = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    fm = (torch.randn(N, 2 * C + 64, generator=g) * 0.3).to(dev) if film else None
    foff = 32 if film else 0
    gr = torch.randn(N, H, H, C, generator=g).to(dev) if gres else None
    slots = ops.conv_gn_slots(N, H, H, Cin, k, C) if tune is None else None
    stats = torch.zeros(N, 32, 2, device=dev)
    y = torch.empty(N, H, H, C, device=dev, dtype=torch.bfloat16)
    ops.gn_forward(x, None, gamma, beta, fm, foff, silu, 0, stats, y)
    res = []
    for fused in (False, True):
        dy = torch.empty(N, H, H, C, device=dev)
        gx = torch.full((N, H, H, C), 0.25, device=dev)
        gx_lo = torch.zeros(N, H, H, C, device=dev, dtype=torch.bfloat16)
        part = None
        if fused:
            if slots is None:      # explicit tune: ask the library for this exact configuration
                import ctypes as C_
                from ishapediting_b200 import _lib
                d = _lib.ConvDesc()
                d.a, d.a_dtype, d.w, d.out, d.out_dtype = 256, _lib.BF16, 256, 256, _lib.F32
                d.N, d.H, d.W, d.Cin, d.ksize, d.Cout, d.gn_cg = N, H, H, Cin, k, C, C // 32
                d.block_n, d.split_k = tune.get("block_n", 0), tune.get("split_k", 0)
                slots = int(ops.lib.isb_conv2d_gn_slots(C_.byref(d)))
            assert slots > 0
            part = torch.full((N, 32, slots, 2), float("nan"), device=dev)
        ops.conv(a, w, None, k, dy, tune=tune, gn_part=part,
                 gn_bwd=(x, gamma, beta, fm, foff, silu, stats) if fused else None)
        ops.gn_backward(x, None, gamma, beta, fm, foff, silu, 0, stats, dy, gr, False, gx, acc, gx_lo, None, False, None,
                        partials=part)
        res.append((gx.clone(), gx_lo.float()))
    assert rel_l2(res[1][0], res[0][0]) < 2e-5
    assert rel_l2(res[1][1], res[0][1]) < 6e-3


@pytest.mark.parametrize("lo", [torch.float32, torch.bfloat16], ids=["fp32_ffma", "bf16_mma"])
@pytest.mark.parametrize("N,S,heads", [(1, 8, 2), (2, 16, 3), (1, 32, 4)])
def test_attention_forward_backward(ops, ref, N, S, heads, lo):
    g = G(5)
    C = heads * 64
    T = S * S
    qkv = torch.randn(N, S, S, 3 * C, generator=g)
    d_out = torch.randn(N, S, S, C, generator=g)
    r = {}
    for name, o in (("ref", ref), ("cuda", ops)):
        dev = "cpu" if name == "ref" else DEV
        probs = torch.zeros(N, heads, T, T, device=dev)
        out = torch.zeros(N, S, S, C, device=dev, dtype=lo)
        o.attention_forward(qkv.to(dev), heads, probs, out)
        tmp = torch.zeros(N, heads, T, T, device=dev)
        dq = torch.zeros(N, S, S, 3 * C, device=dev, dtype=lo)
        o.attention_backward(qkv.to(dev), probs, d_out.to(dev), heads, tmp, dq)
        r[name] = (probs, out.float(), dq.float())
    # bf16 mode rounds q,k,v,P,dS,dO to bf16 before the tensor-core products (softmax stays fp32)
    tol = 2e-5 if lo == torch.float32 else 1.5e-2
    assert rel_l2(r["cuda"][0], r["ref"][0]) < tol
    assert rel_l2(r["cuda"][1], r["ref"][1]) < tol
    assert rel_l2(r["cuda"][2], r["ref"][2]) < tol


@pytest.mark.parametrize("N,S,heads", [(1, 8, 2), (2, 16, 3), (1, 32, 8), (2, 32, 2)])
def test_attention_flash_forward_backward(ops, ref, N, S, heads):
    """Fused attention (no materialised probabilities) against the fp32 torch mirror on the same bf16 inputs."""
    g = G(6)
    C = heads * 64
    T = S * S
    qkv = torch.randn(N, S, S, 3 * C, generator=g).to(torch.bfloat16)
    d_out = (torch.randn(N, S, S, C, generator=g) * 0.5).to(torch.bfloat16)
    r = {}
    for name, o in (("ref", ref), ("cuda", ops)):
        dev = "cpu" if name == "ref" else DEV
        out = torch.zeros(N, S, S, C, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(N, heads, T, device=dev)
        o.attention_flash_forward(qkv.to(dev), heads, out, lse)
        delta = torch.zeros(N, heads, T, device=dev)
        dq = torch.zeros(N, S, S, 3 * C, device=dev, dtype=torch.bfloat16)
        # both sides differentiate around the SAME forward output (the reference's bf16-rounded one)
        o.attention_flash_backward(qkv.to(dev), r["ref"][0].to(dev) if name == "cuda" else out, d_out.to(dev), lse, heads,
                                   delta, dq)
        r[name] = (out, lse, delta, dq)
    f = lambda x: x.float().cpu()
    assert rel_l2(f(r["cuda"][0]), f(r["ref"][0])) < 1e-2          # P is rounded to bf16 before PV
    assert (f(r["cuda"][1]) - f(r["ref"][1])).abs().max() < 1e-3   # log-sum-exp (log2 units), fp32 accumulate
    assert rel_l2(f(r["cuda"][2]), f(r["ref"][2])) < 1e-4
    dq_c, dq_r = f(r["cuda"][3]).reshape(N, T, heads, 3, 64), f(r["ref"][3]).reshape(N, T, heads, 3, 64)
    for part in range(3):                                          # dq, dk, dv separately
        assert rel_l2(dq_c[:, :, :, part], dq_r[:, :, :, part]) < 1.5e-2, part


def test_time_embed(ops, ref):
    g = G(7)
    mc, hid, rows, N = 64, 256, 1000, 3
    t = torch.tensor([0.0, 246.0, 997.5])      # fractional values occur with rescale_timesteps (gaussian_diffusion.py:171-174)
    from ishapediting_b200.guided_diffusion.nn import timestep_freqs

    freqs = timestep_freqs(mc)
    w1, b1 = torch.randn(hid, mc, generator=g) / 8, torch.randn(hid, generator=g) * 0.1
    w2, b2 = torch.randn(hid, hid, generator=g) / 16, torch.randn(hid, generator=g) * 0.1
    wa, ba = torch.randn(rows, hid, generator=g) / 16, torch.randn(rows, generator=g) * 0.1
    out_ref = torch.zeros(N, rows)
    ref.time_embed(t, freqs, w1, b1, w2, b2, wa, ba, None, out_ref)
    c = lambda x: x.to(DEV)
    out = ops.zeros((N, rows))
    ops.time_embed(c(t), c(freqs), c(w1), c(b1), c(w2), c(b2), c(wa), c(ba), ops.empty((N * (mc + 2 * hid),)), out)
    assert rel_l2(out, out_ref) < 2e-5


@pytest.mark.parametrize("nchw", [False, True])
def test_ddpm_step(ops, ref, nchw):
    g = G(9)
    N, C, H, W = 2, 12, 16, 16
    x, noise, grad = (torch.randn(N, C, H, W, generator=g) for _ in range(3))
    mo = torch.randn(N, 2 * C, H, W, generator=g) if nchw else torch.randn(N, H, W, 2 * C + 8, generator=g)
    coef = torch.tensor([1.3, 0.8, 0.4, 0.6, -6.0, -3.0, 1.0, 600.0])
    names = ("x_next", "sample", "mean", "var", "x0", "eps")
    r = {}
    for name, o in (("ref", ref), ("cuda", ops)):
        dev = "cpu" if name == "ref" else DEV
        outs = {k: torch.zeros(N, C, H, W, device=dev) for k in names}
        o.ddpm_step(x.to(dev), mo.to(dev), coef.to(dev), True, noise=noise.to(dev), grad=grad.to(dev),
                    model_out_nhwc=not nchw, **outs)
        r[name] = outs
    for k in names:
        assert rel_l2(r["cuda"][k], r["ref"][k]) < 1e-5, k
    # per-sample schedule rows (batched DDPM inversion)
    coef2 = torch.stack([coef, coef * torch.tensor([0.9, 1.1, 1.2, 0.8, 1.0, 1.0, 0.0, 1.0])])
    r2 = {}
    for name, o in (("ref", ref), ("cuda", ops)):
        dev = "cpu" if name == "ref" else DEV
        outs = {k: torch.zeros(N, C, H, W, device=dev) for k in names}
        o.ddpm_step(x.to(dev), mo.to(dev), coef2.to(dev), True, noise=noise.to(dev), grad=grad.to(dev),
                    model_out_nhwc=not nchw, **outs)
        r2[name] = outs
    for k in names:
        assert rel_l2(r2["cuda"][k], r2["ref"][k]) < 1e-5, k
    assert rel_l2(r2["cuda"]["x_next"][:1], r["cuda"]["x_next"][:1]) == 0.0


def test_drag_loss_grad(ops, ref):
    from ishapediting_b200.drag_utils import DragGeometry, align_maps

    S, Cf = 16, 128
    g = G(11)
    chan_map, inv_map, Ca = align_maps(Cf)
    feat = torch.randn(1, S, S, Cf, generator=g)
    origin = torch.randn(3, S, S, Ca, generator=g)
    sources = (torch.rand(3, 3, generator=g) - 0.5).numpy()
    targets = sources + (torch.rand(3, 3, generator=g).numpy() - 0.5) * 0.4
    targets[2] = [0.97, -0.99, 0.5]     # patch partly outside the plane -> zero padding path
    geo = DragGeometry(sources, targets, r=3, voxel_size=2.0 / 64, S=S, Ca=Ca)
    for loss_type in (0, 1):
        for cof in (0.2, 0.0):
            r = {}
            for name, o in (("ref", ref), ("cuda", ops)):
                dev = "cpu" if name == "ref" else DEV
                t = geo.to(dev)
                npts = t.patch_xy.shape[1]
                gbuf = torch.zeros(3, npts, Ca, device=dev)
                info = torch.zeros(3, npts, 4, device=dev)
                partial = torch.zeros(o.drag_partial_len(S, Cf, npts), dtype=torch.float64, device=dev)
                loss = torch.zeros(1, device=dev)
                d_feat = torch.full((1, S, S, Cf), 7.0, device=dev)
                o.drag_loss_grad(feat.to(dev), origin.to(dev), chan_map.to(dev), inv_map.to(dev), t.patch_xy,
                                 t.shift_xy, t.weight, t.group_size, t.bbox, t.mask, t.mask_count, t.inv_count, cof,
                                 loss_type, gbuf, info, partial, loss, d_feat)
                r[name] = (loss, d_feat)
            assert rel_l2(r["cuda"][0], r["ref"][0]) < 1e-5, (loss_type, cof)
            assert rel_l2(r["cuda"][1], r["ref"][1]) < 2e-5, (loss_type, cof)


def test_resize_feat_align(ops, ref):
    from ishapediting_b200.drag_utils import align_maps

    S, Cf = 8, 128
    chan_map, _, Ca = align_maps(Cf)
    feat = torch.randn(1, S, S, Cf, generator=G(13))
    a = ref.resize_feat_align(feat, chan_map, torch.zeros(3, S, S, Ca))
    b = ops.resize_feat_align(feat.to(DEV), chan_map.to(DEV), ops.empty((3, S, S, Ca)))
    assert torch.equal(a, b.cpu())


def test_triplane_decode(ops, ref):
    g = G(15)
    R = 32
    planes = torch.randn(3, R, R, 32, generator=g) * 0.3
    weights = [torch.randn(32, 64, generator=g)] + [
        t for o_, i_ in ((128, 128), (128, 128), (1, 128))
        for t in ((torch.rand(o_, i_, generator=g) * 2 - 1) / math.sqrt(i_), (torch.rand(o_, generator=g) * 2 - 1) / math.sqrt(i_))]
    coords = torch.rand(1000, 3, generator=g) * 2.2 - 1.1     # some outside [-1,1]: zero padding
    out_ref = ref.decode_points(planes, weights, coords, torch.zeros(1000))
    wd = [w.to(DEV).contiguous() for w in weights]
    out = ops.decode_points(planes.to(DEV), wd, coords.to(DEV), ops.empty((1000,)))
    assert float((out.cpu() - out_ref).abs().max()) < 2e-4 * max(1.0, float(out_ref.abs().max()))
    res = 20
    lin = torch.linspace(-1, 1, res)
    g_ref = ref.decode_grid(planes, weights, lin, 3, 11, torch.zeros(8 * res * res))
    g_out = ops.decode_grid(planes.to(DEV), wd, lin.to(DEV), 3, 11, ops.empty((8 * res * res,)))
    assert float((g_out.cpu() - g_ref).abs().max()) < 2e-4 * max(1.0, float(g_ref.abs().max()))
    assert float(((g_out.cpu() > 0) != (g_ref > 0)).float().mean()) < 1e-3


@pytest.mark.parametrize("npts,b_scale", [(1000, 0.2), (37, 0.2), (4096, 1.0)])
def test_triplane_decode_backward(ops, npts, b_scale):
    """isb_triplane_decode_points_backward against float64 autograd of the reference formula (axisnetworks.py:546-562).
    b_scale = 0.2 keeps the Fourier arguments at a few radians (well-conditioned: tight check of the kernel's math);
    b_scale = 1.0 is the decoder's real regime, where fp32 itself is only good to ~1e-3 (see test_host_logic)."""
    g = G(16)
    R = 32
    planes = torch.randn(3, R, R, 32, generator=g) * 0.3
    weights = [torch.randn(32, 64, generator=g) * b_scale] + [
        t for o_, i_ in ((128, 128), (128, 128), (1, 128))
        for t in ((torch.rand(o_, i_, generator=g) * 2 - 1) / math.sqrt(i_), (torch.rand(o_, generator=g) * 2 - 1) / math.sqrt(i_))]
    coords = torch.rand(npts, 3, generator=g) * 2.2 - 1.1     # some outside [-1,1]: zero padding
    d_logits = torch.randn(npts, generator=g)
    from tests.ref_ops import RefOps

    p64 = planes.double().clone().requires_grad_(True)
    out64 = RefOps("fp32")._decode(p64, [w.double() for w in weights], coords.double())
    (g64,) = torch.autograd.grad((out64 * d_logits.double()).sum(), p64)
    wd = [w.to(DEV).contiguous() for w in weights]
    seed = torch.full((3, R, R, 32), 0.5, device=DEV)          # the kernel ADDS to its output
    got = ops.decode_points_backward(planes.to(DEV), wd, coords.to(DEV), d_logits.to(DEV), seed.clone()) - 0.5
    tol = 2e-5 if b_scale < 1.0 else 5e-3
    assert rel_l2(got.cpu().double(), g64) < tol
    # through the module surface: autograd.Function on MultiTriplane.forward
    from ishapediting_b200.triplane_decoder.axisnetworks import MultiTriplane

    dec = MultiTriplane(1, device=DEV).to(DEV)
    dec.net[0]._B.data.copy_(weights[0])
    for idx, k in ((1, 1), (3, 3), (5, 5)):
        dec.net[idx].weight.data.copy_(weights[k])
        dec.net[idx].bias.data.copy_(weights[k + 1])
    for prm in dec.parameters():
        prm.requires_grad_(False)
    pn = planes.permute(0, 3, 1, 2).contiguous().to(DEV).requires_grad_(True)      # (3,32,R,R)
    for j in range(3):
        dec.embeddings[j] = pn[[j]]
    out = dec(0, coords.to(DEV).unsqueeze(0)).reshape(-1)
    (out * d_logits.to(DEV)).sum().backward()
    assert rel_l2(pn.grad.permute(0, 2, 3, 1).cpu().double(), g64) < tol


def test_abi_rejects_unsupported_arguments(ops):
    """The C ABI reports unsupported shapes / inconsistent descriptors as error codes with a message (surfaced as
    IsbError by the binding) — never a silent fallback, never a launch with garbage parameters."""
    from ishapediting_b200._lib import IsbError

    dev = ops.device
    bf = torch.bfloat16
    a = torch.zeros(1, 16, 16, 48, device=dev, dtype=bf)                 # Cin not a multiple of 64
    with pytest.raises(IsbError, match="Cin"):
        ops.conv(a, torch.zeros(64, 9 * 48, device=dev, dtype=bf), None, 3, torch.zeros(1, 16, 16, 64, device=dev))
    a = torch.zeros(1, 16, 16, 64, device=dev, dtype=bf)
    with pytest.raises(IsbError, match="ksize"):
        ops.conv(a, torch.zeros(64, 25 * 64, device=dev, dtype=bf), None, 5, torch.zeros(1, 16, 16, 64, device=dev))
    with pytest.raises(IsbError, match="gn_slots"):                       # partials buffer sized for another launch
        ops.conv(torch.zeros(1, 16, 16, 256, device=dev, dtype=bf), torch.zeros(256, 9 * 256, device=dev, dtype=bf), None, 3,
                 torch.zeros(1, 16, 16, 256, device=dev), gn_part=torch.zeros(1, 32, 999, 2, device=dev))
    x = torch.zeros(1, 8, 8, 72, device=dev)                             # 72 channels: groups of 2.25
    with pytest.raises(IsbError, match="groups"):
        ops.gn_forward(x, None, torch.ones(72, device=dev), torch.zeros(72, device=dev), None, 0, True, 0,
                       torch.zeros(1, 32, 2, device=dev), torch.zeros(1, 8, 8, 72, device=dev, dtype=bf))
    qkv = torch.zeros(1, 8, 8, 3 * 2 * 32, device=dev, dtype=bf)        # 32-channel heads: fused path is built for 64
    with pytest.raises(IsbError, match="channels per head"):
        ops.attention_flash_forward(qkv, 2, torch.zeros(1, 8, 8, 64, device=dev, dtype=bf), torch.zeros(1, 2, 64, device=dev))
    # and the library is still healthy afterwards
    y = torch.empty(1, 16, 16, 64, device=dev)
    ops.conv(a, torch.zeros(64, 9 * 64, device=dev, dtype=bf), None, 3, y)
    torch.cuda.synchronize()
    assert float(y.abs().max()) == 0.0
