"""GPU parity: isb_conv2d (tcgen05 bf16 path and fp32 FFMA path) vs torch conv2d on CPU.
Replaces nn.Conv2d / nn.Conv1d call sites unet.py:185,211,222,286,294,482,615."""
import pytest
import torch

from tests.conv_cases import CASES
from tests.probe_conv import run_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops_by_mode():
    from ishapediting_b200.ops import CudaOps

    return {m: CudaOps(torch.device("cuda", 0), m) for m in ("bf16", "fp32")}


@pytest.mark.parametrize("packed", [True, False], ids=["panels", "rowmajor"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_bf16_tcgen05(ops_by_mode, case, packed):
    err, mx, _, _ = run_case(ops_by_mode["bf16"], case, "bf16", packed=packed)
    # operands are identical bf16 values on both sides; only the accumulation order (and the bf16
    # output rounding when requested) differs
    assert err < (1e-2 if case[11] else 2e-3), f"rel_l2={err} max_abs={mx}"


@pytest.mark.parametrize("case", [c for c in CASES if not c[11]], ids=[c[0] for c in CASES if not c[11]])
def test_conv_fp32(ops_by_mode, case):
    err, mx, _, _ = run_case(ops_by_mode["fp32"], case, "fp32")
    assert err < 1e-5, f"rel_l2={err} max_abs={mx}"


# (N, H, Cin, Cout, ksize, Cin2, residual, tune): group widths 8 / 16 / 32, every epilogue variant
GN_STAT_CASES = [
    (1, 128, 64, 256, 3, 0, True, None),                                   # CTA pair, direct epilogue, 128 slots
    (1, 64, 128, 256, 3, 0, False, {"two_cta": 2, "block_n": 128, "split_k": 1}),   # single-CTA direct
    (1, 64, 128, 256, 3, 0, True, {"two_cta": 2, "block_n": 128, "split_k": 2}),    # cluster fold, 2 splits
    (1, 32, 128, 512, 3, 64, False, {"block_n": 128, "split_k": 4}),       # fold, 4 splits, fused 1x1 skip source
    (1, 32, 256, 512, 3, 0, True, {"block_n": 256, "split_k": 2, "stages": 3}),     # 256-wide tile (two warps per row)
    (1, 16, 256, 1024, 3, 0, True, {"block_n": 64, "split_k": 8}),         # 64-wide tile (two rows per warp)
    (2, 8, 256, 1024, 3, 0, True, None),                                   # two images per 128-row tile, heuristic
    (3, 8, 128, 256, 1, 0, False, {"block_n": 64, "split_k": 1}),          # odd batch: half-empty last tile, direct
    (3, 8, 512, 512, 3, 0, False, {"block_n": 128, "split_k": 4}),         # odd batch, fold
    (1, 32, 512, 512, 1, 0, True, None),                                   # attention proj_out shape
]


@pytest.mark.parametrize("case", GN_STAT_CASES, ids=[f"n{c[0]}_h{c[1]}_{c[2]}to{c[3]}_k{c[4]}_{i}" for i, c in enumerate(GN_STAT_CASES)])
@pytest.mark.parametrize("rows_mb", [None, "0"], ids=["default", "row-kernel"])
def test_conv_fused_groupnorm_statistics(ops_by_mode, case, rows_mb, monkeypatch):
    """The conv epilogue's (sum, sum of squares) partials per (image, 32 groups) against the tensor it wrote, and
    through gn_forward(partials=...) against the ordinary statistics pass.  rows_mb = "0" forces the large-tensor
    route (statistics fold kernel + whole-row apply kernel) at these sizes."""
    if rows_mb is not None:
        monkeypatch.setenv("ISB_GN_ROWS_MIN_MB", rows_mb)
    ops = ops_by_mode["bf16"]
    N, H, Cin, Cout, k, Cin2, res, tune = case
    dev = ops.device
    g = torch.Generator().manual_seed(11)
    a = torch.randn(N, H, H, Cin, generator=g).to(dev).to(torch.bfloat16)
    a2 = torch.randn(N, H, H, Cin2, generator=g).to(dev).to(torch.bfloat16) if Cin2 else None
    w = (torch.randn(Cout, k * k * Cin + Cin2, generator=g) / (k * k * Cin) ** 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g).to(dev)
    residual = torch.randn(N, H, H, Cout, generator=g).to(dev) if res else None
    out = torch.empty(N, H, H, Cout, device=dev)
    import ctypes as C
    from ishapediting_b200 import _lib

    d = _lib.ConvDesc()
    d.a, d.a_dtype, d.w, d.out, d.out_dtype = 256, _lib.BF16, 256, 256, _lib.F32
    d.N, d.H, d.W, d.Cin, d.ksize, d.Cout = N, H, H, Cin, k, Cout
    if Cin2:
        d.a2, d.Cin2 = 256, Cin2
    d.gn_cg = Cout // 32
    if tune:
        d.block_n, d.split_k, d.stages, d.two_cta = tune.get("block_n", 0), tune.get("split_k", 0), tune.get("stages", 0), tune.get("two_cta", 0)
    slots = int(ops.lib.isb_conv2d_gn_slots(C.byref(d)))
    assert slots > 0
    part = torch.full((N, 32, slots, 2), float("nan"), device=dev)
    ops.conv(a, w, bias, k, out, a2=a2, residual=residual, tune=tune, gn_part=part)
    torch.cuda.synchronize()
    og = out.double().reshape(N, H * H, 32, Cout // 32)
    s_ref, q_ref = og.sum(dim=(1, 3)), (og * og).sum(dim=(1, 3))
    s, q = part[..., 0].double().sum(-1), part[..., 1].double().sum(-1)
    assert torch.isfinite(part).all()
    scale = q_ref.sqrt() * (H * H * Cout // 32) ** 0.5          # ~ n * rms: the natural size of a sum of n terms
    assert ((s - s_ref).abs() / scale).max() < 1e-5
    assert ((q - q_ref).abs() / q_ref).max() < 1e-5
    # consumer side: identical activations with and without the fused statistics
    gamma, beta = torch.randn(Cout, generator=g).to(dev), torch.randn(Cout, generator=g).to(dev)
    ys, sts = [], []
    for p in (None, part):
        st = torch.zeros(N, 32, 2, device=dev)
        y = torch.empty(N, H, H, Cout, device=dev, dtype=torch.bfloat16)
        ops.gn_forward(out, None, gamma, beta, None, 0, True, 0, st, y, partials=p)
        ys.append(y.float())
        sts.append(st)
    assert (sts[0][..., 0] - sts[1][..., 0]).abs().max() < 1e-5
    assert ((sts[0][..., 1] - sts[1][..., 1]).abs() / sts[0][..., 1]).max() < 1e-5
    assert (ys[0] - ys[1]).abs().max() < 0.05      # bf16 outputs: at most a rounding flip


def test_conv_background_occupancy_limit(ops_by_mode):
    """ops.background(): the same convolution with the one-CTA-per-SM shared-memory floor gives the same bits."""
    ops = ops_by_mode["bf16"]
    dev = ops.device
    g = torch.Generator().manual_seed(3)
    a = torch.randn(1, 64, 64, 128, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(256, 9 * 128, generator=g) / 34.0).to(dev).to(torch.bfloat16)
    bias = torch.randn(256, generator=g).to(dev)
    o1, o2 = torch.empty(1, 64, 64, 256, device=dev), torch.empty(1, 64, 64, 256, device=dev)
    ops.conv(a, w, bias, 3, o1)
    with ops.background():
        ops.conv(a, w, bias, 3, o2)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2)
