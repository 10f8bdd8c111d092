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
