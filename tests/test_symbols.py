"""The C-ABI library loads on a CPU-only box and exports every symbol include/ishape_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ishape_b200.h")
LIB = os.path.join(ROOT, "ishapediting_b200", "libishape_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("isb_init", "isb_conv2d", "isb_gn_forward", "isb_gn_backward", "isb_attention_forward",
                 "isb_attention_backward", "isb_time_embed", "isb_ddpm_step", "isb_drag_loss_grad",
                 "isb_triplane_decode_grid", "isb_last_error"):
        assert must in fns


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, f"declared in the header but not exported: {missing}"
    lib.isb_abi_version.restype = ctypes.c_int
    assert lib.isb_abi_version() == 1


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built")
def test_python_prototypes_cover_the_header():
    from ishapediting_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_functions()
    _lib.load()


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: on a box without a GPU the ops layer must raise, not compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ishapediting_b200 import _lib
    from ishapediting_b200.ops import CudaOps

    with pytest.raises(_lib.IsbError):
        CudaOps(None, "bf16")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ishapediting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'""".*?"""', "", text, flags=re.S).replace("# oracle", ""), \
                    f"{f} references the oracle"


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built")
def test_handle_level_abi_refuses_to_run_without_a_device():
    """The handle-level entry points fail loudly, with a message, before isb_init() has bound a B200 — no silent
    host fallback (this box has no GPU; on a GPU box the same calls are exercised by tests/test_gpu_native_unet.py)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ishapediting_b200 import _lib

    lib = _lib.load()
    cfg = _lib.UnetCfg()
    cfg.in_channels, cfg.model_channels, cfg.out_channels, cfg.num_res_blocks = 96, 256, 192, 2
    cfg.n_levels = 1
    cfg.channel_mult[0] = 1
    cfg.num_heads, cfg.num_head_channels, cfg.num_heads_upsample = 1, 64, -1
    cfg.N, cfg.H, cfg.W, cfg.mode, cfg.want_backward = 1, 32, 32, _lib.BF16, 1
    h = ctypes.c_void_p()
    rc = lib.isb_unet_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"isb_init" in lib.isb_last_error()
    assert lib.isb_unet_workspace_bytes(None) == 0 and lib.isb_unet_num_blocks(None) == 0
    lib.isb_unet_destroy(None)          # a null handle is accepted
    with pytest.raises(_lib.IsbError):
        _lib.init(0)                    # no sm_100 device here
