"""ctypes binding of libishape_b200.so (the C ABI declared in include/ishape_b200.h).

The product path has no CPU fallback: if the shared library is missing, or the device is
not a B200 (sm_100a), every op raises.  `load()` only dlopens and declares prototypes (safe on
a CPU-only box — used by the `-m "not gpu"` symbol test); `init()` binds the CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libishape_b200.so")

F32, BF16 = 0, 1

c_void_p, c_int, c_float, c_size_t = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class IsbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("a", c_void_p), ("a_dtype", c_int),
        ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int),
        ("ksize", c_int),
        ("a2", c_void_p), ("Cin2", c_int),
        ("w", c_void_p), ("bias", c_void_p), ("residual", c_void_p),
        ("out", c_void_p), ("out_dtype", c_int), ("Cout", c_int),
        ("accumulate", c_int),
        ("block_n", c_int), ("split_k", c_int), ("stages", c_int), ("w_tiled", c_int), ("two_cta", c_int), ("debug_flags", c_int), ("min_smem_bytes", c_int),
        ("gn_partials", c_void_p), ("gn_cg", c_int), ("gn_slots", c_int),
        ("gn_mode", c_int), ("gb_x", c_void_p), ("gb_gamma", c_void_p), ("gb_beta", c_void_p),
        ("gb_film", c_void_p), ("gb_film_stride", c_int), ("gb_stats", c_void_p), ("gb_silu", c_int),
    ]


class GnDesc(C.Structure):
    _fields_ = [
        ("x1", c_void_p), ("C1", c_int),
        ("x2", c_void_p), ("C2", c_int),
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("groups", c_int), ("eps", c_float),
        ("gamma", c_void_p), ("beta", c_void_p),
        ("film", c_void_p), ("film_stride", c_int),
        ("silu", c_int), ("resample", c_int),
        ("stats", c_void_p),
        ("y", c_void_p), ("y_dtype", c_int),
        ("raw", c_void_p), ("raw_dtype", c_int),
        ("xres", c_void_p),
        ("partials", c_void_p), ("partial_slots", c_int),
    ]


class GnBwdDesc(C.Structure):
    _fields_ = [
        ("f", GnDesc),
        ("dy", c_void_p),
        ("gres", c_void_p), ("gres_at_input", c_int),
        ("gx1", c_void_p), ("acc1", c_int), ("gx1_lo", c_void_p),
        ("gx2", c_void_p), ("acc2", c_int), ("gx2_lo", c_void_p),
        ("lo_dtype", c_int),
        ("partials", c_void_p), ("partial_slots", c_int),
    ]


class DdpmDesc(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("model_out", c_void_p), ("model_out_cstride", c_int), ("model_out_nchw", c_int),
        ("noise", c_void_p), ("grad", c_void_p), ("coef", c_void_p),
        ("N", c_int), ("C", c_int), ("H", c_int), ("W", c_int), ("clip_denoised", c_int),
        ("x_next", c_void_p), ("sample", c_void_p), ("mean", c_void_p), ("var", c_void_p),
        ("x0", c_void_p), ("eps", c_void_p), ("coef_per_sample", c_int),
    ]


class DragDesc(C.Structure):
    _fields_ = [
        ("feat", c_void_p), ("S", c_int), ("Cf", c_int),
        ("origin", c_void_p), ("Ca", c_int),
        ("chan_map", c_void_p), ("inv_map", c_void_p),
        ("patch_xy", c_void_p), ("shift_xy", c_void_p), ("weight", c_void_p),
        ("npts", c_int), ("group_size", c_int),
        ("bbox", c_void_p),
        ("mask", c_void_p), ("mask_count", c_int),
        ("inv_count", c_float), ("cof", c_float), ("loss_type", c_int),
        ("g", c_void_p), ("pt_info", c_void_p), ("partial", c_void_p), ("partial_len", c_int),
        ("loss", c_void_p), ("d_feat", c_void_p), ("dyn_scalars", c_void_p),
    ]


class TriplaneMlp(C.Structure):
    _fields_ = [
        ("fourier_B", c_void_p),
        ("w1", c_void_p), ("b1", c_void_p),
        ("w2", c_void_p), ("b2", c_void_p),
        ("w3", c_void_p), ("b3", c_void_p),
        ("h1_bound", C.c_float),
    ]


class TrackDesc(C.Structure):
    _fields_ = [
        ("feat", c_void_p), ("S", c_int), ("Cf", c_int), ("Ca", c_int),
        ("chan_map", c_void_p), ("f0", c_void_p), ("center", c_void_p),
        ("B", c_int), ("r", c_int), ("voxel", C.c_float),
        ("table", c_void_p), ("out_idx", c_void_p), ("out_dist", c_void_p), ("out_pts", c_void_p),
    ]


class UnetCfg(C.Structure):
    _fields_ = [
        ("in_channels", c_int), ("model_channels", c_int), ("out_channels", c_int), ("num_res_blocks", c_int),
        ("n_levels", c_int), ("channel_mult", c_int * 8),
        ("n_attn", c_int), ("attention_ds", c_int * 8),
        ("num_heads", c_int), ("num_head_channels", c_int), ("num_heads_upsample", c_int),
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("mode", c_int), ("want_backward", c_int), ("side_stream", c_int),
    ]


# name -> (restype, argtypes); also the list the symbol-export test checks against the header
PROTOTYPES = {
    "isb_abi_version": (c_int, []),
    "isb_init": (c_int, [c_int]),
    "isb_last_error": (C.c_char_p, []),
    "isb_launch_count": (C.c_uint64, []),
    "isb_debug_set_trace": (None, [c_void_p]),
    "isb_debug_launch_chain": (c_int, [c_int, c_int, c_int, c_int, c_void_p]),
    "isb_nchw_to_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "isb_nhwc_to_nchw": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "isb_cast_f32_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "isb_conv2d_workspace": (c_size_t, [C.POINTER(ConvDesc)]),
    "isb_conv2d_gn_slots": (c_int, [C.POINTER(ConvDesc)]),
    "isb_conv2d": (c_int, [C.POINTER(ConvDesc), c_void_p, c_size_t, c_void_p]),
    "isb_gn_scratch_bytes": (c_size_t, [c_int, c_int]),
    "isb_gn_forward": (c_int, [C.POINTER(GnDesc), c_void_p, c_void_p]),
    "isb_gn_backward": (c_int, [C.POINTER(GnBwdDesc), c_void_p, c_void_p]),
    "isb_attention_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "isb_attention_flash_forward": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "isb_attention_flash_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                             c_void_p, c_void_p, c_void_p]),
    "isb_attention_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_int, c_void_p]),
    "isb_time_embed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "isb_ddpm_step": (c_int, [C.POINTER(DdpmDesc), c_void_p]),
    "isb_drag_partial_len": (c_size_t, [c_int, c_int, c_int]),
    "isb_drag_loss_grad": (c_int, [C.POINTER(DragDesc), c_void_p]),
    "isb_resize_feat_align": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "isb_track_points": (c_int, [C.POINTER(TrackDesc), c_void_p]),
    "isb_prefetch_l2": (c_int, [c_void_p, c_size_t, c_void_p]),
    "isb_mc_workspace_bytes": (c_size_t, [c_int]),
    "isb_mc_count": (c_int, [c_void_p, c_int, C.c_float, c_void_p, c_size_t, c_void_p, c_void_p]),
    "isb_mc_emit": (c_int, [c_void_p, c_int, C.c_float, c_void_p, c_size_t, C.c_float, c_void_p, c_void_p, c_void_p]),
    "isb_mesh_smooth_workspace_bytes": (c_size_t, [C.c_int64, C.c_int64]),
    "isb_mesh_smooth_simple": (c_int, [c_void_p, C.c_int64, c_void_p, C.c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "isb_unet_create": (c_int, [C.POINTER(UnetCfg), C.POINTER(c_void_p)]),
    "isb_unet_load_weight": (c_int, [c_void_p, C.c_char_p, c_void_p, c_int, C.POINTER(C.c_int64), c_int, c_void_p]),
    "isb_unet_finalize": (c_int, [c_void_p, c_void_p]),
    "isb_unet_workspace_bytes": (c_size_t, [c_void_p]),
    "isb_unet_workspace_init": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "isb_unet_num_blocks": (c_int, [c_void_p]),
    "isb_unet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t,
                                 c_void_p]),
    "isb_unet_forward_tail": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "isb_unet_feat": (c_int, [c_void_p, c_void_p, c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_int * 4)]),
    "isb_unet_backward_input": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "isb_unet_destroy": (None, [c_void_p]),
    "isb_triplane_decode_grid": (c_int, [c_void_p, c_int, C.POINTER(TriplaneMlp), c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_void_p]),
    "isb_triplane_decode_points": (c_int, [c_void_p, c_int, C.POINTER(TriplaneMlp), c_void_p, C.c_int64,
                                           c_void_p, c_void_p]),
    "isb_triplane_decode_points_backward": (c_int, [c_void_p, c_int, C.POINTER(TriplaneMlp), c_void_p, C.c_int64,
                                                    c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited_devices: set[int] = set()


def load():
    """dlopen the library and declare prototypes.  Raises IsbError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise IsbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C ishapediting_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.isb_abi_version() != 1:
            raise IsbError("libishape_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def init(device: int = 0):
    """Bind the library to a CUDA device (idempotent)."""
    lib = load()
    if device in _inited_devices:
        return lib
    with _lock:
        if device not in _inited_devices:
            rc = lib.isb_init(device)
            if rc != 0:
                raise IsbError(f"isb_init({device}) failed ({rc}): {lib.isb_last_error().decode()}")
            _inited_devices.add(device)
    return lib


def check(rc: int, what: str):
    if rc != 0:
        raise IsbError(f"{what} failed ({rc}): {_lib.isb_last_error().decode()}")


def launch_count() -> int:
    return int(load().isb_launch_count())
