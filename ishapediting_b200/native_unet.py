"""The UNet through the handle-level C ABI (isb_unet_*, include/ishape_b200.h).

`NativeUNet` is the thinnest possible host of `isb_unet_create / load_weight / finalize / forward /
backward_input`: it hands the reference-named parameters of a `UNetModel` (or any state_dict with the names of
neural_field_diffusion/guided_diffusion/unet.py) to the library and owns ONE workspace tensor.  Block structure,
weight packing, activation layout and launch schedule are all inside libishape_b200.so — this is what a host written
in another language binds, and what `tests/test_gpu_native_unet.py` holds bit-for-bit against the
Python plan (`guided_diffusion/unet.py::_Plan`).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32
from .guided_diffusion.nn import timestep_freqs


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class NativeUNet:
    def __init__(self, model, N, H, W, mode=None, want_backward=True, side_stream=True):
        """model: a UNetModel (its constructor arguments and state_dict are read; its own plan is not used)."""
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.IsbError("NativeUNet needs the parameters on a CUDA device (B200); there is no CPU fallback")
        self.lib = _lib.init(dev.index or 0)
        self.device = dev
        mode = mode or model._mode
        cfg = _lib.UnetCfg()
        cfg.in_channels, cfg.model_channels, cfg.out_channels = model.in_channels, model.model_channels, model.out_channels
        cfg.num_res_blocks = model.num_res_blocks
        mult = [int(m) for m in model.channel_mult]
        cfg.n_levels = len(mult)
        for i, m in enumerate(mult):
            cfg.channel_mult[i] = m
        ds = sorted(int(a) for a in model.attention_resolutions)
        cfg.n_attn = len(ds)
        for i, a in enumerate(ds):
            cfg.attention_ds[i] = a
        cfg.num_heads, cfg.num_head_channels = model.num_heads, model.num_head_channels
        cfg.num_heads_upsample = model.num_heads_upsample
        cfg.N, cfg.H, cfg.W = N, H, W
        cfg.mode = BF16 if mode == "bf16" else F32
        cfg.want_backward, cfg.side_stream = int(want_backward), int(side_stream)
        self.cfg, self.mode = cfg, mode
        h = C.c_void_p()
        _lib.check(self.lib.isb_unet_create(C.byref(cfg), C.byref(h)), "isb_unet_create")
        self._h = h
        try:
            st = _stream()
            tensors = dict(model.state_dict())
            tensors["time_embed.freqs"] = timestep_freqs(model.model_channels)
            for name, t in tensors.items():
                t = t.detach().to(device=dev, dtype=torch.float32).contiguous()
                shape = (C.c_int64 * t.dim())(*t.shape)
                _lib.check(self.lib.isb_unet_load_weight(h, name.encode(), _p(t), F32, shape, t.dim(), st),
                           f"isb_unet_load_weight({name})")
            _lib.check(self.lib.isb_unet_finalize(h, st), "isb_unet_finalize")
            self.ws_bytes = int(self.lib.isb_unet_workspace_bytes(h))
            self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=dev)
            off = (-self.ws.data_ptr()) % 256
            self._ws_ptr = C.c_void_p(self.ws.data_ptr() + off)
            _lib.check(self.lib.isb_unet_workspace_init(h, self._ws_ptr, self.ws_bytes, st), "isb_unet_workspace_init")
        except Exception:
            self.close()
            raise
        self.num_blocks = int(self.lib.isb_unet_num_blocks(h))
        self.shape_in = (N, model.in_channels, H, W)
        self.shape_out = (N, model.out_channels, H, W)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.isb_unet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------------------------
    def feat(self, feat_layer):
        """(val, grad) fp32 NHWC views INTO the workspace of output_blocks[feat_layer]'s result (grad: None without
        backward support)."""
        val, grad = C.c_void_p(), C.c_void_p()
        dims = (C.c_int * 4)()
        _lib.check(self.lib.isb_unet_feat(self._h, self._ws_ptr, feat_layer, C.byref(val), C.byref(grad), C.byref(dims)),
                   "isb_unet_feat")
        shape = tuple(dims)
        n = shape[0] * shape[1] * shape[2] * shape[3]
        base = self.ws.data_ptr()

        def view(ptr):
            if not ptr.value:
                return None
            off = ptr.value - base
            return self.ws[off:off + 4 * n].view(torch.float32).view(shape)

        return view(val), view(grad)

    def forward(self, x, t, feat_layer=-1, stop_at_feat=False, out=None, out_nhwc=False):
        """x [N,C,H,W] fp32, t [N] timestep values.  Returns out ([N,2C,H,W], or NHWC) — None when stop_at_feat."""
        assert tuple(x.shape) == self.shape_in and x.dtype == torch.float32 and x.is_cuda and x.is_contiguous()
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if out is None and not stop_at_feat:
            N, Co, H, W = self.shape_out
            out = torch.empty((N, H, W, Co) if out_nhwc else (N, Co, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.isb_unet_forward(self._h, _p(x), _p(t), feat_layer, int(stop_at_feat), _p(out), int(out_nhwc),
                                             self._ws_ptr, self.ws_bytes, _stream()), "isb_unet_forward")
        return out

    def forward_tail(self, out=None, out_nhwc=False):
        if out is None:
            N, Co, H, W = self.shape_out
            out = torch.empty((N, H, W, Co) if out_nhwc else (N, Co, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.isb_unet_forward_tail(self._h, _p(out), int(out_nhwc), self._ws_ptr, self.ws_bytes, _stream()),
                   "isb_unet_forward_tail")
        return out

    def backward_input(self, feat_layer=-1, d_feat=None, in_place=False, d_out=None, dx=None):
        """d_feat: fp32 NHWC gradient of the feature (or in_place=True when it was written to feat()[1]); d_out: fp32
        NCHW gradient of the output.  Returns dx [N,C,H,W]."""
        if dx is None:
            dx = torch.empty(self.shape_in, dtype=torch.float32, device=self.device)
        for g in (d_feat, d_out):
            assert g is None or (g.dtype == torch.float32 and g.is_cuda and g.is_contiguous())
        _lib.check(self.lib.isb_unet_backward_input(self._h, feat_layer, _p(d_feat), int(in_place), _p(d_out), _p(dx),
                                                    self._ws_ptr, self.ws_bytes, _stream()), "isb_unet_backward_input")
        return dx
