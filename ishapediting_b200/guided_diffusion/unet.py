"""B200-native UNetModel with the reference's call surface.

Reference: neural_field_diffusion/guided_diffusion/unet.py (UNetModel :396-671, ResBlock :143-256,
AttentionBlock :259-305, QKVAttentionLegacy :328-358, Upsample/Downsample :81-140).

The nn.Module tree below exists only to own parameters under the reference's names (so
`load_state_dict(strict=True)` accepts NFD checkpoints).  The arithmetic is an explicit, static
*plan* of C-ABI kernel calls over preallocated NHWC buffers (see `_Plan`): forward, and the
input-gradient backward that classifier guidance needs (drag_utils.py:383) — dgrad only, no weight
gradients, walking only the layers that lie between the requested gradient roots and the input.
"""
from __future__ import annotations

import math
import os

import torch as th
import torch.nn as nn

from .nn import conv_nd, linear, normalization, timestep_freqs, zero_module


# ---------------------------------------------------------------------------------------------
# parameter containers (same attribute names / creation order as the reference)
# ---------------------------------------------------------------------------------------------
class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    pass


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        if use_conv:
            raise NotImplementedError("conv Upsample (resblock_updown=False) is outside the NFD configuration")
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        if use_conv:
            raise NotImplementedError("conv Downsample (resblock_updown=False) is outside the NFD configuration")
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is outside the NFD configuration")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.updown = up or down
        self.up, self.down = up, down
        if up:
            self.h_upd, self.x_upd = Upsample(channels, False, dims), Upsample(channels, False, dims)
        elif down:
            self.h_upd, self.x_upd = Downsample(channels, False, dims), Downsample(channels, False, dims)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, 2 * self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            raise NotImplementedError("3x3 skip convolution is outside the NFD configuration")
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)


class QKVAttentionLegacy(nn.Module):
    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False):
        super().__init__()
        if use_new_attention_order:
            raise NotImplementedError("use_new_attention_order=True is outside the NFD configuration")
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, \
                f"q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}"
            self.num_heads = channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttentionLegacy(self.num_heads)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))


# ---------------------------------------------------------------------------------------------
# the execution plan
# ---------------------------------------------------------------------------------------------
class _T:
    """A fp32 NHWC 'stream' tensor (block input/output) with its gradient buffers."""
    __slots__ = ("val", "grad", "grad_lo", "has_grad", "name", "gn_slots", "gn_part")

    def __init__(self, val, name, gn_slots=0):
        self.val, self.grad, self.grad_lo, self.has_grad, self.name = val, None, None, False, name
        # GroupNorm statistics fused into the producing conv's epilogue: `gn_slots` = what that conv can deliver
        # (0: nothing); the first consumer whose GroupNorm reads this tensor alone allocates `gn_part`, which
        # switches the producer on.
        self.gn_slots, self.gn_part = gn_slots, None

    def want_gn_part(self, ops):
        if self.gn_slots > 0 and self.gn_part is None:
            N = self.val.shape[0]
            self.gn_part = ops.zeros((N, 32, self.gn_slots, 2))
        return self.gn_part


def _pack3(w, lo):      # [Co,Ci,3,3] -> [Co, 9*Ci], k = (kh*3+kw)*Ci + ci
    return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(lo).contiguous()


def _pack3_dgrad(w, lo):  # -> [Ci, 9*Co] with the taps flipped (backward-data as a forward conv)
    return w.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).to(lo).contiguous()


def _pack1(w, lo):      # [Co,Ci,1(,1)] -> [Co,Ci]
    return w.detach().reshape(w.shape[0], w.shape[1]).to(lo).contiguous()


def _pack1_dgrad(w, lo):
    return w.detach().reshape(w.shape[0], w.shape[1]).t().to(lo).contiguous()


def _f32(p):
    return p.detach().to(th.float32).contiguous()


class _ResLayer:
    def __init__(self, plan, mod: ResBlock, srcs, film_off, name):
        ops, lo = plan.ops, plan.lo
        self.plan, self.name, self.srcs = plan, name, srcs
        x1 = srcs[0].val
        N, H, W = x1.shape[:3]
        Cin = sum(s.val.shape[3] for s in srcs)
        assert Cin == mod.channels, f"{name}: got {Cin} input channels, module expects {mod.channels}"
        Co = mod.out_channels
        self.resample = 1 if mod.down else (2 if mod.up else 0)
        Ho, Wo = (H // 2, W // 2) if mod.down else ((H * 2, W * 2) if mod.up else (H, W))
        self.has_skip_conv = not isinstance(mod.skip_connection, nn.Identity)
        assert not (self.has_skip_conv and self.resample), "up/down ResBlocks keep the channel count"
        assert len(srcs) == 1 or self.has_skip_conv
        self.film_off = film_off
        gn1, conv1 = mod.in_layers[0], mod.in_layers[2]
        gn2, conv2 = mod.out_layers[0], mod.out_layers[3]
        self.g1, self.be1 = _f32(gn1.weight), _f32(gn1.bias)
        self.g2, self.be2 = _f32(gn2.weight), _f32(gn2.bias)
        self.w1, self.b1 = ops.pack_weight(_pack3(conv1.weight, lo)), _f32(conv1.bias)
        w2 = _pack3(conv2.weight, lo)
        b2 = _f32(conv2.bias)
        if self.has_skip_conv:
            w2 = th.cat([w2, _pack1(mod.skip_connection.weight, lo)], dim=1).contiguous()
            b2 = (b2 + _f32(mod.skip_connection.bias)).contiguous()
        self.w2, self.b2 = ops.pack_weight(w2), b2
        if plan.want_backward:
            self.w1_d = ops.pack_weight(_pack3_dgrad(conv1.weight, lo))
            self.w2_d = ops.pack_weight(_pack3_dgrad(conv2.weight, lo))
            self.wskip_d = ops.pack_weight(_pack1_dgrad(mod.skip_connection.weight, lo)) if self.has_skip_conv else None
        # buffers
        self.stats1, self.stats2 = ops.empty((N, 32, 2)), ops.empty((N, 32, 2))
        self.a1 = plan.scratch("a", (N, Ho, Wo, Cin), lo)
        self.xraw = ops.empty((N, H, W, Cin), lo) if self.has_skip_conv else None
        self.xres = ops.empty((N, Ho, Wo, Cin)) if self.resample else None
        self.h1 = ops.empty((N, Ho, Wo, Co))
        s1 = ops.conv_gn_slots(N, Ho, Wo, Cin, 3, Co)          # GN2 reads conv1's output: statistics from its epilogue
        self.h1_part = ops.zeros((N, 32, s1, 2)) if s1 > 0 else None
        self.a2 = plan.scratch("a", (N, Ho, Wo, Co), lo)
        self.out = _T(ops.empty((N, Ho, Wo, Co)), name,
                      ops.conv_gn_slots(N, Ho, Wo, Co, 3, Co, Cin if self.has_skip_conv else 0))
        # GN1 over a single, un-resampled source whose producer can deliver the statistics
        self.x_part = srcs[0].want_gn_part(ops) if (len(srcs) == 1 and self.resample == 0) else None
        # backward: the dgrad convs that produce dy of GN2 / GN1 accumulate those layers' reduction terms
        self.b2_part = self.b1_part = None
        if plan.want_backward and plan.fuse_gn_bwd:
            s2 = ops.conv_gn_slots(N, Ho, Wo, Co, 3, Co)
            self.b2_part = ops.zeros((N, 32, s2, 2)) if s2 > 0 else None
            s1b = ops.conv_gn_slots(N, H, W, Co, 3, Cin) if (len(srcs) == 1 and self.resample == 0) else 0
            self.b1_part = ops.zeros((N, 32, s1b, 2)) if s1b > 0 else None
        self.dims = (N, H, W, Cin, Ho, Wo, Co)

    def forward(self):
        ops, film = self.plan.ops, self.plan.film_all
        x1 = self.srcs[0].val
        x2 = self.srcs[1].val if len(self.srcs) > 1 else None
        ops.gn_forward(x1, x2, self.g1, self.be1, None, 0, True, self.resample, self.stats1, self.a1,
                       raw=self.xraw, xres=self.xres, partials=self.x_part)
        ops.conv(self.a1, self.w1, self.b1, 3, self.h1, gn_part=self.h1_part)
        ops.gn_forward(self.h1, None, self.g2, self.be2, film, self.film_off, True, 0, self.stats2, self.a2,
                       partials=self.h1_part)
        if self.has_skip_conv:
            ops.conv(self.a2, self.w2, self.b2, 3, self.out.val, a2=self.xraw, gn_part=self.out.gn_part)
        else:
            ops.conv(self.a2, self.w2, self.b2, 3, self.out.val, residual=self.xres if self.resample else x1,
                     gn_part=self.out.gn_part)

    def backward(self):
        plan, ops, lo = self.plan, self.plan.ops, self.plan.lo
        N, H, W, Cin, Ho, Wo, Co = self.dims
        g_out, g_out_lo = self.out.grad, self.out.grad_lo
        g_a2 = plan.scratch("g", (N, Ho, Wo, Co), th.float32)
        joined = None
        if self.has_skip_conv:
            gres = plan.scratch("g2", (N, H, W, Cin), th.float32)
            if plan.side is not None:
                # the 1x1 skip backward only needs g_out: run it beside the conv2 -> GN2 -> conv1 chain
                side, ev_fork, ev_join = plan.side
                ev_fork.record(th.cuda.current_stream())
                with th.cuda.stream(side), ops.workspace_slot(2):
                    side.wait_event(ev_fork)
                    ops.conv(g_out_lo, self.wskip_d, None, 1, gres)
                    ev_join.record(side)
                joined = ev_join
            else:
                ops.conv(g_out_lo, self.wskip_d, None, 1, gres)
            at_input = True
        else:
            gres, at_input = g_out, False
        gn2 = (self.h1, self.g2, self.be2, plan.film_all, self.film_off, True, self.stats2)
        ops.conv(g_out_lo, self.w2_d, None, 3, g_a2, gn_part=self.b2_part, gn_bwd=gn2 if self.b2_part is not None else None)
        g_h1_lo = plan.scratch("glo", (N, Ho, Wo, Co), lo)
        ops.gn_backward(self.h1, None, self.g2, self.be2, plan.film_all, self.film_off, True, 0, self.stats2,
                        g_a2, None, False, None, False, g_h1_lo, None, False, None, partials=self.b2_part)
        g_a1 = plan.scratch("g", (N, Ho, Wo, Cin), th.float32)
        gn1 = (self.srcs[0].val, self.g1, self.be1, None, 0, True, self.stats1)
        ops.conv(g_h1_lo, self.w1_d, None, 3, g_a1, gn_part=self.b1_part, gn_bwd=gn1 if self.b1_part is not None else None)
        if joined is not None:
            th.cuda.current_stream().wait_event(joined)
        s1 = self.srcs[0]
        s2 = self.srcs[1] if len(self.srcs) > 1 else None
        plan.ensure_grad(s1)
        if s2 is not None:
            plan.ensure_grad(s2)
        ops.gn_backward(s1.val, s2.val if s2 is not None else None, self.g1, self.be1, None, 0, True, self.resample,
                        self.stats1, g_a1, gres, at_input,
                        s1.grad, s1.has_grad, s1.grad_lo,
                        s2.grad if s2 is not None else None, s2.has_grad if s2 is not None else False,
                        s2.grad_lo if s2 is not None else None, partials=self.b1_part)
        s1.has_grad = True
        if s2 is not None:
            s2.has_grad = True


class _AttnLayer:
    def __init__(self, plan, mod: AttentionBlock, src, name):
        ops, lo = plan.ops, plan.lo
        self.plan, self.name, self.src = plan, name, src
        N, H, W, Cc = src.val.shape
        assert Cc == mod.channels
        self.heads = mod.num_heads
        self.g, self.be = _f32(mod.norm.weight), _f32(mod.norm.bias)
        self.wqkv, self.bqkv = ops.pack_weight(_pack1(mod.qkv.weight, lo)), _f32(mod.qkv.bias)
        self.wproj, self.bproj = ops.pack_weight(_pack1(mod.proj_out.weight, lo)), _f32(mod.proj_out.bias)
        if plan.want_backward:
            self.wqkv_d = ops.pack_weight(_pack1_dgrad(mod.qkv.weight, lo))
            self.wproj_d = ops.pack_weight(_pack1_dgrad(mod.proj_out.weight, lo))
        T = H * W
        self.stats = ops.empty((N, 32, 2))
        self.a = plan.scratch("a", (N, H, W, Cc), lo)
        # bf16 mode with 64-channel heads: fused attention, the [heads,T,T] probabilities are never materialised
        self.flash = lo == th.bfloat16 and Cc // self.heads == 64 and T % 64 == 0
        if self.flash:
            self.qkv = ops.empty((N, H, W, 3 * Cc), lo)
            self.lse = ops.empty((N, self.heads, T))
            self.o = ops.empty((N, H, W, Cc), lo) if plan.want_backward else plan.scratch("o", (N, H, W, Cc), lo)
        else:
            self.qkv = ops.empty((N, H, W, 3 * Cc))
            self.probs = ops.empty((N, self.heads, T, T))
            self.o = plan.scratch("o", (N, H, W, Cc), lo)
        self.out = _T(ops.empty((N, H, W, Cc)), name, ops.conv_gn_slots(N, H, W, Cc, 1, Cc))
        self.x_part = src.want_gn_part(ops)
        sb = ops.conv_gn_slots(N, H, W, 3 * Cc, 1, Cc) if (plan.want_backward and plan.fuse_gn_bwd) else 0
        self.b_part = ops.zeros((N, 32, sb, 2)) if sb > 0 else None
        self.dims = (N, H, W, Cc, T)

    def forward(self):
        ops, x = self.plan.ops, self.src.val
        ops.gn_forward(x, None, self.g, self.be, None, 0, False, 0, self.stats, self.a, partials=self.x_part)
        ops.conv(self.a, self.wqkv, self.bqkv, 1, self.qkv)
        if self.flash:
            ops.attention_flash_forward(self.qkv, self.heads, self.o, self.lse)
        else:
            ops.attention_forward(self.qkv, self.heads, self.probs, self.o)
        ops.conv(self.o, self.wproj, self.bproj, 1, self.out.val, residual=x, gn_part=self.out.gn_part)

    def backward(self):
        plan, ops, lo = self.plan, self.plan.ops, self.plan.lo
        N, H, W, Cc, T = self.dims
        g_qkv = plan.scratch("glo", (N, H, W, 3 * Cc), lo)
        if self.flash:
            g_o = plan.scratch("golo", (N, H, W, Cc), lo)
            ops.conv(self.out.grad_lo, self.wproj_d, None, 1, g_o)
            delta = plan.scratch("fadelta", (N, self.heads, T), th.float32)
            ops.attention_flash_backward(self.qkv, self.o, g_o, self.lse, self.heads, delta, g_qkv)
        else:
            g_o = plan.scratch("g", (N, H, W, Cc), th.float32)
            ops.conv(self.out.grad_lo, self.wproj_d, None, 1, g_o)
            tmp = plan.scratch("ptmp", (N, self.heads, T, T), th.float32)
            ops.attention_backward(self.qkv, self.probs, g_o, self.heads, tmp, g_qkv)
        g_a = plan.scratch("g", (N, H, W, Cc), th.float32)
        s = self.src
        gnb = (s.val, self.g, self.be, None, 0, False, self.stats)
        ops.conv(g_qkv, self.wqkv_d, None, 1, g_a, gn_part=self.b_part, gn_bwd=gnb if self.b_part is not None else None)
        plan.ensure_grad(s)
        ops.gn_backward(s.val, None, self.g, self.be, None, 0, False, 0, self.stats, g_a, self.out.grad, False,
                        s.grad, s.has_grad, s.grad_lo, None, False, None, partials=self.b_part)
        s.has_grad = True


class _Plan:
    """Static schedule of kernel calls for one (batch, resolution, precision mode)."""

    def __init__(self, model: "UNetModel", ops, N, H, W, want_backward=True):
        self.model, self.ops, self.lo = model, ops, ops.lo
        self.want_backward = want_backward
        self.N, self.H, self.W = N, H, W
        self._scratch = {}
        self.layers = []          # execution order, objects with forward()/backward()/out
        self.generation = 0
        self.film_external = False
        # GroupNorm-backward reduction terms from the dgrad conv's epilogue (ISB_GN_FUSE_BWD=1).  Measured on B200: the
        # step gets 4 % SLOWER — the terms need x and a SiLU derivative per element, which the four epilogue warps of a
        # conv CTA do far more slowly than the wide reduction kernel they replace.  Kept behind the switch, tested.
        self.fuse_gn_bwd = os.environ.get("ISB_GN_FUSE_BWD", "0") == "1"
        self.side = None          # (stream, fork_event, join_event) for branch-level concurrency in the backward
        lo = self.lo
        Cin = model.in_channels
        self.cin_pad = ((Cin + 63) // 64) * 64
        mc = model.model_channels

        # --- timestep-embedding path (unet.py:471-475 + every emb_layers Linear) ---
        te = model.time_embed
        self.te_w1, self.te_b1 = _f32(te[0].weight), _f32(te[0].bias)
        self.te_w2, self.te_b2 = _f32(te[2].weight), _f32(te[2].bias)
        self.freqs = timestep_freqs(mc).to(ops.device)
        res_mods = [m for m in model.modules() if isinstance(m, ResBlock)]
        offs, off = {}, 0
        for m in res_mods:
            offs[id(m)] = off
            off += 2 * m.out_channels
        self.film_rows = off
        self.w_all = th.cat([_f32(m.emb_layers[1].weight) for m in res_mods], dim=0).contiguous()
        self.b_all = th.cat([_f32(m.emb_layers[1].bias) for m in res_mods], dim=0).contiguous()
        hidden = self.te_w1.shape[0]
        self.te_scratch = ops.empty((N * (mc + 2 * hidden),))
        self.film_all = ops.empty((N, self.film_rows))
        self.t_dev = ops.zeros((N,), th.float32)     # timestep values as the reference embeds them: floats (nn.py:116)

        # --- input conv (input_blocks[0], unet.py:482) ---
        conv0 = model.input_blocks[0][0]
        w0 = th.zeros(conv0.weight.shape[0], self.cin_pad, 3, 3, device=conv0.weight.device)
        w0[:, :Cin] = conv0.weight.detach()
        self.w_in, self.b_in = ops.pack_weight(_pack3(w0, lo)), _f32(conv0.bias)
        if want_backward:
            self.w_in_d = ops.pack_weight(_pack3_dgrad(w0, lo))
        self.x_nchw = ops.empty((N, Cin, H, W))
        self.x_lo = ops.empty((N, H, W, self.cin_pad), lo)
        self.h0 = _T(ops.empty((N, H, W, conv0.weight.shape[0])), "input_blocks.0",
                     ops.conv_gn_slots(N, H, W, self.cin_pad, 3, conv0.weight.shape[0]))

        def add_block(seq, srcs, prefix):
            cur = srcs
            for li, mod in enumerate(seq):
                name = f"{prefix}.{li}"
                if isinstance(mod, ResBlock):
                    layer = _ResLayer(self, mod, cur, offs[id(mod)], name)
                elif isinstance(mod, AttentionBlock):
                    assert len(cur) == 1
                    layer = _AttnLayer(self, mod, cur[0], name)
                else:
                    raise NotImplementedError(f"{name}: {type(mod).__name__} is not supported by the B200 plan")
                self.layers.append(layer)
                cur = [layer.out]
            return cur[0]

        hs = [self.h0]
        h = self.h0
        for i in range(1, len(model.input_blocks)):
            h = add_block(model.input_blocks[i], [h], f"input_blocks.{i}")
            hs.append(h)
        h = add_block(model.middle_block, [h], "middle_block")
        self.block_out = []
        for i, blk in enumerate(model.output_blocks):
            h = add_block(blk, [h, hs.pop()], f"output_blocks.{i}")
            self.block_out.append(h)
        self.h_last = h
        self.out_part = h.want_gn_part(ops)

        # --- out (unet.py:612-616) ---
        gn, conv = model.out[0], model.out[2]
        self.out_g, self.out_b = _f32(gn.weight), _f32(gn.bias)
        self.w_out, self.b_out = ops.pack_weight(_pack3(conv.weight, lo)), _f32(conv.bias)
        if want_backward:
            self.w_out_d = ops.pack_weight(_pack3_dgrad(conv.weight, lo))
        self.out_stats = ops.empty((N, 32, 2))
        self.out_a = self.scratch("a", (N, H, W, h.val.shape[3]), lo)
        self.out_nhwc = ops.empty((N, H, W, conv.weight.shape[0]))

    def compute_film(self):
        """film_all[n] = every ResBlock's emb_layers(time_embed(sinusoid(t_dev[n]))) — depends only on the timestep."""
        self.ops.time_embed(self.t_dev, self.freqs, self.te_w1, self.te_b1, self.te_w2, self.te_b2, self.w_all,
                            self.b_all, self.te_scratch, self.film_all)

    # -- scratch buffers shared by shape-compatible transient tensors --------------------------
    def scratch(self, tag, shape, dtype):
        n = 1
        for s in shape:
            n *= s
        key = (tag, dtype)
        buf = self._scratch.get(key)
        if buf is None or buf.numel() < n:
            if buf is not None and th.cuda.is_available() and th.cuda.is_current_stream_capturing():
                raise RuntimeError("scratch growth during CUDA-graph capture; run an eager step first")
            buf = self.ops.empty((n,), dtype)
            self._scratch[key] = buf
        return buf[:n].view(shape)

    def ensure_grad(self, t: _T):
        if t.grad is None:
            t.grad = self.ops.empty(tuple(t.val.shape))
            t.grad_lo = self.ops.empty(tuple(t.val.shape), self.lo)

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, x_nchw, t_orig, feat_layer=-1, upto_feat_only=False):
        """x_nchw fp32 [N,C,H,W] (device), t_orig fp32 [N] timestep values (device).
        Fills self.out_nhwc (unless upto_feat_only) and returns the inter-feature _T (or None)."""
        ops = self.ops
        self.generation += 1
        if x_nchw.data_ptr() != self.x_nchw.data_ptr():
            self.x_nchw.copy_(x_nchw)
        if t_orig.data_ptr() != self.t_dev.data_ptr():
            self.t_dev.copy_(t_orig)
        if not self.film_external:     # GuidedStepper caches the per-timestep FiLM rows and fills film_all itself
            self.compute_film()
        ops.to_nhwc(self.x_nchw, self.x_lo)
        ops.conv(self.x_lo, self.w_in, self.b_in, 3, self.h0.val, gn_part=self.h0.gn_part)
        inter = self.block_out[feat_layer] if feat_layer >= 0 else None
        stop_after = inter if (upto_feat_only and inter is not None) else None
        self._tail_from = len(self.layers)
        for li, layer in enumerate(self.layers):
            layer.forward()
            if stop_after is not None and layer.out is stop_after:
                self._tail_from = li + 1
                return inter
        self.forward_out_layer()
        return inter

    def forward_out_layer(self):
        ops = self.ops
        ops.gn_forward(self.h_last.val, None, self.out_g, self.out_b, None, 0, True, 0, self.out_stats, self.out_a,
                       partials=self.out_part)
        ops.conv(self.out_a, self.w_out, self.b_out, 3, self.out_nhwc)

    def forward_tail(self):
        """The rest of the forward pass after `forward(..., upto_feat_only=True)`: the layers behind the
        intermediate feature and the output layer.  They do not lie on the guidance-gradient path, so the
        guided step runs them on a second stream, concurrently with the backward pass."""
        for layer in self.layers[self._tail_from:]:
            layer.forward()
        self.forward_out_layer()

    # -- backward ---------------------------------------------------------------------------------
    def begin_backward(self):
        assert self.want_backward, "plan was built without backward support"
        self.h0.has_grad = False
        for layer in self.layers:
            layer.out.has_grad = False

    def seed_grad(self, t: _T, g_f32_nhwc=None):
        """Mark t as a gradient root.  If g is given it is copied in; otherwise the caller has
        already written t.grad (e.g. the drag kernel).  The low-precision copy is refreshed."""
        self.ensure_grad(t)
        if g_f32_nhwc is not None:
            if t.has_grad:
                t.grad.add_(g_f32_nhwc)
            else:
                t.grad.copy_(g_f32_nhwc)
        self.ops.cast_lo(t.grad, t.grad_lo)
        t.has_grad = True

    def backward_out_layer(self, g_out_nchw):
        """Gradient arriving at the UNet output (N,2C,H,W) -> h_last (unet.py:612-616 backward)."""
        ops = self.ops
        N, H, W, Co = self.out_nhwc.shape
        g_lo = self.scratch("glo", (N, H, W, Co), self.lo)
        ops.to_nhwc(g_out_nchw.contiguous(), g_lo)
        Cl = self.h_last.val.shape[3]
        g_a = self.scratch("g", (N, H, W, Cl), th.float32)
        ops.conv(g_lo, self.w_out_d, None, 3, g_a)
        s = self.h_last
        self.ensure_grad(s)
        ops.gn_backward(s.val, None, self.out_g, self.out_b, None, 0, True, 0, self.out_stats, g_a, None, False,
                        s.grad, s.has_grad, s.grad_lo, None, False, None)
        s.has_grad = True

    def backward(self, dx_nchw):
        """Walk the layers in reverse from whatever roots were seeded; writes d/dx into dx_nchw."""
        ops = self.ops
        for layer in reversed(self.layers):
            if layer.out.has_grad:
                layer.backward()
        assert self.h0.has_grad, "no gradient reached the input (no roots seeded?)"
        N, H, W = self.N, self.H, self.W
        g_x = self.scratch("g", (N, H, W, self.cin_pad), th.float32)
        ops.conv(self.h0.grad_lo, self.w_in_d, None, 3, g_x)
        ops.to_nchw(g_x, dx_nchw)
        return dx_nchw


class _UNetFn(th.autograd.Function):
    """torch.autograd bridge so `loss.backward(); img.grad` (drag_utils.py:383-384) keeps working."""

    @staticmethod
    def forward(ctx, x, model, plan, t_orig, feat_layer):
        if hasattr(ctx, "set_materialize_grads"):
            ctx.set_materialize_grads(False)      # an unused output gets None, not a zero tensor that drives a backward branch
        inter = plan.forward(x.detach().to(th.float32).contiguous(), t_orig, feat_layer)
        ctx.plan, ctx.gen, ctx.inter, ctx.shape = plan, plan.generation, inter, x.shape
        N, C2, H, W = plan.N, plan.out_nhwc.shape[3], plan.H, plan.W
        out = plan.ops.to_nchw(plan.out_nhwc, plan.ops.empty((N, C2, H, W)))
        if inter is None:
            return out
        Ni, Hi, Wi, Ci = inter.val.shape
        feat = plan.ops.to_nchw(inter.val, plan.ops.empty((Ni, Ci, Hi, Wi)))
        return out, feat

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        if plan.generation != ctx.gen:
            raise RuntimeError("UNetModel.backward: the plan's activations were overwritten by a later forward; "
                               "call backward before the next forward (the reference loop does)")
        g_out = grads[0]
        g_feat = grads[1] if len(grads) > 1 else None
        plan.begin_backward()
        if g_out is not None:
            plan.backward_out_layer(g_out.to(th.float32))
        if g_feat is not None and ctx.inter is not None:
            Ni, Hi, Wi, Ci = ctx.inter.val.shape
            g_nhwc = plan.scratch("g2", (Ni, Hi, Wi, Ci), th.float32)
            plan.ops.to_nhwc(g_feat.to(th.float32).contiguous(), g_nhwc)
            plan.seed_grad(ctx.inter, g_nhwc)
        dx = plan.ops.empty(tuple(ctx.shape))
        plan.backward(dx)
        return dx, None, None, None, None


class _NativeUNetFn(th.autograd.Function):
    """The same bridge over the handle-level C ABI (native_unet.NativeUNet / csrc/unet.cu): the whole pass is ONE
    library call, so the eager public API is GPU-bound (5.4 ms per NFD forward + backward where the per-operator
    Python plan needs 9 ms of host time).  Bit-identical to `_UNetFn` (tests/test_gpu_native_unet.py)."""

    @staticmethod
    def forward(ctx, x, model, nat, t_orig, feat_layer):
        if hasattr(ctx, "set_materialize_grads"):
            ctx.set_materialize_grads(False)
        ops = model._get_ops()
        nat.generation = getattr(nat, "generation", 0) + 1
        out = nat.forward(x.detach().to(th.float32).contiguous(), t_orig, feat_layer=feat_layer)
        ctx.nat, ctx.gen, ctx.feat_layer, ctx.shape, ctx.ops = nat, nat.generation, feat_layer, x.shape, ops
        if feat_layer < 0:
            return out
        val, _ = nat.feat(feat_layer)
        Ni, Hi, Wi, Ci = val.shape
        feat = ops.to_nchw(val, ops.empty((Ni, Ci, Hi, Wi)))
        return out, feat

    @staticmethod
    def backward(ctx, *grads):
        nat, ops = ctx.nat, ctx.ops
        if nat.generation != ctx.gen:
            raise RuntimeError("UNetModel.backward: the plan's activations were overwritten by a later forward; "
                               "call backward before the next forward (the reference loop does)")
        g_out = grads[0]
        g_feat = grads[1] if len(grads) > 1 else None
        d_feat = None
        if g_feat is not None and ctx.feat_layer >= 0:
            Ni, Ci, Hi, Wi = g_feat.shape
            d_feat = ops.to_nhwc(g_feat.to(th.float32).contiguous(), ops.empty((Ni, Hi, Wi, Ci)))
        d_out = g_out.to(th.float32).contiguous() if g_out is not None else None
        dx = nat.backward_input(ctx.feat_layer if d_feat is not None else -1, d_feat=d_feat, d_out=d_out)
        return dx, None, None, None, None


# ---------------------------------------------------------------------------------------------
# UNetModel
# ---------------------------------------------------------------------------------------------
class UNetModel(nn.Module):
    """Same constructor and `forward(x, timesteps, y=None, feat_layer=-1)` as unet.py:396-671.

    `use_fp16=True` selects the tensor-core path (bf16 operands, fp32 accumulation and fp32
    GroupNorm/softmax/residual stream — the B200 analogue of the reference's fp16 torso,
    fp16_util.py:14-21); `use_fp16=False` selects the all-fp32 FFMA path used for 1e-4 parity.
    """

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False):
        super().__init__()
        if dims != 2:
            raise NotImplementedError("the B200 path implements the 2-D (triplane) UNet only")
        if num_classes is not None:
            raise NotImplementedError("class conditioning is outside the NFD configuration (class_cond=False)")
        if not resblock_updown:
            raise NotImplementedError("resblock_updown=False is outside the NFD configuration")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = th.float16 if use_fp16 else th.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))

        def res(cin, cout, **kw):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, **kw)

        def attn(c, heads):
            return AttentionBlock(c, use_checkpoint=use_checkpoint, num_heads=heads,
                                  num_head_channels=num_head_channels,
                                  use_new_attention_order=use_new_attention_order)

        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, ch, 3, padding=1))])
        skip_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, int(mult * model_channels))]
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(res(ch, ch, down=True)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch, num_heads), res(ch, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = skip_chans.pop()
                layers = [res(ch + ich, int(model_channels * mult))]
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, up=True))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(conv_nd(dims, input_ch, out_channels, 3, padding=1)))
        self._plans = {}
        self._native = {}
        self._ops = None
        self._mode = "bf16" if use_fp16 else "fp32"
        self.weights_generation = 0      # bumped whenever packed weight panels (plans) become stale

    # -- precision / backend selection ------------------------------------------------------------
    def convert_to_fp16(self):
        """Reference: cast the torso convs to half (unet.py:618-624).  Here: switch the plan to the
        bf16 tensor-core kernels; parameters stay fp32 masters and are packed to bf16 panels."""
        self.dtype = th.float16
        self._mode = "bf16"
        self.invalidate()

    def convert_to_fp32(self):
        self.dtype = th.float32
        self._mode = "fp32"
        self.invalidate()

    def set_ops(self, ops):
        """Inject an operator backend (tests use a pure-torch mirror on CPU)."""
        self._ops = ops
        self._mode = "bf16" if ops.lo == th.bfloat16 else "fp32"
        self.invalidate()

    def invalidate(self):
        """Parameters, precision mode or backend changed: drop every plan (their weight panels are packed COPIES) and
        tell whoever cached one (GuidedStepper and its captured graph, FiLM rows) through `weights_generation`."""
        self._plans = {}
        self._native = {}
        self.weights_generation = getattr(self, "weights_generation", 0) + 1

    def load_state_dict(self, *a, **kw):
        r = super().load_state_dict(*a, **kw)
        self.invalidate()
        return r

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        if hasattr(self, "_plans"):
            self.invalidate()
        return r

    def _get_ops(self):
        if self._ops is None or (getattr(self._ops, "name", "") == "cuda" and self._ops.mode != self._mode):
            from ..ops import CudaOps
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("UNetModel parameters are on %s: the B200 path has no CPU fallback "
                                   "(move the model to cuda)" % dev)
            self._ops = CudaOps(dev, self._mode)
        return self._ops

    def plan(self, N, H, W, want_backward=True) -> _Plan:
        if self.training and self.dropout > 0:
            raise NotImplementedError("dropout>0 in training mode is not implemented: call model.eval() "
                                      "(the editor always does, drag_utils.py:187)")
        key = (N, H, W, self._mode, want_backward)
        p = self._plans.get(key)
        if p is None:
            p = _Plan(self, self._get_ops(), N, H, W, want_backward)
            self._plans[key] = p
        return p

    def _native_plan(self, N, H, W):
        """The handle-level plan for the eager API (None when switched off with ISB_NATIVE_EAGER=0 or when a non-CUDA
        operator backend was injected by a test).  The steppers (GuidedStepper, ReconStepper, the no-grad step graph)
        keep the per-operator plan, which they capture into CUDA graphs."""
        if os.environ.get("ISB_NATIVE_EAGER", "1") == "0" or getattr(self._get_ops(), "name", "") != "cuda":
            return None
        if self.training and self.dropout > 0:
            raise NotImplementedError("dropout>0 in training mode is not implemented: call model.eval() "
                                      "(the editor always does, drag_utils.py:187)")
        key = (N, H, W, self._mode)
        nat = self._native.get(key)
        if nat is None:
            from ..native_unet import NativeUNet
            nat = NativeUNet(self, N, H, W, mode=self._mode, want_backward=True)
            self._native[key] = nat
        return nat

    def forward(self, x, timesteps, y=None, feat_layer=-1):
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        N, _, H, W = x.shape
        need_grad = th.is_grad_enabled() and x.requires_grad
        t_orig = timesteps.to(device=x.device, dtype=th.float32).contiguous()   # may be fractional (rescale_timesteps)
        nat = self._native_plan(N, H, W)
        if nat is not None:          # eager public API: the whole pass is one call into the library (csrc/unet.cu)
            if need_grad:
                return _NativeUNetFn.apply(x, self, nat, t_orig, feat_layer)
            with th.no_grad():
                return _NativeUNetFn.forward(_NoCtx(), x, self, nat, t_orig, feat_layer)
        plan = self.plan(N, H, W, want_backward=True)
        if need_grad:
            return _UNetFn.apply(x, self, plan, t_orig, feat_layer)
        with th.no_grad():
            return _UNetFn.forward(_NoCtx(), x, self, plan, t_orig, feat_layer)


class _NoCtx:
    pass
