"""Pure-torch mirror of ishapediting_b200.ops.CudaOps — TEST CODE ONLY.

Same method signatures, NHWC buffers, written with torch library calls (and autograd for the
backward ops) independently of the CUDA kernels.  Two uses:
  * on CPU: validate the UNet plan's graph / manual-backward logic against the oracle without a GPU;
  * on the GPU box: per-op reference the CUDA kernels are compared with.
`mode="bf16"` emulates the tensor-core path's operand rounding (bf16 operands, fp32 accumulate).
"""
import contextlib
import math

import torch
import torch.nn.functional as F


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


class RefOps:
    name = "ref"

    @contextlib.contextmanager
    def workspace_slot(self, slot):
        yield

    @contextlib.contextmanager
    def background(self, on=True):
        yield

    def __init__(self, mode="fp32", device="cpu"):
        self.mode = mode
        self.lo = torch.bfloat16 if mode == "bf16" else torch.float32
        self.device = torch.device(device)

    def empty(self, shape, dtype=torch.float32):
        # NaN-filled so that reads of never-written buffers show up in tests
        t = torch.empty(shape, dtype=dtype, device=self.device)
        if t.is_floating_point():
            t.fill_(float("nan"))
        return t

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    # ---- layout ----
    def to_nhwc(self, x_nchw, out):
        C = x_nchw.shape[1]
        out.zero_()
        out[..., :C] = _nhwc(x_nchw).to(out.dtype)
        return out

    def to_nchw(self, x_nhwc, out_nchw):
        C = out_nchw.shape[1]
        out_nchw.copy_(_nchw(x_nhwc[..., :C].float()))
        return out_nchw

    def cast_lo(self, src, dst):
        dst.copy_(src.to(dst.dtype))
        return dst

    # ---- conv ----
    def pack_weight(self, w2d):
        return w2d.contiguous()

    def conv_gn_slots(self, N, H, W, Cin, ksize, Cout, Cin2=0, groups=32):
        # mirror of CudaOps.conv_gn_slots: fusable in bf16 mode for group widths 8/16/32 (3 slots here)
        return 3 if (self.lo == torch.bfloat16 and Cout % groups == 0 and Cout // groups in (8, 16, 32)) else 0

    def conv(self, a, w, bias, ksize, out, a2=None, residual=None, accumulate=False, tune=None, gn_part=None,
             gn_bwd=None):
        N, H, W, Cin = a.shape
        Cout = out.shape[3]
        kk = ksize * ksize
        w = w.float()
        w3 = w[:, :kk * Cin].reshape(Cout, ksize, ksize, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(_nchw(a.float()), w3, None, padding=ksize // 2)
        if a2 is not None:
            y = y + F.conv2d(_nchw(a2.float()), w[:, kk * Cin:][:, :, None, None])
        y = _nhwc(y)
        if bias is not None:
            y = y + bias
        if residual is not None:
            y = y + residual
        if accumulate:
            y = y + out.float()
        out.copy_(y.to(out.dtype))
        if gn_part is not None and gn_bwd is None:
            # (sum, sum of squares) per (image, group): everything in slot 0, zeros elsewhere
            G = gn_part.shape[1]
            yg = out.float().reshape(N, H * W, G, Cout // G)
            gn_part.zero_()
            gn_part[:, :, 0, 0] = yg.sum(dim=(1, 3))
            gn_part[:, :, 0, 1] = (yg * yg).sum(dim=(1, 3))
        elif gn_part is not None:
            gn_part.zero_()
            t = self._gn_bwd_terms(out.float(), *gn_bwd)
            gn_part[:, :, 0, 0], gn_part[:, :, 0, 1] = t[0], t[1]
        return out

    @staticmethod
    def _gn_bwd_terms(dy, x, gamma, beta, film, film_off, silu, stats):
        """sum(dz g') and sum(dz g' xhat) per (image, 32 groups) of GroupNorm(+FiLM)(+SiLU) backward (NHWC tensors)."""
        N, H, W, C = x.shape
        ga, be = gamma.reshape(1, 1, 1, C), beta.reshape(1, 1, 1, C)
        if film is not None:
            sc = film[:, film_off:film_off + C].reshape(N, 1, 1, C)
            sh = film[:, film_off + C:film_off + 2 * C].reshape(N, 1, 1, C)
            ga, be = ga * (1 + sc), be * (1 + sc) + sh
        mean = stats[..., 0].repeat_interleave(C // 32, dim=1).reshape(N, 1, 1, C)
        rstd = stats[..., 1].repeat_interleave(C // 32, dim=1).reshape(N, 1, 1, C)
        xhat = (x - mean) * rstd
        d = dy
        if silu:
            z = xhat * ga + be
            sg = torch.sigmoid(z)
            d = d * (sg * (1 + z * (1 - sg)))
        dzg = d * ga
        g = lambda t: t.reshape(N, H * W, 32, C // 32).sum(dim=(1, 3))
        return g(dzg), g(dzg * xhat)

    # ---- group norm ----
    @staticmethod
    def _gn_fn(x, gamma, beta, film, film_off, silu, resample):
        """x: NCHW fp32 (requires_grad ok) -> (y, xres) NCHW."""
        C = x.shape[1]
        z = F.group_norm(x, 32, gamma, beta, eps=1e-5)
        if film is not None:
            sc = film[:, film_off:film_off + C][:, :, None, None]
            sh = film[:, film_off + C:film_off + 2 * C][:, :, None, None]
            z = z * (1 + sc) + sh
        a = F.silu(z) if silu else z
        xr = x
        if resample == 1:
            a, xr = F.avg_pool2d(a, 2), F.avg_pool2d(x, 2)
        elif resample == 2:
            a, xr = F.interpolate(a, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
        return a, xr

    def gn_forward(self, x1, x2, gamma, beta, film, film_off, silu, resample, stats, y, raw=None, xres=None,
                   partials=None):
        x = x1 if x2 is None else torch.cat([x1, x2], dim=3)
        xn = _nchw(x)
        a, xr = self._gn_fn(xn, gamma, beta, film, film_off, silu, resample)
        N, C = x.shape[0], x.shape[3]
        xg = xn.reshape(N, 32, -1)
        stats[..., 0] = xg.mean(-1)
        stats[..., 1] = 1.0 / torch.sqrt(xg.var(-1, unbiased=False) + 1e-5)
        if partials is not None:     # the producer's partials must describe THIS tensor (catches stale buffers)
            assert x2 is None and resample == 0
            m = xg.shape[-1]
            mean_p = partials[..., 0].sum(-1) / m
            var_p = partials[..., 1].sum(-1) / m - mean_p * mean_p
            assert torch.allclose(mean_p, stats[..., 0], atol=1e-4, rtol=1e-4), "stale GroupNorm partials"
            assert torch.allclose(var_p, xg.var(-1, unbiased=False), atol=1e-4, rtol=1e-3), "stale GroupNorm partials"
        y.copy_(_nhwc(a).to(y.dtype))
        if raw is not None:
            raw.copy_(x.to(raw.dtype))
        if xres is not None:
            xres.copy_(_nhwc(xr))
        return y

    def gn_backward(self, x1, x2, gamma, beta, film, film_off, silu, resample, stats, dy, gres, gres_at_input,
                    gx1, acc1, gx1_lo, gx2, acc2, gx2_lo, partials=None):
        if partials is not None:     # the dgrad conv's reduction terms must describe THIS (x, dy) pair
            assert x2 is None and resample == 0
            t0, t1 = self._gn_bwd_terms(dy.float(), x1, gamma, beta, film, film_off, silu, stats)
            for got, want in ((partials[..., 0].sum(-1), t0), (partials[..., 1].sum(-1), t1)):
                assert torch.allclose(got, want, atol=1e-4 * float(want.abs().max()) + 1e-12, rtol=1e-3), "stale GN-backward partials"
        x = (x1 if x2 is None else torch.cat([x1, x2], dim=3)).detach().clone()
        xn = _nchw(x).contiguous().requires_grad_(True)
        with torch.enable_grad():
            a, xr = self._gn_fn(xn, gamma, beta, film, film_off, silu, resample)
            total = (a * _nchw(dy)).sum()
            if gres is not None:
                total = total + ((xn if gres_at_input else xr) * _nchw(gres)).sum()
            (g,) = torch.autograd.grad(total, xn)
        g = _nhwc(g)
        C1 = x1.shape[3]
        for gx, acc, lo, sl in ((gx1, acc1, gx1_lo, slice(0, C1)), (gx2, acc2, gx2_lo, slice(C1, None))):
            if gx is None and lo is None:
                continue
            v = g[..., sl]
            if acc and gx is not None:
                v = v + gx
            if gx is not None:
                gx.copy_(v)
            if lo is not None:
                lo.copy_(v.to(lo.dtype))

    # ---- attention (unet.py:337-354 on NHWC-flattened qkv) ----
    @staticmethod
    def _attn_fn(qkv, heads):
        N, H, W, C3 = qkv.shape
        T = H * W
        ch = C3 // (3 * heads)
        q, k, v = qkv.reshape(N, T, heads, 3, ch).unbind(3)          # each [N,T,heads,ch]
        s = torch.einsum("nthc,nshc->nhts", q, k) / math.sqrt(ch)
        p = torch.softmax(s.float(), dim=-1)
        o = torch.einsum("nhts,nshc->nthc", p, v).reshape(N, H, W, heads * ch)
        return p, o

    def attention_forward(self, qkv, heads, probs, out):
        p, o = self._attn_fn(qkv, heads)
        probs.copy_(p)
        out.copy_(o.to(out.dtype))
        return out

    def attention_backward(self, qkv, probs, d_out, heads, tmp, d_qkv):
        q = qkv.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            _, o = self._attn_fn(q, heads)
            (g,) = torch.autograd.grad((o * d_out).sum(), q)
        d_qkv.copy_(g.to(d_qkv.dtype))
        return d_qkv

    def attention_flash_forward(self, qkv, heads, out, lse):
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        x = qkv.float().reshape(N, T, heads, 3, ch)
        q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]
        s = torch.einsum("nthc,nshc->nhts", q, k) / math.sqrt(ch)
        lse.copy_(torch.logsumexp(s, dim=-1) / math.log(2.0))
        o = torch.einsum("nhts,nshc->nthc", torch.softmax(s, dim=-1), v)
        out.copy_(o.reshape(N, H, W, heads * ch).to(out.dtype))
        return out

    def attention_flash_backward(self, qkv, out, d_out, lse, heads, delta, d_qkv):
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        x = qkv.detach().float().clone().requires_grad_(True)
        with torch.enable_grad():
            y = x.reshape(N, T, heads, 3, ch)
            s = torch.einsum("nthc,nshc->nhts", y[:, :, :, 0], y[:, :, :, 1]) / math.sqrt(ch)
            o = torch.einsum("nhts,nshc->nthc", torch.softmax(s, dim=-1), y[:, :, :, 2]).reshape(N, H, W, heads * ch)
            (g,) = torch.autograd.grad((o * d_out.float()).sum(), x)
        delta.copy_((d_out.float() * out.float()).reshape(N, T, heads, ch).sum(-1).permute(0, 2, 1))
        d_qkv.copy_(g.to(d_qkv.dtype))
        return d_qkv

    # ---- timestep embedding ----
    def time_embed(self, t, freqs, w1, b1, w2, b2, w_all, b_all, scratch, film_all):
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        h = F.linear(F.silu(F.linear(emb, w1, b1)), w2, b2)
        film_all.copy_(F.linear(F.silu(h), w_all, b_all))
        return film_all

    # ---- DDPM ----
    def ddpm_step(self, x, model_out, coef, clip_denoised, noise=None, grad=None, x_next=None, sample=None, mean=None,
                  var=None, x0=None, eps=None, *, model_out_nhwc=False):
        C = x.shape[1]
        mo = _nchw(model_out) if model_out_nhwc else model_out
        e, v = mo[:, :C], mo[:, C:2 * C]
        c = coef if coef.dim() == 1 else [coef[:, k].reshape(-1, 1, 1, 1) for k in range(8)]
        frac = (v + 1) / 2
        variance = torch.exp(frac * c[5] + (1 - frac) * c[4])
        xs = c[0] * x - c[1] * e
        if clip_denoised:
            xs = xs.clamp(-1, 1)
        mu = c[2] * xs + c[3] * x
        smp = mu if noise is None else mu + c[6] * torch.sqrt(variance) * noise
        nxt = smp if grad is None else smp + variance * (c[7] * grad)
        for dst, val in ((x_next, nxt), (sample, smp), (mean, mu), (var, variance), (x0, xs), (eps, e)):
            if dst is not None:
                dst.copy_(val)

    # ---- drag ----
    def resize_feat_align(self, feat, chan_map, out):
        S, Ca = feat.shape[1], out.shape[3]
        out.copy_(feat[0][:, :, chan_map.long()].reshape(S, S, 3, Ca).permute(2, 0, 1, 3))
        return out

    def drag_partial_len(self, S, Cf, npts):
        return 1

    def drag_loss_grad(self, feat, origin, chan_map, inv_map, patch_xy, shift_xy, weight, group_size, bbox, mask,
                       mask_count, inv_count, cof, loss_type, g, pt_info, partial, loss, d_feat, dyn=None):
        if dyn is not None:
            inv_count = float(dyn[0])
        S, Cf, Ca = feat.shape[1], feat.shape[3], origin.shape[3]
        f = feat.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            edit = f[0][:, :, chan_map.long()].reshape(S, S, 3, Ca).permute(2, 3, 0, 1)      # [3,Ca,S,S]
            orig = origin.permute(0, 3, 1, 2)
            pf = F.grid_sample(orig, patch_xy[:, None], mode="bilinear", padding_mode="zeros", align_corners=True)
            sf = F.grid_sample(edit, shift_xy[:, None], mode="bilinear", padding_mode="zeros", align_corners=True)
            diff = (sf - pf.detach())[:, :, 0, :]                                            # [3,Ca,npts]
            wsum = weight[None, None, :]
            if loss_type == 0:
                motion = (diff ** 2 * wsum).sum() * inv_count
            else:
                motion = (diff.abs() * wsum).sum() * inv_count
            total = -motion
            if cof > 0:
                m = mask.float()[:, None]
                dm = (edit - orig) * m
                mnorm = float(dyn[1]) if dyn is not None else 1.0 / (Ca * mask_count)
                ml = ((dm ** 2).sum() if loss_type == 0 else dm.abs().sum()) * mnorm
                total = total - cof * ml
            (gf,) = torch.autograd.grad(total, f)
        loss.copy_(total.detach().reshape(1))
        d_feat.copy_(gf)

    # ---- decoder ----
    @staticmethod
    def _decode(planes_hwc, weights, coords):
        B, w1, b1, w2, b2, w3, b3 = weights
        planes = planes_hwc.permute(0, 3, 1, 2)
        c = coords[None]

        def sample(c2, plane):
            s = F.grid_sample(plane[None], c2.reshape(1, 1, -1, 2), mode="bilinear", padding_mode="zeros",
                              align_corners=True)
            return s.reshape(plane.shape[0], -1).t()

        f = sample(c[..., 0:2], planes[0]) + sample(c[..., 1:3], planes[1]) + sample(c[..., :3:2], planes[2])
        x = 2 * math.pi * (f @ B)
        x = torch.cat([torch.sin(x), torch.cos(x)], dim=-1)
        x = F.relu(F.linear(x, w1, b1))
        x = F.relu(F.linear(x, w2, b2))
        return F.linear(x, w3, b3).reshape(-1)

    def decode_points(self, planes_hwc, weights, coords, out):
        out.copy_(self._decode(planes_hwc, weights, coords))
        return out

    def decode_points_backward(self, planes_hwc, weights, coords, d_logits, d_planes_hwc):
        p = planes_hwc.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (g,) = torch.autograd.grad((self._decode(p, weights, coords) * d_logits).sum(), p)
        d_planes_hwc.add_(g)
        return d_planes_hwc

    def decode_grid(self, planes_hwc, weights, lin, x_begin, x_end, out):
        xs, ys, zs = torch.meshgrid([lin[x_begin:x_end], lin, lin], indexing="ij")
        coords = torch.stack([xs, ys, zs], -1).reshape(-1, 3)
        out.copy_(self._decode(planes_hwc, weights, coords))
        return out
