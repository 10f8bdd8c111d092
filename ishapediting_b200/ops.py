"""Operator layer: torch CUDA tensors in, C-ABI calls on torch's current stream out.

Everything is NHWC and preallocated by the caller (the UNet plan); PyTorch is used only for
device memory and streams.  `CudaOps` is the only backend the product ships — the pure-torch
mirror of this interface used to validate the graph logic on CPU lives under tests/ and is never
imported from here.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import torch

from . import _lib
from ._lib import BF16, F32

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "ops expect contiguous CUDA tensors"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


class PackedWeight:
    """A GEMM weight panel in the backend's preferred memory layout (see CudaOps.pack_weight)."""
    __slots__ = ("data", "cout", "k", "tiled")

    def __init__(self, data, cout, k, tiled):
        self.data, self.cout, self.k, self.tiled = data, cout, k, tiled

    @property
    def shape(self):
        return (self.cout, self.k)


class CudaOps:
    name = "cuda"

    def __init__(self, device=None, mode: str = "bf16"):
        if not torch.cuda.is_available():
            raise _lib.IsbError("ishapediting_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        assert mode in ("bf16", "fp32")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mode = mode
        self.lo = torch.bfloat16 if mode == "bf16" else torch.float32
        self.lib = _lib.init(self.device.index or 0)
        self.tile_weights = os.environ.get("ISB_TILED_WEIGHTS", "0") == "1"
        self._ws = {}
        self._ws_slot = 0
        self._bg = False
        self._gn_scratch = {}
        # Buffers that were handed to launches and later outgrown.  Captured CUDA graphs keep RAW pointers into them
        # (GuidedStepper / ReconStepper / the no-grad step graph share this ops object per model), and their arrival
        # counters must stay 0 between launches, so they are never returned to the caching allocator.
        self._retired = []
        self._mlp_bounds = {}
        self._pf_rec = None          # weight-prefetch recording: [(ptr, bytes)] of the conv launches of one pass
        self._pf = None              # active prefetch schedule (inside weight_prefetch())

    # ---- memory -----------------------------------------------------------
    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def _workspace(self, nbytes):
        """Split-K partial tiles of the conv kernel.  One buffer per concurrency SLOT (see workspace_slot): the
        guided step runs convolutions on up to three streams at once (forward tail, ResBlock skip dgrad), and a
        launch owns its workspace from its first partial store to its last fold load."""
        if nbytes == 0:
            return None, 0
        key = self._ws_slot
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.IsbError("conv workspace must be sized by an eager warm-up before graph capture")
            if ws is not None:
                self._retired.append(ws)      # graphs captured earlier still point at it
            # zero-filled: the split-K arrival counters at its head must start (and are left) at 0
            ws = torch.zeros(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws, ws.numel()

    @contextlib.contextmanager
    def workspace_slot(self, slot):
        """Launches made inside this context (on a side stream) use their own conv workspace."""
        prev, self._ws_slot = self._ws_slot, slot
        try:
            yield
        finally:
            self._ws_slot = prev

    @contextlib.contextmanager
    def background(self, on=True):
        """Convolutions launched inside this context belong to work that runs BESIDE a latency-critical stream (the
        forward tail next to the backward pass): they ask for more than half an SM's shared memory, so only one of
        their CTAs is resident per SM and the critical stream's CTAs always find room next to it."""
        prev, self._bg = self._bg, on
        try:
            yield
        finally:
            self._bg = prev

    # ---- background L2 prefetch of the weight panels ------------------------------------------------
    # At batch 1 every launch of the step is latency-bound and HBM sits idle ~95 % of the time, while each conv starts by
    # pulling COLD weights (1.57 GB per step against 126 MB of L2) through a 3-4 stage ring: the small-spatial layers run
    # at ~10 % of their weight-streaming roofline (profiles/r02_ncu_step_traffic.md).  Inside a captured step the launch
    # sequence is static, so a side stream can pull the panels of the next layers into L2 while the current ones run.
    @contextlib.contextmanager
    def record_weight_sequence(self):
        """Record (pointer, bytes) of every conv weight panel launched inside the context, in issue order."""
        self._pf_rec = []
        try:
            yield self._pf_rec
        finally:
            self._pf_rec = None

    @contextlib.contextmanager
    def weight_prefetch(self, seq, stream, window_bytes=48 << 20, min_bytes=1 << 20):
        """Inside the context (a graph capture of the pass that produced `seq`) the k-th conv launch first gates the L2
        prefetch of the panels of the following launches — as many as fit `window_bytes` — on its own start, on
        `stream`.  The caller's stream joins `stream` when the context ends."""
        st = self._pf = {"seq": seq, "k": 0, "upto": 0, "ahead": 0, "stream": stream, "window": window_bytes,
                         "min": min_bytes, "used": False}
        try:
            yield
        finally:
            self._pf = None
            if st["used"]:
                ev = torch.cuda.Event()
                ev.record(stream)
                torch.cuda.current_stream().wait_event(ev)

    def _pf_step(self, ptr):
        st = self._pf
        seq, k = st["seq"], st["k"]
        st["k"] = k + 1
        if k >= len(seq) or seq[k][0] != ptr:        # the pass diverged from the recording: stop prefetching
            st["k"] = len(seq) + 1
            return
        if st["upto"] <= k:                           # this launch was not prefetched (first launches, window jumps)
            st["upto"], st["ahead"] = k + 1, 0
        else:
            st["ahead"] -= seq[k][1]
        todo = []
        while st["upto"] < len(seq) and (st["ahead"] + seq[st["upto"]][1] <= st["window"] or st["upto"] == k + 1):
            j = st["upto"]
            st["upto"] += 1
            st["ahead"] += seq[j][1]
            if seq[j][1] >= st["min"]:
                todo.append(seq[j])
        if todo:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())    # = "launch k is about to start"
            st["stream"].wait_event(ev)
            sp = C.c_void_p(st["stream"].cuda_stream)
            for ptr_j, nbytes in todo:
                _lib.check(self.lib.isb_prefetch_l2(C.c_void_p(ptr_j), nbytes, sp), "isb_prefetch_l2")
            st["used"] = True

    def _scratch(self, N, groups=32, which="fwd"):
        """GroupNorm reduction scratch (arrival counters + partials).  Forward and backward kernels get
        separate buffers: the guided step runs the tail of the forward pass concurrently with the backward
        pass on a second stream."""
        need = N * groups
        cur = self._gn_scratch.get(which)
        if cur is None or cur[1] < need:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.IsbError("GroupNorm scratch must be sized by an eager warm-up before graph capture")
            if cur is not None:
                self._retired.append(cur[0])  # graphs captured at a smaller batch still point at it
            nbytes = self.lib.isb_gn_scratch_bytes(N, groups)
            cur = (torch.zeros(nbytes, dtype=torch.uint8, device=self.device), need)
            self._gn_scratch[which] = cur
        return cur[0]

    # ---- layout -------------------------------------------------------------
    def to_nhwc(self, x_nchw, out):
        N, Cc, H, W = x_nchw.shape
        _chk(x_nchw, torch.float32); _chk(out)
        assert out.shape[0] == N and out.shape[1] == H and out.shape[2] == W
        _lib.check(self.lib.isb_nchw_to_nhwc(_p(x_nchw), _p(out), _DT[out.dtype], N, Cc, H, W, out.shape[3], _stream()),
                   "isb_nchw_to_nhwc")
        return out

    def to_nchw(self, x_nhwc, out_nchw):
        N, Cc, H, W = out_nchw.shape
        _chk(x_nhwc); _chk(out_nchw, torch.float32)
        _lib.check(self.lib.isb_nhwc_to_nchw(_p(x_nhwc), _DT[x_nhwc.dtype], _p(out_nchw), N, Cc, H, W,
                                             x_nhwc.shape[3], _stream()), "isb_nhwc_to_nchw")
        return out_nchw

    def cast_lo(self, src, dst):
        _chk(src, torch.float32); _chk(dst)
        if dst.dtype == torch.float32:
            dst.copy_(src)
        else:
            _lib.check(self.lib.isb_cast_f32_bf16(_p(src), _p(dst), src.numel(), _stream()), "isb_cast_f32_bf16")
        return dst

    # ---- conv / linear --------------------------------------------------------
    def pack_weight(self, w2d):
        """[Cout, K] row-major -> the layout the conv kernel streams best.  bf16 mode with Cout, K multiples
        of 64: 64x64 panels, [Cout/64][K/64][64][64], so that every TMA box is a run of contiguous 8 KiB
        blocks and the K loop walks HBM sequentially (the row-major layout scatters each box over `bn`
        rows that are K*2 bytes apart)."""
        cout, k = w2d.shape
        # measured (profiles/r01_ablation.md): panels are NOT faster than row-major on B200 — the 128-byte
        # rows of a box already spread over HBM channels — so the layout is opt-in (ISB_TILED_WEIGHTS=1)
        if self.tile_weights and self.mode == "bf16" and cout % 64 == 0 and k % 64 == 0:
            t = w2d.reshape(cout // 64, 64, k // 64, 64).permute(0, 2, 1, 3).contiguous()
            return PackedWeight(t, cout, k, True)
        return PackedWeight(w2d.contiguous(), cout, k, False)

    def conv_gn_slots(self, N, H, W, Cin, ksize, Cout, Cin2=0, groups=32):
        """Contributions per (image, group) a bf16 conv of this shape writes when the GroupNorm statistics of its
        output are fused into its epilogue; 0 = not fusable (fp32 mode, group width other than 8/16/32, ...)."""
        if self.lo != torch.bfloat16 or Cout % groups or os.environ.get("ISB_GN_FUSE", "1") == "0":
            return 0
        d = _lib.ConvDesc()
        d.a, d.a_dtype, d.w, d.out, d.out_dtype = 256, BF16, 256, 256, F32      # pointers only need to be non-null
        d.N, d.H, d.W, d.Cin, d.ksize, d.Cout = N, H, W, Cin, ksize, Cout
        if Cin2:
            d.a2, d.Cin2 = 256, Cin2
        d.gn_cg = Cout // groups
        return int(self.lib.isb_conv2d_gn_slots(C.byref(d)))

    def conv(self, a, w, bias, ksize, out, a2=None, residual=None, accumulate=False, tune=None, gn_part=None,
             gn_bwd=None):
        """a [N,H,W,Cin], w PackedWeight (or a plain [Cout, k*k*Cin (+Cin2)] tensor), out [N,H,W,Cout].
        gn_part: optional fp32 [N, groups, slots, 2] buffer (slots = conv_gn_slots(...)) that receives the GroupNorm
        statistics partials of `out`.  gn_bwd = (x, gamma, beta, film, film_off, silu, stats): `out` is dy of that
        GroupNorm layer and gn_part receives the two reduction terms of its backward instead."""
        tiled = False
        if isinstance(w, PackedWeight):
            wshape, tiled, w = (w.cout, w.k), w.tiled, w.data
        else:
            wshape = tuple(w.shape)
        _chk(a); _chk(w, a.dtype); _chk(out)
        N, H, W, Cin = a.shape
        d = _lib.ConvDesc()
        d.a, d.a_dtype = _p(a), _DT[a.dtype]
        d.N, d.H, d.W, d.Cin, d.ksize = N, H, W, Cin, ksize
        if a2 is not None:
            _chk(a2, a.dtype)
            assert a2.shape[:3] == a.shape[:3]
            d.a2, d.Cin2 = _p(a2), a2.shape[3]
        d.w = _p(w)
        d.bias = _p(_chk(bias, torch.float32)) if bias is not None else None
        if residual is not None:
            _chk(residual, torch.float32)
            assert residual.shape == out.shape
            d.residual = _p(residual)
        d.out, d.out_dtype, d.Cout = _p(out), _DT[out.dtype], out.shape[3]
        assert wshape[0] == d.Cout and wshape[1] == ksize * ksize * Cin + (a2.shape[3] if a2 is not None else 0), \
            f"packed weight {wshape} does not match conv {Cin}->{d.Cout} k{ksize}"
        d.accumulate = int(accumulate)
        d.w_tiled = int(tiled)
        if tune:
            d.block_n, d.split_k, d.stages = tune.get("block_n", 0), tune.get("split_k", 0), tune.get("stages", 0)
            d.two_cta = tune.get("two_cta", 0)
            d.debug_flags = tune.get("debug", 0)
        if self._bg:
            d.min_smem_bytes = 116 * 1024
        if gn_part is not None:
            _chk(gn_part, torch.float32)
            assert gn_part.shape[0] == N and gn_part.shape[3] == 2 and d.Cout % gn_part.shape[1] == 0
            d.gn_partials, d.gn_cg, d.gn_slots = _p(gn_part), d.Cout // gn_part.shape[1], gn_part.shape[2]
            if gn_bwd is not None:
                gx, gamma, beta, film, film_off, silu, stats = gn_bwd
                assert gx.shape == out.shape and out.dtype == torch.float32
                d.gn_mode, d.gb_x = 2, _p(_chk(gx, torch.float32))
                d.gb_gamma, d.gb_beta = _p(_chk(gamma, torch.float32)), _p(_chk(beta, torch.float32))
                d.gb_stats, d.gb_silu = _p(_chk(stats, torch.float32)), int(silu)
                if film is not None:
                    d.gb_film = C.c_void_p(_chk(film, torch.float32).data_ptr() + 4 * film_off)
                    d.gb_film_stride = film.shape[1]
        if tune and tune.get("trace") is not None:      # profiling: debug bit 2 -> phase stamps land in this tensor
            self.lib.isb_debug_set_trace(_p(tune["trace"]))
        ws, ws_bytes = self._workspace(self.lib.isb_conv2d_workspace(C.byref(d)))
        if self._pf_rec is not None:
            self._pf_rec.append((w.data_ptr(), w.numel() * w.element_size()))
        if self._pf is not None:
            self._pf_step(w.data_ptr())
        _lib.check(self.lib.isb_conv2d(C.byref(d), _p(ws), ws_bytes, _stream()), "isb_conv2d")
        return out

    # ---- group norm -------------------------------------------------------------
    def _gn_desc(self, x1, x2, gamma, beta, film, film_off, silu, resample, stats):
        d = _lib.GnDesc()
        _chk(x1, torch.float32)
        N, H, W, C1 = x1.shape
        d.x1, d.C1 = _p(x1), C1
        if x2 is not None:
            _chk(x2, torch.float32)
            assert x2.shape[:3] == x1.shape[:3]
            d.x2, d.C2 = _p(x2), x2.shape[3]
        d.N, d.H, d.W = N, H, W
        d.groups, d.eps = 32, 1e-5
        d.gamma, d.beta = _p(_chk(gamma, torch.float32)), _p(_chk(beta, torch.float32))
        if film is not None:
            _chk(film, torch.float32)
            d.film = C.c_void_p(film.data_ptr() + 4 * film_off)
            d.film_stride = film.shape[1]
        d.silu, d.resample = int(silu), int(resample)
        d.stats = _p(_chk(stats, torch.float32))
        return d

    def gn_forward(self, x1, x2, gamma, beta, film, film_off, silu, resample, stats, y, raw=None, xres=None,
                   partials=None):
        d = self._gn_desc(x1, x2, gamma, beta, film, film_off, silu, resample, stats)
        if partials is not None:     # statistics already accumulated by the producer conv (conv(..., gn_part=))
            _chk(partials, torch.float32)
            assert x2 is None and resample == 0 and partials.shape[:2] == (x1.shape[0], 32)
            d.partials, d.partial_slots = _p(partials), partials.shape[2]
        d.y, d.y_dtype = _p(_chk(y)), _DT[y.dtype]
        if raw is not None:
            d.raw, d.raw_dtype = _p(_chk(raw)), _DT[raw.dtype]
        if xres is not None:
            d.xres = _p(_chk(xres, torch.float32))
        _lib.check(self.lib.isb_gn_forward(C.byref(d), _p(self._scratch(x1.shape[0])), _stream()), "isb_gn_forward")
        return y

    def gn_backward(self, x1, x2, gamma, beta, film, film_off, silu, resample, stats, dy, gres, gres_at_input,
                    gx1, acc1, gx1_lo, gx2, acc2, gx2_lo, partials=None):
        b = _lib.GnBwdDesc()
        if partials is not None:     # reduction terms already accumulated by the dgrad conv (conv(..., gn_bwd=))
            _chk(partials, torch.float32)
            assert x2 is None and resample == 0 and partials.shape[:2] == (x1.shape[0], 32)
            b.partials, b.partial_slots = _p(partials), partials.shape[2]
        b.f = self._gn_desc(x1, x2, gamma, beta, film, film_off, silu, resample, stats)
        b.dy = _p(_chk(dy, torch.float32))
        if gres is not None:
            b.gres, b.gres_at_input = _p(_chk(gres, torch.float32)), int(gres_at_input)
        lo_dtype = None
        if gx1 is not None:
            b.gx1, b.acc1 = _p(_chk(gx1, torch.float32)), int(acc1)
        if gx1_lo is not None:
            b.gx1_lo, lo_dtype = _p(_chk(gx1_lo)), gx1_lo.dtype
        if gx2 is not None:
            b.gx2, b.acc2 = _p(_chk(gx2, torch.float32)), int(acc2)
        if gx2_lo is not None:
            assert lo_dtype is None or lo_dtype == gx2_lo.dtype
            b.gx2_lo, lo_dtype = _p(_chk(gx2_lo)), gx2_lo.dtype
        b.lo_dtype = _DT[lo_dtype] if lo_dtype is not None else F32
        _lib.check(self.lib.isb_gn_backward(C.byref(b), _p(self._scratch(x1.shape[0], which="bwd")), _stream()),
                   "isb_gn_backward")

    # ---- attention ----------------------------------------------------------------
    def attention_forward(self, qkv, heads, probs, out):
        _chk(qkv, torch.float32); _chk(probs, torch.float32); _chk(out)
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        _lib.check(self.lib.isb_attention_forward(_p(qkv), N, T, heads, ch, _p(probs), _p(out), _DT[out.dtype],
                                                  _stream()), "isb_attention_forward")
        return out

    def attention_backward(self, qkv, probs, d_out, heads, tmp, d_qkv):
        _chk(qkv, torch.float32); _chk(probs, torch.float32); _chk(d_out, torch.float32); _chk(tmp, torch.float32)
        _chk(d_qkv)
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        _lib.check(self.lib.isb_attention_backward(_p(qkv), _p(probs), _p(d_out), N, T, heads, ch, _p(tmp), _p(d_qkv),
                                                   _DT[d_qkv.dtype], _stream()), "isb_attention_backward")
        return d_qkv

    def attention_flash_forward(self, qkv, heads, out, lse):
        """Fused attention (bf16 mode, 64 channels per head): qkv [N,H,W,3C] bf16 -> out [N,H,W,C] bf16, lse [N,heads,T]."""
        _chk(qkv, torch.bfloat16); _chk(out, torch.bfloat16); _chk(lse, torch.float32)
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        _lib.check(self.lib.isb_attention_flash_forward(_p(qkv), N, T, heads, ch, _p(out), _p(lse), _stream()),
                   "isb_attention_flash_forward")
        return out

    def attention_flash_backward(self, qkv, out, d_out, lse, heads, delta, d_qkv):
        for x in (qkv, out, d_out, d_qkv):
            _chk(x, torch.bfloat16)
        _chk(lse, torch.float32); _chk(delta, torch.float32)
        N, H, W, C3 = qkv.shape
        T, ch = H * W, C3 // (3 * heads)
        _lib.check(self.lib.isb_attention_flash_backward(_p(qkv), _p(out), _p(d_out), _p(lse), N, T, heads, ch,
                                                         _p(delta), _p(d_qkv), _stream()),
                   "isb_attention_flash_backward")
        return d_qkv

    # ---- timestep embedding ---------------------------------------------------------
    def time_embed(self, t, freqs, w1, b1, w2, b2, w_all, b_all, scratch, film_all):
        _chk(t, torch.float32)
        N = t.shape[0]
        model_ch, hidden = w1.shape[1], w1.shape[0]
        assert scratch.numel() >= N * (model_ch + 2 * hidden)
        _lib.check(self.lib.isb_time_embed(_p(t), _p(freqs), N, model_ch, hidden, _p(w1), _p(b1), _p(w2), _p(b2),
                                           _p(w_all), _p(b_all), w_all.shape[0], _p(scratch), _p(film_all), _stream()),
                   "isb_time_embed")
        return film_all

    # ---- DDPM update -------------------------------------------------------------------
    def ddpm_step(self, x, model_out, coef, clip_denoised, noise=None, grad=None, x_next=None, sample=None, mean=None,
                  var=None, x0=None, eps=None, *, model_out_nhwc=False):
        """model_out: the UNet output (eps | v).  Its layout is stated by the caller, never guessed from the shape:
        NCHW [N,2C,H,W] (what UNetModel.forward returns) by default, `model_out_nhwc=True` for the plan's own
        channels-last buffer [N,H,W,>=2C]."""
        _chk(x, torch.float32); _chk(model_out, torch.float32); _chk(coef, torch.float32)
        N, Cc, H, W = x.shape
        d = _lib.DdpmDesc()
        d.x, d.model_out = _p(x), _p(model_out)
        if model_out_nhwc:
            assert tuple(model_out.shape[:3]) == (N, H, W) and model_out.shape[3] >= 2 * Cc, \
                f"NHWC model output {tuple(model_out.shape)} does not match x {tuple(x.shape)}"
            d.model_out_nchw, d.model_out_cstride = 0, model_out.shape[3]
        else:
            assert tuple(model_out.shape) == (N, 2 * Cc, H, W), \
                f"NCHW model output {tuple(model_out.shape)} does not match x {tuple(x.shape)}"
            d.model_out_nchw, d.model_out_cstride = 1, 0
        d.noise, d.grad, d.coef = _p(noise), _p(grad), _p(coef)
        d.N, d.C, d.H, d.W, d.clip_denoised = N, Cc, H, W, int(clip_denoised)
        d.x_next, d.sample, d.mean, d.var, d.x0, d.eps = _p(x_next), _p(sample), _p(mean), _p(var), _p(x0), _p(eps)
        if coef.dim() == 2:
            assert coef.shape == (N, 8)
            d.coef_per_sample = 1
        _lib.check(self.lib.isb_ddpm_step(C.byref(d), _stream()), "isb_ddpm_step")

    # ---- drag guidance --------------------------------------------------------------------
    def resize_feat_align(self, feat, chan_map, out):
        _chk(feat, torch.float32); _chk(chan_map, torch.int32); _chk(out, torch.float32)
        S, Cf, Ca = feat.shape[1], feat.shape[3], out.shape[3]
        _lib.check(self.lib.isb_resize_feat_align(_p(feat), S, Cf, _p(chan_map), _p(out), Ca, _stream()),
                   "isb_resize_feat_align")
        return out

    def drag_partial_len(self, S, Cf, npts):
        return int(self.lib.isb_drag_partial_len(S, Cf, npts))

    def drag_loss_grad(self, feat, origin, chan_map, inv_map, patch_xy, shift_xy, weight, group_size, bbox, mask,
                       mask_count, inv_count, cof, loss_type, g, pt_info, partial, loss, d_feat, dyn=None):
        d = _lib.DragDesc()
        _chk(feat, torch.float32); _chk(origin, torch.float32); _chk(d_feat, torch.float32)
        d.feat, d.S, d.Cf = _p(feat), feat.shape[1], feat.shape[3]
        d.origin, d.Ca = _p(origin), origin.shape[3]
        d.chan_map, d.inv_map = _p(_chk(chan_map, torch.int32)), _p(_chk(inv_map, torch.int32))
        d.patch_xy, d.shift_xy, d.weight = _p(_chk(patch_xy, torch.float32)), _p(_chk(shift_xy, torch.float32)), \
            _p(_chk(weight, torch.float32))
        d.npts, d.group_size = patch_xy.shape[1], group_size
        d.bbox = _p(_chk(bbox, torch.int32))
        d.mask, d.mask_count = (_p(_chk(mask, torch.uint8)) if mask is not None else None), int(mask_count)
        d.inv_count, d.cof, d.loss_type = float(inv_count), float(cof), int(loss_type)
        d.g, d.pt_info = _p(_chk(g, torch.float32)), _p(_chk(pt_info, torch.float32))
        d.partial, d.partial_len = _p(_chk(partial, torch.float64)), partial.numel()
        d.loss, d.d_feat = _p(_chk(loss, torch.float32)), _p(d_feat)
        d.dyn_scalars = _p(_chk(dyn, torch.float32)) if dyn is not None else None
        _lib.check(self.lib.isb_drag_loss_grad(C.byref(d), _stream()), "isb_drag_loss_grad")

    def track_points(self, feat, chan_map, f0, center, r, voxel, table=None):
        """Opt-in nearest-feature tracking (isb_track_points): feat [1,S,S,Cf] fp32 NHWC, f0 [B,3,Ca], center [B,3].
        Returns (idx int32 [B], dist [B], pts [B,3], table [B,3,(2r+1)^2])."""
        _chk(feat, torch.float32); _chk(chan_map, torch.int32); _chk(f0, torch.float32); _chk(center, torch.float32)
        B, side = center.shape[0], 2 * r + 1
        assert f0.shape[0] == B and f0.shape[1] == 3 and chan_map.numel() == 3 * f0.shape[2]
        if table is None:
            table = self.empty((B, 3, side * side))
        idx = torch.empty((B,), dtype=torch.int32, device=self.device)
        dist, pts = self.empty((B,)), self.empty((B, 3))
        d = _lib.TrackDesc()
        d.feat, d.S, d.Cf, d.Ca = _p(feat), feat.shape[1], feat.shape[3], f0.shape[2]
        d.chan_map, d.f0, d.center = _p(chan_map), _p(f0), _p(center)
        d.B, d.r, d.voxel = B, int(r), float(voxel)
        d.table, d.out_idx, d.out_dist, d.out_pts = _p(_chk(table, torch.float32)), _p(idx), _p(dist), _p(pts)
        _lib.check(self.lib.isb_track_points(C.byref(d), _stream()), "isb_track_points")
        return idx, dist, pts, table

    # ---- meshing (marching cubes + Laplacian smoothing) ------------------------------------------
    def marching_cubes(self, vol, iso=0.0, scale_div=0.0):
        """vol (res,res,res) fp32 -> (verts [nv,3] fp32, tris [nt,3] int32) on the device (isb_mc_count / isb_mc_emit).
        One host read-back of the two counts (the output size is data-dependent)."""
        _chk(vol, torch.float32)
        res = vol.shape[0]
        assert tuple(vol.shape) == (res, res, res)
        nbytes = int(self.lib.isb_mc_workspace_bytes(res))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        _lib.check(self.lib.isb_mc_count(_p(vol), res, float(iso), _p(ws), nbytes, _p(counts), _stream()), "isb_mc_count")
        nv, nt = (int(v) for v in counts.tolist())
        verts = torch.empty((nv, 3), dtype=torch.float32, device=self.device)
        tris = torch.empty((nt, 3), dtype=torch.int32, device=self.device)
        if nv or nt:
            _lib.check(self.lib.isb_mc_emit(_p(vol), res, float(iso), _p(ws), nbytes, float(scale_div), _p(verts), _p(tris),
                                            _stream()), "isb_mc_emit")
        return verts, tris

    def smooth_simple(self, verts, tris, iterations):
        """In-place uniform Laplacian smoothing (isb_mesh_smooth_simple)."""
        _chk(verts, torch.float32); _chk(tris, torch.int32)
        nv, nt = verts.shape[0], tris.shape[0]
        if nv == 0 or nt == 0 or iterations <= 0:
            return verts
        nbytes = int(self.lib.isb_mesh_smooth_workspace_bytes(nv, nt))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.isb_mesh_smooth_simple(_p(verts), nv, _p(tris), nt, int(iterations), _p(ws), nbytes, _stream()),
                   "isb_mesh_smooth_simple")
        return verts

    # ---- triplane decoder ---------------------------------------------------------------------
    def _mlp(self, weights):
        m = _lib.TriplaneMlp()
        m.fourier_B, m.w1, m.b1, m.w2, m.b2, m.w3, m.b3 = (_p(_chk(w, torch.float32)) for w in weights)
        # bound of the hidden pre-activations (layer-1 inputs are sin / cos): selects the split-fp16 tcgen05 decoder
        # when it is fp16-safe.  One device reduction + read-back per weight version, never inside a graph capture.
        w1, b1 = weights[1], weights[2]
        key = (w1.data_ptr(), w1._version, b1.data_ptr(), b1._version)
        bound = self._mlp_bounds.get(key)
        if bound is None:
            if torch.cuda.is_current_stream_capturing() or os.environ.get("ISB_DECODE_TC", "1") == "0":
                bound = 0.0
            else:
                bound = float((w1.abs().sum(dim=1) + b1.abs()).max().item())
                if len(self._mlp_bounds) > 64:
                    self._mlp_bounds.clear()
                self._mlp_bounds[key] = bound
        m.h1_bound = bound
        return m

    def decode_grid(self, planes_hwc, weights, lin, x_begin, x_end, out):
        _chk(planes_hwc, torch.float32); _chk(lin, torch.float32); _chk(out, torch.float32)
        m = self._mlp(weights)
        _lib.check(self.lib.isb_triplane_decode_grid(_p(planes_hwc), planes_hwc.shape[1], C.byref(m), _p(lin),
                                                     lin.numel(), x_begin, x_end, _p(out), _stream()),
                   "isb_triplane_decode_grid")
        return out

    def decode_points_backward(self, planes_hwc, weights, coords, d_logits, d_planes_hwc):
        """d_planes_hwc (3,R,R,32) += d(sum_i d_logits[i] * logit_i) / d planes   (zero-fill it first)."""
        if coords.shape[0] == 0:
            return d_planes_hwc
        for t in (planes_hwc, coords, d_logits, d_planes_hwc):
            _chk(t, torch.float32)
        assert d_planes_hwc.shape == planes_hwc.shape and d_logits.numel() == coords.shape[0]
        m = self._mlp(weights)
        _lib.check(self.lib.isb_triplane_decode_points_backward(_p(planes_hwc), planes_hwc.shape[1], C.byref(m),
                                                                _p(coords), coords.shape[0], _p(d_logits),
                                                                _p(d_planes_hwc), _stream()),
                   "isb_triplane_decode_points_backward")
        return d_planes_hwc

    def decode_points(self, planes_hwc, weights, coords, out):
        if coords.shape[0] == 0:
            return out
        _chk(planes_hwc, torch.float32); _chk(coords, torch.float32); _chk(out, torch.float32)
        m = self._mlp(weights)
        _lib.check(self.lib.isb_triplane_decode_points(_p(planes_hwc), planes_hwc.shape[1], C.byref(m), _p(coords),
                                                       coords.shape[0], _p(out), _stream()),
                   "isb_triplane_decode_points")
        return out
