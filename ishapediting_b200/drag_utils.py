"""Drag editing logic with the reference's surface (/root/reference/drag_utils.py).

Kept: get_args (:23-58), make_offsets (:134-138), resize_feat_align (:141-159), DragStuff with
update_latent_params (:252-280), get_mesh (:282-300), training (:302-399, a generator yielding
progress), latent_inversion (:552-566), clear_params / reset_params / set_offset1 and the attributes
the GUI reads (w, w0, feature_guidance, variance, variance_noise, train_flag, r1, offset1,
voxel_size).  Out of scope here: checkpoint discovery on disk (update_model_params :210-250 — weights
are loaded with load_state_dict), Open3D meshing and the mesh -> occupancy sampling inside train_triplane (:411-440).

What differs from the reference, by design:
  * the guided step never enters torch.autograd: UNet forward plan -> isb_drag_loss_grad -> UNet
    input-gradient plan -> isb_ddpm_step, all on one stream, optionally replayed as one CUDA graph;
  * only the input gradient is computed (the reference also computes and accumulates weight
    gradients for 406 tensors, :383, and never reads them);
  * the feature cache stays on the device (the reference moves 8.4 MB per step to and from the CPU,
    :276,352);
  * the mask index sets (:322-334) are built with vectorised integer ops instead of Python sets —
    same sets, checked bit-exactly in tests.
"""
from __future__ import annotations

import argparse
import os
import copy
from argparse import Namespace

import numpy as np
import torch as th
import torch.nn.functional as F

from .guided_diffusion.script_util import args_to_dict, create_model_and_diffusion, model_and_diffusion_defaults
from .triplane_decoder.axisnetworks import MultiTriplane
from .triplane_decoder.visualize import create_obj_o3d, query_volume


def get_args(argv=None):
    """Reference :23-58.  Unlike the reference this does not parse sys.argv at import time; pass argv
    explicitly (None -> defaults)."""
    parser = argparse.ArgumentParser(description="Generate a set of triplane and their corresponding meshes")
    parser.add_argument("--resolution", type=int, default=128)
    parser.add_argument("--num_steps", type=int, default=200)
    parser.add_argument("--shape_resolution", type=int, default=256)
    parser.add_argument("--w_time", type=int, default=170)
    parser.add_argument("--feat_layer", type=int, default=8)
    parser.add_argument("--loss_type", type=str, default="l2")
    parser.add_argument("--points_size", type=int, default=200000)
    parser.add_argument("--points_uniform_ratio", type=float, default=0.5)
    args = parser.parse_args([] if argv is None else argv)
    return Namespace(
        clip_denoised=True, num_samples=1, batch_size=1, use_ddim=False, model_path=None, stats_dir=None,
        num_steps=args.num_steps, explicit_normalization=True, save_dir=None, save_intermediate=False,
        save_timestep_interval=20, image_size=args.resolution, num_channels=256, num_res_blocks=2, num_heads=4,
        num_heads_upsample=-1, num_head_channels=64, attention_resolutions="32,16,8", channel_mult="", dropout=0.1,
        class_cond=False, shape_resolution=args.shape_resolution, use_checkpoint=False, use_scale_shift_norm=True,
        resblock_updown=True, use_fp16=True, use_new_attention_order=False, in_out_channels=96, learn_sigma=True,
        diffusion_steps=1000, noise_schedule="linear", timestep_respacing=str(args.num_steps), w_time=args.w_time,
        feat_layer=args.feat_layer, points_size=args.points_size, points_uniform_ratio=args.points_uniform_ratio,
        loss_type=args.loss_type, use_kl=False, predict_xstart=False, rescale_timesteps=False, decoder_ckpt=None,
        rescale_learned_sigmas=False)


def _copy_mesh(m):
    """The reference deep-copies its Open3D meshes (:279,564); device meshes / volumes are cloned."""
    return m.clone() if hasattr(m, "clone") else copy.deepcopy(m)


def make_offsets(r, device):
    p = th.arange(-r, r + 1, device=device)
    px, py, pz = th.meshgrid(p, p, p, indexing="ij")
    return th.stack([px.reshape(-1), py.reshape(-1), pz.reshape(-1)], dim=-1)


# ------------------------------------------------------------------------------------------------------
# resize_feat_align as an index map
# ------------------------------------------------------------------------------------------------------
def align_maps(channel_num):
    """Index maps equivalent to resize_feat_align(cat_var=True) (reference :141-159) for a feature with
    `channel_num` channels: returns (chan_map int32 [3*Ca], inv_map int32 [channel_num], Ca) where
    aligned[pl, a] = feature[chan_map[pl*Ca + a]] and inv_map is its inverse (-1 = dropped channel).
    The nearest-neighbour channel resampling is taken from F.interpolate itself, so the index rule is
    torch's, not a re-derivation."""
    assert channel_num % 2 == 0
    half = channel_num // 2
    expect = half - half % 3
    if half % 3:
        idx = th.arange(half, dtype=th.float32).reshape(1, 1, 1, half)
        src = F.interpolate(idx, (1, expect)).reshape(-1).long()      # nearest, as reference :149-151
    else:
        src = th.arange(half)
    per = expect // 3
    Ca = 2 * per
    chan_map = th.empty(3, Ca, dtype=th.int64)
    for pl in range(3):
        chan_map[pl, :per] = src[pl * per:(pl + 1) * per]              # mean half reshaped to (3, per, H, W)
        chan_map[pl, per:] = half + src[pl * per:(pl + 1) * per]       # var half
    inv = th.full((channel_num,), -1, dtype=th.int64)
    inv[chan_map.reshape(-1)] = th.arange(3 * Ca)
    return chan_map.reshape(-1).to(th.int32), inv.to(th.int32), Ca


def resize_feat_align(feature, cat_var=True):
    """Reference :141-159: (1, C, H, W) -> (3, Ca, H, W) fp32.  Gather kernel on the device."""
    assert cat_var, "cat_var=False is never used by the editor"
    b, c, h, w = feature.shape
    assert b == 1 and h == w
    from .ops import CudaOps
    ops = CudaOps(feature.device, "fp32")
    chan_map, _, Ca = align_maps(c)
    nhwc = ops.to_nhwc(feature.detach().to(th.float32).contiguous(), ops.empty((1, h, w, c)))
    out = ops.resize_feat_align(nhwc, chan_map.to(feature.device), ops.empty((3, h, w, Ca)))
    return out.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------------
# point tracking — OPT-IN EXTENSION, parity unpinned (the reference has no tracking function: SURVEY.md §0.3)
# ------------------------------------------------------------------------------------------------------
def handle_features(origin_feature, points):
    """f0 [B,3,Ca]: the aligned triplane feature of each 3-D point (bilinear, align_corners=True, zeros — the sampling
    of reference :355-358) taken from a cached origin feature (3,S,S,Ca channels-last, a feature_guidance entry)."""
    pts = points if th.is_tensor(points) else th.as_tensor(np.asarray(points))
    pts = pts.to(device=origin_feature.device, dtype=th.float32).reshape(-1, 3)
    planes = origin_feature.permute(0, 3, 1, 2)                                   # (3, Ca, S, S) view
    axes = ((0, 1), (1, 2), (0, 2))
    grid = th.stack([pts[:, list(a)] for a in axes], dim=0).unsqueeze(2)          # (3, B, 1, 2): (u -> W, v -> H)
    f = F.grid_sample(planes, grid, mode="bilinear", padding_mode="zeros", align_corners=True)   # (3, Ca, B, 1)
    return f[..., 0].permute(2, 0, 1).contiguous()


def track_points(feature_nhwc, f0, points, r, voxel_size, ops=None):
    """Nearest-feature search with a warp-level argmin (isb_track_points): for every point, the lattice point within
    r voxels whose aligned triplane feature in `feature_nhwc` ([1,S,S,Cf] raw intermediate feature, NHWC fp32) is
    nearest in L1 to f0.  Returns (new_points [B,3], lattice index int32 [B], distance [B]).
    This is an extension named by BASELINE.json's north_star; the reference keeps handles and targets fixed for a
    whole edit (:305-321) and has nothing to compare against, so parity is unpinned."""
    if ops is None:
        from .ops import CudaOps
        ops = CudaOps(feature_nhwc.device, "fp32")
    chan_map, _, Ca = align_maps(feature_nhwc.shape[3])
    center = th.as_tensor(np.asarray(points) if not th.is_tensor(points) else points, dtype=th.float32,
                          device=ops.device).reshape(-1, 3).contiguous()
    idx, dist, pts, _ = ops.track_points(feature_nhwc.contiguous(), chan_map.to(ops.device), f0.contiguous(), center,
                                         r, voxel_size)
    return pts, idx, dist


# ------------------------------------------------------------------------------------------------------
# geometry of one edit (host side, once per training() call)
# ------------------------------------------------------------------------------------------------------
class DragGeometry:
    """Everything training() derives from (sources, targets) before its loop (reference :305-334).

    The (2r+1)^3 patch lattice around a handle projects, on each of the three planes, onto a
    (2r+1)^2 lattice of distinct 2-D points, each hit (2r+1) times with identical coordinates (the
    projected-out offset does not enter).  Only the distinct points are kept, with that multiplicity.
    """

    def __init__(self, sources, targets, r, voxel_size, S, Ca):
        src = th.as_tensor(np.asarray(sources), dtype=th.float32).reshape(-1, 3)
        tgt = th.as_tensor(np.asarray(targets), dtype=th.float32).reshape(-1, 3)
        assert src.shape == tgt.shape
        B = src.shape[0]
        side = 2 * r + 1
        off1d = voxel_size * th.arange(-r, r + 1)            # int64 * python float -> fp32, as :316-317
        axes = ((0, 1), (1, 2), (0, 2))                      # plane grids: xy, yz, xz (:318-321)
        self.group_size = side * side
        self.npts = B * self.group_size

        def plane_pts(p):
            out = th.empty(3, B, side, side, 2)
            for pl, (au, av) in enumerate(axes):
                u = p[:, au, None] + off1d[None, :]          # [B, side] -> W coordinate
                v = p[:, av, None] + off1d[None, :]          # [B, side] -> H coordinate
                out[pl, ..., 0] = u[:, :, None]
                out[pl, ..., 1] = v[:, None, :]
            return out.reshape(3, self.npts, 2).contiguous()

        self.patch_xy = plane_pts(src)
        self.shift_xy = plane_pts(tgt)
        self.weight = th.full((self.npts,), float(side))
        self.inv_count = 1.0 / (3.0 * Ca * B * side ** 3)
        # conservative bounding boxes (floor index of the shift samples, padded by one pixel)
        ix = th.floor(((self.shift_xy + 1.0) / 2.0) * (S - 1)).to(th.int64).reshape(3, B, self.group_size, 2)
        self.bbox = th.stack([ix[..., 0].amin(-1) - 1, ix[..., 0].amax(-1) + 1,
                              ix[..., 1].amin(-1) - 1, ix[..., 1].amax(-1) + 1], dim=-1).to(th.int32).contiguous()
        # mask regulariser: complement of the rounded content pixels (:322-334), [plane, row, col]
        content = th.zeros(3, S, S, dtype=th.bool)
        for pts in (self.patch_xy, self.shift_xy):
            rc = th.round((pts + 1) * (S - 1) / 2).to(th.int16).long()   # [..., 0] = col (W), [..., 1] = row (H)
            for pl in range(3):
                col, row = rc[pl, :, 0], rc[pl, :, 1]
                ok = (col >= 0) & (col < S) & (row >= 0) & (row < S)
                content[pl, row[ok], col[ok]] = True
        self.mask = (~content).to(th.uint8).contiguous()
        self.mask_count = int(self.mask.sum())

    def mask_index_sets(self):
        """The three (K,2) [row, col] index lists the reference builds (order-insensitive)."""
        return [th.nonzero(self.mask[pl].bool()) for pl in range(3)]

    def to(self, device):
        g = copy.copy(self)
        for k in ("patch_xy", "shift_xy", "weight", "bbox", "mask"):
            setattr(g, k, getattr(self, k).to(device))
        return g


# ------------------------------------------------------------------------------------------------------
# the guided step (fast path)
# ------------------------------------------------------------------------------------------------------
class GuidedStepper:
    """One drag-guided DDPM step = UNet forward + drag loss/grad + UNet dgrad + fused update, issued
    as C-ABI calls on the current stream (reference loop body :340-392).  With `use_graph` the whole
    sequence is captured once and replayed; per-step inputs (step scalars, timestep, origin feature,
    noise) are refreshed in static device buffers before each replay."""

    def __init__(self, model, diffusion, geometry, feat_layer, cof, loss_type, scale,
                 clip_denoised=True, use_graph=True, overlap_tail=True):
        """`geometry`: one DragGeometry (the reference's batch-1 edit) or a list of B geometries with the same
        number of handles — B independent edits advanced together as one batch-B UNet pass (the "batched"
        variant of SURVEY.md §8d config 5; the reference itself refuses num_samples > 1, drag_utils.py:303)."""
        self.model, self.diffusion = model, diffusion
        geos = list(geometry) if isinstance(geometry, (list, tuple)) else [geometry]
        B = self.batch = len(geos)
        assert all(g.npts == geos[0].npts and g.group_size == geos[0].group_size for g in geos)
        self.plan = model.plan(B, model.image_size, model.image_size, want_backward=True)
        self.weights_generation = model.weights_generation      # the plan's weight panels are packed copies
        ops = self.ops = self.plan.ops
        dev = ops.device
        self.feat_layer, self.cof, self.scale, self.clip = feat_layer, float(cof), float(scale), clip_denoised
        self.loss_type = 1 if loss_type == "l1" else 0
        inter = self.plan.block_out[feat_layer]
        _, S, _, Cf = inter.val.shape
        chan_map, inv_map, Ca = align_maps(Cf)
        self.chan_map, self.inv_map, self.S, self.Cf, self.Ca = chan_map.to(dev), inv_map.to(dev), S, Cf, Ca
        self.geos = [g.to(dev) for g in geos]
        self.dyns = [th.tensor([g.inv_count, 1.0 / (max(g.mask_count, 1) * Ca)], dtype=th.float32, device=dev)
                     for g in geos]
        self.geo, self.dyn = self.geos[0], self.dyns[0]
        C, R = model.in_channels, model.image_size
        self.img = ops.empty((B, C, R, R))
        self.img_next = ops.empty((B, C, R, R))
        self.noise = ops.empty((B, C, R, R))
        self.grad = ops.empty((B, C, R, R))
        self.variance = ops.empty((B, C, R, R))
        self.sample = ops.empty((B, C, R, R))
        self.origin = ops.empty((B, 3, S, S, Ca)) if B > 1 else ops.empty((3, S, S, Ca))
        self.coef = ops.empty((8,))
        self.coef_table = diffusion.coef_table(dev, guide_scale=self.scale)
        tmap = getattr(diffusion, "timestep_map", list(range(diffusion.num_timesteps)))
        self.t_table = th.tensor(tmap, device=dev, dtype=th.int64)
        npts = self.geo.npts
        self.g = ops.empty((B, 3, npts, Ca))
        self.pt_info = ops.empty((B, 3, npts, 4))
        self.partial = ops.zeros((B, ops.drag_partial_len(S, Cf, npts)), th.float64)
        self.loss = ops.zeros((B,))
        self.plan.ensure_grad(inter)
        self.use_graph = use_graph and dev.type == "cuda"
        self.overlap_tail = overlap_tail
        if dev.type == "cuda":
            # The forward tail is off the critical path: it gets the LOW priority (0), the captured main stream and
            # the backward side stream the HIGH one, so that whenever an SM frees a slot the block scheduler hands it
            # to the backward chain first (ISB_STREAM_PRIO=0 puts everything back on equal footing).
            hi = -1 if os.environ.get("ISB_STREAM_PRIO", "1") != "0" else 0
            self._side = th.cuda.Stream(device=dev, priority=0)
            self._cap_stream = th.cuda.Stream(device=dev, priority=hi)
            self._ev_fork, self._ev_join = th.cuda.Event(), th.cuda.Event()
            self._side_bwd = (th.cuda.Stream(device=dev, priority=hi), th.cuda.Event(), th.cuda.Event())
        self._film = {}
        # background L2 prefetch of the next layers' weight panels inside the captured step (ops.weight_prefetch).
        # MEASURED AND OFF: 4.90 ms with it, 4.77 ms without (tools/ab_step.py on B200).  The conv launches take the same
        # time with L2-warm and HBM-cold weights (profiles/r01_conv_auto_{warm,cold}_weights.txt: 15.4 vs 15.7 us for the
        # 8x8 1024->1024 layer): PDL already streams the first weight tiles under the predecessor's tail, and the
        # mainloop is bound by the producer thread's per-iteration latency, not by HBM.  The ~200 extra graph nodes and
        # cross-stream edges only add launch overhead.  Kept behind ISB_WEIGHT_PREFETCH=1 with its test.
        self._prefetch = os.environ.get("ISB_WEIGHT_PREFETCH", "0") == "1" and dev.type == "cuda"
        self._pf_window = int(float(os.environ.get("ISB_PREFETCH_WINDOW_MB", "48")) * (1 << 20))
        self._pf_seq = None
        self._pf_stream = th.cuda.Stream(device=dev, priority=0) if dev.type == "cuda" else None
        self._film_cache = os.environ.get("ISB_FILM_CACHE", "1") != "0"
        self._bwd_branches = os.environ.get("ISB_SIDE_BWD", "1") != "0"
        self._graph = None
        self._warm = 0

    def compatible(self, geometry, cof, loss_type):
        """True if a new edit (other handles/targets, other scale) can reuse this stepper and its captured graph:
        the model's weights are still the ones this stepper's plan packed (load_state_dict / convert_to_fp16 / .to()
        on the same model object make it stale — the reference swaps checkpoints that way, drag_utils.py:229-232),
        same number of sample points and the by-value kernel scalars (cof, loss type) unchanged."""
        return (self.weights_generation == self.model.weights_generation and geometry.npts == self.geo.npts and geometry.group_size == self.geo.group_size
                and float(cof) == self.cof and (1 if loss_type == "l1" else 0) == self.loss_type)

    def retarget(self, geometry, scale, b=0):
        """Load a new edit's geometry into the static device buffers of batch slot b (no re-capture)."""
        g, geo = geometry, self.geos[b]
        for k in ("patch_xy", "shift_xy", "weight", "bbox", "mask"):
            getattr(geo, k).copy_(getattr(g, k))
        geo.mask_count, geo.inv_count = g.mask_count, g.inv_count
        self.dyns[b].copy_(th.tensor([g.inv_count, 1.0 / (max(g.mask_count, 1) * self.Ca)], dtype=th.float32))
        if float(scale) != self.scale:
            self.scale = float(scale)
            self.coef_table = self.diffusion.coef_table(self.ops.device, guide_scale=self.scale)

    def _body(self):
        plan, ops, geo = self.plan, self.ops, self.geo
        overlap = self.overlap_tail and ops.device.type == "cuda"
        plan.film_external = self._film_cache   # FiLM rows of this timestep were loaded into plan.film_all by step()
        plan.side = self._side_bwd if (overlap and self._bwd_branches) else None
        try:
            self._body_inner(plan, ops, overlap)
        finally:
            plan.film_external = False
            plan.side = None

    def _body_inner(self, plan, ops, overlap):
        inter = plan.forward(self.img, plan.t_dev, self.feat_layer, upto_feat_only=overlap)
        if overlap:
            # The layers behind the intermediate feature (output_blocks[9..14] + out: 292 of the 635 forward
            # GFLOP, all at 64^2/128^2) only feed the DDPM update; the backward pass that starts here is a chain
            # of small, latency-bound kernels.  Run the two concurrently and join before the update.
            main = th.cuda.current_stream()
            self._ev_fork.record(main)
            bg = os.environ.get("ISB_TAIL_BG", "0") == "1"    # measured neutral on B200 (profiles/README.md): off
            with th.cuda.stream(self._side), ops.workspace_slot(1), ops.background(bg):
                self._side.wait_event(self._ev_fork)
                plan.forward_tail()
                self._ev_join.record(self._side)
        for b in range(self.batch):           # the drag loss couples nothing across edits: one small launch set each
            geo = self.geos[b]
            origin = self.origin[b] if self.batch > 1 else self.origin
            ops.drag_loss_grad(inter.val[b:b + 1], origin, self.chan_map, self.inv_map, geo.patch_xy, geo.shift_xy,
                               geo.weight, geo.group_size, geo.bbox, geo.mask, geo.mask_count, geo.inv_count,
                               self.cof, self.loss_type, self.g[b], self.pt_info[b], self.partial[b],
                               self.loss[b:b + 1], inter.grad[b:b + 1], dyn=self.dyns[b])
        plan.begin_backward()
        plan.seed_grad(inter)
        plan.backward(self.grad)
        if overlap:
            th.cuda.current_stream().wait_event(self._ev_join)
        ops.ddpm_step(self.img, plan.out_nhwc, self.coef, self.clip, noise=self.noise, grad=self.grad,
                      x_next=self.img_next, sample=self.sample, var=self.variance, model_out_nhwc=True)
        self.img.copy_(self.img_next)

    def step(self, i, origin_feature, noise=None):
        """Advance self.img from respaced step i to i-1.  origin_feature: (3,S,S,Ca) device tensor."""
        self.coef.copy_(self.coef_table[i])
        self.plan.t_dev.copy_(self.t_table[i:i + 1].expand(self.batch))
        film = self._film.get(i)           # the timestep-embedding path depends only on t: computed once per step index
        if not self._film_cache:
            pass
        elif film is None:
            self.plan.compute_film()
            self._film[i] = self.plan.film_all.clone()
        else:
            self.plan.film_all.copy_(film)
        self.origin.copy_(origin_feature)
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        if not self.use_graph:
            self._body()
            return
        if self._graph is None:
            if self._warm < 1:          # eager warm-up sizes every workspace/scratch buffer
                self._warm += 1
                if self._prefetch:      # ... and records the order in which the weight panels are touched
                    with self.ops.record_weight_sequence() as seq:
                        self._body()
                    self._pf_seq = list(seq)
                else:
                    self._body()
                return
            img_keep = self.img.clone()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g, stream=self._cap_stream):
                if self._prefetch and self._pf_seq:
                    with self.ops.weight_prefetch(self._pf_seq, self._pf_stream, self._pf_window):
                        self._body()
                else:
                    self._body()
            self._graph = g
            self.img.copy_(img_keep)    # capture does not execute; restore and replay for real
        self._graph.replay()


class HostStepPipeline:
    """Drives a GuidedStepper from HOST buffers (the shape of the reference's loop, drag_utils.py:341-385, where the
    cached origin feature of every step lives in host memory and the latent is read back): the origin feature and
    noise of step k+1 travel host->device on a copy stream while step k computes, and the new latent and loss of step
    k travel device->host while step k+1 computes.  Every step still pays its copies; they just no longer sit on the
    critical path.  Results are read one step late: result(k) blocks until step k's read-back has landed.

        pipe.prefetch(0, origin_host, noise_host)
        for k in range(n):
            if k + 1 < n: pipe.prefetch(k + 1, next_origin_host, next_noise_host)
            pipe.run(k, i_k)
            if k: img, loss = pipe.result(k - 1)
        img, loss = pipe.result(n - 1)
    """

    def __init__(self, stepper: "GuidedStepper"):
        st = stepper
        dev = st.img.device
        assert dev.type == "cuda", "HostStepPipeline needs a CUDA stepper"
        self.st = st
        self.copy_in, self.copy_out = th.cuda.Stream(device=dev), th.cuda.Stream(device=dev)
        two = range(2)
        self.origin_dev = [th.empty_like(st.origin) for _ in two]
        self.noise_dev = [th.empty_like(st.noise) for _ in two]
        self.img_stage = [th.empty_like(st.img) for _ in two]
        self.loss_stage = [th.empty_like(st.loss) for _ in two]
        self.img_host = [th.empty(st.img.shape, dtype=st.img.dtype).pin_memory() for _ in two]
        self.loss_host = [th.empty(st.loss.shape, dtype=st.loss.dtype).pin_memory() for _ in two]
        ev = lambda: [th.cuda.Event() for _ in two]
        self.in_ready, self.in_free, self.snap, self.out_done = ev(), ev(), ev(), ev()
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.origin_dev[0], self.noise_dev[0]))
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in (self.img_stage[0], self.loss_stage[0]))

    def prefetch(self, k, origin_host, noise_host):
        """Start the host->device copy of step k's inputs (pinned host tensors)."""
        s = k & 1
        with th.cuda.stream(self.copy_in):
            self.copy_in.wait_event(self.in_free[s])      # step k-2 has consumed this slot
            self.origin_dev[s].copy_(origin_host, non_blocking=True)
            self.noise_dev[s].copy_(noise_host, non_blocking=True)
            self.in_ready[s].record(self.copy_in)

    def run(self, k, i):
        """Launch step k (respaced index i) on the current stream and start the read-back of its result."""
        s = k & 1
        main = th.cuda.current_stream()
        main.wait_event(self.in_ready[s])
        self.st.step(i, self.origin_dev[s], self.noise_dev[s])
        self.in_free[s].record(main)
        main.wait_event(self.out_done[s])                 # the read-back of step k-2 has left this staging slot
        self.img_stage[s].copy_(self.st.img)              # snapshot: step k+1 updates st.img in place
        self.loss_stage[s].copy_(self.st.loss)
        self.snap[s].record(main)
        with th.cuda.stream(self.copy_out):
            self.copy_out.wait_event(self.snap[s])
            self.img_host[s].copy_(self.img_stage[s], non_blocking=True)
            self.loss_host[s].copy_(self.loss_stage[s], non_blocking=True)
            self.out_done[s].record(self.copy_out)

    def result(self, k):
        """(latent, loss) of step k in pinned host memory; valid until step k+2 is run."""
        s = k & 1
        self.out_done[s].synchronize()
        return self.img_host[s], self.loss_host[s]


class ReconStepper:
    """The loop body of train_triplane (drag_utils.py:445-463) without torch.autograd, as one static launch sequence
    (optionally a CUDA graph): UNet forward (whole network) -> fused posterior (pred_xstart, sample, variance) ->
    triplane decoder at the sample points -> BCE gradient -> decoder backward into the planes -> clamp / eps chain rule
    -> UNet input-gradient backward from the OUTPUT layer -> img <- sample + variance * scale * grad.
    Same arithmetic as recon_guided_step (which goes through the autograd bridge and launches eagerly)."""

    def __init__(self, model, diffusion, decoder, n_points, scale=600.0, rng=1.0, middle=0.0, clip_denoised=True,
                 use_graph=True):
        self.model, self.diffusion, self.decoder = model, diffusion, decoder
        C, R = model.in_channels, model.image_size
        assert C == 96, "the triplane decoder reads 3 planes x 32 features"
        self.plan = model.plan(1, R, R, want_backward=True)
        ops = self.ops = self.plan.ops
        dev = ops.device
        self.scale, self.clip, self.P, self.R = float(scale), clip_denoised, n_points, R
        # range / middle: scalars, or the per-channel (1,96,1,1) statistics of explicit_normalization
        # (drag_utils.py:236-245); decided once, applied at (1,96,R,R) BEFORE the (3,32,R,R) reshape like get_mesh
        self.identity_norm = (not th.is_tensor(rng) and not th.is_tensor(middle) and float(rng) == 1.0
                              and float(middle) == 0.0)
        as_dev = lambda v: v.to(device=dev, dtype=th.float32) if th.is_tensor(v) else float(v)  # noqa: E731
        self.rng, self.middle = as_dev(rng), as_dev(middle)
        self.weights = decoder.mlp_weights()
        e = ops.empty
        self.img, self.noise, self.grad = e((1, C, R, R)), e((1, C, R, R)), e((1, C, R, R))
        self.sample, self.variance, self.x0 = e((1, C, R, R)), e((1, C, R, R)), e((1, C, R, R))
        self.g_out = ops.zeros((1, 2 * C, R, R))             # d loss / d model_output: eps half only
        self.coords, self.gt = e((n_points, 3)), e((n_points, 1))
        self.logits, self.d_logits = e((n_points,)), e((n_points,))
        self.planes_hwc, self.d_planes_hwc = e((3, R, R, 32)), e((3, R, R, 32))
        self.g_x0 = e((3, 32, R, R))
        self.loss = ops.zeros((1,))
        self.coef = e((8,))
        self.coef_table = diffusion.coef_table(dev, guide_scale=self.scale)
        tmap = getattr(diffusion, "timestep_map", list(range(diffusion.num_timesteps)))
        self.t_table = th.tensor(tmap, device=dev, dtype=th.int64)
        self.use_graph = use_graph and dev.type == "cuda"
        self._graph, self._warm = None, 0
        self._cap_stream = th.cuda.Stream(device=dev) if dev.type == "cuda" else None

    def _body(self):
        plan, ops = self.plan, self.ops
        plan.forward(self.img, plan.t_dev, -1)
        ops.ddpm_step(self.img, plan.out_nhwc, self.coef, self.clip, noise=self.noise, sample=self.sample,
                      var=self.variance, x0=self.x0, model_out_nhwc=True)
        planes = self.x0 if self.identity_norm else self.x0 * self.rng + self.middle      # (1,96,R,R)
        ops.to_nhwc(planes.reshape(3, 32, self.R, self.R).contiguous(), self.planes_hwc)
        ops.decode_points(self.planes_hwc, self.weights, self.coords, self.logits)
        # loss = -BCEWithLogits(mean):  d loss / d logit = -(sigmoid(logit) - gt) / P
        th.sigmoid(self.logits, out=self.d_logits)
        self.d_logits.sub_(self.gt.reshape(-1)).mul_(-1.0 / self.P)
        self.loss.copy_(-F.binary_cross_entropy_with_logits(self.logits, self.gt.reshape(-1)).reshape(1))
        self.d_planes_hwc.zero_()
        ops.decode_points_backward(self.planes_hwc, self.weights, self.coords, self.d_logits, self.d_planes_hwc)
        ops.to_nchw(self.d_planes_hwc, self.g_x0)
        g_x0 = self.g_x0.reshape(1, 96, self.R, self.R)
        if not self.identity_norm:
            g_x0 = g_x0 * self.rng
        if self.clip:                                        # x0 = clamp(c0 x - c1 eps): no gradient where it saturates
            g_x0 = g_x0 * (self.x0.abs() < 1.0)
        self.g_out[:, :96].copy_(g_x0 * (-self.coef[1]))     # d x0 / d eps = -sqrt(1/abar - 1)
        plan.begin_backward()
        plan.backward_out_layer(self.g_out)
        plan.backward(self.grad)
        self.grad.add_(g_x0 * self.coef[0])                  # d x0 / d x = sqrt(1/abar)
        ops.ddpm_step(self.img, plan.out_nhwc, self.coef, self.clip, noise=self.noise, grad=self.grad, x_next=self.sample,
                      model_out_nhwc=True)
        self.img.copy_(self.sample)

    def step(self, i, coords, gt, noise=None):
        """Advance self.img from respaced step i to i-1 under the occupancy samples (coords (P,3), gt (P,1))."""
        self.coef.copy_(self.coef_table[i])
        self.plan.t_dev.copy_(self.t_table[i:i + 1])
        self.coords.copy_(coords)
        self.gt.copy_(gt.reshape(self.P, 1))
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise)
        if not self.use_graph:
            self._body()
            return
        if self._graph is None:
            if self._warm < 1:
                self._warm += 1
                self._body()
                return
            keep = self.img.clone()
            g = th.cuda.CUDAGraph()
            with th.cuda.graph(g, stream=self._cap_stream):
                self._body()
            self._graph = g
            self.img.copy_(keep)
        self._graph.replay()


def recon_guided_step(model, diffusion, decoder, img, i, coords, gt, scale=600.0, noise=None, rng=1.0, middle=0.0):
    """One iteration of the reference's real-shape reconstruction guidance (drag_utils.py:445-463, SURVEY.md §8f
    rank 2): classifier guidance on the predicted x_start through the triplane decoder.  img (1,96,R,R) latent at
    respaced step i; coords (P,3) in [-1,1]; gt (P,1) occupancies in {0,1}.  Returns (next latent, loss).
    The UNet forward and its FULL input-gradient backward (out layer and output blocks 9-14 included) run on the
    kernel plan behind the autograd bridge; the decoder forward / backward are isb_triplane_decode_points(_backward);
    the posterior algebra in between is elementwise torch (gaussian_diffusion.p_mean_variance, grad route)."""
    dev = img.device
    img = img.detach().clone().requires_grad_(True)
    outs = diffusion.p_sample_guidance(model, img, th.tensor([i], device=dev), noise=noise)
    R = img.shape[-1]
    predict_x0 = (outs["pred_xstart"] * rng + middle).reshape(3, 32, R, R)
    for j in range(3):
        decoder.embeddings[j] = predict_x0[[j]]
    prediction = decoder(0, coords.to(dev).unsqueeze(0)).squeeze(0)
    gt = gt.to(dev)
    assert gt.shape == prediction.shape
    loss = -th.nn.BCEWithLogitsLoss()(prediction, gt)
    loss.backward()
    with th.no_grad():
        nxt = (outs["sample"] + outs["variance"] * (scale * img.grad)).clone().detach()
    for j in range(3):          # do not keep the autograd graph alive through the decoder's plane list
        decoder.embeddings[j] = decoder.embeddings[j].detach()
    return nxt, loss.detach()


# ------------------------------------------------------------------------------------------------------
# DragStuff
# ------------------------------------------------------------------------------------------------------
class DragStuff:
    args = None   # set lazily: the reference parses sys.argv at class-definition time (:176)

    def __init__(self, args=None, device=None, use_graph=True):
        self.args = args if args is not None else (DragStuff.args or get_args())
        self.device = th.device(device) if device is not None else th.device("cuda", th.cuda.current_device())
        self.model, self.diffusion = create_model_and_diffusion(
            **args_to_dict(self.args, model_and_diffusion_defaults().keys()))
        self.model.to(self.device)
        self.model.eval()
        self.decoder = MultiTriplane(1, input_dim=3, output_dim=1).to(self.device)
        self.decoder.eval()
        self.range = 1.
        self.middle = 0.
        self.latent_code = None
        self.w0 = None
        self.w = None
        self.r1 = 12
        self.offset1 = make_offsets(self.r1, self.device)
        self.voxel_size = 2. / self.args.shape_resolution
        self.train_flag = True
        self.targets = None
        self.sources = None
        self.mesh = None
        self.mesh0 = None
        self.noise = []
        self.variance = []
        self.variance_noise = []
        self.feature_guidance = []       # device tensors, channels-last (3,S,S,Ca); see feature_guidance_nchw()
        self.use_graph = use_graph
        self.last_volume = None
        # True (default): get_mesh returns the smoothed triangle mesh like the reference (marching cubes + 10 Laplacian
        # iterations, on the device); False: it stops at the logit volume (benchmark legs that time the decode alone)
        self.mesh_on_device = True

    def set_offset1(self, r1):
        self.r1 = r1
        self.offset1 = make_offsets(r1, self.device)

    def feature_guidance_nchw(self, k):
        """k-th cached feature in the reference's (3, Ca, S, S) layout."""
        return self.feature_guidance[k].permute(0, 3, 1, 2)

    # ---- no-grad trajectory + feature cache (reference :252-280) --------------------------------
    def _nograd_step(self, plan, img, i, feat_layer, noise=None):
        """One unguided step (reference p_sample_guidance under no_grad, :267-271 / :289-293); returns
        (img_next, inter _T).  On the GPU the step (UNet forward + fused posterior/sample) is captured once into
        a CUDA graph per plan and replayed; step scalars, timestep, latent and noise live in static buffers."""
        ops = plan.ops
        dev = self.device
        st = getattr(plan, "_nograd_state", None)
        if st is None:
            st = plan._nograd_state = {
                "img": ops.empty(tuple(img.shape)), "next": ops.empty(tuple(img.shape)),
                "noise": ops.empty(tuple(img.shape)), "coef": ops.empty((8,)), "graph": None, "warm": 0,
                "feat_layer": feat_layer}
        tmap = self.diffusion.timestep_map_tensor(dev)
        st["img"].copy_(img)
        st["coef"].copy_(self.diffusion.coef_table(dev)[i])
        plan.t_dev.copy_(tmap[i:i + 1].expand(plan.N))
        if noise is None:
            st["noise"].normal_()
        else:
            st["noise"].copy_(noise)

        def body():
            plan.forward(st["img"], plan.t_dev, feat_layer)
            ops.ddpm_step(st["img"], plan.out_nhwc, st["coef"], self.args.clip_denoised, noise=st["noise"],
                          x_next=st["next"], model_out_nhwc=True)

        use_graph = self.use_graph and dev.type == "cuda" and st["feat_layer"] == feat_layer
        if not use_graph:
            body()
        elif st["graph"] is None:
            if st["warm"] < 1:
                st["warm"] += 1
                body()
            else:
                g = th.cuda.CUDAGraph()
                with th.cuda.graph(g):
                    body()
                st["graph"] = g
                g.replay()
        else:
            st["graph"].replay()
        inter = plan.block_out[feat_layer] if feat_layer >= 0 else None
        return st["next"].clone(), inter

    def update_latent_params(self, img=None, **kwargs):
        dev = self.device
        R = self.args.image_size
        if img is None:
            img = th.randn((1, 96, R, R), device=dev)
        elif th.is_tensor(img):
            img = img.to(device=dev, dtype=th.float32)
        elif isinstance(img, np.ndarray):
            img = th.tensor(img, dtype=th.float32, device=dev)
        else:
            raise NotImplementedError("Unknown data type!")
        self.latent_code = img.clone().detach()
        plan = self.model.plan(1, R, R, want_backward=True)
        chan_map = None
        self.feature_guidance = []
        with th.no_grad():
            for i in range(self.args.num_steps - 1, -1, -1):
                img, inter = self._nograd_step(plan, img.contiguous(), i, self.args.feat_layer, kwargs.get("noise"))
                if i == self.args.w_time:
                    self.w = img.clone().detach()
                    self.w0 = self.w.clone().detach()
                if i < self.args.w_time:
                    if chan_map is None:
                        cm, _, Ca = align_maps(inter.val.shape[3])
                        chan_map = cm.to(dev)
                    S = inter.val.shape[1]
                    self.feature_guidance.append(plan.ops.resize_feat_align(inter.val, chan_map,
                                                                            plan.ops.empty((3, S, S, Ca))))
            assert len(self.feature_guidance) == self.args.w_time
            self.mesh0 = self.get_mesh(tri_feat=img)
            self.mesh = _copy_mesh(self.mesh0)
            return img

    # ---- decode (reference :282-300) -----------------------------------------------------------------
    def get_mesh(self, tri_feat=None, img=None, t=0):
        R = self.args.image_size
        with th.no_grad():
            if tri_feat is None:
                img = img if img is not None else th.randn((1, 96, R, R), device=self.device)
                plan = self.model.plan(1, R, R, want_backward=True)
                for i in range(t - 1, -1, -1):
                    img, _ = self._nograd_step(plan, img.contiguous(), i, self.args.feat_layer)
                tri_feat = img
            tri_feat = (tri_feat * self.range + self.middle).reshape(3, 32, R, R)
            for i in range(3):
                self.decoder.embeddings[i] = tri_feat[[i]]
            res = self.args.shape_resolution
            self.last_volume = query_volume(self.decoder, 0, res=res)           # decoded ONCE; meshed from this volume
            if not self.mesh_on_device:
                return self.last_volume
            mesh = create_obj_o3d(self.decoder, 0, res=res, volume=self.last_volume)
            return mesh.filter_smooth_simple(number_of_iterations=10)           # reference :300

    # ---- real-shape reconstruction guidance (reference :400-471) ---------------------------
    def recon_guided_step(self, img, i, coords, gt, scale=600.0, noise=None):
        """Loop body of train_triplane (:445-463), see the module-level recon_guided_step."""
        return recon_guided_step(self.model, self.diffusion, self.decoder, img, i, coords, gt, scale=scale, noise=noise,
                                 rng=self.range, middle=self.middle)

    def train_triplane(self, points=None, occupancies=None, tri_feat=None, scale=600, batch_size=40000, seed=None):
        """Reference :400-471.  The mesh -> (points, occupancies) sampling of :411-440 is Open3D / CPU code outside
        the path: pass the samples directly (`points` (M,3) float32 in [-1,1], `occupancies` (M,) or (M,1) in {0,1}),
        or a ready `tri_feat` latent as the reference's `tri_feat_path` branch does.  Returns the reconstructed
        latent; sets self.mesh / self.mesh0 like the reference."""
        if tri_feat is not None:
            img = th.as_tensor(tri_feat, device=self.device, dtype=th.float32)
        else:
            pts = th.as_tensor(np.asarray(points), dtype=th.float32)
            occ = th.as_tensor(np.asarray(occupancies), dtype=th.float32).reshape(-1, 1)
            g = th.Generator().manual_seed(seed) if seed is not None else None
            R = self.args.image_size
            n = min(batch_size, pts.shape[0])
            st = ReconStepper(self.model, self.diffusion, self.decoder, n, scale=scale, rng=self.range,
                              middle=self.middle, use_graph=self.use_graph)
            st.img.copy_(th.randn((1, 96, R, R), generator=g).to(self.device))
            pts_d, occ_d = pts.to(self.device), occ.to(self.device)
            for i in range(self.args.num_steps - 1, -1, -1):
                idx = th.randperm(pts.shape[0], generator=g)[:n].to(self.device)   # DataLoader(shuffle=True) batch, :442,454
                noise = th.randn((1, 96, R, R), generator=g).to(self.device)
                st.step(i, pts_d[idx], occ_d[idx], noise=noise)
            img = st.img.clone()
        self.clear_params()
        self.mesh = self.get_mesh(tri_feat=img)
        self.mesh0 = _copy_mesh(self.mesh)
        return img

    # ---- the guided edit (reference :302-399) -----------------------------------------------------------
    def training(self, sources=None, targets=None, scale=600, cof=0.2, noises=None, track=False, track_radius=None):
        """Reference :302-399.  `track=True` (opt-in extension, off by default; parity unpinned) additionally follows
        the handles: after every step the point of the current feature nearest to each handle's ORIGINAL feature is
        searched around its last tracked position (track_points); the positions are kept in self.tracked and the
        L-inf distance to the targets, in voxels, in self.track_error — nothing of the guided update changes."""
        if self.args.num_samples > 1:
            raise NotImplementedError("We can handle only one shape at each time!")
        self.sources = th.tensor(np.asarray(sources), device=self.device, dtype=th.float32)
        self.targets = th.tensor(np.asarray(targets), device=self.device, dtype=th.float32)
        assert self.sources.shape[0] == self.targets.shape[0]
        S, Ca = self.feature_guidance[0].shape[1], self.feature_guidance[0].shape[3]
        geo = DragGeometry(np.asarray(sources), np.asarray(targets), self.r1, self.voxel_size, S, Ca)
        stepper = getattr(self, "stepper", None)
        if stepper is not None and stepper.model is self.model and stepper.compatible(geo, cof, self.args.loss_type):
            stepper.retarget(geo, scale)          # same captured graph, new handles
        else:
            stepper = GuidedStepper(self.model, self.diffusion, geo, self.args.feat_layer, cof, self.args.loss_type,
                                    scale, clip_denoised=True, use_graph=self.use_graph)
        stepper.img.copy_(self.w.detach())
        self.stepper = stepper
        stop_time = 0
        self.train_flag = True
        w_time = self.args.w_time
        self.tracked, self.track_error = None, None
        if track:
            f0 = handle_features(self.feature_guidance[0], self.sources)
            self.tracked = self.sources.clone()
            inter = stepper.plan.block_out[self.args.feat_layer]
        for i in range(w_time - 1, -1, -1):
            if not self.train_flag:
                stop_time = i + 1
                break
            stepper.step(i, self.feature_guidance[w_time - 1 - i], None if noises is None else noises[w_time - 1 - i])
            if track:
                self.tracked, _, _ = track_points(inter.val, f0, self.tracked, track_radius or self.r1, self.voxel_size,
                                                  ops=stepper.ops)
                self.track_error = (self.tracked - self.targets).abs().amax(dim=1) / self.voxel_size
            yield 1 - i / (w_time - 1.) if w_time > 1 else 1.0
        self.mesh = self.get_mesh(img=stepper.img.clone(), t=stop_time)

    def training_batch(self, edits, scale=600, cof=0.2, noises=None, decode=True):
        """Throughput mode (BASELINE configs[4], "batched" variant of SURVEY.md §8d config 5): B independent drags of
        the CURRENT shape — `edits` = [(sources, targets), ...] with the same number of handles — advanced together as
        one batch-B guided step per timestep (the reference handles one edit at a time, :303-304; each edit here is
        exactly its own `training()` run: same start latent self.w, same cached origin features, its own handles).
        Returns (latents (B,96,R,R), [volume or mesh per edit] if `decode`)."""
        S, Ca = self.feature_guidance[0].shape[1], self.feature_guidance[0].shape[3]
        geos = [DragGeometry(np.asarray(s_), np.asarray(t_), self.r1, self.voxel_size, S, Ca) for s_, t_ in edits]
        B = len(geos)
        st = getattr(self, "batch_stepper", None)
        reuse = (st is not None and st.batch == B and st.weights_generation == self.model.weights_generation
                 and all(st.compatible(g, cof, self.args.loss_type) for g in geos))
        if reuse:
            for b, g in enumerate(geos):
                st.retarget(g, scale, b)
        else:
            st = GuidedStepper(self.model, self.diffusion, geos, self.args.feat_layer, cof, self.args.loss_type, scale,
                               clip_denoised=True, use_graph=self.use_graph)
        self.batch_stepper = st
        st.img.copy_(self.w.detach().expand(B, -1, -1, -1))
        w_time = self.args.w_time
        for i in range(w_time - 1, -1, -1):
            st.step(i, self.feature_guidance[w_time - 1 - i], None if noises is None else noises[w_time - 1 - i])
        lat = st.img.clone()
        if not decode:
            return lat, None
        return lat, [self.get_mesh(tri_feat=lat[b:b + 1]) for b in range(B)]

    # ---- real-shape inversion (reference :552-566) -----------------------------------------------------------
    def latent_inversion(self, tri_feat):
        with th.no_grad():
            outs = self.diffusion.ddpm_inversion(self.model, tri_feat, self.args.w_time,
                                                 clip_denoised=self.args.clip_denoised,
                                                 feat_layer=self.args.feat_layer)
            noise = outs["latent"]
        self.w = noise.clone().detach()
        self.w0 = self.w.clone().detach()
        self.feature_guidance = [resize_feat_align(f).permute(0, 2, 3, 1).contiguous() for f in outs["inter_feat"]]
        self.mesh = self.get_mesh(tri_feat=outs["sample"])
        self.mesh0 = _copy_mesh(self.mesh)
        self.variance = [v.clone().detach() for v in outs["variance"]]
        self.variance_noise = [v.clone().detach() for v in outs["variance_noise"]]

    def clear_params(self):
        self.mesh0 = None
        self.mesh = None
        self.latent_code = None
        self.w0 = None
        self.w = None
        self.feature_guidance.clear()
        self.noise.clear()
        self.variance.clear()
        self.variance_noise.clear()

    def reset_params(self):
        if self.mesh is not None:
            self.mesh = _copy_mesh(self.mesh0)
        if self.w0 is not None:
            self.w = self.w0.clone().detach()
