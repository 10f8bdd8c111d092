"""MultiTriplane with the reference's surface (triplane_decoder/axisnetworks.py:78-90,517-575), the
whole per-point pipeline (3 bilinear plane samples, sum, Fourier features, 3-layer MLP) running as
one CUDA kernel (isb_triplane_decode_points / _grid).  The other experimental networks of that file
are never instantiated by the editor and are out of scope."""
import torch
import torch.nn as nn


class FourierFeatureTransform(nn.Module):
    """Parameter holder: `_B` (num_input_channels, mapping_size) * scale (axisnetworks.py:78-90)."""

    def __init__(self, num_input_channels, mapping_size, scale=10):
        super().__init__()
        self._num_input_channels = num_input_channels
        self._mapping_size = mapping_size
        self._B = nn.Parameter(torch.randn((num_input_channels, mapping_size)) * scale, requires_grad=False)


class _DecodePoints(torch.autograd.Function):
    """MultiTriplane.forward with a gradient for the planes (the decoder MLP is frozen, drag_utils.py:248-249):
    the reference's reconstruction guidance differentiates a BCE loss on sample points w.r.t. pred_xstart
    (drag_utils.py:443-463)."""

    @staticmethod
    def forward(ctx, planes_nchw, coords, ops, weights):
        R = planes_nchw.shape[-1]
        hwc = ops.to_nhwc(planes_nchw.detach().to(torch.float32).contiguous(), ops.empty((3, R, R, 32)))
        ctx.ops, ctx.weights, ctx.R = ops, weights, R
        ctx.save_for_backward(hwc, coords)
        return ops.decode_points(hwc, weights, coords, ops.empty((coords.shape[0],)))

    @staticmethod
    def backward(ctx, d_logits):
        hwc, coords = ctx.saved_tensors
        ops = ctx.ops
        d_hwc = ops.zeros((3, ctx.R, ctx.R, 32))
        ops.decode_points_backward(hwc, ctx.weights, coords, d_logits.detach().to(torch.float32).contiguous(), d_hwc)
        return ops.to_nchw(d_hwc, ops.empty((3, 32, ctx.R, ctx.R))), None, None, None


class MultiTriplane(nn.Module):
    def __init__(self, num_objs, input_dim=3, output_dim=1, noise_val=None, device="cuda"):
        super().__init__()
        if input_dim != 3 or output_dim != 1:
            raise NotImplementedError("the B200 decoder kernel implements input_dim=3, output_dim=1 (occupancy)")
        self.device = device
        self.num_objs = num_objs
        self.embeddings = [torch.randn(1, 32, 128, 128) * 0.001 for _ in range(3 * num_objs)]
        self.noise_val = noise_val
        self.net = nn.Sequential(
            FourierFeatureTransform(32, 64, scale=1),
            nn.Linear(128, 128), nn.ReLU(inplace=True),
            nn.Linear(128, 128), nn.ReLU(inplace=True),
            nn.Linear(128, output_dim),
        )
        self._ops = None
        self._planes_key = None
        self._planes_hwc = None

    # ---- kernel plumbing ---------------------------------------------------------------------
    def set_ops(self, ops):
        self._ops = ops

    def _get_ops(self):
        if self._ops is None:
            from ..ops import CudaOps
            dev = self.net[1].weight.device
            if dev.type != "cuda":
                raise RuntimeError("MultiTriplane parameters are on %s: the B200 path has no CPU fallback" % dev)
            self._ops = CudaOps(dev, "fp32")
        return self._ops

    def mlp_weights(self):
        f32 = lambda p: p.detach().to(torch.float32).contiguous()  # noqa: E731
        n = self.net
        return [f32(n[0]._B), f32(n[1].weight), f32(n[1].bias), f32(n[3].weight), f32(n[3].bias),
                f32(n[5].weight), f32(n[5].bias)]

    def planes_hwc(self, obj_idx):
        """The three (1,32,R,R) embeddings of `obj_idx` as one channels-last (3,R,R,32) tensor."""
        ops = self._get_ops()
        embs = [self.embeddings[3 * obj_idx + i] for i in range(3)]
        key = tuple((e.data_ptr(), e._version) for e in embs)
        if key != self._planes_key:
            nchw = torch.cat([e.detach().to(device=ops.device, dtype=torch.float32) for e in embs], dim=0).contiguous()
            R = nchw.shape[-1]
            self._planes_hwc = ops.to_nhwc(nchw, ops.empty((3, R, R, 32)))
            self._planes_key = key
        return self._planes_hwc

    # ---- reference surface ----------------------------------------------------------------------
    def forward(self, obj_idx, coordinates, debug=False):
        """coordinates (1, N, 3) in [-1,1] -> logits (1, N, 1)   (axisnetworks.py:546-562)."""
        if self.noise_val is not None and self.training:
            raise NotImplementedError("training-time feature noise is not implemented (decoder is frozen, eval)")
        batch, n, _ = coordinates.shape
        assert batch == 1
        ops = self._get_ops()
        coords = coordinates.detach().reshape(n, 3).to(device=ops.device, dtype=torch.float32).contiguous()
        embs = [self.embeddings[3 * obj_idx + i] for i in range(3)]
        if torch.is_grad_enabled() and any(e.requires_grad for e in embs):
            # planes that carry a graph (pred_xstart of a guided step): differentiable route
            planes = torch.cat([e.to(device=ops.device) for e in embs], dim=0)
            return _DecodePoints.apply(planes, coords, ops, self.mlp_weights()).reshape(1, n, 1)
        out = ops.decode_points(self.planes_hwc(obj_idx), self.mlp_weights(), coords, ops.empty((n,)))
        return out.reshape(1, n, 1)

    def tvreg(self):
        l = 0
        for e in self.embeddings:
            l += ((e[:, :, 1:] - e[:, :, :-1]) ** 2).sum() ** 0.5
            l += ((e[:, :, :, 1:] - e[:, :, :, :-1]) ** 2).sum() ** 0.5
        return l / self.num_objs

    def l2reg(self):
        l = 0
        for e in self.embeddings:
            l += (e ** 2).sum() ** 0.5
        return l / self.num_objs
