"""Opt-in point tracking (isb_track_points; BASELINE.json north_star: "nearest-feature search with a warp-level
argmin", "tracked point indices must be bit-exact in fp32 mode").  The reference has no tracking function
(SURVEY.md §0.3), so PARITY IS UNPINNED; these are self-consistency tests:
  * the separable distance tables against a float64 torch restatement (grid_sample on the aligned planes);
  * the argmin lattice index BIT-EXACTLY against a brute-force search (torch.argmin over the full (2r+1)^3 lattice,
    first occurrence on ties) over the same fp32 distances;
  * a planted-feature case where the answer is known."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _brute_force(table, side):
    """table [B,3,side*side] fp32 -> (index [B], distance [B]) of min over (i,j,k) of (Txy[i,j] + Tyz[j,k]) + Txz[i,k]."""
    B = table.shape[0]
    t = table.reshape(B, 3, side, side)
    d = (t[:, 0][:, :, :, None] + t[:, 1][:, None, :, :]) + t[:, 2][:, :, None, :]          # [B, i, j, k], fp32
    flat = d.reshape(B, -1)
    idx = torch.argmin(flat, dim=1)
    return idx.to(torch.int32), flat.gather(1, idx[:, None])[:, 0]


@pytest.mark.parametrize("S,Cf,r,B", [(64, 512, 12, 4), (32, 128, 3, 5), (16, 64, 1, 2)])
def test_track_tables_and_bit_exact_argmin(S, Cf, r, B):
    from ishapediting_b200.drag_utils import align_maps, handle_features
    from ishapediting_b200.ops import CudaOps

    ops = CudaOps(DEV, "fp32")
    g = torch.Generator().manual_seed(S + r)
    feat = torch.randn(1, S, S, Cf, generator=g).to(DEV)
    chan_map, _, Ca = align_maps(Cf)
    cm = chan_map.to(DEV)
    aligned = ops.resize_feat_align(feat, cm, ops.empty((3, S, S, Ca)))
    voxel = 2.0 / (4 * S)
    center = ((torch.rand(B, 3, generator=g) - 0.5) * 1.9).to(DEV)            # some lattices cross the border
    f0 = handle_features(aligned, (torch.rand(B, 3, generator=g) - 0.5).numpy())
    idx, dist, pts, table = ops.track_points(feat, cm, f0, center, r, voxel)
    side = 2 * r + 1
    # (1) tables vs float64 torch
    planes = aligned.permute(0, 3, 1, 2).double()
    off = voxel * torch.arange(-r, r + 1, device=DEV, dtype=torch.float32)
    for pl, (au, av) in enumerate(((0, 1), (1, 2), (0, 2))):
        u = (center[:, au, None] + off[None, :])[:, :, None].expand(B, side, side)
        v = (center[:, av, None] + off[None, :])[:, None, :].expand(B, side, side)
        grid = torch.stack([u, v], dim=-1).double().reshape(1, B, side * side, 2)
        smp = F.grid_sample(planes[pl:pl + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        ref = (smp[0].permute(1, 2, 0) - f0[:, pl].double()[:, None, :]).abs().sum(-1)      # [B, side*side]
        err = float((table[:, pl].double() - ref).abs().max() / ref.abs().max())
        assert err < 2e-5, (pl, err)
    # (2) the index: bit-exact against brute force over the same fp32 distances
    bi, bd = _brute_force(table, side)
    assert torch.equal(idx, bi)
    assert torch.equal(dist, bd)
    # (3) the reported point is that lattice point
    i, j, k = bi // (side * side), (bi // side) % side, bi % side
    want = center + voxel * torch.stack([i - r, j - r, k - r], dim=1).float()
    assert torch.equal(pts, want)


def test_track_ties_go_to_the_lowest_index():
    from ishapediting_b200.drag_utils import align_maps
    from ishapediting_b200.ops import CudaOps

    ops = CudaOps(DEV, "fp32")
    S, Cf, r = 16, 64, 2
    chan_map, _, Ca = align_maps(Cf)
    feat = torch.zeros(1, S, S, Cf, device=DEV)                     # constant feature: every lattice point ties
    f0 = torch.zeros(3, 3, Ca, device=DEV)
    center = torch.zeros(3, 3, device=DEV)
    idx, dist, pts, _ = ops.track_points(feat, chan_map.to(DEV), f0, center, r, 0.01)
    assert idx.tolist() == [0, 0, 0] and dist.tolist() == [0.0, 0.0, 0.0]


def test_track_finds_a_planted_feature():
    """Move a smooth feature field by a known 3-D lattice offset (each aligned plane rolled by its projection of the
    offset): the tracker must find exactly that offset, at (almost) zero distance."""
    from ishapediting_b200.drag_utils import align_maps, handle_features, track_points
    from ishapediting_b200.ops import CudaOps

    ops = CudaOps(DEV, "fp32")
    S, Cf, r = 64, 512, 6
    voxel = 2.0 / (S - 1)                                            # one voxel = one feature pixel
    g = torch.Generator().manual_seed(5)
    base = F.interpolate(torch.randn(1, Cf, 8, 8, generator=g), size=(S, S), mode="bicubic", align_corners=True)
    feat0 = base.permute(0, 2, 3, 1).contiguous().to(DEV)            # [1,S,S,Cf]
    chan_map, _, Ca = align_maps(Cf)
    cm = chan_map.to(DEV).long()
    dx, dy, dz = -2, 1, 3
    feat1 = feat0.clone()
    # plane pl samples (u -> column, v -> row): xy = (x, y), yz = (y, z), xz = (x, z)
    for pl, (du, dv) in enumerate(((dx, dy), (dy, dz), (dx, dz))):
        ch = cm[pl * Ca:(pl + 1) * Ca]
        feat1[..., ch] = torch.roll(feat0[..., ch], shifts=(dv, du), dims=(1, 2))
    aligned0 = ops.resize_feat_align(feat0, chan_map.to(DEV), ops.empty((3, S, S, Ca)))
    p = torch.tensor([[-1.0 + 20 * voxel, -1.0 + 30 * voxel, -1.0 + 25 * voxel]])      # on the pixel lattice
    f0 = handle_features(aligned0, p.numpy())
    new_pts, idx, dist = track_points(feat1.contiguous(), f0, p, r, voxel, ops=ops)
    d = ((new_pts.cpu() - p) / voxel).round().int()[0].tolist()
    assert d == [dx, dy, dz], d
    side = 2 * r + 1
    assert int(idx) == ((dx + r) * side + (dy + r)) * side + (dz + r)
    assert float(dist) < 1e-3 * float(f0.abs().sum())


def test_training_with_tracking_runs_and_reports_lattice_points():
    """DragStuff.training(track=True): the guided update is unchanged (same latents as track=False), the tracked
    handles stay on the voxel lattice around their previous position, and track_error is reported in voxels."""
    from oracle import nfd_oracle as O
    from ishapediting_b200.drag_utils import DragStuff, get_args

    cfg = O.mid_cfg()
    cfg.update(in_out_channels=96, timestep_respacing="20")
    a = get_args(["--num_steps", "20", "--w_time", "3", "--shape_resolution", "16", "--feat_layer", "5", "--resolution", "32"])
    a.channel_mult, a.attention_resolutions, a.use_fp16 = "1,2,4", "16,8", True
    ds = DragStuff(args=a, device=DEV, use_graph=True)
    ds.mesh_on_device = False
    ds.model.load_state_dict(O.synth_state_dict(cfg))
    ds.model.to(DEV).eval()
    ds.set_offset1(3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 96, 32, 32, generator=g).to(DEV)
    noise = torch.randn(1, 96, 32, 32, generator=g).to(DEV)
    ds.update_latent_params(x, noise=noise)
    src = (torch.rand(2, 3, generator=g) - 0.5).numpy()
    tgt = src + 0.05
    list(ds.training(src, tgt, scale=600, cof=0.2, noises=[noise] * 3))
    plain = ds.stepper.img.clone()
    list(ds.training(src, tgt, scale=600, cof=0.2, noises=[noise] * 3, track=True, track_radius=3))
    assert torch.equal(ds.stepper.img, plain)                       # tracking only observes
    assert ds.tracked.shape == (2, 3) and ds.track_error.shape == (2,)
    steps = (ds.tracked.cpu() - torch.from_numpy(src)) / ds.voxel_size
    assert float((steps - steps.round()).abs().max()) < 1e-3         # on the lattice
    assert float(steps.abs().max()) <= 3 * 3 + 1e-3                  # at most r voxels per step, 3 steps
