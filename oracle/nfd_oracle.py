"""CPU oracle for the drag-guided triplane-diffusion editing step.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this module, and only as the checker / reported baseline —
the product path (ishapediting_b200/) never does and fails loudly without its CUDA library.

This is a plain-PyTorch fp32 (CPU) restatement of the reference algorithm, written functionally
over a state_dict (no nn.Module tree).  The arithmetic itself lives in PyTorch (the reference pins
torch==1.12.0, /root/reference/README.md:29; this image has 2.11) — conv2d, group_norm, softmax,
grid_sample, interpolate and autograd semantics used here are unchanged between those versions.

Parity pinning: the reference ships NO tests, golden vectors or fixtures.  The
oracle is therefore pinned against outputs of the reference's own modules imported in the build
container (oracle/ref_import.py + tests/golden/make_golden.py, fixtures under tests/golden/), and
tests/test_oracle_vs_reference.py re-checks it live whenever /root/reference is present.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# configuration (drag_utils.py:44-57) and architecture walk (unet.py:428-616, script_util.py:132-187)
# ------------------------------------------------------------------------------------------------
NFD_CFG = dict(
    image_size=128, in_out_channels=96, num_channels=256, num_res_blocks=2, num_head_channels=64,
    attention_resolutions="32,16,8", channel_mult=(1, 1, 2, 3, 4), learn_sigma=True,
    diffusion_steps=1000, timestep_respacing="200", feat_layer=8,
)


def small_cfg():
    """A shrunken NFD-shaped UNet (same topology: 5 levels, attention at the 3 coarsest) used for
    the committed golden fixtures and fast parity tests."""
    c = dict(NFD_CFG)
    c.update(image_size=32, in_out_channels=12, num_channels=64, attention_resolutions="8,4,2")
    return c


def mid_cfg():
    """NFD-width (256 base channels, so every GroupNorm group is a multiple of 8 channels and every
    GEMM K a multiple of 64 — the shapes the sm_100a kernels are written for) but 3 levels at 32x32:
    small enough for the CPU oracle to run forward+backward in about a second."""
    c = dict(NFD_CFG)
    c.update(image_size=32, in_out_channels=32, channel_mult=(1, 2, 4), attention_resolutions="16,8", feat_layer=5)
    return c


def unet_structure(cfg):
    """Block list of UNetModel.__init__ (unet.py:480-616): returns (input, middle, output) where
    each block is a list of ('res', cin, cout, updown) / ('attn', ch, heads) entries."""
    mc = cfg["num_channels"]
    mult = cfg["channel_mult"]
    att_ds = [cfg["image_size"] // int(r) for r in cfg["attention_resolutions"].split(",")]
    nrb = cfg["num_res_blocks"]
    hc = cfg["num_head_channels"]
    ch = int(mult[0] * mc)
    inp = [[("conv", cfg["in_out_channels"], ch)]]
    chans = [ch]
    ds = 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            blk = [("res", ch, int(m * mc), 0)]
            ch = int(m * mc)
            if ds in att_ds:
                blk.append(("attn", ch, ch // hc))
            inp.append(blk)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("res", ch, ch, 1)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch, 0), ("attn", ch, ch // hc), ("res", ch, ch, 0)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            blk = [("res", ch + ich, int(mc * m), 0)]
            ch = int(mc * m)
            if ds in att_ds:
                blk.append(("attn", ch, ch // hc))
            if level and i == nrb:
                blk.append(("res", ch, ch, 2))
                ds //= 2
            out.append(blk)
    return inp, mid, out


def unet_param_shapes(cfg):
    """Names and shapes of every parameter, in the reference's registration order."""
    inp, mid, out = unet_structure(cfg)
    mc = cfg["num_channels"]
    emb = 4 * mc
    shapes = [("time_embed.0.weight", (emb, mc)), ("time_embed.0.bias", (emb,)),
              ("time_embed.2.weight", (emb, emb)), ("time_embed.2.bias", (emb,))]

    def res(prefix, cin, cout):
        s = [(f"{prefix}.in_layers.0.weight", (cin,)), (f"{prefix}.in_layers.0.bias", (cin,)),
             (f"{prefix}.in_layers.2.weight", (cout, cin, 3, 3)), (f"{prefix}.in_layers.2.bias", (cout,)),
             (f"{prefix}.emb_layers.1.weight", (2 * cout, emb)), (f"{prefix}.emb_layers.1.bias", (2 * cout,)),
             (f"{prefix}.out_layers.0.weight", (cout,)), (f"{prefix}.out_layers.0.bias", (cout,)),
             (f"{prefix}.out_layers.3.weight", (cout, cout, 3, 3)), (f"{prefix}.out_layers.3.bias", (cout,))]
        if cin != cout:
            s += [(f"{prefix}.skip_connection.weight", (cout, cin, 1, 1)), (f"{prefix}.skip_connection.bias", (cout,))]
        return s

    def attn(prefix, ch):
        return [(f"{prefix}.norm.weight", (ch,)), (f"{prefix}.norm.bias", (ch,)),
                (f"{prefix}.qkv.weight", (3 * ch, ch, 1)), (f"{prefix}.qkv.bias", (3 * ch,)),
                (f"{prefix}.proj_out.weight", (ch, ch, 1)), (f"{prefix}.proj_out.bias", (ch,))]

    def block(prefix, blk):
        s = []
        for j, e in enumerate(blk):
            if e[0] == "conv":
                s += [(f"{prefix}.{j}.weight", (e[2], e[1], 3, 3)), (f"{prefix}.{j}.bias", (e[2],))]
            elif e[0] == "res":
                s += res(f"{prefix}.{j}", e[1], e[2])
            else:
                s += attn(f"{prefix}.{j}", e[1])
        return s

    for i, blk in enumerate(inp):
        shapes += block(f"input_blocks.{i}", blk)
    shapes += block("middle_block", mid)
    for i, blk in enumerate(out):
        shapes += block(f"output_blocks.{i}", blk)
    ch0 = int(cfg["channel_mult"][0] * mc)
    oc = cfg["in_out_channels"] * (2 if cfg["learn_sigma"] else 1)
    shapes += [("out.0.weight", (ch0,)), ("out.0.bias", (ch0,)), ("out.2.weight", (oc, ch0, 3, 3)), ("out.2.bias", (oc,))]
    return shapes


def synth_state_dict(cfg, seed=1234, branch_gain=0.5):
    """Deterministic synthetic weights (checkpoints are unavailable offline).  The reference's
    random init is degenerate — zero_module zeroes the 2nd conv of every ResBlock, every attention
    proj_out and the out conv (unet.py:210-212,294,615), so eps == 0 and no gradient flows
     — hence every tensor is drawn here: variance-preserving N(0, 1/fan_in) for
    weights (x branch_gain on the residual-branch output convs so the stream stays O(1) through 51
    residual adds), N(0, 0.02^2) biases, GroupNorm gamma ~ 1 + N(0, 0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in unet_param_shapes(cfg):
        if name.endswith("bias"):
            t = torch.randn(shape, generator=g) * 0.02
        elif len(shape) == 1:  # GroupNorm weight
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = int(np.prod(shape[1:]))
            t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
            if name.endswith("out_layers.3.weight") or name.endswith("proj_out.weight"):
                t = t * branch_gain
            if name.startswith("out.2"):
                t = t * 0.5
        sd[name] = t
    return sd


# ------------------------------------------------------------------------------------------------
# UNet forward (unet.py:634-671)
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim, max_period=10000):
    """nn.py:102-120."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, p):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)  # nn.py:16-18,92-99


def _resblock(sd, p, x, emb, updown):
    """ResBlock._forward with use_scale_shift_norm (unet.py:236-256)."""
    h = F.silu(_gn(x, sd, p + ".in_layers.0"))
    if updown == 1:      # Downsample without conv = AvgPool2d(2) on both branches (unet.py:136,190-191)
        h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    elif updown == 2:    # Upsample without conv = nearest 2x (unet.py:107,187-188)
        h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])[:, :, None, None]
    scale, shift = torch.chunk(e, 2, dim=1)
    h = _gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)  # dropout: eval
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attention(sd, p, x, heads):
    """AttentionBlock._forward + QKVAttentionLegacy.forward (unet.py:299-305,337-354)."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, sd, p + ".norm"), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    length = xf.shape[-1]
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, 3 * ch, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, length)
    h = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def _run_block(sd, prefix, blk, h, emb):
    for j, e in enumerate(blk):
        p = f"{prefix}.{j}"
        if e[0] == "conv":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif e[0] == "res":
            h = _resblock(sd, p, h, emb, e[3])
        else:
            h = _attention(sd, p, h, e[2])
    return h


def unet_forward(sd, cfg, x, t_orig, feat_layer=-1):
    """UNetModel.forward (unet.py:634-671).  t_orig: original (un-respaced) timesteps, int64 [N]."""
    inp, mid, out = unet_structure(cfg)
    emb = timestep_embedding(t_orig, cfg["num_channels"])
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    hs = []
    h = x.float()
    for i, blk in enumerate(inp):
        h = _run_block(sd, f"input_blocks.{i}", blk, h, emb)
        hs.append(h)
    h = _run_block(sd, "middle_block", mid, h, emb)
    inter = None
    for i, blk in enumerate(out):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"output_blocks.{i}", blk, h, emb)
        if i == feat_layer:
            inter = h.clone()
    h = F.conv2d(F.silu(_gn(h, sd, "out.0")), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return (h, inter) if feat_layer >= 0 else h


# ------------------------------------------------------------------------------------------------
# diffusion schedule (gaussian_diffusion.py:118-169, respace.py:6-86)
# ------------------------------------------------------------------------------------------------
class Schedule:
    def __init__(self, diffusion_steps=1000, respacing="200"):
        scale = 1000 / diffusion_steps
        base_betas = np.linspace(scale * 0.0001, scale * 0.02, diffusion_steps, dtype=np.float64)  # :27-34
        acp = np.cumprod(1.0 - base_betas)
        n = int(respacing)
        # respace.py:36-59 with a single section
        stride = (diffusion_steps - 1) / (n - 1) if n > 1 else 1
        cur, use = 0.0, []
        for _ in range(n):                    # accumulated float stride + round(), as the reference
            use.append(round(cur))
            cur += stride
        use = sorted(set(use))
        self.timestep_map = use
        last, betas = 1.0, []
        for i in use:                         # respace.py:71-79
            betas.append(1 - acp[i] / last)
            last = acp[i]
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = len(betas)
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)                 # :151
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)           # :156


def _f(a, i):
    return torch.tensor(float(np.float32(a[i])))  # _extract_into_tensor: float64 table -> .float() (:1045)


def p_sample_guidance(sd, cfg, sched, x, i, noise, feat_layer=8, clip_denoised=True):
    """p_mean_variance + p_sample_guidance for EPSILON / LEARNED_RANGE
    (gaussian_diffusion.py:232-331, 446-510).  `i` is the respaced step index."""
    t_orig = torch.full((x.shape[0],), sched.timestep_map[i], dtype=torch.int64, device=x.device)   # respace.py:122-124
    if feat_layer >= 0:
        model_output, inter = unet_forward(sd, cfg, x, t_orig, feat_layer)
    else:                                                   # unet.py:668-671: no intermediate feature requested
        model_output, inter = unet_forward(sd, cfg, x, t_orig, feat_layer), None
    C = x.shape[1]
    eps, v = torch.split(model_output, C, dim=1)
    min_log = _f(sched.posterior_log_variance_clipped, i)
    max_log = _f(np.log(sched.betas), i)
    frac = (v + 1) / 2
    log_var = frac * max_log + (1 - frac) * min_log
    var = torch.exp(log_var)
    x0 = _f(sched.sqrt_recip_alphas_cumprod, i) * x - _f(sched.sqrt_recipm1_alphas_cumprod, i) * eps
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    mean = _f(sched.posterior_mean_coef1, i) * x0 + _f(sched.posterior_mean_coef2, i) * x
    nonzero = 0.0 if i == 0 else 1.0
    sample = mean + nonzero * torch.sqrt(var) * noise
    return dict(sample=sample, pred_xstart=x0, inter_feat=inter, model_output=eps, noise=noise, variance=var, mean=mean)


# ------------------------------------------------------------------------------------------------
# DDIM variants and cond_fn conditioning (gaussian_diffusion.py:364-398, 400-444, 654-761)
# ------------------------------------------------------------------------------------------------
def _eps_from_xstart(sched, x, i, x0):
    """_predict_eps_from_xstart (:351-355)."""
    return (_f(sched.sqrt_recip_alphas_cumprod, i) * x - x0) / _f(sched.sqrt_recipm1_alphas_cumprod, i)


def p_sample_cond(sd, cfg, sched, x, i, noise, cond_fn):
    """p_sample with cond_fn (:400-444): condition_mean (:364-377) shifts the posterior mean by variance * grad;
    the sample uses exp(0.5 * log_variance) (:443), unlike p_sample_guidance's sqrt(variance)."""
    o = p_sample_guidance(sd, cfg, sched, x, i, noise, feat_layer=-1)
    t = torch.full((x.shape[0],), sched.timestep_map[i], dtype=torch.int64)     # respace.py:97-101: original t
    mean = o["mean"].float() + o["variance"] * cond_fn(x, t).float()
    nonzero = 0.0 if i == 0 else 1.0
    return dict(sample=mean + nonzero * torch.exp(0.5 * torch.log(o["variance"])) * noise, pred_xstart=o["pred_xstart"],
                mean=mean)


def ddim_sample(sd, cfg, sched, x, i, noise, eta=0.0, cond_fn=None, feat_layer=-1):
    """ddim_sample (:654-705) with optional condition_score (:379-398)."""
    o = p_sample_guidance(sd, cfg, sched, x, i, noise, feat_layer=feat_layer)
    x0 = o["pred_xstart"]
    abar, abar_prev = _f(sched.alphas_cumprod, i), _f(sched.alphas_cumprod_prev, i)
    if cond_fn is not None:
        t = torch.full((x.shape[0],), sched.timestep_map[i], dtype=torch.int64)     # respace.py:97-101: original t
        eps = _eps_from_xstart(sched, x, i, x0) - (1 - abar).sqrt() * cond_fn(x, t)
        x0 = _f(sched.sqrt_recip_alphas_cumprod, i) * x - _f(sched.sqrt_recipm1_alphas_cumprod, i) * eps
    eps = _eps_from_xstart(sched, x, i, x0)
    sigma = eta * torch.sqrt((1 - abar_prev) / (1 - abar)) * torch.sqrt(1 - abar / abar_prev)
    mean_pred = x0 * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev - sigma ** 2) * eps
    nonzero = 0.0 if i == 0 else 1.0
    return dict(sample=mean_pred + nonzero * sigma * noise, pred_xstart=x0, inter_feat=o["inter_feat"],
                model_output=o["model_output"])


def ddim_reverse_sample(sd, cfg, sched, x, i):
    """ddim_reverse_sample (:718-761): x_{t+1} along the deterministic DDIM ODE."""
    o = p_sample_guidance(sd, cfg, sched, x, i, torch.zeros_like(x), feat_layer=-1)
    eps = _eps_from_xstart(sched, x, i, o["pred_xstart"])
    abar_next = _f(sched.alphas_cumprod_next, i)
    return dict(sample=o["pred_xstart"] * torch.sqrt(abar_next) + torch.sqrt(1 - abar_next) * eps,
                pred_xstart=o["pred_xstart"])


def ddim_guidance_sample(sched, eps, grads, xt, i, clip_denoised=True):
    """ddim_guidance_sample (:707-716)."""
    eps = eps - _f(sched.sqrt_one_minus_alphas_cumprod, i) * grads
    x0 = _f(sched.sqrt_recip_alphas_cumprod, i) * xt - _f(sched.sqrt_recipm1_alphas_cumprod, i) * eps
    if clip_denoised:
        x0 = x0.clamp(-1, 1)
    eps = _eps_from_xstart(sched, xt, i, x0)
    abar_prev = _f(sched.alphas_cumprod_prev, i)
    return x0 * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev) * eps


# ------------------------------------------------------------------------------------------------
# drag guidance (drag_utils.py:134-159, 302-392)
# ------------------------------------------------------------------------------------------------
def make_offsets(r):
    p = torch.arange(-r, r + 1)
    px, py, pz = torch.meshgrid(p, p, p, indexing="ij")
    return torch.stack([px.reshape(-1), py.reshape(-1), pz.reshape(-1)], dim=-1)


def resize_feat_align(feature):
    """drag_utils.py:141-159 with cat_var=True."""
    b, c2 = feature.shape[:2]
    c = c2 // 2
    mean, var = torch.split(feature, c, dim=1)
    if c % 3:
        e = c - c % 3
        mean = F.interpolate(mean.permute(2, 3, 0, 1), (b, e)).permute(2, 3, 0, 1)
        var = F.interpolate(var.permute(2, 3, 0, 1), (b, e)).permute(2, 3, 0, 1)
    hh, ww = mean.shape[2], mean.shape[3]
    return torch.cat((mean.reshape(3, -1, hh, ww), var.reshape(3, -1, hh, ww)), dim=1).float()


def drag_setup(sources, targets, r1, voxel_size, img_width):
    """Point sets, plane grids and mask index sets of drag_utils.py:305-334.  Returns the grids and
    the three complement masks as boolean (3, W, W) arrays indexed [row, col] like the reference's
    `feature[pl, :, idx[:,0], idx[:,1]]`."""
    src = torch.as_tensor(sources, dtype=torch.float32)
    tgt = torch.as_tensor(targets, dtype=torch.float32)
    off = make_offsets(r1)
    patch = src.unsqueeze(1) + voxel_size * off.unsqueeze(0)
    shift = tgt.unsqueeze(1) + voxel_size * off.unsqueeze(0)
    patch_grid = torch.cat((patch[..., :2].unsqueeze(0), patch[..., 1:].unsqueeze(0), patch[..., :3:2].unsqueeze(0)), 0)
    shift_grid = torch.cat((shift[..., :2].unsqueeze(0), shift[..., 1:].unsqueeze(0), shift[..., :3:2].unsqueeze(0)), 0)
    pi = torch.round((patch + 1) * (img_width - 1) / 2).to(torch.int16).reshape(-1, 3)
    si = torch.round((shift + 1) * (img_width - 1) / 2).to(torch.int16).reshape(-1, 3)
    content = torch.cat((pi, si), 0).long()
    masks = np.ones((3, img_width, img_width), dtype=bool)
    for pl, cols in enumerate(([1, 0], [2, 1], [2, 0])):        # xy, yz, xz (drag_utils.py:329-334)
        rc = content[:, cols]
        ok = (rc[:, 0] >= 0) & (rc[:, 0] < img_width) & (rc[:, 1] >= 0) & (rc[:, 1] < img_width)
        rc = rc[ok].numpy()
        masks[pl, rc[:, 0], rc[:, 1]] = False
    return patch_grid, shift_grid, masks


def drag_loss(edit_feature, origin_feature, patch_grid, shift_grid, masks, cof=0.2, loss_type="l2"):
    """Motion supervision + mask regulariser (drag_utils.py:355-382)."""
    patch = F.grid_sample(origin_feature, patch_grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    shift = F.grid_sample(edit_feature, shift_grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    C = origin_feature.shape[1]
    m = torch.from_numpy(masks).to(edit_feature.device)
    cnt = int(masks.sum())
    diff = (edit_feature - origin_feature) * m[:, None].float()
    if loss_type == "l1":
        mask_loss = diff.abs().sum() / (C * cnt) if cof > 0 else 0.0
        return -F.l1_loss(shift, patch.detach()) - cof * mask_loss
    mask_loss = (diff ** 2).sum() / (C * cnt) if cof > 0 else 0.0
    return -((shift.reshape(-1) - patch.detach().reshape(-1)) ** 2).mean() - cof * mask_loss


def guided_step(sd, cfg, sched, img, i, origin_feature, noise, patch_grid, shift_grid, masks,
                scale=600.0, cof=0.2, loss_type="l2"):
    """One iteration of the DragStuff.training loop body (drag_utils.py:340-392)."""
    img = img.detach().clone().requires_grad_(True)
    outs = p_sample_guidance(sd, cfg, sched, img, i, noise, feat_layer=cfg["feat_layer"])
    edit = resize_feat_align(outs["inter_feat"])
    loss = drag_loss(edit, origin_feature, patch_grid, shift_grid, masks, cof, loss_type)
    (grad,) = torch.autograd.grad(loss, img)
    with torch.no_grad():
        nxt = outs["sample"] + outs["variance"] * (scale * grad)
    return dict(img=nxt.detach(), grad=grad.detach(), loss=loss.detach(),
                **{k: (v.detach() if torch.is_tensor(v) else v) for k, v in outs.items()})


# ------------------------------------------------------------------------------------------------
# triplane decoder (axisnetworks.py:78-90,517-562; visualize.py:79-98)
# ------------------------------------------------------------------------------------------------
def synth_decoder(seed=7, plane_seed=6, R=128):
    """Decoder MLP with torch's default Linear init + Fourier matrix randn(32,64) (axisnetworks.py:84,526-535)
    and smooth synthetic planes whose zero level set is non-trivial."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        b = 1 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b, (torch.rand(o, generator=g) * 2 - 1) * b

    w = {"B": torch.randn(32, 64, generator=g)}
    w["w1"], w["b1"] = lin(128, 128)
    w["w2"], w["b2"] = lin(128, 128)
    w["w3"], w["b3"] = lin(1, 128)
    gp = torch.Generator().manual_seed(plane_seed)
    planes = torch.randn(3, 32, R, R, generator=gp)
    k = torch.ones(1, 1, 9, 9) / 81.0
    for _ in range(2):   # low-pass so the field is smooth at the decode resolution
        planes = F.conv2d(planes.reshape(96, 1, R, R), k, padding=4).reshape(3, 32, R, R)
    planes = planes / planes.std() * 0.05
    return w, planes.contiguous()


def triplane_forward(w, planes, coords):
    """MultiTriplane.forward (axisnetworks.py:546-562): coords (N,3) -> logits (N,)."""
    c = coords[None]

    def sample(c2, plane):
        s = F.grid_sample(plane[None], c2.reshape(1, 1, -1, 2), mode="bilinear", padding_mode="zeros", align_corners=True)
        return s.reshape(1, plane.shape[0], -1).permute(0, 2, 1)

    f = sample(c[..., 0:2], planes[0]) + sample(c[..., 1:3], planes[1]) + sample(c[..., :3:2], planes[2])
    x = 2 * np.pi * (f.reshape(-1, 32) @ w["B"])
    x = torch.cat([torch.sin(x), torch.cos(x)], dim=-1)
    x = F.relu(F.linear(x, w["w1"], w["b1"]))
    x = F.relu(F.linear(x, w["w2"], w["b2"]))
    return F.linear(x, w["w3"], w["b3"]).reshape(-1)


def recon_guided_step(sd, cfg, sched, img, i, noise, w_dec, coords, gt, scale=600.0, rng=1.0, middle=0.0):
    """One iteration of the reference's real-shape reconstruction guidance, drag_utils.py:445-463 (train_triplane):
    classifier guidance on the predicted x_start — BCE of the decoder's occupancy logits at sampled points against
    the mesh occupancies, differentiated w.r.t. the noisy latent.  coords (P,3), gt (P,1) in {0,1}."""
    img = img.detach().clone().requires_grad_(True)
    outs = p_sample_guidance(sd, cfg, sched, img, i, noise, feat_layer=-1)
    R = img.shape[-1]
    planes = (outs["pred_xstart"] * rng + middle).reshape(3, 32, R, R)        # :449-453
    prediction = triplane_forward(w_dec, planes, coords).reshape(-1, 1)       # :456
    loss = -F.binary_cross_entropy_with_logits(prediction, gt)               # :458
    (grad,) = torch.autograd.grad(loss, img)
    nxt = (outs["sample"] + outs["variance"] * (scale * grad)).detach()       # :460-463
    return dict(img=nxt, grad=grad.detach(), loss=loss.detach(), pred_xstart=outs["pred_xstart"].detach(),
                logits=prediction.detach().reshape(-1))


def dense_grid_coords(res, x_begin=0, x_end=None):
    """visualize.py:79-86: index = x*res^2 + y*res + z."""
    x_end = res if x_end is None else x_end
    lin = torch.linspace(-1, 1, res)
    xs, ys, zs = torch.meshgrid([lin[x_begin:x_end], lin, lin], indexing="ij")
    return torch.cat([xs.unsqueeze(-1), ys.unsqueeze(-1), zs.unsqueeze(-1)], -1).reshape(-1, 3)


def decode_grid(w, planes, res, x_begin=0, x_end=None, max_batch=50000):
    """The chunk loop of create_obj_o3d (visualize.py:89-98) up to the logit volume."""
    coords = dense_grid_coords(res, x_begin, x_end)
    out = torch.empty(coords.shape[0])
    with torch.no_grad():
        for h in range(0, coords.shape[0], max_batch):
            out[h:h + max_batch] = triplane_forward(w, planes, coords[h:h + max_batch])
    return out
