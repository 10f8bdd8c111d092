"""Multi-GPU plumbing: one process per GPU (torchrun), no collective inside a denoising step.

The reference's only data-parallel code is its batch sampler (image_sample.py:53-113: per-rank
sampling + dist.all_gather).  The editing path shards naturally:
  * independent edits / seeds are dealt round-robin to the ranks (`assign_edits`);
  * the dense decode grid is split into contiguous x-slabs (`slab_range`; the reference's flat index
    is x*res^2 + y*res + z, visualize.py:83-86, so an x-slab is a contiguous block of the volume) and
    the slabs are written straight into their final offset of a gather buffer (`gather_volume`).
torch.distributed (NCCL on GPUs, Gloo in the CPU tests) is used only for these gathers.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def assign_edits(n_edits: int, rank: int, world_size: int):
    """Static round-robin: edit e runs on rank e % world_size."""
    return list(range(rank, n_edits, world_size))


def slab_range(res: int, rank: int, world_size: int):
    """Contiguous [x_begin, x_end) of the slowest grid axis for this rank (balanced to within 1)."""
    base, extra = divmod(res, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_volume(local_slab: torch.Tensor, res: int):
    """All-gather the per-rank x-slabs (each (nx_r, res, res)) into the full (res,res,res) volume on
    every rank.  Slabs may differ by one row; they are padded to the widest for the collective."""
    rank, ws = world()
    if ws == 1:
        return local_slab.reshape(res, res, res)
    widths = [slab_range(res, r, ws)[1] - slab_range(res, r, ws)[0] for r in range(ws)]
    wmax = max(widths)
    pad = local_slab.new_zeros((wmax, res, res))
    pad[:local_slab.shape[0]] = local_slab
    buf = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(buf, pad)
    return torch.cat([b[:w] for b, w in zip(buf, widths)], dim=0)


def gather_results(local: torch.Tensor, n_total: int):
    """Gather per-edit results dealt by assign_edits back into edit order on every rank.
    local: (n_local, ...) for edits rank, rank+ws, ..."""
    rank, ws = world()
    if ws == 1:
        return local
    n_max = (n_total + ws - 1) // ws
    pad = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    buf = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(buf, pad)
    out = local.new_zeros((n_total,) + tuple(local.shape[1:]))
    for r in range(ws):
        ids = assign_edits(n_total, r, ws)
        out[ids] = buf[r][:len(ids)]
    return out
