"""Multi-GPU plumbing: one process per GPU (torchrun), no collective inside a denoising step.

The reference's only data-parallel code is its batch sampler (image_sample.py:53-113: per-rank
sampling + dist.all_gather).  The editing path shards naturally:
  * independent edits / seeds are dealt round-robin to the ranks (`assign_edits`) and their results
    collected with ONE all-gather (`gather_results`);
  * the dense decode grid is split into contiguous x-slabs (`slab_range`; the reference's flat index
    is x*res^2 + y*res + z, visualize.py:83-86, so an x-slab is a contiguous block of the volume).
    `decode_volume_sharded` lets the decode kernel write each rank's slab straight into its final
    offset of the full (res,res,res) buffer and completes the buffer with ONE in-place
    `all_gather_into_tensor` (equal slabs) — no padding, no staging copy, no concatenation.
torch.distributed (NCCL over NVLink on GPUs, Gloo in the CPU tests) is used only for these gathers.
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


def complete_volume_(vol: torch.Tensor):
    """vol: (res,res,res) on every rank, each rank having filled ITS x-slab (slab_range).  Fills in the other ranks'
    slabs in place.  Equal slabs: one in-place all-gather whose send buffer is this rank's slice of the receive
    buffer (ncclAllGather's in-place form).  Ragged slabs (res % world != 0): one broadcast per rank, still straight
    into the final offsets."""
    rank, ws = world()
    if ws == 1:
        return vol
    res = vol.shape[0]
    assert vol.is_contiguous()
    if res % ws == 0:
        b, e = slab_range(res, rank, ws)
        dist.all_gather_into_tensor(vol.view(-1), vol[b:e].reshape(-1))
    else:
        for r in range(ws):
            b, e = slab_range(res, r, ws)
            dist.broadcast(vol[b:e], src=r)
    return vol


def decode_volume_sharded(decode_slab, res: int, device, out: torch.Tensor = None):
    """The slab-sharded dense decode (BASELINE configs[3]).  `decode_slab(x_begin, x_end, out_slab)` must write the
    logits of grid rows [x_begin, x_end) into `out_slab` ((x_end-x_begin), res, res) — e.g.
    `lambda b, e, o: query_volume(decoder, 0, res, b, e, out=o)`.  Returns the full (res,res,res) volume, identical
    on every rank and bit-identical to the single-GPU decode (the slabs are disjoint: no reduction)."""
    rank, ws = world()
    vol = out if out is not None else torch.empty((res, res, res), dtype=torch.float32, device=device)
    b, e = slab_range(res, rank, ws)
    decode_slab(b, e, vol[b:e])
    return complete_volume_(vol)


def gather_volume(local_slab: torch.Tensor, res: int):
    """All-gather per-rank x-slabs that were decoded into separate tensors ((nx_r, res, res) each) into the full
    volume on every rank.  One copy of the local slab into its final offset, then `complete_volume_`; prefer
    `decode_volume_sharded`, which avoids even that copy."""
    rank, ws = world()
    if ws == 1:
        return local_slab.reshape(res, res, res)
    vol = torch.empty((res, res, res), dtype=local_slab.dtype, device=local_slab.device)
    b, e = slab_range(res, rank, ws)
    vol[b:e].copy_(local_slab.reshape(e - b, res, res))
    return complete_volume_(vol)


def gather_results(local: torch.Tensor, n_total: int):
    """Gather per-edit results dealt by assign_edits back into edit order on every rank (the reference pattern:
    image_sample.py:104-105 all_gather of the per-rank samples).  local: (n_local, ...) for edits rank, rank+ws, ...
    One all_gather_into_tensor; edit e sits at [e % ws, e // ws] of the gathered buffer."""
    rank, ws = world()
    if ws == 1:
        return local
    n_max = (n_total + ws - 1) // ws
    tail = tuple(local.shape[1:])
    if local.shape[0] == n_max:
        send = local.contiguous()
    else:
        send = local.new_zeros((n_max,) + tail)
        send[:local.shape[0]] = local
    buf = local.new_empty((ws, n_max) + tail)
    dist.all_gather_into_tensor(buf.view(-1), send.view(-1))
    return buf.transpose(0, 1).reshape((ws * n_max,) + tail)[:n_total]
