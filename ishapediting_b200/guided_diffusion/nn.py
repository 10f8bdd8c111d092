"""Host-side mirror of neural_field_diffusion/guided_diffusion/nn.py (reference).

Only parameter containers and host helpers live here; the arithmetic (GroupNorm32 nn.py:16-18,
SiLU, timestep_embedding nn.py:102-120) runs in libishape_b200.so.  The classes keep the
reference names so NFD checkpoints load with strict=True.
"""
import math

import torch as th
import torch.nn as nn


class GroupNorm32(nn.GroupNorm):
    """Parameter holder for a 32-group GroupNorm (computed in fp32 by isb_gn_forward)."""


def normalization(channels):
    return GroupNorm32(32, channels)


def conv_nd(dims, *args, **kwargs):
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    if dims == 2:
        return nn.Conv2d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims} (the B200 path implements the 2-D triplane UNet)")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def timestep_freqs(dim, max_period=10000):
    """The frequency table of nn.py:112-115, evaluated with the same torch expression so the
    device kernel multiplies bit-identical fp32 values."""
    half = dim // 2
    return th.exp(-math.log(max_period) * th.arange(start=0, end=half, dtype=th.float32) / half)
