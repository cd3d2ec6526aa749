"""Time embeddings (reference: stribor/net/time_net.py:18-28).  Only ``TimeLinear`` is on the
hot path; it is fused into the continuous-affine coupling kernel as ``scale * t``."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _epoch

__all__ = ['TimeLinear']


class TimeLinear(_epoch.Tracked, nn.Module):
    def __init__(self, out_dim: int, **kwargs):
        super().__init__()
        self.scale = nn.Parameter(torch.randn(1, out_dim))
        nn.init.xavier_uniform_(self.scale)

    def forward(self, t):
        return self.scale * t

    def derivative(self, t):
        return self.scale * torch.ones_like(t)
