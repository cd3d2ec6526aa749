"""Base densities (reference: stribor/dist/normal.py:8-54).

``UnitNormal`` is what the hot path uses: inside ``NormalizingFlow.log_prob`` its log-density
is fused into the last layer's kernel.  Like the reference it is a plain
``torch.distributions`` wrapper (not an ``nn.Module``, not in the state-dict); unlike the
reference its ``log_prob`` / ``sample`` follow the device of the data / the flow.
"""
from __future__ import annotations

import math
from numbers import Number

import torch
import torch.distributions as td

__all__ = ['Normal', 'UnitNormal']


class Normal(td.Independent):
    def __init__(self, loc, scale, **kwargs):
        self.loc = loc
        self.scale = scale
        rbd = 0 if isinstance(self.loc, float) else 1
        super().__init__(td.Normal(self.loc, self.scale, **kwargs), reinterpreted_batch_ndims=rbd)


class UnitNormal(Normal):
    def __init__(self, dim: int, **kwargs):
        self.dim = dim
        super().__init__(torch.zeros(self.dim), torch.ones(self.dim), **kwargs)

    def log_prob(self, x):
        # sum_j -(x_j^2)/2 - log(sqrt(2 pi)), evaluated where x lives (differentiable)
        return (-(x ** 2) / 2 - math.log(math.sqrt(2 * math.pi))).sum(-1)

    def sample(self, sample_shape=torch.Size(), device=None):
        if isinstance(sample_shape, Number):
            sample_shape = (sample_shape,)
        return torch.randn(*sample_shape, self.dim, device=device)

    def rsample(self, sample_shape=torch.Size(), device=None):
        return self.sample(sample_shape, device=device)
