"""Time embeddings phi(t) with phi(0) = 0 for ``ContinuousAffineCoupling`` (reference: stribor/net/time_net.py).

``TimeLinear`` is the one on the hot path: it is fused into the continuous-affine coupling kernels as ``scale * t``
(time_net.py:18-28).  The other embeddings of the reference are provided with the same constructor arguments,
parameter names and ``forward`` / ``derivative`` methods, so that code written against the reference keeps working; a
coupling that uses one of them evaluates the embedding and the conditioner as PyTorch modules and hands the resulting
per-row affine parameters to the element-wise CUDA kernels (``flows/coupling.py:_run_autograd``)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _epoch

__all__ = ['TimeIdentity', 'TimeLinear', 'TimeTanh', 'TimeLog', 'TimeFourier', 'TimeFourierBounded']


class TimeIdentity(_epoch.Tracked, nn.Module):
    """phi(t) = t on every output channel (time_net.py:7-16)."""

    def __init__(self, out_dim: int, **kwargs):
        super().__init__()
        self.out_dim = out_dim

    def forward(self, t):
        return t.expand(*t.shape[:-1], self.out_dim) if t.shape[-1] == 1 else t.repeat_interleave(self.out_dim, dim=-1)

    def derivative(self, t):
        return torch.ones_like(self.forward(t))


class TimeLinear(_epoch.Tracked, nn.Module):
    """phi(t) = scale * t with a learned per-channel scale (time_net.py:18-28)."""

    def __init__(self, out_dim: int, **kwargs):
        super().__init__()
        self.scale = nn.Parameter(torch.randn(1, out_dim))
        nn.init.xavier_uniform_(self.scale)

    def forward(self, t):
        return self.scale * t

    def derivative(self, t):
        return self.scale * torch.ones_like(t)


class TimeTanh(TimeLinear):
    """phi(t) = tanh(scale * t): bounded embedding (time_net.py:31-36)."""

    def forward(self, t):
        return torch.tanh(self.scale * t)

    def derivative(self, t):
        return self.scale * (1 - torch.tanh(self.scale * t) ** 2)


class TimeLog(TimeLinear):
    """phi(t) = log(exp(scale) * t + 1): slow growth in t (time_net.py:38-43)."""

    def forward(self, t):
        return torch.log(self.scale.exp() * t + 1)

    def derivative(self, t):
        a = self.scale.exp()
        return a / (a * t + 1)


class TimeFourier(_epoch.Tracked, nn.Module):
    """phi(t) = sum_k a_k sin(s_k t) with learned amplitudes and frequencies (time_net.py:45-83).

    Args:
        out_dim: output channels
        hidden_dim: number of Fourier features per channel
        lmbd: rate of the exponential distribution the frequencies are drawn from
        bounded: amplitudes softmax-normalised to sum to 1/2, i.e. |phi| <= 1/2 (``TimeFourierBounded``)
    """

    def __init__(self, out_dim, hidden_dim, lmbd=0.5, bounded=False, **kwargs):
        super().__init__()
        self.bounded = bounded
        self.hidden_dim = hidden_dim
        self.shift = nn.Parameter(torch.empty(out_dim, hidden_dim).exponential_(lmbd))
        self.weight = nn.Parameter(torch.empty(out_dim, hidden_dim))
        nn.init.xavier_normal_(self.weight)

    def get_scale(self):
        return torch.softmax(self.weight, -1) / 2 if self.bounded else self.weight / self.hidden_dim

    def forward(self, t):
        return (self.get_scale() * torch.sin(self.shift * t.unsqueeze(-1))).sum(-1)

    def derivative(self, t):
        return (self.shift * self.get_scale() * torch.cos(self.shift * t.unsqueeze(-1))).sum(-1)


class TimeFourierBounded(TimeFourier):
    """``TimeFourier`` with values in [-1/2, 1/2] (time_net.py:85-88)."""

    def __init__(self, out_dim, hidden_dim, lmbd=0.5, **kwargs):
        super().__init__(out_dim, hidden_dim, lmbd, True)
