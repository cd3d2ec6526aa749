"""Parameter-free layers that sit between couplings (scope table "next" rows):
``Flip`` / ``Permute`` (reference: stribor/flows/permute.py:11-82) and ``Sigmoid`` / ``Logit``
(stribor/flows/sigmoid.py:9-56).  Without gradients they run as CUDA kernels of the same C ABI
(`STB_PERMUTE / STB_SIGMOID / STB_LOGIT`) and take part in fused chains; with gradients enabled
they are plain differentiable tensor expressions (they own no parameters).
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from .. import _lib
from ..flow import ElementwiseTransform, run_layer
from ._native import layer_is_plain, needs_autograd

__all__ = ['Flip', 'Permute', 'Sigmoid', 'Logit']


def _meta(kind, dim, n_params):
    return [kind, dim, 0, 0, 0, 0, 0, 0, 0, 0, 0, n_params, 0, 0]


class _Pointwise(ElementwiseTransform):
    kind = None
    in_place_ok = True

    def _params(self, dim, device):
        return []

    def chainable(self):
        return self.in_place_ok and layer_is_plain(self)

    def describe(self, dim, latent_dim, device):
        p = self._params(dim, device)
        return {'meta': _meta(self.kind, dim, len(p)) + self._meta_tail(dim), 'fmeta': [0., 1.] * 3, 'mask': None,
                'params': p, 'packed': None}

    def _meta_tail(self, dim):
        return []

    def _run(self, x, direction, want_ldj):
        return run_layer(self.describe(x.shape[-1], 0, x.device), x, None, None, direction, want_ldj)

    def forward(self, x, **kwargs):
        if needs_autograd(self, x):
            return self._torch(x, False)[0]
        return self._run(x, _lib.FORWARD, False)[0]

    def inverse(self, y, **kwargs):
        if needs_autograd(self, y):
            return self._torch(y, True)[0]
        return self._run(y, _lib.INVERSE, False)[0]

    def forward_and_log_det_jacobian(self, x, **kwargs):
        if needs_autograd(self, x):
            return self._torch(x, False)
        return self._run(x, _lib.FORWARD, True)

    def inverse_and_log_det_jacobian(self, y, **kwargs):
        if needs_autograd(self, y):
            return self._torch(y, True)
        return self._run(y, _lib.INVERSE, True)

    def log_det_jacobian(self, x, y=None, **kwargs):
        return self.forward_and_log_det_jacobian(x)[1]


class Permute(_Pointwise):
    """Fixed random permutation of the last dimension (permute.py:47-82); log-det 0."""
    kind = _lib.PERMUTE
    in_place_ok = True       # the kernel reads a whole row before it writes it; between chained couplings the
                             # permutation is folded into their index lists and costs nothing

    def _meta_tail(self, dim):
        """host copies of the index arrays (stb_layer.perm_host / perm_inv_host)"""
        return [int(v) for v in self.permutation.tolist()] + [int(v) for v in self.inverse_permutation.tolist()]

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self.permutation = torch.randperm(dim)
        self.inverse_permutation = torch.empty(dim).long()
        self.inverse_permutation[self.permutation] = torch.arange(dim)
        self._dev = {}

    def _params(self, dim, device):
        if dim != self.permutation.numel():
            raise ValueError(f'permutation of size {self.permutation.numel()} applied to dim {dim}')
        k = str(device)
        if k not in self._dev:
            self._dev[k] = [self.permutation.to(torch.int32).to(device).contiguous(),
                            self.inverse_permutation.to(torch.int32).to(device).contiguous()]
        return self._dev[k]

    def _torch(self, v, inverse):
        idx = (self.inverse_permutation if inverse else self.permutation).to(v.device)
        return v[..., idx], torch.zeros_like(v[..., :1])

    def log_diag_jacobian(self, x, y=None, **kwargs):
        return torch.eye(self.dim, device=x.device)[self.permutation.to(x.device)].diag().log().expand_as(x)


class Flip(Permute):
    """Reverses the order of the last dimension (permute.py:11-44, ``dims=[-1]`` only)."""

    def __init__(self, dims: List[int] = [-1]):
        ElementwiseTransform.__init__(self)
        if list(dims) != [-1]:
            raise NotImplementedError('Flip along dims other than the last is not built')
        self.dims = dims
        self.dim = None
        self._dev = {}

    def _params(self, dim, device):
        if self.dim != dim:
            self.dim = dim
            self.permutation = torch.arange(dim - 1, -1, -1)
            self.inverse_permutation = self.permutation.clone()
            self._dev = {}
        return super()._params(dim, device)

    def _torch(self, v, inverse):
        return torch.flip(v, [-1]), torch.zeros_like(v[..., :1])

    def log_diag_jacobian(self, x, y=None, **kwargs):
        return torch.eye(x.shape[-1], device=x.device).flip([-1]).diag().log().expand_as(x)


class Sigmoid(_Pointwise):
    """y = clamp(sigmoid(x), tiny, 1 - eps); log-diag = -softplus(-x) - softplus(x) (sigmoid.py:9-41)."""
    kind = _lib.SIGMOID

    def __init__(self, **kwargs):
        super().__init__()

    @staticmethod
    def _sig(x):
        fi = torch.finfo(x.dtype)
        return torch.clamp(torch.sigmoid(x), min=fi.tiny, max=1. - fi.eps)

    @staticmethod
    def _logit(y):
        fi = torch.finfo(y.dtype)
        y = y.clamp(min=fi.tiny, max=1. - fi.eps)
        return y.log() - (-y).log1p()

    @staticmethod
    def _ld(u):
        return (-F.softplus(-u) - F.softplus(u)).sum(-1, keepdim=True)

    def _torch(self, v, inverse):
        if not inverse:
            return self._sig(v), self._ld(v)
        x = self._logit(v)
        return x, -self._ld(x)

    def log_diag_jacobian(self, x, y=None, **kwargs):
        return -F.softplus(-x) - F.softplus(x)


class Logit(Sigmoid):
    """Inverse of ``Sigmoid`` (sigmoid.py:44-56)."""
    kind = _lib.LOGIT

    def _torch(self, v, inverse):
        if not inverse:
            y = self._logit(v)
            return y, -self._ld(y)
        return self._sig(v), self._ld(v)

    def log_diag_jacobian(self, x, y=None, **kwargs):
        return F.softplus(-y) + F.softplus(y)
