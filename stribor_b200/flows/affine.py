"""Element-wise affine flow (reference: stribor/flows/affine.py:13-123).

``y = x * exp(log_scale) + shift`` with parameters that are learned ``[1, dim]`` tensors, fixed
``scale`` / ``shift`` values, or the output of ``latent_net(latent)`` split as
``[log_scale(dim) | shift(dim)]`` (affine.py:59-67).
"""
from __future__ import annotations

from numbers import Number

import torch
import torch.nn as nn

from .. import _lib
from ..flow import ElementwiseTransform, run_layer, run_layer_diag
from ._native import build_meta, fusable, needs_autograd, row_params_from_net

__all__ = ['Affine']


class Affine(ElementwiseTransform):
    kind = _lib.AFFINE
    n_bins = 0

    def __init__(self, dim: int, *, latent_net=None, scale=None, shift=None, **kwargs):
        super().__init__()
        self.dim = dim
        self.latent_net = latent_net
        if latent_net is None:
            if scale is None:
                self.log_scale = nn.Parameter(torch.empty(1, dim))
                self.shift = nn.Parameter(torch.empty(1, dim))
                nn.init.xavier_uniform_(self.log_scale)
                nn.init.xavier_uniform_(self.shift)
            else:
                if isinstance(scale, Number):
                    scale = torch.as_tensor([scale], dtype=torch.get_default_dtype())
                    shift = torch.as_tensor([shift], dtype=torch.get_default_dtype())
                scale, shift = scale.float(), shift.float()
                assert torch.all(scale > 0), '`scale` mush have positive values'
                # fixed values: buffers (so .to(device) moves them) kept out of the state-dict
                self.register_buffer('log_scale', scale.log(), persistent=False)
                self.register_buffer('shift', shift.clone(), persistent=False)

    def plain(self) -> bool:
        """True when a subclass overrides none of the transform's methods (nor adds the reference's ``_get_params``
        hook, affine.py:59-67): only then may the fused kernels stand in for them."""
        cls = type(self)
        own = ('forward', 'inverse', 'log_det_jacobian', 'forward_and_log_det_jacobian', 'inverse_and_log_det_jacobian',
               'log_diag_jacobian')
        return all(getattr(cls, n) is getattr(Affine, n) for n in own) and not hasattr(cls, '_get_params')

    def chainable(self):
        return self.plain() and (self.latent_net is None or fusable(self.latent_net))

    def params_per_dim(self):
        return 2

    def const_out(self, device=None):
        """[log_scale(dim) | shift(dim)] when there is no latent_net."""
        ls = self.log_scale.reshape(-1).expand(self.dim) if self.log_scale.numel() == 1 else self.log_scale.reshape(-1)
        sh = self.shift.reshape(-1).expand(self.dim) if self.shift.numel() == 1 else self.shift.reshape(-1)
        out = torch.cat([ls, sh.to(ls.device)]).contiguous()
        if device is not None and out.device != device and not isinstance(self.log_scale, nn.Parameter):
            out = out.to(device)          # fixed scale / shift given on another device (the reference keeps them
        return out                        # as plain attributes: affine.py:50-57)

    def fmeta(self):
        return [0., 1.] * 3

    def describe(self, dim, latent_dim, device):
        """Stand-alone element-wise layer: every dim transformed, network input = latent."""
        net = self.latent_net
        meta, params = build_meta(self.kind, dim, latent_dim if net is not None else 0, 0, 0, self.n_bins,
                                  1, 0, net, 0)
        if net is None:
            params = [self.const_out(device)]
        return {'meta': meta, 'fmeta': self.fmeta(), 'mask': None, 'params': list(params), 'packed': None}

    def _run(self, x, latent, direction, want_ldj=True):
        lat = latent if self.latent_net is not None else None
        if lat is not None and (needs_autograd(self, x, lat) or not fusable(self.latent_net)):
            return run_layer(self._describe_rows(x, lat), x, None, None, direction, want_ldj)
        d = self.describe(x.shape[-1], 0 if lat is None else lat.shape[-1], x.device)
        return run_layer(d, x, lat, None, direction, want_ldj)

    def _describe_rows(self, x, lat):
        """Training path: conditioner through autograd, its output handed over per row."""
        lead = x.shape[:-1]
        if lat.shape[:-1] != lead:
            lat = lat.expand(*lead, lat.shape[-1])
        prm = row_params_from_net(self.latent_net, lat.reshape(-1, lat.shape[-1]))
        meta, _ = build_meta(self.kind, x.shape[-1], 0, 0, 0, self.n_bins, 1, 0, None, 0)
        meta[13] = 1
        return {'meta': meta, 'fmeta': self.fmeta(), 'mask': None, 'params': [prm.contiguous()], 'packed': None}

    def forward(self, x, latent=None, **kwargs):
        return self._run(x, latent, _lib.FORWARD, False)[0]

    def inverse(self, y, latent=None, **kwargs):
        return self._run(y, latent, _lib.INVERSE, False)[0]

    def forward_and_log_det_jacobian(self, x, latent=None, *, reverse=False, **kwargs):
        if reverse:      # affine.py:97-109: the inverse map together with the FORWARD log-det (+sum log_scale)
            y, ldj = self._run(x, latent, _lib.INVERSE)
            return y, -ldj
        return self._run(x, latent, _lib.FORWARD)

    def inverse_and_log_det_jacobian(self, y, latent=None, **kwargs):
        return self._run(y, latent, _lib.INVERSE)

    def log_det_jacobian(self, x, y=None, latent=None, **kwargs):
        return self._run(x, latent, _lib.FORWARD)[1]

    def log_diag_jacobian(self, x, y=None, latent=None, **kwargs):
        lat = latent if self.latent_net is not None else None
        d = self.describe(x.shape[-1], 0 if lat is None else lat.shape[-1], x.device)
        return run_layer_diag(d, x, lat, None, _lib.FORWARD)[1]
