"""Element-wise spline flow (reference: stribor/flows/spline.py:11-143).

Monotone rational-quadratic ('quadratic', util/rational_quadratic_spline.py) or cubic
('cubic', util/cubic_spline.py) spline on ``[lower, upper]`` with identity tails.  Parameters per
dim are ``[widths(K) | heights(K) | derivatives(K-1 or 2)]`` (spline.py:76-87), learned or
produced by ``latent_net(latent)``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib
from ..flow import ElementwiseTransform, run_layer, run_layer_diag
from ._native import build_meta, fusable, needs_autograd, row_params_from_net

__all__ = ['Spline']


class Spline(ElementwiseTransform):
    def __init__(self, dim: int, n_bins: int, latent_net=None, lower=0, upper=1,
                 spline_type='cubic', **kwargs):
        super().__init__()
        self.lower = lower
        self.upper = upper
        self.dim = dim
        self.n_bins = n_bins
        self.latent_net = latent_net
        if spline_type == 'quadratic':
            self.kind = _lib.RQS
            self.derivative_dim = n_bins - 1
        elif spline_type == 'cubic':
            self.kind = _lib.CUBIC
            self.derivative_dim = 2
        else:
            raise ValueError('spline_type must be either `quadratic` or `cubic`')
        self.spline_type = spline_type
        if self.latent_net is None:
            self.width = nn.Parameter(torch.empty(self.dim, n_bins))
            self.height = nn.Parameter(torch.empty(self.dim, n_bins))
            self.derivative = nn.Parameter(torch.empty(self.dim, self.derivative_dim))
            self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.width)
        nn.init.xavier_uniform_(self.height)
        nn.init.xavier_uniform_(self.derivative)

    def chainable(self):
        return self.latent_net is None or fusable(self.latent_net)

    def params_per_dim(self):
        return 2 * self.n_bins + self.derivative_dim

    def const_out(self):
        return torch.cat([self.width, self.height, self.derivative], -1).reshape(-1).contiguous()

    def fmeta(self):
        return [float(self.lower), float(self.upper)] * 3

    def describe(self, dim, latent_dim, device):
        net = self.latent_net
        meta, params = build_meta(self.kind, dim, latent_dim if net is not None else 0, 0, 0, self.n_bins,
                                  1, 0, net, 0)
        if net is None:
            params = [self.const_out()]
        return {'meta': meta, 'fmeta': self.fmeta(), 'mask': None, 'params': list(params), 'packed': None}

    def _desc(self, x, latent):
        lat = latent if self.latent_net is not None else None
        return self.describe(x.shape[-1], 0 if lat is None else lat.shape[-1], x.device), lat

    def _run(self, x, latent, direction, want_ldj=True):
        lat = latent if self.latent_net is not None else None
        if lat is not None and (needs_autograd(self, x, lat) or not fusable(self.latent_net)):
            lead = x.shape[:-1]
            if lat.shape[:-1] != lead:
                lat = lat.expand(*lead, lat.shape[-1])
            prm = row_params_from_net(self.latent_net, lat.reshape(-1, lat.shape[-1]))
            meta, _ = build_meta(self.kind, x.shape[-1], 0, 0, 0, self.n_bins, 1, 0, None, 0)
            meta[13] = 1
            d = {'meta': meta, 'fmeta': self.fmeta(), 'mask': None, 'params': [prm.contiguous()], 'packed': None}
            return run_layer(d, x, None, None, direction, want_ldj)
        d, lat = self._desc(x, latent)
        return run_layer(d, x, lat, None, direction, want_ldj)

    def forward(self, x, latent=None, **kwargs):
        return self._run(x, latent, _lib.FORWARD, False)[0]

    def inverse(self, y, latent=None, **kwargs):
        return self._run(y, latent, _lib.INVERSE, False)[0]

    def forward_and_log_diag_jacobian(self, x, latent=None, *, reverse=False, **kwargs):
        d, lat = self._desc(x, latent)
        return run_layer_diag(d, x, lat, None, _lib.INVERSE if reverse else _lib.FORWARD)

    def inverse_and_log_diag_jacobian(self, y, latent=None, **kwargs):
        return self.forward_and_log_diag_jacobian(y, latent, reverse=True)

    def forward_and_log_det_jacobian(self, x, latent=None, **kwargs):
        return self._run(x, latent, _lib.FORWARD)

    def inverse_and_log_det_jacobian(self, y, latent=None, **kwargs):
        return self._run(y, latent, _lib.INVERSE)       # the inverse map's own log-derivative

    def log_det_jacobian(self, x, y=None, latent=None, **kwargs):
        return self._run(x, latent, _lib.FORWARD)[1]

    def log_diag_jacobian(self, x, y=None, latent=None, **kwargs):
        return self.forward_and_log_diag_jacobian(x, latent)[1]
