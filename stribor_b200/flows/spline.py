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
from ..util.splines import unconstrained_cubic_spline, unconstrained_rational_quadratic_spline
from ._native import build_meta, fusable, needs_autograd, row_params_from_net

__all__ = ['Spline']


class Spline(ElementwiseTransform):
    """Same hooks as the reference class (spline.py:39-143): ``self.spline`` is the functional spline,
    ``_get_params(latent)`` returns ``(width, height, derivative)``, and every method funnels into
    ``forward_and_log_diag_jacobian`` -- so a subclass that overrides one of them (the reference's own
    ``test_spline.py:43-46`` does) changes all of them.  When nothing is overridden the calls go to the fused
    kernels directly (same values, one launch, conditioner evaluated in the kernel)."""

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
            self.spline = unconstrained_rational_quadratic_spline
            self.derivative_dim = n_bins - 1
        elif spline_type == 'cubic':
            self.kind = _lib.CUBIC
            self.spline = unconstrained_cubic_spline
            self.derivative_dim = 2
        else:
            raise ValueError('spline_type must be either `quadratic` or `cubic`')
        self._default_spline = self.spline
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

    # -- reference hooks (spline.py:76-105) ---------------------------------------------------------
    def _get_params(self, latent=None):
        if latent is None or self.latent_net is None:
            return self.width, self.height, self.derivative
        params = self.latent_net(latent)
        params = params.view(*params.shape[:-1], self.dim, self.n_bins * 2 + self.derivative_dim)
        return (params[..., :self.n_bins], params[..., self.n_bins:2 * self.n_bins],
                params[..., 2 * self.n_bins:])

    def plain(self) -> bool:
        """True when no hook is overridden: the fused kernels may evaluate this transform."""
        cls = type(self)
        return (cls.forward_and_log_diag_jacobian is Spline.forward_and_log_diag_jacobian
                and cls._get_params is Spline._get_params and self.spline is self._default_spline)

    def chainable(self):
        return self.plain() and (self.latent_net is None or fusable(self.latent_net))

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

    def _kernel_ok(self, x, lat):
        """May the kernels evaluate the conditioner themselves for this call?"""
        return lat is None or (fusable(self.latent_net) and not needs_autograd(self, x, lat))

    def _run(self, x, latent, direction, want_ldj=True):
        """fused: y (+ summed log-derivative) in one launch"""
        lat = latent if self.latent_net is not None else None
        if not self._kernel_ok(x, lat):
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
        if self.plain():
            return self._run(x, latent, _lib.FORWARD, False)[0]
        return self.forward_and_log_diag_jacobian(x, latent)[0]

    def inverse(self, y, latent=None, **kwargs):
        if self.plain():
            return self._run(y, latent, _lib.INVERSE, False)[0]
        return self.inverse_and_log_diag_jacobian(y, latent)[0]

    def forward_and_log_diag_jacobian(self, x, latent=None, *, reverse=False, **kwargs):
        lat = latent if self.latent_net is not None else None
        if self.plain() and self._kernel_ok(x, lat):
            d, lat = self._desc(x, latent)
            return run_layer_diag(d, x, lat, None, _lib.INVERSE if reverse else _lib.FORWARD)
        # the reference's composition (spline.py:101-105): differentiable in x, the learned parameters and --
        # through autograd over latent_net -- the conditioner; also the path of any nn.Module conditioner
        w, h, d = self._get_params(lat)
        return self.spline(x, w, h, d, inverse=reverse, lower=self.lower, upper=self.upper)

    def inverse_and_log_diag_jacobian(self, y, latent=None, **kwargs):
        return self.forward_and_log_diag_jacobian(y, latent, reverse=True)

    def forward_and_log_det_jacobian(self, x, latent=None, **kwargs):
        if self.plain():
            return self._run(x, latent, _lib.FORWARD)
        y, ld = self.forward_and_log_diag_jacobian(x, latent)
        return y, ld.sum(-1, keepdim=True)

    def inverse_and_log_det_jacobian(self, y, latent=None, **kwargs):
        if self.plain():
            return self._run(y, latent, _lib.INVERSE)       # the inverse map's own log-derivative
        x, ld = self.forward_and_log_diag_jacobian(y, latent, reverse=True)
        return x, ld.sum(-1, keepdim=True)

    def log_det_jacobian(self, x, y=None, latent=None, **kwargs):
        return self.forward_and_log_det_jacobian(x, latent)[1]

    def log_diag_jacobian(self, x, y=None, latent=None, **kwargs):
        return self.forward_and_log_diag_jacobian(x, latent)[1]
