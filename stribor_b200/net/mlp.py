"""Latent network (reference: stribor/net/mlp.py:6-65).

Same constructor, same ``self.net = nn.Sequential(...)`` layout, hence the same state-dict
keys (``net.{0,2,4,...}.{weight,bias}``) so reference weights load directly.  Inside a
coupling layer the module is never *called*: its weights are handed to the fused CUDA kernel
(``describe()``).  Calling it directly evaluates the plain ``nn.Sequential``.
"""
from __future__ import annotations

from typing import Callable, List, Union

import torch.nn as nn

from .. import _epoch, _lib

__all__ = ['MLP']


class MLP(_epoch.Tracked, nn.Module):
    def __init__(self, in_dim: int, hidden_dims: List[int], out_dim: int,
                 activation: Union[str, Callable] = 'Tanh', final_activation: str = None,
                 nn_linear_wrapper_func: Callable = None, **kwargs):
        super().__init__()
        act = getattr(nn, activation)() if isinstance(activation, str) else activation
        fact = getattr(nn, final_activation)() if isinstance(final_activation, str) else final_activation
        wrap = nn_linear_wrapper_func if nn_linear_wrapper_func else (lambda m: m)
        self._wrapped = nn_linear_wrapper_func is not None

        widths = [in_dim] + list(hidden_dims) + [out_dim]
        mods = []
        for i in range(len(widths) - 1):
            lin = nn.Linear(widths[i], widths[i + 1])
            if i > 0:
                mods.append(act)          # one shared activation instance, as in the reference
                lin = wrap(lin)
            mods.append(lin)
        mods[-1].bias.data.fill_(0.0)     # mlp.py:53
        if fact is not None:
            mods.append(fact)
        self.net = nn.Sequential(*mods)

    def forward(self, x, **kwargs):
        return self.net(x)

    # -- what the fused kernels need ---------------------------------------------------------
    def describe(self):
        """-> (dims, activation enum, final-activation enum, [W0, b0, W1, b1, ...]).
        Raises NotImplementedError for configurations the kernels do not implement."""
        if self._wrapped:
            raise NotImplementedError('nn_linear_wrapper_func conditioners are not fused')
        mods = list(self.net)
        linears = [m for m in mods if isinstance(m, nn.Linear)]
        others = [m for m in mods if not isinstance(m, nn.Linear)]
        if len(linears) > _lib.STB_MAX_LINEAR:
            raise NotImplementedError(f'more than {_lib.STB_MAX_LINEAR} linear layers')
        final = None
        if not isinstance(mods[-1], nn.Linear):
            final = type(mods[-1]).__name__
            others = others[:-1]
        act_names = {type(m).__name__ for m in others}
        if len(act_names) > 1:
            raise NotImplementedError('mixed activations')
        act = act_names.pop() if act_names else None
        for name in (act, final):
            if name not in _lib.ACTIVATIONS:
                raise NotImplementedError(f'activation {name} is not built into the kernels')
        for m in others:
            if isinstance(m, nn.LeakyReLU) and m.negative_slope != 0.01:
                raise NotImplementedError('LeakyReLU slope other than 0.01')
            if isinstance(m, nn.ELU) and m.alpha != 1.0:
                raise NotImplementedError('ELU alpha other than 1')
            if isinstance(m, nn.Softplus) and (m.beta != 1 or m.threshold != 20):
                raise NotImplementedError('Softplus with non-default beta / threshold')
            if isinstance(m, nn.GELU) and m.approximate != 'none':
                raise NotImplementedError('approximate GELU')
        dims = [linears[0].in_features] + [l.out_features for l in linears]
        params = []
        for l in linears:
            if l.bias is None:
                raise NotImplementedError('Linear without bias')
            params += [l.weight, l.bias]
        return dims, _lib.ACTIVATIONS[act], _lib.ACTIVATIONS[final], params
