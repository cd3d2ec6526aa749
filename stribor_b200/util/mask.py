"""Coupling masks by name (reference: stribor/util/mask.py:6-57).

A mask generator maps ``dim`` to a float 0/1 vector: 1 = the coordinate passes through and
conditions the transform, 0 = the coordinate is transformed.  ``random_half`` is drawn ONCE
per (generator, dim) and then frozen -- the reference redraws it from numpy's global RNG on
every call, which makes even ``forward`` and ``log_det_jacobian`` of one layer disagree
(hard part 7); freezing is a deliberate, documented deviation.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ['get_mask']

_ORDERED = {'ordered_right_half': False, 'ordered_0': False, 'ordered_left_half': True, 'ordered_1': True}
_PARITY = {'parity_even': False, 'parity_odd': True}


def _n_zero(dim: int, ratio: float = 0.5) -> int:
    return int(np.clip(int(dim * ratio), 1, dim - 1))


def get_mask(mask: str):
    if mask == 'none':
        return lambda dim: torch.zeros(1)
    if mask in _ORDERED:
        flip = _ORDERED[mask]

        def ordered(dim: int) -> torch.Tensor:
            if dim == 1:
                return torch.ones(1)
            m = torch.ones(dim)
            m[:_n_zero(dim)] = 0.
            return 1. - m if flip else m
        return ordered
    if mask in _PARITY:
        flip = _PARITY[mask]

        def parity(dim: int) -> torch.Tensor:
            if dim == 1:
                return torch.ones(1)
            m = torch.ones(dim)
            m[0::2] = 0.
            return 1. - m if flip else m
        return parity
    if mask == 'random_half':
        frozen = {}

        def random_half(dim: int) -> torch.Tensor:
            if dim == 1:
                return torch.ones(1)
            if dim not in frozen:
                m = np.zeros(dim, dtype=np.float32)
                m[np.random.choice(np.arange(dim), _n_zero(dim), replace=False)] = 1.
                frozen[dim] = torch.from_numpy(m)
            return frozen[dim].clone()
        return random_half
    raise NotImplementedError()
