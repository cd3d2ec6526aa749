"""Deterministic case table + portable weights for the golden fixtures.

Shared by ``make_golden.py`` (which runs the UNMODIFIED reference in the build container)
and by the tests (which rebuild the same weights/inputs without the reference).  All
numbers come from ``numpy.random.RandomState`` (legacy MT19937 stream, stable across
numpy versions), never from torch's RNG, so the fixtures only need to store outputs.
"""
from __future__ import annotations

import numpy as np
import torch


def _rs(seed):
    return np.random.RandomState(seed)


def linear_params(rs, out_dim, in_dim):
    """nn.Linear-like init: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    b = 1.0 / np.sqrt(max(in_dim, 1))
    w = rs.uniform(-b, b, size=(out_dim, in_dim)).astype(np.float32)
    bias = rs.uniform(-b, b, size=(out_dim,)).astype(np.float32)
    return torch.from_numpy(w), torch.from_numpy(bias)


def mlp_spec(rs, in_dim, hidden, out_dim, activation='Tanh', final_activation=None):
    dims = [in_dim] + list(hidden) + [out_dim]
    ws, bs = [], []
    for i in range(len(dims) - 1):
        w, b = linear_params(rs, dims[i + 1], dims[i])
        ws.append(w)
        bs.append(b)
    return {'weights': ws, 'biases': bs, 'activation': activation,
            'final_activation': final_activation}


def n_params(kind, n_bins):
    return {'affine': 2, 'quadratic': 3 * n_bins - 1, 'cubic': 2 * n_bins + 2}[kind]


def coupling_spec(rs, kind, dim, hidden, mask, n_bins=0, lower=0., upper=1., latent_dim=0,
                  activation='Tanh'):
    tr = {'kind': kind, 'dim': dim, 'n_bins': n_bins, 'lower': lower, 'upper': upper,
          'net': mlp_spec(rs, dim + latent_dim, hidden, dim * n_params(kind, n_bins), activation)}
    return {'type': 'coupling', 'mask': mask, 'transform': tr}


def elementwise_spec(rs, kind, dim, hidden, n_bins=0, lower=0., upper=1., latent_dim=0):
    tr = {'kind': kind, 'dim': dim, 'n_bins': n_bins, 'lower': lower, 'upper': upper}
    if latent_dim:
        tr['net'] = mlp_spec(rs, latent_dim, hidden, dim * n_params(kind, n_bins))
    else:
        tr['net'] = None
        if kind == 'affine':
            shapes = [(1, dim), (1, dim)]
        else:
            shapes = [(dim, n_bins), (dim, n_bins), (dim, n_params(kind, n_bins) - 2 * n_bins)]
        tr['params'] = [torch.from_numpy(rs.uniform(-1, 1, size=s).astype(np.float32)) for s in shapes]
    return {'type': 'elementwise', 'transform': tr}


def cont_affine_spec(rs, dim, hidden, mask, latent_dim=0, concatenate_time=True):
    in_dim = dim + latent_dim + (1 if concatenate_time else 0)
    return {'type': 'cont_affine_coupling', 'mask': mask, 'concatenate_time': concatenate_time,
            'net': mlp_spec(rs, in_dim, hidden, 2 * dim),
            'time_scale': torch.from_numpy(rs.uniform(-0.5, 0.5, size=(1, 2 * dim)).astype(np.float32))}


ALT = ('ordered_right_half', 'ordered_left_half')


def build_case(name):
    """-> dict(spec=[layers], inputs={x, latent?, t?, t0?}, ops=[...])  (all torch fp32, CPU)."""
    c = CASES[name]
    return c()


def _x(rs, shape, scale=1.0, shift=0.0, uniform=False):
    v = rs.uniform(0, 1, size=shape) if uniform else rs.standard_normal(size=shape)
    return torch.from_numpy((v * scale + shift).astype(np.float32))


SHAPES = [(1, 1), (2, 10), (10, 2), (7, 4, 5)]
CASES = {}


def _reg(name):
    def deco(fn):
        CASES[name] = fn
        return fn
    return deco


def _mk_affine_coupling(shape, latent_dim, seed):
    def fn():
        rs = _rs(seed)
        dim = shape[-1]
        spec = [coupling_spec(rs, 'affine', dim, [13], 'ordered_left_half', latent_dim=latent_dim)]
        inp = {'x': _x(rs, shape)}
        if latent_dim:
            inp['latent'] = _x(rs, shape[:-1] + (latent_dim,))
        return {'spec': spec, 'inputs': inp, 'ops': ['forward_ldj', 'inverse_ldj']}
    return fn


def _mk_cont_affine(shape, latent_dim, seed):
    def fn():
        rs = _rs(seed)
        dim = shape[-1]
        spec = [cont_affine_spec(rs, dim, [13], 'ordered_left_half', latent_dim=latent_dim)]
        inp = {'x': _x(rs, shape), 't': _x(rs, shape[:-1] + (1,))}
        if latent_dim:
            inp['latent'] = _x(rs, shape[:-1] + (latent_dim,))
        return {'spec': spec, 'inputs': inp, 'ops': ['forward_ldj', 'inverse_ldj']}
    return fn


def _mk_spline(shape, n_bins, latent_dim, kind, seed):
    def fn():
        rs = _rs(seed)
        dim = shape[-1]
        spec = [elementwise_spec(rs, kind, dim, [12], n_bins=n_bins, lower=0., upper=2.,
                                 latent_dim=latent_dim)]
        # mostly inside [0, 2], some in the identity tails, plus the exact end points
        x = _x(rs, shape, scale=2.6, shift=-0.3, uniform=True)
        flat = x.view(-1)
        if flat.numel() >= 4:
            flat[0], flat[1] = 0.0, 2.0
        inp = {'x': x}
        if latent_dim:
            inp['latent'] = _x(rs, shape[:-1] + (latent_dim,))
        return {'spec': spec, 'inputs': inp, 'ops': ['forward_ldj', 'inverse_ldj']}
    return fn


_seed = 1000
for _s in SHAPES:
    for _l in (0, 1, 13):
        _seed += 1
        CASES[f'affine_coupling_{"x".join(map(str, _s))}_l{_l}'] = _mk_affine_coupling(_s, _l, _seed)
        _seed += 1
        CASES[f'cont_affine_{"x".join(map(str, _s))}_l{_l}'] = _mk_cont_affine(_s, _l, _seed)
for _s in SHAPES:
    for _k in (1, 3, 10):
        for _l in (0, 13):
            for _kind in ('quadratic', 'cubic'):
                _seed += 1
                CASES[f'spline_{_kind}_{"x".join(map(str, _s))}_k{_k}_l{_l}'] = _mk_spline(_s, _k, _l, _kind, _seed)


@_reg('readme_affine_2d')
def _c1():
    """BASELINE.json configs[0]: README 2-D flow (README.md:59-82 minus the CNF layer)."""
    rs = _rs(11)
    spec = [coupling_spec(rs, 'affine', 2, [64], 'ordered_right_half')]
    return {'spec': spec, 'inputs': {'x': _x(rs, (10, 2), uniform=True)},
            'ops': ['forward_ldj', 'inverse_ldj', 'log_prob']}


def _mk_flow(kind, dim, hidden, n_layers, n_bins, rows, seed, masks=ALT, lower=-4., upper=4.,
             scale=1.0, ops=('forward_ldj', 'inverse_ldj', 'log_prob'), activation='Tanh'):
    def fn():
        rs = _rs(seed)
        spec = [coupling_spec(rs, kind, dim, hidden, masks[i % len(masks)], n_bins=n_bins,
                              lower=lower, upper=upper, activation=activation)
                for i in range(n_layers)]
        return {'spec': spec, 'inputs': {'x': _x(rs, (rows, dim), scale=scale)}, 'ops': list(ops)}
    return fn


# BASELINE.json configs[1..2] at fixture-sized batches
CASES['affine_d64_L8_h256x256'] = _mk_flow('affine', 64, [256, 256], 8, 0, 24, 21)
CASES['quadratic_d64_L8_k16_h64'] = _mk_flow('quadratic', 64, [64], 8, 16, 24, 22, scale=1.5)
CASES['cubic_d64_L8_k16_h64'] = _mk_flow('cubic', 64, [64], 8, 16, 24, 23, scale=1.5)
CASES['quadratic_d64_L2_k16_h256x256'] = _mk_flow('quadratic', 64, [256, 256], 2, 16, 16, 24, scale=1.5)
CASES['quadratic_d128_L2_k16_h64'] = _mk_flow('quadratic', 128, [64], 2, 16, 16, 25, scale=1.5)
# masks / odd sizes / other activations
CASES['quadratic_d5_parity'] = _mk_flow('quadratic', 5, [16], 4, 8, 33, 26,
                                         masks=('parity_even', 'parity_odd'), lower=-3., upper=3.)
CASES['cubic_d7_ordered'] = _mk_flow('cubic', 7, [16, 8], 3, 5, 33, 27, lower=-3., upper=3.)
CASES['cubic_d6_parity_relu'] = _mk_flow('cubic', 6, [16], 2, 4, 17, 28,
                                          masks=('parity_odd', 'parity_even'), lower=-2., upper=2.,
                                          activation='ReLU')
CASES['affine_d3_none_mask'] = _mk_flow('affine', 3, [8], 2, 0, 9, 29, masks=('none',))
CASES['quadratic_d1_none_mask'] = _mk_flow('quadratic', 1, [8], 2, 6, 19, 30, masks=('none',),
                                            lower=-3., upper=3.)
CASES['quadratic_d16_default_box'] = _mk_flow('quadratic', 16, [32], 3, 16, 40, 31, lower=0., upper=1.,
                                               scale=0.6)
CASES['cubic_d16_default_box'] = _mk_flow('cubic', 16, [32], 3, 16, 40, 32, lower=0., upper=1.,
                                           scale=0.6)


@_reg('sigmoid_cubic_logit_d2')
def _c_sandwich():
    """test_normalizing_flow.py:8-34 minus the CNF layer: affine coupling, Flip, Sigmoid,
    cubic-spline coupling on the default [0, 1] box, Logit."""
    rs = _rs(61)
    spec = [coupling_spec(rs, 'affine', 2, [13], 'ordered_1'), {'type': 'flip'}, {'type': 'sigmoid'},
            coupling_spec(rs, 'cubic', 2, [13], 'ordered_0', n_bins=3), {'type': 'logit'}]
    return {'spec': spec, 'inputs': {'x': _x(rs, (3, 4, 2))}, 'ops': ['forward_ldj', 'inverse_ldj', 'log_prob']}


@_reg('permute_quadratic_d16')
def _c_perm():
    rs = _rs(62)
    spec = []
    for i in range(3):
        spec.append(coupling_spec(rs, 'quadratic', 16, [32], 'ordered_right_half', n_bins=8, lower=-3., upper=3.))
        spec.append({'type': 'permute', 'perm': rs.permutation(16).tolist()} if i % 2 == 0 else {'type': 'flip'})
    return {'spec': spec, 'inputs': {'x': _x(rs, (29, 16))}, 'ops': ['forward_ldj', 'inverse_ldj', 'log_prob']}


@_reg('sigmoid_logit_only_d5')
def _c_sig():
    rs = _rs(63)
    return {'spec': [{'type': 'sigmoid'}, {'type': 'flip'}, {'type': 'logit'}, {'type': 'sigmoid'}],
            'inputs': {'x': _x(rs, (11, 5), scale=3.0)}, 'ops': ['forward_ldj', 'inverse_ldj_unit']}


@_reg('neural_flow_d16_L4')
def _c4():
    """BASELINE.json configs[3] at fixture size: 4x ContinuousAffineCoupling, dim 16."""
    rs = _rs(41)
    spec = [cont_affine_spec(rs, 16, [64], ('ordered_0', 'ordered_1')[i % 2]) for i in range(4)]
    return {'spec': spec,
            'inputs': {'x': _x(rs, (3, 5, 16)), 't': _x(rs, (3, 5, 1), uniform=True),
                       't0': _x(rs, (3, 5, 1), uniform=True)},
            'ops': ['neural_flow', 'neural_flow_t0']}


@_reg('neural_flow_d2_notime')
def _c4b():
    """test_neural_flow.py:8-14: concatenate_time=False (TimeLinear(2*dim) here)."""
    rs = _rs(42)
    spec = [cont_affine_spec(rs, 2, [32], 'ordered_0', concatenate_time=False)]
    return {'spec': spec,
            'inputs': {'x': _x(rs, (10, 4, 2)), 't': _x(rs, (10, 4, 1)), 't0': _x(rs, (10, 4, 1))},
            'ops': ['neural_flow', 'neural_flow_t0']}


# gradient cases (NLL training step): loss = -log_prob(y).mean()
CASES['grad_quadratic_d8_L2'] = _mk_flow('quadratic', 8, [16], 2, 8, 12, 51, lower=-3., upper=3.,
                                          ops=('log_prob', 'nll_grad'))
CASES['grad_affine_d8_L2'] = _mk_flow('affine', 8, [16, 16], 2, 0, 12, 52, ops=('log_prob', 'nll_grad'))
CASES['grad_cubic_d8_L2'] = _mk_flow('cubic', 8, [16], 2, 8, 12, 53, lower=-3., upper=3.,
                                      ops=('log_prob', 'nll_grad'))
CASES['grad_quadratic_d64_L2_k16'] = _mk_flow('quadratic', 64, [64], 2, 16, 8, 54,
                                               ops=('log_prob', 'nll_grad'), scale=1.5)


# ---- bin-index fixtures (tests/golden/make_golden_bins.py -> reference_bins.npz) -------------------------------
# Per spline layer the reference's own searchsorted results are recorded: forward search on the layer input,
# inverse search on the layer output and the forward re-search on the recovered point.  Besides every CASES
# entry that contains a spline, these larger batches exercise the tensor-core kernels' search on many rows
# (partial last tile included: 300 = 256 + 44 rows on the 256-row kernel, 128 + 128 + 44 on the 128-row one).
BIN_CASES = {
    'bins_quadratic_d64_k16_h64': _mk_flow('quadratic', 64, [64], 2, 16, 300, 71, scale=1.5),
    'bins_cubic_d64_k16_h64': _mk_flow('cubic', 64, [64], 2, 16, 300, 72, scale=1.5),
    'bins_quadratic_d128_k16_h64': _mk_flow('quadratic', 128, [64], 2, 16, 300, 73, scale=1.5),
    'bins_cubic_d128_k16_h64': _mk_flow('cubic', 128, [64], 2, 16, 300, 74, scale=1.5),
    'bins_quadratic_d48_parity': _mk_flow('quadratic', 48, [64], 2, 16, 200, 75, masks=('parity_even', 'parity_odd'),
                                          lower=-2., upper=2., scale=1.5),
}


def spline_cases():
    """names (CASES and BIN_CASES) whose spec holds at least one spline layer"""
    out = []
    for name, fn in list(CASES.items()) + list(BIN_CASES.items()):
        spec = fn()['spec']
        if any(l.get('transform', {}).get('kind') in ('quadratic', 'cubic') for l in spec):
            out.append(name)
    return out


def build_bin_case(name):
    return (CASES.get(name) or BIN_CASES[name])()
