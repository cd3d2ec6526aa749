"""Plain-data description of a flow ("spec") <-> modules.

A spec is a list of layer dicts holding CPU tensors (see oracle/coupling_flow_oracle.py for the
schema).  It is the neutral format the parity tests and the benchmark use to give the CUDA
modules and the CPU oracle the SAME weights; nothing here evaluates a flow.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from . import _lib
from .flows.affine import Affine
from .flows.coupling import ContinuousAffineCoupling, Coupling
from .flows.spline import Spline
from .flows.pointwise import Flip, Permute, Sigmoid, Logit
from .net.mlp import MLP
from .net.time_net import TimeLinear

_KIND_NAME = {_lib.AFFINE: 'affine', _lib.RQS: 'quadratic', _lib.CUBIC: 'cubic'}


def _mlp_from_spec(net: Dict) -> MLP:
    ws = net['weights']
    m = MLP(ws[0].shape[1], [w.shape[0] for w in ws[:-1]], ws[-1].shape[0],
            activation=net.get('activation', 'Tanh'), final_activation=net.get('final_activation'))
    lin = [l for l in m.net if isinstance(l, nn.Linear)]
    for l, w, b in zip(lin, ws, net['biases']):
        l.weight.data = w.detach().clone().float()
        l.bias.data = b.detach().clone().float()
    return m


def _transform_from_spec(tr: Dict):
    net = _mlp_from_spec(tr['net']) if tr.get('net') is not None else None
    if tr['kind'] == 'affine':
        f = Affine(tr['dim'], latent_net=net)
        if net is None:
            f.log_scale.data = tr['params'][0].detach().clone().float()
            f.shift.data = tr['params'][1].detach().clone().float()
        return f
    f = Spline(tr['dim'], tr['n_bins'], latent_net=net, lower=tr['lower'], upper=tr['upper'],
               spline_type=tr['kind'])
    if net is None:
        for name, p in zip(('width', 'height', 'derivative'), tr['params']):
            getattr(f, name).data = p.detach().clone().float()
    return f


def layer_from_spec(layer: Dict) -> nn.Module:
    typ = layer['type']
    if typ == 'elementwise':
        return _transform_from_spec(layer['transform'])
    if typ == 'coupling':
        return Coupling(_transform_from_spec(layer['transform']), mask=layer['mask'])
    if typ == 'sigmoid':
        return Sigmoid()
    if typ == 'logit':
        return Logit()
    if typ == 'flip':
        return Flip()
    if typ == 'permute':
        f = Permute(len(layer['perm']))
        f.permutation = torch.as_tensor(layer['perm']).long()
        f.inverse_permutation = torch.empty_like(f.permutation)
        f.inverse_permutation[f.permutation] = torch.arange(len(layer['perm']))
        return f
    if typ == 'cont_affine_coupling':
        tn = TimeLinear(layer['time_scale'].shape[-1])
        tn.scale.data = layer['time_scale'].detach().clone().float()
        return ContinuousAffineCoupling(_mlp_from_spec(layer['net']), tn, layer['mask'],
                                        concatenate_time=layer.get('concatenate_time', True))
    raise ValueError(typ)


def layers_from_spec(spec: List[Dict]) -> List[nn.Module]:
    return [layer_from_spec(l) for l in spec]


def _mlp_to_spec(m: MLP) -> Dict:
    lin = [l for l in m.net if isinstance(l, nn.Linear)]
    others = [type(l).__name__ for l in m.net if not isinstance(l, nn.Linear)]
    final = others[-1] if not isinstance(list(m.net)[-1], nn.Linear) else None
    act = others[0] if (others and not (final and len(others) == 1)) else 'Tanh'
    return {'weights': [l.weight.detach().cpu().clone() for l in lin],
            'biases': [l.bias.detach().cpu().clone() for l in lin],
            'activation': act, 'final_activation': final}


def _transform_to_spec(f) -> Dict:
    tr = {'kind': _KIND_NAME[f.kind], 'dim': f.dim, 'n_bins': f.n_bins,
          'net': _mlp_to_spec(f.latent_net) if f.latent_net is not None else None}
    if isinstance(f, Spline):
        tr['lower'], tr['upper'] = f.lower, f.upper
        if f.latent_net is None:
            tr['params'] = [p.detach().cpu().clone() for p in (f.width, f.height, f.derivative)]
    elif f.latent_net is None:
        tr['params'] = [f.log_scale.detach().cpu().clone(), f.shift.detach().cpu().clone()]
    return tr


def spec_from_layers(layers) -> List[Dict]:
    out = []
    for f in layers:
        if isinstance(f, Coupling):
            out.append({'type': 'coupling', 'mask': f.mask_name, 'transform': _transform_to_spec(f.transform)})
        elif isinstance(f, ContinuousAffineCoupling):
            out.append({'type': 'cont_affine_coupling', 'mask': f.mask_name,
                        'concatenate_time': bool(f.concatenate_time), 'net': _mlp_to_spec(f.latent_net),
                        'time_scale': f.time_net.scale.detach().cpu().clone()})
        elif isinstance(f, Flip):
            out.append({'type': 'flip'})
        elif isinstance(f, Permute):
            out.append({'type': 'permute', 'perm': f.permutation.tolist()})
        elif isinstance(f, Logit):
            out.append({'type': 'logit'})
        elif isinstance(f, Sigmoid):
            out.append({'type': 'sigmoid'})
        elif isinstance(f, (Affine, Spline)):
            out.append({'type': 'elementwise', 'transform': _transform_to_spec(f)})
        else:
            raise ValueError(f'no spec for {type(f).__name__}')
    return out
