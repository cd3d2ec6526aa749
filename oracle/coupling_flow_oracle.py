"""CPU oracle for stribor's coupling-flow hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement, on torch CPU tensors, of the algorithm the
reference (mbilos/stribor 0.2.0, pure PyTorch) runs for
``NormalizingFlow.log_prob / forward / inverse / sample`` over ``Coupling`` layers that
wrap ``Affine`` / ``Spline`` (rational-quadratic or cubic) with an ``MLP`` conditioner,
plus ``ContinuousAffineCoupling`` / ``NeuralFlow``.  Every function cites the reference
file:line it follows (paths relative to the reference checkout).

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- as the CHECKER (or the timed CPU baseline),
never as part of the product path.  ``stribor_b200`` itself never imports ``oracle``.

Parity pin: ``tests/golden/make_golden.py`` imports the UNMODIFIED reference in the build
container and stores its inputs/outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this file against every one of them, and against
the reference's own golden vector (``stribor/test/test_normalizing_flow.py:45-55``) and
exact mask vectors (``stribor/test/test_mask.py:4-30``).

Deliberate differences from the reference (all value-preserving):
* no boolean gather/scatter of the "inside" elements -- every element is evaluated and
  the identity tail is selected with ``torch.where`` (reference:
  ``util/rational_quadratic_spline.py:161-164,250``, ``util/cubic_spline.py:55-67``);
* the reference's accidental ``[N] < [N,1]`` broadcast in its domain re-check
  (``util/rational_quadratic_spline.py:167-178``) is O(N^2) and can never fire after the
  inside filter, so it is not restated;
* the conditioner is evaluated once per layer, not twice (``flow.py:35-47`` calls it in
  ``forward`` and again in ``log_det_jacobian`` with identical inputs).

The functions work in the dtype of their inputs (fp32, or fp64 when used as arbiter).
A flow is described by a plain "spec": a list of layer dicts, see ``layer_spec`` helpers
in ``stribor_b200`` (``module.export_spec()``) or ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# masks  (util/mask.py:6-57)
# --------------------------------------------------------------------------------------

def make_mask(name: str, dim: int) -> Tensor:
    """0/1 float vector; 1 = passes through and conditions, 0 = is transformed.

    util/mask.py:6-20 (name dispatch), :35-45 (ordered), :47-57 (parity), :22-23 (none).
    ``random_half`` (:25-33) draws from numpy's global RNG on every call in the
    reference; it is not restated here (non-deterministic).
    """
    if name == 'none':
        return torch.zeros(1)
    if dim == 1:
        if name in ('ordered_right_half', 'ordered_0', 'ordered_left_half', 'ordered_1',
                    'parity_even', 'parity_odd'):
            return torch.ones(1)
        raise NotImplementedError(name)
    if name in ('ordered_right_half', 'ordered_0', 'ordered_left_half', 'ordered_1'):
        n_zero = int(np.clip(int(dim * 0.5), 1, dim - 1))          # mask.py:40
        m = torch.ones(dim)
        m[:n_zero] = 0
        if name in ('ordered_left_half', 'ordered_1'):
            m = 1 - m
        return m
    if name in ('parity_even', 'parity_odd'):
        m = torch.ones(dim)
        m[::2] = 0                                                    # mask.py:52
        if name == 'parity_odd':
            m = 1 - m
        return m
    raise NotImplementedError(name)


# --------------------------------------------------------------------------------------
# conditioner  (net/mlp.py:46-58, net/time_net.py:18-28)
# --------------------------------------------------------------------------------------

_ACT = {
    'Tanh': torch.tanh, 'ReLU': torch.relu, 'Sigmoid': torch.sigmoid, 'ELU': F.elu,
    'Softplus': F.softplus, 'LeakyReLU': F.leaky_relu, 'SiLU': F.silu, 'GELU': F.gelu,
    'Identity': lambda v: v,
}


def mlp(net: Dict, z: Tensor) -> Tensor:
    """Linear -> (act -> Linear)* [-> final act]; weights are nn.Linear layout [out,in]."""
    act = _ACT[net.get('activation', 'Tanh')]
    ws, bs = net['weights'], net['biases']
    h = F.linear(z, ws[0].to(z.dtype), bs[0].to(z.dtype))
    for w, b in zip(ws[1:], bs[1:]):
        h = F.linear(act(h), w.to(z.dtype), b.to(z.dtype))
    fa = net.get('final_activation')
    if fa is not None:
        h = _ACT[fa](h)
    return h


# --------------------------------------------------------------------------------------
# bin search  (util/search_sorted.py:3-5)
# --------------------------------------------------------------------------------------

def bin_search(knots: Tensor, x: Tensor, eps: float = 1e-6) -> Tensor:
    """``sum(x >= knots) - 1`` with the LAST knot nudged up by ``eps`` first.

    The reference mutates its argument in place; the nudge never reaches a value that is
    gathered later on this path, so a copy is used here.
    """
    k = knots.clone()
    k[..., -1] += eps
    return torch.sum(x[..., None] >= k, dim=-1) - 1


def _take(t: Tensor, idx: Tensor) -> Tensor:
    return t.gather(-1, idx[..., None])[..., 0]


# --------------------------------------------------------------------------------------
# rational-quadratic spline  (util/rational_quadratic_spline.py)
# --------------------------------------------------------------------------------------

RQS_MIN_W = RQS_MIN_H = RQS_MIN_D = 1e-3


def rqs_knots(uw: Tensor, uh: Tensor, ud: Tensor, left, right, bottom, top
              ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Unconstrained params -> (cumwidths, widths, cumheights, heights, derivatives).

    rational_quadratic_spline.py:79-83 (edge-derivative padding constant),
    :101-107 (softmax / softplus + minimums), :180-192 (knots forced to the box, widths
    and heights RE-derived from the knots).
    """
    K = uw.shape[-1]
    if RQS_MIN_W * K > 1.0 or RQS_MIN_H * K > 1.0:
        raise ValueError('Minimal bin width too large for the number of bins')
    if ud.shape[-1] == K - 1:
        const = np.log(np.exp(1 - RQS_MIN_D) - 1)
        ud = F.pad(ud, pad=(1, 1))
        ud[..., 0] = const
        ud[..., -1] = const
    w = RQS_MIN_W + (1 - RQS_MIN_W * K) * F.softmax(uw, dim=-1)
    h = RQS_MIN_H + (1 - RQS_MIN_H * K) * F.softmax(uh, dim=-1)
    der = RQS_MIN_D + F.softplus(ud)

    def knots(v, lo, hi):
        lo_t = torch.ones(v.shape[:-1] + (1,), dtype=v.dtype) * lo
        hi_t = torch.ones(v.shape[:-1] + (1,), dtype=v.dtype) * hi
        c = F.pad(torch.cumsum(v, dim=-1), pad=(1, 0), mode='constant', value=0.0)
        c = (hi_t - lo_t) * c + lo_t
        c[..., 0, None] = lo_t
        c[..., -1, None] = hi_t
        return c, c[..., 1:] - c[..., :-1]

    cw, w = knots(w, left, right)
    ch, h = knots(h, bottom, top)
    return cw, w, ch, h, der


def rqs(x: Tensor, uw: Tensor, uh: Tensor, ud: Tensor, inverse: bool, lower, upper,
        left=None, right=None, bottom=None, top=None, return_bins: bool = False):
    """Element-wise monotone RQ spline with identity tails.

    x [..., d]; uw, uh [..., d, K]; ud [..., d, K-1] or [..., d, K+1].
    Returns (out [..., d], log|d out/d x| [..., d]) (+ bin index if asked).
    rational_quadratic_spline.py:55-64 (boxes), :71 (inside, inclusive both ends),
    :86-87 (identity tails), :194-248 (bin search, inverse root, forward map, log-deriv).
    """
    if all(v is not None for v in (left, right, top, bottom)):
        lower, upper = (bottom, top) if inverse else (left, right)
    else:
        left = bottom = lower
        right = top = upper
    shape = x.shape
    uw = uw.expand(*shape, -1)
    uh = uh.expand(*shape, -1)
    ud = ud.expand(*shape, -1)
    inside = (x >= lower) & (x <= upper)
    x_in = x
    # tail elements still flow through the arithmetic below (their result is discarded);
    # give them an in-box value so no NaN reaches autograd through torch.where.
    x = torch.where(inside, x, torch.zeros_like(x) + lower)

    cw, w, ch, h, der = rqs_knots(uw, uh, ud, left, right, bottom, top)
    K = w.shape[-1]
    idx = bin_search(ch if inverse else cw, x).clamp(0, K - 1)
    x_k, w_k = _take(cw, idx), _take(w, idx)
    y_k, h_k = _take(ch, idx), _take(h, idx)
    delta = _take(h / w, idx)
    d0, d1 = _take(der, idx), _take(der[..., 1:], idx)

    if inverse:
        dy = x - y_k
        s = d0 + d1 - 2 * delta
        a = dy * s + h_k * (delta - d0)
        b = h_k * d0 - dy * s
        c = -delta * dy
        disc = b.pow(2) - 4 * a * c
        if inside.any() and not bool((disc[inside] >= 0).all()):
            raise AssertionError('negative discriminant')          # :223
        root = (2 * c) / (-b - torch.sqrt(disc))
        out = root * w_k + x_k
        tt = root * (1 - root)
        den = delta + s * tt
        num = delta.pow(2) * (d1 * root.pow(2) + 2 * delta * tt + d0 * (1 - root).pow(2))
        ld = -torch.log(num) + 2 * torch.log(den)
    else:
        theta = (x - x_k) / w_k
        tt = theta * (1 - theta)
        num = h_k * (delta * theta.pow(2) + d0 * tt)
        den = delta + (d0 + d1 - 2 * delta) * tt
        out = y_k + num / den
        dnum = delta.pow(2) * (d1 * theta.pow(2) + 2 * delta * tt + d0 * (1 - theta).pow(2))
        ld = torch.log(dnum) - 2 * torch.log(den)

    out = torch.where(inside, out, x_in)
    ld = torch.where(inside, ld, torch.zeros_like(ld))
    if return_bins:
        return out, ld, torch.where(inside, idx, torch.full_like(idx, -1))
    return out, ld


# --------------------------------------------------------------------------------------
# cubic spline  (util/cubic_spline.py)
# --------------------------------------------------------------------------------------

CUB_MIN_W = CUB_MIN_H = 1e-2
CUB_EPS = 1e-5
CUB_QUAD_THRESHOLD = 1e-3


def _cbrt(v: Tensor) -> Tensor:
    return torch.sign(v) * torch.exp(torch.log(torch.abs(v)) / 3.0)     # cubic_spline.py:18-20


def cubic_coefficients(uw: Tensor, uh: Tensor, ul: Tensor, ur: Tensor):
    """Per-bin cubic ``a s^3 + b s^2 + c s + d`` on the unit box.

    cubic_spline.py:104-116 (softmax, cumsum, LAST knot := 1, widths NOT re-derived),
    :118-133 (monotone interior derivatives, sigmoid end derivatives), :135-138.
    ul, ur: [..., 1] unconstrained left / right end derivative.
    """
    K = uw.shape[-1]
    if CUB_MIN_W * K > 1.0 or CUB_MIN_H * K > 1.0:
        raise ValueError('Minimal bin width too large for the number of bins')
    w = CUB_MIN_W + (1 - CUB_MIN_W * K) * F.softmax(uw, dim=-1)
    cw = torch.cumsum(w, dim=-1)
    cw[..., -1] = 1
    cw = F.pad(cw, pad=(1, 0), mode='constant', value=0.0)
    h = CUB_MIN_H + (1 - CUB_MIN_H * K) * F.softmax(uh, dim=-1)
    ch = torch.cumsum(h, dim=-1)
    ch[..., -1] = 1
    ch = F.pad(ch, pad=(1, 0), mode='constant', value=0.0)

    s = h / w
    m1 = torch.min(torch.abs(s[..., :-1]), torch.abs(s[..., 1:]))
    m2 = 0.5 * (w[..., 1:] * s[..., :-1] + w[..., :-1] * s[..., 1:]) / (w[..., :-1] + w[..., 1:])
    inner = torch.min(m1, m2) * (torch.sign(s[..., :-1]) + torch.sign(s[..., 1:]))
    d_left = torch.sigmoid(ul) * 3 * s[..., 0][..., None]
    d_right = torch.sigmoid(ur) * 3 * s[..., -1][..., None]
    der = torch.cat([d_left, inner, d_right], dim=-1)

    a = (der[..., :-1] + der[..., 1:] - 2 * s) / w.pow(2)
    b = (3 * s - 2 * der[..., :-1] - der[..., 1:]) / w
    c = der[..., :-1]
    d = ch[..., :-1]
    return cw, ch, a, b, c, d


def cubic(x: Tensor, uw: Tensor, uh: Tensor, ud: Tensor, inverse: bool, lower, upper,
          return_bins: bool = False):
    """Element-wise monotone cubic spline with identity tails.

    x [..., d]; uw, uh [..., d, K]; ud [..., d, 2] = (left, right) end derivatives.
    cubic_spline.py:35-48 (inside / tails), :99-102 (normalise), :140-151 (bin search,
    gathers), :153-228 (inverse: one-root Cardano, three-root trigonometric, almost-
    quadratic override), :229-238 (forward), :240-245 (rescale).
    """
    shape = x.shape
    uw = uw.expand(*shape, -1)
    uh = uh.expand(*shape, -1)
    ul = ud[..., 0, None].expand(*shape, -1)
    ur = ud[..., 1, None].expand(*shape, -1)
    inside = (x >= lower) & (x <= upper)
    x_in = x
    x = torch.where(inside, x, torch.zeros_like(x) + lower)      # see rqs()
    left = bottom = lower
    right = top = upper

    u = (x - bottom) / (top - bottom) if inverse else (x - left) / (right - left)
    cw, ch, a, b, c, d = cubic_coefficients(uw, uh, ul, ur)
    K = uw.shape[-1]
    idx = bin_search(ch if inverse else cw, u).clamp(0, K - 1)
    a_k, b_k, c_k, d_k = _take(a, idx), _take(b, idx), _take(c, idx), _take(d, idx)
    x_l, x_r = _take(cw, idx), _take(cw, idx + 1)

    if inverse:
        b_ = (b_k / a_k) / 3.
        c_ = (c_k / a_k) / 3.
        d_ = (d_k - u) / a_k
        delta_1 = -b_.pow(2) + c_
        delta_2 = -c_ * b_ + d_
        delta_3 = b_ * d_ - c_.pow(2)
        disc = 4. * delta_1 * delta_3 - delta_2.pow(2)
        dep_1 = -2. * b_ * delta_1 + delta_2
        dep_2 = delta_1
        three = disc > 0

        # one real root (disc <= 0): Cardano                                   :175-180
        one = torch.ones_like(disc)          # placeholders keep unselected branches NaN-free
        sq = torch.sqrt(torch.where(three, one, -disc))
        p = _cbrt((-dep_1 + sq) / 2.)
        q = _cbrt((-dep_1 - sq) / 2.)
        one_root = (p + q) - b_ + x_l

        # three real roots (disc > 0): trigonometric form                      :184-213
        theta = torch.atan2(torch.sqrt(torch.where(three, disc, one)), -dep_1) / 3.
        cr1, cr2 = torch.cos(theta), torch.sin(theta)
        scale = 2 * torch.sqrt(torch.where(three, -dep_2, one))
        shift = -b_ + x_l
        r1 = cr1 * scale + shift
        r2 = (-0.5 * cr1 - 0.5 * math.sqrt(3) * cr2) * scale + shift
        r3 = (-0.5 * cr1 + 0.5 * math.sqrt(3) * cr2) * scale + shift
        ok = lambda r: ((x_l - CUB_EPS) < r) & (r < (x_r + CUB_EPS))
        # argsort(descending)[..., 0] on the 0/1 masks = first root whose mask is 1,
        # or root 1 when none is.
        three_root = torch.where(ok(r1), r1, torch.where(ok(r2), r2, torch.where(ok(r3), r3, r1)))

        out = torch.where(three, three_root, one_root)

        # |a| -> 0: quadratic formula override                                 :217-223
        quad = a_k.abs() < CUB_QUAD_THRESHOLD
        qc = d_k - u
        alpha = (-c_k + torch.sqrt(torch.where(quad, c_k.pow(2) - 4 * b_k * qc, one))) / (2 * b_k)
        out = torch.where(quad, alpha + x_l, out)

        sft = out - x_l
        ld = -torch.log(3 * a_k * sft.pow(2) + 2 * b_k * sft + c_k)
        out = out * (right - left) + left
        ld = ld - math.log(top - bottom) + math.log(right - left)
    else:
        sft = u - x_l
        out = a_k * sft.pow(3) + b_k * sft.pow(2) + c_k * sft + d_k
        ld = torch.log(3 * a_k * sft.pow(2) + 2 * b_k * sft + c_k)
        out = out * (top - bottom) + bottom
        ld = ld + math.log(top - bottom) - math.log(right - left)

    out = torch.where(inside, out, x_in)
    ld = torch.where(inside, ld, torch.zeros_like(ld))
    if return_bins:
        return out, ld, torch.where(inside, idx, torch.full_like(idx, -1))
    return out, ld


# --------------------------------------------------------------------------------------
# element-wise transforms with (optional) conditioner   (flows/affine.py, flows/spline.py)
# --------------------------------------------------------------------------------------

def transform_params(tr: Dict, latent: Optional[Tensor], dtype) -> Tuple[Tensor, ...]:
    """affine.py:59-67 (chunk(2): first d = log-scale, last d = shift);
    spline.py:76-87 (view(..., d, 2K+der); [:K]=widths, [K:2K]=heights, [2K:]=derivs)."""
    kind, d = tr['kind'], tr['dim']
    if tr.get('net') is None:
        return tuple(p.to(dtype) for p in tr['params'])
    out = mlp(tr['net'], latent)
    if kind == 'affine':
        return out[..., :d], out[..., d:]
    K = tr['n_bins']
    P = 2 * K + (K - 1 if kind == 'quadratic' else 2)
    out = out.view(*out.shape[:-1], d, P)
    return out[..., :K], out[..., K:2 * K], out[..., 2 * K:]


def elementwise(tr: Dict, x: Tensor, latent: Optional[Tensor], inverse: bool
                ) -> Tuple[Tensor, Tensor]:
    """(T(x) or T^-1(x), log-diag-Jacobian OF THE DIRECTION EVALUATED) [..., d] each.

    affine.py:97-109 ; spline.py:101-105.
    """
    kind = tr['kind']
    p = transform_params(tr, latent, x.dtype)
    if kind == 'affine':
        log_scale, shift = p
        if inverse:
            return (x - shift) * torch.exp(-log_scale), (-log_scale).expand_as(x)
        return x * torch.exp(log_scale) + shift, log_scale.expand_as(x)
    fn = rqs if kind == 'quadratic' else cubic
    return fn(x, p[0], p[1], p[2], inverse, tr.get('lower', 0), tr.get('upper', 1))


# --------------------------------------------------------------------------------------
# coupling layers  (flows/coupling.py)
# --------------------------------------------------------------------------------------

def _conditioning(x: Tensor, mask: Tensor, latent: Optional[Tensor], t: Optional[Tensor] = None
                  ) -> Tensor:
    """coupling.py:55-67 and :140-154: z = x*mask (*0 if d == 1) [| latent] [| t]."""
    z = x * mask
    if x.shape[-1] == 1:
        z = z * 0
    if latent is not None:
        z = torch.cat([z, latent], -1)
    if t is not None:
        z = torch.cat([z, t], -1)
    return z


def layer_apply(layer: Dict, x: Tensor, *, inverse: bool, latent: Optional[Tensor] = None,
                t: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """One layer in one direction -> (out [..., d], ldj [..., 1]).

    ldj follows the reference's inherited defaults (flow.py:35-47):
      forward:  log_det_jacobian(x, y)           = sum of FORWARD log-diag at x
      inverse:  -log_det_jacobian(x_rec, y)      = -(FORWARD log-diag at the recovered x)
    which for splines is NOT the inverse branch's own log-derivative.
    """
    typ = layer['type']
    if typ in ('flip', 'permute'):      # flows/permute.py:11-82: index shuffle, log-det 0
        if typ == 'flip':
            out = torch.flip(x, [-1])
        else:
            perm = torch.as_tensor(layer['perm']).long()
            if inverse:
                inv = torch.empty_like(perm)
                inv[perm] = torch.arange(perm.numel())
                perm = inv
            out = x[..., perm]
        return out, torch.zeros_like(x[..., :1])
    if typ in ('sigmoid', 'logit'):     # flows/sigmoid.py:9-56 with the inherited log-det defaults
        fi = torch.finfo(x.dtype)
        sig = lambda v: torch.clamp(torch.sigmoid(v), min=fi.tiny, max=1. - fi.eps)
        def logit(v):
            v = v.clamp(min=fi.tiny, max=1. - fi.eps)
            return v.log() - (-v).log1p()
        ld = lambda u: (-F.softplus(-u) - F.softplus(u)).sum(-1, keepdim=True)
        to_unit = (typ == 'sigmoid') != inverse
        if to_unit:                     # sigmoid forward / logit inverse: log-diag at the INPUT
            return sig(x), ld(x)
        out = logit(x)                  # sigmoid inverse / logit forward: -log-diag at the OUTPUT
        return out, -ld(out)
    if typ == 'elementwise':            # stand-alone Affine / Spline (affine.py, spline.py)
        tr = layer['transform']
        out, ld = elementwise(tr, x, latent, inverse)
        return out, ld.sum(-1, keepdim=True)

    mask = make_mask(layer['mask'], x.shape[-1]).to(x.dtype).expand_as(x)

    if typ == 'cont_affine_coupling':   # coupling.py:188-213
        z = _conditioning(x, mask, latent, t if layer.get('concatenate_time', True) else None)
        out = mlp(layer['net'], z)
        d = x.shape[-1]
        ls, sh = out[..., :d], out[..., d:]
        tn = layer['time_scale'].to(x.dtype) * t                     # time_net.py:24-25
        td_ = tn.shape[-1] // 2
        tls, tsh = tn[..., :td_], tn[..., td_:]
        if inverse:
            y = (x - sh * tsh) * torch.exp(-ls * tls)
        else:
            y = x * torch.exp(ls * tls) + sh * tsh
        ld = (ls * tls * (1 - mask)).sum(-1, keepdim=True)
        y = y * (1 - mask) + x * mask
        return y, (-ld if inverse else ld)

    assert typ == 'coupling'            # coupling.py:69-95
    tr = layer['transform']
    z = _conditioning(x, mask, latent)
    out, ld = elementwise(tr, x, z, inverse)
    out = out * (1 - mask) + x * mask
    if inverse and tr['kind'] != 'affine':
        # flow.py:42-47 + coupling.py:84-95: forward log-diag at the recovered point,
        # with z recomputed from it (identical: masked coordinates pass through).
        _, ld_f = elementwise(tr, out, _conditioning(out, mask, latent), False)
        ld = -ld_f
    return out, (ld * (1 - mask)).sum(-1, keepdim=True)


def layer_bins(layer: Dict, x: Tensor, *, inverse: bool, latent: Optional[Tensor] = None) -> Tensor:
    """Bin index the knot search picks for every element of a spline layer: int64 [..., d], -1 for
    out-of-box elements and for a coupling's pass-through dims (search_sorted.py:3-5 called at
    rational_quadratic_spline.py:194-197 / cubic_spline.py:140-143: cumulative widths when
    ``inverse`` is False, cumulative heights otherwise)."""
    tr = layer['transform']
    assert tr['kind'] in ('quadratic', 'cubic')
    if layer['type'] == 'coupling':
        mask = make_mask(layer['mask'], x.shape[-1]).to(x.dtype).expand_as(x)
        cond = _conditioning(x, mask, latent)
    else:
        mask, cond = torch.zeros_like(x), latent
    p = transform_params(tr, cond, x.dtype)
    fn = rqs if tr['kind'] == 'quadratic' else cubic
    bins = fn(x, p[0], p[1], p[2], inverse, tr.get('lower', 0), tr.get('upper', 1), return_bins=True)[2]
    return torch.where(mask != 0, torch.full_like(bins, -1), bins)


def layer_knots(layer: Dict, x: Tensor, *, inverse: bool, latent: Optional[Tensor] = None) -> Tensor:
    """The knot positions ``layer_bins`` searches, in data units: [..., d, K + 1] (cumulative widths, or
    heights when ``inverse``).  Test helper: distance of an element to the knot between two candidate bins."""
    tr = layer['transform']
    if layer['type'] == 'coupling':
        mask = make_mask(layer['mask'], x.shape[-1]).to(x.dtype).expand_as(x)
        cond = _conditioning(x, mask, latent)
    else:
        cond = latent
    uw, uh, ud = transform_params(tr, cond, x.dtype)
    lo, hi = tr.get('lower', 0), tr.get('upper', 1)
    uw, uh, ud = (v.expand(*x.shape, -1) for v in (uw, uh, ud))
    if tr['kind'] == 'quadratic':
        cw, _, ch, _, _ = rqs_knots(uw, uh, ud, lo, hi, lo, hi)
        return ch if inverse else cw
    cw, ch = cubic_coefficients(uw, uh, ud[..., 0, None], ud[..., 1, None])[:2]
    return (ch if inverse else cw) * (hi - lo) + lo


def layer_log_det(layer: Dict, x: Tensor, **kw) -> Tensor:
    return layer_apply(layer, x, inverse=False, **kw)[1]


# --------------------------------------------------------------------------------------
# whole flow  (flow.py:99-152, :172-184 ; dist/normal.py:8-54)
# --------------------------------------------------------------------------------------

def unit_normal_log_prob(x: Tensor) -> Tensor:
    """td.Independent(td.Normal(0,1),1).log_prob: sum_j(-(x_j^2)/2 - log(1) - log(sqrt(2pi)))."""
    return (-(x ** 2) / 2 - math.log(1.0) - math.log(math.sqrt(2 * math.pi))).sum(-1)


def flow_forward(spec: Sequence[Dict], x: Tensor, with_ldj: bool = False, **kw):
    ldj = 0
    for layer in spec:
        x, l = layer_apply(layer, x, inverse=False, **kw)
        ldj = ldj + l
    return (x, ldj) if with_ldj else x


def flow_inverse(spec: Sequence[Dict], y: Tensor, with_ldj: bool = False, **kw):
    ldj = 0
    for layer in reversed(spec):
        y, l = layer_apply(layer, y, inverse=True, **kw)
        ldj = ldj + l
    return (y, ldj) if with_ldj else y


def flow_log_prob(spec: Sequence[Dict], y: Tensor, **kw) -> Tensor:
    """flow.py:127-130 with a UnitNormal base."""
    x, ldj = flow_inverse(spec, y, with_ldj=True, **kw)
    return unit_normal_log_prob(x).unsqueeze(-1) + ldj


def neural_flow_forward(spec: Sequence[Dict], x: Tensor, t: Tensor, t0: Optional[Tensor] = None,
                        **kw) -> Tensor:
    """flow.py:172-184."""
    if t0 is not None:
        for layer in reversed(spec):
            x, _ = layer_apply(layer, x, inverse=True, t=t0, **kw)
    for layer in spec:
        x, _ = layer_apply(layer, x, inverse=False, t=t, **kw)
    return x


def spec_to(spec: Sequence[Dict], dtype) -> List[Dict]:
    """Deep-copy a spec with every tensor cast to ``dtype`` (fp64 arbiter runs)."""
    def conv(v):
        if isinstance(v, torch.Tensor):
            return v.detach().to(dtype).clone()
        if isinstance(v, dict):
            return {k: conv(u) for k, u in v.items()}
        if isinstance(v, (list, tuple)):
            return type(v)(conv(u) for u in v)
        return v
    return [conv(l) for l in spec]
