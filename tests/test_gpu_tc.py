"""GPU: the tcgen05 path -- building blocks, and the tensor-core layer kernel against the
generic CUDA-core kernel and the oracle."""
import ctypes

import pytest
import torch

import cases
from golden_util import close_or_arbitrated
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200 import _lib, _ops
from stribor_b200.spec import layers_from_spec, spec_from_layers

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
@pytest.mark.parametrize('K,N', [(16, 16), (32, 64), (64, 96), (64, 256)])
def test_umma_building_blocks(mode, K, N):
    """D = A B^T on one CTA: fp16 / tf32 single pass, and the 3-pass hi/lo split (fp32-grade)."""
    torch.manual_seed(K * 1000 + N + mode)
    A = torch.rand(128, K, device=DEV) * 2 - 1
    B = torch.rand(N, K, device=DEV) * 2 - 1
    D = torch.full((128, N), float('nan'), device=DEV)
    rc = _lib.lib().stb_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, mode, 0,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < (2e-5 if mode >= 2 else 1e-2), err
    if mode >= 2:      # as good as an fp32 GEMM
        err32 = ((A @ B.t()).double() - ref).abs().max().item()
        assert err < 8 * err32 + 1e-6


@pytest.mark.parametrize('variant', [4, 8, 12])
@pytest.mark.parametrize('mode', [0, 2])
@pytest.mark.parametrize('K,N', [(16, 64), (96, 64), (128, 64), (128, 96)])
def test_umma_mn_major_operands(mode, K, N, variant):
    """Transposed operands consumed in place (MN-major descriptors): the gradient products of the
    training kernel reuse the K-major buffers of the forward recompute this way."""
    torch.manual_seed(K * 1000 + N + mode + variant)
    A = torch.rand(128, K, device=DEV) * 2 - 1
    B = torch.rand(N, K, device=DEV) * 2 - 1
    Ain = A.t().contiguous() if variant & 4 else A
    Bin = B.t().contiguous() if variant & 8 else B
    D = torch.full((128, N), float('nan'), device=DEV)
    rc = _lib.lib().stb_tc_selftest(Ain.data_ptr(), Bin.data_ptr(), D.data_ptr(), K, N, mode, variant,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < (2e-5 if mode >= 2 else 1e-2) * max(1.0, K / 32), err


def _flows(kind, d, masks, seed, n_layers=3, lower=-4., upper=4.):
    case = cases._mk_flow(kind, d, [64], n_layers, 16, 700, seed, masks=masks, lower=lower, upper=upper,
                          scale=1.7)()
    return case


@pytest.mark.parametrize('kind', ['quadratic', 'cubic'])
@pytest.mark.parametrize('d,masks', [(64, cases.ALT), (64, ('parity_even', 'parity_odd')),
                                     (30, cases.ALT), (63, ('ordered_left_half', 'parity_odd')), (2, cases.ALT),
                                     # dim > 64: the 128-row kernel (tc_wide.cu), BASELINE configs[4] width
                                     (128, cases.ALT), (100, ('parity_even', 'parity_odd')),
                                     (65, ('ordered_left_half', 'parity_odd'))])
def test_tensor_path_matches_generic_and_oracle(kind, d, masks, monkeypatch):
    case = _flows(kind, d, masks, seed=900 + d)
    x = case['inputs']['x'].to(DEV)
    x[2, 1] = 5.5                                          # an identity-tail value
    spec = case['spec']

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NormalizingFlow(st.UnitNormal(d), layers), layers

    tflow, tl = build()
    with torch.no_grad():
        desc = tl[0].describe(d, 0, torch.device(DEV))
        assert desc['packed'] is not None, 'tensor path was not selected'
        L = _ops.make_struct(desc['meta'], desc['fmeta'], desc['mask'], desc['params'], desc['packed'])
        assert _lib.lib().stb_layer_uses_tensor_path(ctypes.byref(L)) == 1
        lp_t = tflow.log_prob(x)
        xi_t, li_t = tflow.inverse_and_log_det_jacobian(x)
        yf_t, lf_t = tflow.forward_and_log_det_jacobian(x)
        xi_only = tflow.inverse(x)
        yf_only = tflow.forward(x)
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gflow, gl = build()
    with torch.no_grad():
        assert gl[0].describe(d, 0, torch.device(DEV))['packed'] is None
        lp_g = gflow.log_prob(x)
        xi_g, li_g = gflow.inverse_and_log_det_jacobian(x)
        yf_g, lf_g = gflow.forward_and_log_det_jacobian(x)
    assert torch.equal(xi_only, xi_t) and torch.equal(yf_only, yf_t)

    xc = x.cpu()
    s64 = O.spec_to(spec, torch.float64)
    o32 = {'lp': O.flow_log_prob(spec, xc), 'inv': O.flow_inverse(spec, xc, with_ldj=True),
           'fwd': O.flow_forward(spec, xc, with_ldj=True)}
    o64 = {'lp': O.flow_log_prob(s64, xc.double()), 'inv': O.flow_inverse(s64, xc.double(), with_ldj=True),
           'fwd': O.flow_forward(s64, xc.double(), with_ldj=True)}

    def chk(got, a32, a64, what, frac=0.0):
        fail, _, mx = close_or_arbitrated(got, a32, a64, 1e-5, 1e-5)
        assert fail <= frac, f'{what}: {fail:.3%} outside tolerance (max abs err {mx:.3e})'

    def noise_ratio(got, a32, a64):
        """median |got - ref64| over median |ref32 - ref64|: how much noisier than the reference's own fp32."""
        e_new = (got.detach().cpu().double() - a64).abs().flatten()
        e_ref = (a32.double() - a64).abs().flatten()
        return (e_new.median() / e_ref.median().clamp_min(1e-12)).item()

    for tag, (lp, xi, li, yf, lf) in (('tensor', (lp_t, xi_t, li_t, yf_t, lf_t)),
                                      ('generic', (lp_g, xi_g, li_g, yf_g, lf_g))):
        # log_prob -- the quantity BASELINE.json's tolerance is stated for: every row, except the
        # cubic one-root rows whose reference values are themselves ill-conditioned
        # (d = 2: one transformed element per layer, nothing averages out -> a 0.5 % tail at the
        # reference's own noise level, see profiles/r01_tc_accuracy.txt)
        # (cubic: measured on the d = 63 case, tools/tc_debug3.py -- the failing row's one-root Cardano
        # element is off by 3e-2 in the reference's OWN fp32 run and both CUDA paths agree with each
        # other to 1e-4 there; 1-2 such rows in 350 is the level to allow)
        chk(lp, o32['lp'], o64['lp'], f'{tag} log_prob', 6e-3 if kind == 'cubic' else (5e-3 if d < 8 else 0.0))
        # intermediate quantities have |value| ~ 1-10, so rtol/atol = 1e-5 sits AT the fp32 noise
        # floor of a 3-layer flow (the reference's own fp32-vs-fp64 error reaches 1e-4): allow a
        # 2 % tail, and require the error level to stay within 2.5x of the reference's own.
        for got, key, idx in ((xi, 'inv', 0), (li, 'inv', 1), (yf, 'fwd', 0), (lf, 'fwd', 1)):
            # idx 1 = log-det: a sum over >= 45 element log-derivatives, each good to atol 1e-5
            fail, _, mx = close_or_arbitrated(got, o32[key][idx], o64[key][idx], 1e-5, 1e-4 if idx else 1e-5)
            assert fail <= 2e-2, f'{tag} {key}[{idx}]: {fail:.3%} outside tolerance (max abs err {mx:.3e})'
            r = noise_ratio(got, o32[key][idx], o64[key][idx])
            assert r < 2.5, f'{tag} {key}[{idx}]: {r:.2f}x the reference fp32 noise'
    # the two CUDA paths agree with each other at fp32 noise level
    if kind == 'cubic':      # a handful of one-root Cardano elements amplify 1-ulp differences
        assert ((xi_t - xi_g).abs() > 1e-4).float().mean().item() < 5e-3
    else:
        assert (xi_t - xi_g).abs().max().item() < 1e-4
    assert (yf_t - yf_g).abs().max().item() < 1e-4


@pytest.mark.parametrize('d,hidden,masks', [(64, [256, 256], cases.ALT), (64, [64], ('parity_even', 'parity_odd')),
                                            (30, [128, 128], cases.ALT), (63, [192], ('ordered_left_half', 'parity_odd')),
                                            (2, [64, 64], cases.ALT)])
def test_affine_tensor_path_matches_generic_and_oracle(d, hidden, masks, monkeypatch):
    """tc_mlp.cu (affine coupling, wide conditioner on tcgen05) vs the generic kernel and the oracle."""
    case = cases._mk_flow('affine', d, hidden, 3, 0, 700, 700 + d + len(hidden), masks=masks, scale=1.3)()
    spec = case['spec']
    x = case['inputs']['x'].to(DEV)

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NormalizingFlow(st.UnitNormal(d), layers), layers

    tflow, tl = build()
    with torch.no_grad():
        desc = tl[0].describe(d, 0, torch.device(DEV))
        assert desc['packed'] is not None, 'tensor path was not selected'
        lp_t = tflow.log_prob(x)
        xi_t, li_t = tflow.inverse_and_log_det_jacobian(x)
        yf_t, lf_t = tflow.forward_and_log_det_jacobian(x)
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gflow, _ = build()
    with torch.no_grad():
        lp_g = gflow.log_prob(x)
        xi_g, li_g = gflow.inverse_and_log_det_jacobian(x)
    xc = x.cpu()
    s64 = O.spec_to(spec, torch.float64)
    for got, f32, f64, what in (
            (lp_t, O.flow_log_prob(spec, xc), O.flow_log_prob(s64, xc.double()), 'log_prob'),
            (xi_t, O.flow_inverse(spec, xc), O.flow_inverse(s64, xc.double()), 'inverse x'),
            (li_t, O.flow_inverse(spec, xc, with_ldj=True)[1], O.flow_inverse(s64, xc.double(), with_ldj=True)[1], 'inverse ldj'),
            (yf_t, O.flow_forward(spec, xc), O.flow_forward(s64, xc.double()), 'forward y'),
            (lf_t, O.flow_forward(spec, xc, with_ldj=True)[1], O.flow_forward(s64, xc.double(), with_ldj=True)[1], 'forward ldj')):
        fail, _, mx = close_or_arbitrated(got, f32, f64, 1e-5, 1e-5)
        assert fail <= 2e-3, f'{what}: {fail:.3%} outside tolerance (max abs err {mx:.3e})'
    # the two CUDA paths agree at fp32 noise level; log_prob reaches |lp| ~ 200 through 0.5 * |x|^2, so its bound is
    # relative (rtol = 1e-5, the parity tolerance) on top of the absolute one
    assert ((lp_t - lp_g).abs() <= 2e-4 + 1e-5 * lp_g.abs()).all().item() and (xi_t - xi_g).abs().max().item() < 2e-4
    assert (li_t - li_g).abs().max().item() < 2e-4


@pytest.mark.parametrize('kind,d,hidden,latent_dim,n_layers,rows', [
    ('quadratic', 32, [64], 8, 3, 700),        # spline, whole-flow kernel (n_cond 16 + 8 latent columns of GEMM1)
    ('cubic', 40, [64], 12, 2, 300),           # 20 + 12 = all 32 K columns in use
    ('quadratic', 32, [64], 16, 1, 130),       # one layer: the per-layer kernel
    ('quadratic', 20, [64], 5, 9, 40000),      # more than 8 layers: two chained launches, many tiles
    ('quadratic', 64, [64], 16, 2, 300),       # 32 conditioning + 16 latent columns > 32: the 128-row kernel (64 K columns)
    ('cubic', 100, [64], 10, 2, 200),          # d > 64 with a latent input (tc_wide.cu)
    ('quadratic', 32, [128, 128], 8, 2, 300),  # spline with a wide conditioner (tc_hwide.cu)
    ('affine', 32, [128, 128], 8, 2, 500),     # affine, pipelined kernel
    ('affine', 24, [64], 7, 3, 300),           # affine, phase-by-phase kernel (H = 64: no chain with a latent input)
])
def test_latent_input_on_the_tensor_path(kind, d, hidden, latent_dim, n_layers, rows, monkeypatch):
    """`latent=` (coupling.py:64-65): the conditioner reads [x * mask | latent]; on the tensor-core kernels the latent
    columns are extra K columns of GEMM1, read from the caller's [rows, latent_dim] tensor.  Against the CUDA-core kernel
    and the oracle, both directions, chained and layer by layer."""
    rs = cases._rs(900 + d + latent_dim)
    spec = [cases.coupling_spec(rs, kind, d, hidden, cases.ALT[i % 2], n_bins=16 if kind != 'affine' else 0,
                                lower=-4., upper=4., latent_dim=latent_dim) for i in range(n_layers)]
    torch.manual_seed(rows)
    x = torch.randn(rows, d, device=DEV) * 1.3
    lat = torch.randn(rows, latent_dim, device=DEV)

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NormalizingFlow(st.UnitNormal(d), layers), layers

    tflow, tl = build()
    with torch.no_grad():
        desc = tl[0].describe(d, latent_dim, torch.device(DEV))
        assert desc['packed'] is not None, 'tensor path was not selected'
        L = _ops.make_struct(desc['meta'], desc['fmeta'], desc['mask'], [p.detach() for p in desc['params']], desc['packed'])
        assert _lib.lib().stb_layer_uses_tensor_path(ctypes.byref(L)) == 1
        tflow.log_prob(x[:4], latent=lat[:4])
        n0 = _ops.launch_count()
        lp_t = tflow.log_prob(x, latent=lat)
        launches = _ops.launch_count() - n0
        xi_t, li_t = tflow.inverse_and_log_det_jacobian(x, latent=lat)
        yf_t, lf_t = tflow.forward_and_log_det_jacobian(x, latent=lat)
        # layer by layer (the per-layer kernels): bit-identical to the chained launch for the spline kernels
        cur, tot = x, 0
        for f in reversed(tflow.transforms):
            cur, l = f.inverse_and_log_det_jacobian(cur, latent=lat)
            tot = tot + l
    if kind != 'affine':
        # MLP[64] spline layers chain (<= 8 per launch) while [conditioning | latent] fits 32 K columns at d <= 64;
        # the d <= 128 kernel and wide conditioners go out one launch per layer
        chained = hidden == [64] and d <= 64 and (d + 1) // 2 + latent_dim <= 32
        assert launches == ((n_layers + 7) // 8 if chained else n_layers)
        assert torch.equal(cur, xi_t)
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gflow, _ = build()
    with torch.no_grad():
        lp_g = gflow.log_prob(x, latent=lat)
        xi_g, li_g = gflow.inverse_and_log_det_jacobian(x, latent=lat)
        yf_g, lf_g = gflow.forward_and_log_det_jacobian(x, latent=lat)
    tol = 5e-4 if n_layers > 3 else 2e-4
    if kind == 'cubic':      # a handful of one-root Cardano elements amplify 1-ulp differences (as in the test above)
        assert ((xi_t - xi_g).abs() > 1e-4).float().mean().item() < 5e-3
        assert ((li_t - li_g).abs() > 1e-3).float().mean().item() < 2e-2
    else:
        assert (xi_t - xi_g).abs().max().item() < tol
        assert (li_t - li_g).abs().max().item() < 5 * tol
        assert ((lp_t - lp_g).abs() <= 5 * tol + 1e-5 * lp_g.abs()).all().item()
    assert (yf_t - yf_g).abs().max().item() < tol and (lf_t - lf_g).abs().max().item() < 5 * tol
    if rows <= 1000:
        xc, lc = x.cpu(), lat.cpu()
        s64 = O.spec_to(spec, torch.float64)
        for got, f32, f64, what in (
                (xi_t, O.flow_inverse(spec, xc, latent=lc), O.flow_inverse(s64, xc.double(), latent=lc.double()), 'inverse x'),
                (yf_t, O.flow_forward(spec, xc, latent=lc), O.flow_forward(s64, xc.double(), latent=lc.double()), 'forward y'),
                (lp_t, O.flow_log_prob(spec, xc, latent=lc), O.flow_log_prob(s64, xc.double(), latent=lc.double()), 'log_prob')):
            fail, _, mx = close_or_arbitrated(got, f32, f64, 1e-5, 1e-5)
            assert fail <= 2e-2, f'{what}: {fail:.3%} outside tolerance (max abs err {mx:.3e})'


def test_latent_input_with_folded_permutations():
    """`latent=` couplings with Flip / Permute layers between them: still ONE launch (the permutations are folded into the
    gather lists of the conditioning columns only -- the latent columns are not tile columns)."""
    d, latent_dim, rows = 32, 8, 600
    rs = cases._rs(77)
    spec = []
    for i in range(3):
        spec.append(cases.coupling_spec(rs, 'quadratic', d, [64], 'ordered_right_half', n_bins=16, lower=-4., upper=4.,
                                        latent_dim=latent_dim))
        spec.append({'type': 'flip'} if i != 1 else {'type': 'permute', 'perm': rs.permutation(d).tolist()})
    torch.manual_seed(5)
    x = torch.randn(rows, d, device=DEV) * 1.3
    lat = torch.randn(rows, latent_dim, device=DEV)
    flow = st.NormalizingFlow(st.UnitNormal(d), [l.to(DEV) for l in layers_from_spec(spec)])
    with torch.no_grad():
        flow.log_prob(x[:4], latent=lat[:4])
        n0 = _ops.launch_count()
        lp = flow.log_prob(x, latent=lat)
        assert _ops.launch_count() - n0 == 1
        yf, lf = flow.forward_and_log_det_jacobian(x, latent=lat)
    xc, lc = x.cpu(), lat.cpu()
    s64 = O.spec_to(spec, torch.float64)
    for got, f32, f64, what in (
            (lp, O.flow_log_prob(spec, xc, latent=lc), O.flow_log_prob(s64, xc.double(), latent=lc.double()), 'log_prob'),
            (yf, O.flow_forward(spec, xc, latent=lc), O.flow_forward(s64, xc.double(), latent=lc.double()), 'forward y')):
        fail, _, mx = close_or_arbitrated(got, f32, f64, 1e-5, 1e-5)
        assert fail <= 2e-2, f'{what}: {fail:.3%} outside tolerance (max abs err {mx:.3e})'


@pytest.mark.parametrize('kind,d,K,n_layers,rows,masks', [
    ('quadratic', 64, 8, 3, 700, cases.ALT),
    ('quadratic', 64, 5, 8, 300, cases.ALT),
    ('quadratic', 30, 2, 2, 257, ('parity_even', 'parity_odd')),
    ('quadratic', 64, 15, 1, 130, cases.ALT),
    ('quadratic', 32, 10, 3, 50000, cases.ALT),           # many tiles per CTA
    ('cubic', 64, 8, 3, 700, cases.ALT),
    ('cubic', 33, 3, 2, 300, ('ordered_left_half', 'parity_odd')),
    ('cubic', 64, 12, 1, 129, cases.ALT),
])
def test_fewer_than_16_bins_on_the_tensor_path(kind, d, K, n_layers, rows, masks, monkeypatch):
    _check_few_bins(kind, d, K, n_layers, rows, masks, [64], 1, monkeypatch)


@pytest.mark.parametrize('kind,d,K,hidden,rows', [
    ('quadratic', 100, 8, [64], 300),            # d > 64: the 128-row kernel (tc_wide.cu)
    ('cubic', 128, 5, [64], 200),
    ('quadratic', 48, 10, [128, 128], 300),      # wide conditioners (tc_hwide.cu)
    ('cubic', 32, 6, [256], 130),
])
def test_fewer_than_16_bins_on_the_other_spline_kernels(kind, d, K, hidden, rows, monkeypatch):
    _check_few_bins(kind, d, K, 2, rows, cases.ALT, hidden, 2, monkeypatch)


def _check_few_bins(kind, d, K, n_layers, rows, masks, hidden, expect_launches, monkeypatch):
    """Spline couplings with 2 <= n_bins < 16 run on the tensor-core path too: the packed image keeps its 16-bin
    layout, padded bins get weight 0 / bias -inf (numerator exactly 0) and the search never selects them
    (tc_spline16.cuh, FULL = false).  Against the CUDA-core kernel and the oracle, both directions."""
    rs = cases._rs(1300 + d + K)
    spec = [cases.coupling_spec(rs, kind, d, hidden, masks[i % 2], n_bins=K, lower=-4., upper=4.) for i in range(n_layers)]
    torch.manual_seed(rows + K)
    x = torch.randn(rows, d, device=DEV) * 1.6                   # a good share of elements outside the box, and inside
    x[0, :] = 4.0                                               # exactly on the upper box edge: the last REAL bin
    x[1 % rows, :] = -4.0

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NormalizingFlow(st.UnitNormal(d), layers), layers

    tflow, tl = build()
    with torch.no_grad():
        desc = tl[0].describe(d, 0, torch.device(DEV))
        assert desc['packed'] is not None, 'tensor path was not selected'
        L = _ops.make_struct(desc['meta'], desc['fmeta'], desc['mask'], [p.detach() for p in desc['params']], desc['packed'])
        assert _lib.lib().stb_layer_uses_tensor_path(ctypes.byref(L)) == 1
        tflow.log_prob(x[:4])
        n0 = _ops.launch_count()
        lp_t = tflow.log_prob(x)
        assert _ops.launch_count() - n0 == expect_launches
        xi_t, li_t = tflow.inverse_and_log_det_jacobian(x)
        yf_t, lf_t = tflow.forward_and_log_det_jacobian(x)
        back = tflow.inverse(yf_t)
    assert torch.isfinite(lp_t).all() and torch.isfinite(xi_t).all() and torch.isfinite(yf_t).all()
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gflow, _ = build()
    with torch.no_grad():
        lp_g = gflow.log_prob(x)
        xi_g, li_g = gflow.inverse_and_log_det_jacobian(x)
        yf_g, lf_g = gflow.forward_and_log_det_jacobian(x)
    tol = 5e-4 if n_layers > 3 else 2e-4
    # the two rows that sit exactly on the box edges: a layer maps the edge to the edge +- 1 ulp, and which side of it
    # the NEXT layer sees decides "inside / identity" for that element -- compare their log-dets only for one layer
    body = slice(0, None) if n_layers == 1 else slice(2, None)
    if kind == 'cubic':
        assert ((xi_t - xi_g).abs() > 1e-4).float().mean().item() < 5e-3
        assert ((li_t - li_g).abs() > 1e-3).float().mean().item() < 2e-2
        assert ((back - x).abs() > 1e-3).float().mean().item() < 5e-3
    else:
        assert (xi_t - xi_g).abs().max().item() < tol
        assert (li_t - li_g)[body].abs().max().item() < 5 * tol
        assert ((lp_t - lp_g).abs() <= 5 * tol + 1e-5 * lp_g.abs())[body].all().item()
        assert (back - x).abs().max().item() < 1e-3
    assert (yf_t - yf_g).abs().max().item() < tol and (lf_t - lf_g)[body].abs().max().item() < 5 * tol
    if rows <= 1000:
        xc = x.cpu()
        s64 = O.spec_to(spec, torch.float64)
        for got, f32, f64, what in (
                (xi_t, O.flow_inverse(spec, xc), O.flow_inverse(s64, xc.double()), 'inverse x'),
                (yf_t, O.flow_forward(spec, xc), O.flow_forward(s64, xc.double()), 'forward y'),
                (lf_t, O.flow_forward(spec, xc, with_ldj=True)[1], O.flow_forward(s64, xc.double(), with_ldj=True)[1], 'forward ldj'),
                (lp_t, O.flow_log_prob(spec, xc), O.flow_log_prob(s64, xc.double()), 'log_prob')):
            fail, _, mx = close_or_arbitrated(got, f32, f64, 1e-5, 1e-4 if 'ldj' in what or what == 'log_prob' else 1e-5)
            assert fail <= 2e-2, f'{what}: {fail:.3%} outside tolerance (max abs err {mx:.3e})'


@pytest.mark.parametrize('hidden', [[64], [128, 128]])
def test_continuous_affine_with_latent_on_the_tensor_path(hidden, monkeypatch):
    """ContinuousAffineCoupling with `latent=`: the conditioner reads [x * mask | latent | t] (coupling.py:140-154)."""
    d, latent_dim, rows = 16, 6, 777
    rs = cases._rs(431 + len(hidden))
    spec = [cases.cont_affine_spec(rs, d, hidden, ('ordered_0', 'ordered_1')[i % 2], latent_dim=latent_dim) for i in range(3)]
    torch.manual_seed(3)
    x = torch.randn(rows, d, device=DEV)
    t = torch.rand(rows, 1, device=DEV)
    lat = torch.randn(rows, latent_dim, device=DEV)

    def run():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        with torch.no_grad():
            desc = layers[0].describe(d, latent_dim, torch.device(DEV))
            return st.NeuralFlow(layers)(x, t=t, latent=lat), desc['packed'] is not None

    y_t, packed = run()
    assert packed, 'tensor path was not selected'
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    y_g, packed_g = run()
    assert not packed_g
    assert (y_t - y_g).abs().max().item() < 2e-4
    s64 = O.spec_to(spec, torch.float64)
    want32 = O.flow_forward(spec, x.cpu(), t=t.cpu(), latent=lat.cpu())
    want64 = O.flow_forward(s64, x.cpu().double(), t=t.cpu().double(), latent=lat.cpu().double())
    fail, _, mx = close_or_arbitrated(y_t, want32, want64, 1e-5, 1e-5)
    assert fail <= 2e-3, f'{fail:.3%} outside tolerance (max abs err {mx:.3e})'


@pytest.mark.parametrize('d,kind', [(16, 'cont'), (30, 'affine')])
def test_small_conditioner_two_ctas_per_sm(d, kind, monkeypatch):
    """Many rows of a small conditioner (dim <= 32, MLP[64]): the 8-warp configuration of tc_mlp.cu runs with two
    CTAs per SM and resident weights.  Agreement with the CUDA-core kernel over every tile, ragged tail included."""
    import numpy as np
    rs = np.random.RandomState(977 + d)
    rows = 2 * 148 * 128 + 12345
    if kind == 'cont':
        spec = [cases.cont_affine_spec(rs, d, [64], ('ordered_0', 'ordered_1')[i % 2]) for i in range(2)]
    else:
        spec = cases._mk_flow('affine', d, [64], 2, 0, 8, 1234)()['spec']
    x = cases._x(rs, (rows, d)).to(DEV)
    t = cases._x(rs, (rows, 1), uniform=True).to(DEV)
    out = {}
    for force in ('0', '1'):
        monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', force)
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        with torch.no_grad():
            if kind == 'cont':
                assert (layers[0].describe(d, 0, torch.device(DEV))['packed'] is not None) == (force == '0')
                out[force] = (st.NeuralFlow(layers)(x, t=t), None)
            else:
                flow = st.NormalizingFlow(st.UnitNormal(d), layers)
                out[force] = (flow.inverse(x), flow.log_prob(x))
    assert (out['0'][0] - out['1'][0]).abs().max().item() < 5e-5
    if kind == 'affine':
        assert (out['0'][1] - out['1'][1]).abs().max().item() < 2e-4


@pytest.mark.parametrize('concat', [True, False])
def test_continuous_affine_tensor_path(concat, monkeypatch):
    """NeuralFlow of ContinuousAffineCoupling layers on the tcgen05 path: oracle parity (with and
    without t0), agreement with the generic kernel, exact identity at t = 0 (test_neural_flow.py:22-27)."""
    import numpy as np
    rs = np.random.RandomState(4242 + concat)
    d = 16
    spec = [cases.cont_affine_spec(rs, d, [64], ('ordered_0', 'ordered_1', 'parity_even')[i % 3], concatenate_time=concat)
            for i in range(4)]
    x = cases._x(rs, (37, 11, d)).to(DEV)
    t = cases._x(rs, (37, 11, 1), uniform=True).to(DEV)
    t0 = cases._x(rs, (37, 11, 1), uniform=True).to(DEV)

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NeuralFlow(layers), layers

    nf, layers = build()
    with torch.no_grad():
        assert layers[0].describe(d, 0, torch.device(DEV))['packed'] is not None, 'tensor path was not selected'
        y = nf(x, t=t)
        y0 = nf(x, t=t, t0=t0)
        assert (nf(x, t=torch.zeros_like(t)) == x).all()
        assert torch.allclose(nf(x, t=t0, t0=t0), x, atol=1e-5)
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gnf, gl = build()
    with torch.no_grad():
        assert gl[0].describe(d, 0, torch.device(DEV))['packed'] is None
        yg = gnf(x, t=t)
    assert (y - yg).abs().max().item() < 2e-5
    s64 = O.spec_to(spec, torch.float64)
    xc, tc_, t0c = x.cpu(), t.cpu(), t0.cpu()
    for got, a32, a64 in ((y, O.neural_flow_forward(spec, xc, tc_), O.neural_flow_forward(s64, xc.double(), tc_.double())),
                          (y0, O.neural_flow_forward(spec, xc, tc_, t0c),
                           O.neural_flow_forward(s64, xc.double(), tc_.double(), t0c.double()))):
        fail, _, mx = close_or_arbitrated(got, a32, a64, 1e-5, 1e-5)
        assert fail <= 1e-3, (fail, mx)


@pytest.mark.parametrize('kind', ['quadratic', 'cubic'])
def test_tensor_path_box_ends_forward(kind):
    """x exactly on the box ends (bin 0 / bin K-1 through the nudged last knot, search_sorted.py:4)
    and just outside: forward direction, where the map is continuous, against the fp64 oracle."""
    d = 64
    case = _flows(kind, d, cases.ALT, seed=77, n_layers=1)
    spec = case['spec']
    x = case['inputs']['x'][:256].clone()
    x[:, 0] = 4.0
    x[:, 1] = -4.0
    x[:, 2] = 4.000001
    x[:, 3] = -4.000001
    x[:, 40] = 4.0
    f = layers_from_spec(spec)[0].to(DEV)
    with torch.no_grad():
        y, ldj = f.forward_and_log_det_jacobian(x.to(DEV))
    y64, l64 = O.layer_apply(O.spec_to(spec, torch.float64)[0], x.double(), inverse=False)
    y32, l32 = O.layer_apply(spec[0], x, inverse=False)
    for got, a32, a64, at in ((y, y32, y64, 1e-5), (ldj, l32, l64, 1e-4)):
        fail, _, mx = close_or_arbitrated(got, a32, a64, 1e-5, at)
        assert fail == 0.0, (fail, mx)
    assert torch.equal(y[:, 2].cpu(), x[:, 2]) and torch.equal(y[:, 3].cpu(), x[:, 3])   # identity tails
    assert (y[:, 0].cpu() - 4.0).abs().max() < 1e-5 and (y[:, 1].cpu() + 4.0).abs().max() < 1e-5


def test_tensor_path_repacks_when_weights_change():
    d = 64
    c = st.Coupling(st.Spline(d, 16, latent_net=st.net.MLP(d, [64], d * 47), lower=-4, upper=4,
                              spline_type='quadratic'), mask='ordered_0').to(DEV)
    x = torch.randn(300, d, device=DEV)
    with torch.no_grad():
        y0 = c(x)
        p0 = c.describe(d, 0, x.device)['packed']
        assert p0 is not None and c.describe(d, 0, x.device)['packed'] is p0      # cached
        c.transform.latent_net.net[2].weight.mul_(0.5)                            # in-place update
        y1 = c(x)
        assert not torch.equal(y0, y1)
        ref = st.Coupling(st.Spline(d, 16, latent_net=st.net.MLP(d, [64], d * 47), lower=-4, upper=4,
                                    spline_type='quadratic'), mask='ordered_0').to(DEV)
        ref.load_state_dict(c.state_dict())
        assert torch.equal(ref(x), y1)


@pytest.mark.parametrize('kind', ['quadratic', 'cubic'])
@pytest.mark.parametrize('n_layers,d,rows', [(12, 64, 1000), (3, 30, 70000), (9, 63, 513)])
def test_whole_flow_kernel_matches_layer_by_layer(kind, n_layers, d, rows):
    """stb_flow_log_prob / stb_flow_apply launch runs of <= 8 layers as ONE kernel (tile resident in shared
    memory between layers, CHAIN variant).  Same arithmetic as one launch per layer: results must agree to
    the last bit for x and to fp32 summation order for the log-dets."""
    case = _flows(kind, d, cases.ALT if d % 2 == 0 else ('ordered_left_half', 'parity_odd'), seed=1500 + d, n_layers=n_layers)
    torch.manual_seed(rows)
    x = (torch.randn(rows, d) * 1.7).to(DEV)
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    flow = st.NormalizingFlow(st.UnitNormal(d), layers)
    n0 = _ops.launch_count()
    with torch.no_grad():
        lp = flow.log_prob(x)
        n_lp = _ops.launch_count() - n0
        xi, li = flow.inverse_and_log_det_jacobian(x)
        yf, lf = flow.forward_and_log_det_jacobian(x)
        # layer by layer through the single-layer entry point
        cur, tot = x, torch.zeros(rows, 1, device=DEV)
        for l in reversed(layers):
            cur, ld = l.inverse_and_log_det_jacobian(cur)
            tot = tot + ld
        cur_f, tot_f = x, torch.zeros(rows, 1, device=DEV)
        for l in layers:
            cur_f, ld = l.forward_and_log_det_jacobian(cur_f)
            tot_f = tot_f + ld
    packs = 6 * n_layers                                   # first use packs both images of every layer
    assert n_lp - packs == (n_layers + 7) // 8, f'{n_lp - packs} launches for {n_layers} layers'
    assert torch.equal(xi, cur), (xi - cur).abs().max().item()
    assert torch.equal(yf, cur_f), (yf - cur_f).abs().max().item()
    assert (li - tot).abs().max().item() < 2e-4 and (lf - tot_f).abs().max().item() < 2e-4
    base = (-0.5 * cur * cur - 0.9189385332046727).sum(-1, keepdim=True)
    assert (lp - (tot + base)).abs().max().item() < 5e-4


def _cont_flow(d, n_layers, seed, hidden=(64,), latent=False):
    rs = cases._rs(seed)
    spec = [cases.cont_affine_spec(rs, d, list(hidden), ('ordered_0', 'ordered_1')[i % 2]) for i in range(n_layers)]
    return spec


@pytest.mark.parametrize('n_layers,d,rows,hidden', [(4, 16, 70000, (64,)), (4, 16, 129, (64,)), (7, 12, 5000, (64,)),
                                                    (2, 32, 1000, (64,)), (3, 6, 300, (64,)), (3, 16, 5000, (64, 64)),
                                                    (9, 8, 40000, (64,))])
def test_affine_chain_kernel_matches_layer_by_layer(n_layers, d, rows, hidden):
    """NeuralFlow.forward (flow.py:172-184) over ContinuousAffineCoupling + TimeLinear layers with small
    conditioners is ONE launch (tc_mlp.cu, CHAIN kernel: weights of all layers resident in shared memory, the tile
    stays on chip between layers).  Same arithmetic as one launch per layer: bit-identical outputs; and both
    against the oracle."""
    # (64,) with <= 16 inputs: tc_mlp_chain4_kernel (h in TMEM, four tiles in flight); 32 dims (17 inputs) or two hidden
    # layers: tc_mlp_chain_kernel<2> (h through shared memory)
    spec = _cont_flow(d, n_layers, 4000 + d, hidden=hidden)
    layers = [l.to(DEV) for l in layers_from_spec(spec)]
    torch.manual_seed(rows)
    x = torch.randn(rows, d, device=DEV)
    t = torch.rand(rows, 1, device=DEV)
    t0 = torch.rand(rows, 1, device=DEV)
    nf = st.NeuralFlow(layers)
    with torch.no_grad():
        nf(x[:8], t=t[:8])                                  # packs the weight images
        n0 = _ops.launch_count()
        y = nf(x, t=t)
        n_fwd = _ops.launch_count() - n0
        y0 = nf(x, t=t, t0=t0)
        cur = x
        for l in layers:
            cur = l(cur, t=t)
        cur0 = x
        for l in reversed(layers):
            cur0 = l.inverse(cur0, t=t0)
        for l in layers:
            cur0 = l(cur0, t=t)
    per_launch = 2 if len(hidden) > 1 else (4 if d <= 16 else 3)    # layers whose weights fit next to the tiles (at least)
    assert n_fwd <= (n_layers + per_launch - 1) // per_launch + 1, f'{n_fwd} launches for {n_layers} layers'
    assert torch.equal(y, cur), (y - cur).abs().max().item()
    assert torch.equal(y0, cur0), (y0 - cur0).abs().max().item()
    want = O.neural_flow_forward(O.spec_to(spec, torch.float64), x.cpu().double(), t.cpu().double())
    assert ((y.cpu().double() - want).abs() <= 2e-5 + 1e-5 * want.abs()).all()
    # identity at t = 0 stays exact through the chain
    with torch.no_grad():
        assert torch.equal(nf(x, t=torch.zeros_like(t)), x)


@pytest.mark.parametrize('direction', ['log_prob', 'forward_ldj'])
def test_affine_chain_kernel_log_det(direction):
    """an affine-coupling NormalizingFlow with MLP[64] conditioners on the chain kernel: log-dets / log_prob
    equal the layer-by-layer composition and the oracle"""
    d, n_layers, rows = 16, 5, 3000
    case = cases._mk_flow('affine', d, [64], n_layers, 0, rows, 4100)()
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    flow = st.NormalizingFlow(st.UnitNormal(d), layers)
    x = case['inputs']['x'].to(DEV)
    with torch.no_grad():
        if direction == 'log_prob':
            got = flow.log_prob(x)
            cur, tot = x, torch.zeros(rows, 1, device=DEV)
            for l in reversed(layers):
                cur, ld = l.inverse_and_log_det_jacobian(cur)
                tot = tot + ld
            ref_v = tot + (-0.5 * cur * cur - 0.9189385332046727).sum(-1, keepdim=True)
            want = O.flow_log_prob(O.spec_to(case['spec'], torch.float64), x.cpu().double())
        else:
            y, got = flow.forward_and_log_det_jacobian(x)
            cur, tot = x, torch.zeros(rows, 1, device=DEV)
            for l in layers:
                cur, ld = l.forward_and_log_det_jacobian(cur)
                tot = tot + ld
            ref_v = tot
            assert torch.equal(y, cur)
            want = O.flow_forward(O.spec_to(case['spec'], torch.float64), x.cpu().double(), with_ldj=True)[1]
    assert (got - ref_v).abs().max().item() < 2e-4
    assert ((got.cpu().double() - want).abs() <= 1e-4 + 1e-5 * want.abs()).all()


def test_log_prob_without_latent_rows():
    """stb_flow_log_prob(x_out = NULL): allowed exactly when the flow is one chained launch; same values"""
    import ctypes as C
    case = _flows('quadratic', 64, cases.ALT, seed=77, n_layers=8)
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    flow = st.NormalizingFlow(st.UnitNormal(64), layers)
    x = torch.randn(5000, 64, device=DEV) * 1.5
    with torch.no_grad():
        lp = flow.log_prob(x)                                             # CHAIN_LOG_PROB_ONLY: no latent rows written
        xi, li = flow.inverse_and_log_det_jacobian(x)
    want = li + (-0.5 * xi * xi - 0.9189385332046727).sum(-1, keepdim=True)
    assert (lp - want).abs().max().item() < 5e-4
    # through the raw ABI: 8 chainable layers need no x_out, 9 layers (two launches) do
    def arr_of(ls):
        descs = [l.describe(64, 0, torch.device(DEV)) for l in ls]
        arr = (_lib.StbLayer * len(ls))()
        keep = [_ops._fill_struct(arr[i], d['meta'], d['fmeta'], d['mask'], [p.detach() for p in d['params']], d['packed'])
                for i, d in enumerate(descs)]
        return arr, keep, descs
    arr, keep, descs = arr_of(layers)
    lib = _lib.lib()
    assert lib.stb_flow_log_prob_needs_x_out(arr, 8) == 0
    out = torch.empty(5000, device=DEV)
    rc = lib.stb_flow_log_prob(arr, 8, x.data_ptr(), None, None, None, out.data_ptr(), 5000,
                               torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    assert torch.equal(out.view(-1, 1), lp)
    case9 = _flows('quadratic', 64, cases.ALT, seed=78, n_layers=9)
    layers9 = [l.to(DEV) for l in layers_from_spec(case9['spec'])]
    arr9, keep9, descs9 = arr_of(layers9)
    assert lib.stb_flow_log_prob_needs_x_out(arr9, 9) == 1
    rc = lib.stb_flow_log_prob(arr9, 9, x.data_ptr(), None, None, None, out.data_ptr(), 5000,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.E_INVAL
    with torch.no_grad():
        lp9 = st.NormalizingFlow(st.UnitNormal(64), layers9).log_prob(x)      # the module allocates the scratch itself
    assert torch.isfinite(lp9).all()


@pytest.mark.parametrize('family', ['quadratic', 'cubic', 'cont_affine'])
def test_permutations_fold_into_the_chain(family):
    """Flip / Permute (flows/permute.py:11-82) between chained couplings cost no launch: they are folded into the
    chain kernels' gather / scatter lists.  Result == layer-by-layer application == oracle."""
    rs = cases._rs(515)
    if family == 'cont_affine':
        d, rows = 16, 3000
        spec = []
        for i in range(3):
            spec.append(cases.cont_affine_spec(rs, d, [64], ('ordered_0', 'ordered_1')[i % 2]))
            spec.append({'type': 'permute', 'perm': rs.permutation(d).tolist()} if i % 2 == 0 else {'type': 'flip'})
    else:
        d, rows = 64, 3000
        spec = [{'type': 'flip'}]
        for i in range(4):
            spec.append(cases.coupling_spec(rs, family, d, [64], 'ordered_right_half', n_bins=16, lower=-4., upper=4.))
            spec.append({'type': 'permute', 'perm': rs.permutation(d).tolist()} if i % 2 == 0 else {'type': 'flip'})
    layers = [l.to(DEV) for l in layers_from_spec(spec)]
    torch.manual_seed(5)
    x = torch.randn(rows, d, device=DEV) * 1.5
    kw = {'t': torch.rand(rows, 1, device=DEV)} if family == 'cont_affine' else {}
    flow = st.NormalizingFlow(st.UnitNormal(d), layers)
    with torch.no_grad():
        flow.forward(x[:4], **{k: v[:4] for k, v in kw.items()})         # pack
        n0 = _ops.launch_count()
        yf, lf = flow.forward_and_log_det_jacobian(x, **kw)
        n_fwd = _ops.launch_count() - n0
        xi, li = flow.inverse_and_log_det_jacobian(x, **kw)
        lp = flow.log_prob(x, **kw)
        cur, tot = x, torch.zeros(rows, 1, device=DEV)
        for l in layers:
            cur, ld = l.forward_and_log_det_jacobian(cur, **kw)
            tot = tot + ld
        cur_i, tot_i = x, torch.zeros(rows, 1, device=DEV)
        for l in reversed(layers):
            cur_i, ld = l.inverse_and_log_det_jacobian(cur_i, **kw)
            tot_i = tot_i + ld
    assert n_fwd == 1, f'{n_fwd} launches'
    assert torch.equal(yf, cur), (yf - cur).abs().max().item()
    assert torch.equal(xi, cur_i), (xi - cur_i).abs().max().item()
    assert (lf - tot).abs().max().item() < 2e-4 and (li - tot_i).abs().max().item() < 2e-4
    base = (-0.5 * cur_i * cur_i - 0.9189385332046727).sum(-1, keepdim=True)
    assert (lp - (tot_i + base)).abs().max().item() < 5e-4
    okw = {k: v.cpu().double() for k, v in kw.items()}
    want = O.flow_forward(O.spec_to(spec, torch.float64), x.cpu().double(), **okw)
    err = (yf.cpu().double() - want).abs()
    if family == 'cubic':
        assert (err > 1e-3).double().mean().item() < 0.01
    else:
        assert err.max().item() < 2e-4, err.max().item()


def test_permutation_in_place_between_unchained_layers():
    """a permutation that is NOT absorbed into a chain (generic-path couplings around it) runs as its own kernel on
    the flow's single output buffer: the row is gathered into registers before it is written"""
    case = cases.build_case('permute_quadratic_d16')
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    flow = st.NormalizingFlow(st.UnitNormal(16), layers)
    x = case['inputs']['x'].to(DEV)
    with torch.no_grad():
        y = flow.forward(x)
        xr = flow.inverse(y)
    want = O.flow_forward(case['spec'], x.cpu())
    torch.testing.assert_close(y.cpu(), want, rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(xr, x, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('variant', [16, 18])
@pytest.mark.parametrize('mode', [0, 2])
@pytest.mark.parametrize('K,N', [(16, 16), (64, 16), (64, 64), (64, 96), (128, 64)])
def test_umma_a_operand_from_tmem(mode, K, N, variant):
    """The A operand written to TMEM with tcgen05.st (lane = row, two fp16 K elements per column) and consumed by
    TMEM-sourced UMMAs: the building block that keeps a layer's activations on the tensor-core side."""
    torch.manual_seed(K * 1000 + N + mode + variant)
    A = torch.rand(128, K, device=DEV) * 2 - 1
    B = torch.rand(N, K, device=DEV) * 2 - 1
    D = torch.full((128, N), float('nan'), device=DEV)
    rc = _lib.lib().stb_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, mode, variant,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < (2e-5 if mode >= 2 else 1e-2) * max(1.0, K / 32), err


@pytest.mark.parametrize('kind', ['quadratic', 'cubic'])
@pytest.mark.parametrize('d,hidden,rows', [(64, [256, 256], 700), (64, [256], 300), (64, [128, 128], 257), (64, [128], 129),
                                           (30, [192], 300), (64, [64, 64], 300), (17, [256, 256], 1)])
def test_wide_conditioner_spline_kernel(kind, d, hidden, rows, monkeypatch):
    """tc_hwide.cu: spline couplings with MLP[H] / MLP[H,H] conditioners (SURVEY 8d's secondary configs[2] shape) on the
    tensor cores -- hidden activations written back into the consumed accumulator columns (tcgen05.st) and read as the
    next GEMM's A operand from TMEM.  Against the CUDA-core kernel and the fp64 oracle; bins against the oracle."""
    masks = cases.ALT if d % 2 == 0 else ('ordered_left_half', 'parity_odd')
    case = cases._mk_flow(kind, d, hidden, 2, 16, max(rows, 2), 7000 + d + len(hidden), masks=masks, lower=-4., upper=4., scale=1.6)()
    spec = case['spec']
    x = case['inputs']['x'][:rows].to(DEV)

    def build():
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        return st.NormalizingFlow(st.UnitNormal(d), layers), layers

    tflow, tl = build()
    with torch.no_grad():
        desc = tl[0].describe(d, 0, torch.device(DEV))
        assert desc['packed'] is not None, 'tensor path was not selected'
        L = _ops.make_struct(desc['meta'], desc['fmeta'], desc['mask'], desc['params'], desc['packed'])
        assert _lib.lib().stb_layer_uses_tensor_path(ctypes.byref(L)) == 1
        lp_t = tflow.log_prob(x)
        xi_t, li_t = tflow.inverse_and_log_det_jacobian(x)
        yf_t, lf_t = tflow.forward_and_log_det_jacobian(x)
        rt = (tflow.forward(xi_t) - x).abs().max().item()
        _, _, bins = _ops.layer_apply_bins(x, None, None, desc['mask'], [p.detach() for p in desc['params']], desc['packed'],
                                           desc['meta'], desc['fmeta'], _lib.FORWARD)
    monkeypatch.setenv('STRIBOR_B200_FORCE_GENERIC', '1')
    gflow, _ = build()
    with torch.no_grad():
        lp_g = gflow.log_prob(x)
        xi_g, li_g = gflow.inverse_and_log_det_jacobian(x)
    xc = x.cpu()
    s64 = O.spec_to(spec, torch.float64)
    lp32, lp64 = O.flow_log_prob(spec, xc), O.flow_log_prob(s64, xc.double())
    fail, _, mx = close_or_arbitrated(lp_t, lp32, lp64, 1e-5, 1e-5)
    assert fail <= (6e-3 if kind == 'cubic' else 0.0), f'log_prob: {fail:.3%} outside tolerance (max abs err {mx:.3e})'
    f64 = O.flow_forward(s64, xc.double(), with_ldj=True)
    f32 = O.flow_forward(spec, xc, with_ldj=True)
    for got, a32, a64, atol in ((yf_t, f32[0], f64[0], 1e-5), (lf_t, f32[1], f64[1], 1e-4)):
        fail, _, mx = close_or_arbitrated(got, a32, a64, 1e-5, atol)
        assert fail <= 2e-2, f'forward: {fail:.3%} outside tolerance (max abs err {mx:.3e})'
    if kind == 'quadratic':
        assert rt < 1e-4, rt
        assert (xi_t - xi_g).abs().max().item() < 1e-4 and (li_t - li_g).abs().max().item() < 5e-4
        assert (lp_t - lp_g).abs().max().item() < 1e-3
    else:
        assert ((xi_t - xi_g).abs() > 1e-4).float().mean().item() < 5e-3
    want_bins = O.layer_bins(spec[0], xc, inverse=False).to(torch.int32)
    flips = int((bins.cpu() != want_bins).sum())
    assert flips <= max(1, int(1e-4 * want_bins.numel())), flips
