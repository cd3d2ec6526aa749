"""GPU: host-side contracts added in round 2 -- argument checks, weight-image invalidation, activation gating of
the tensor-core paths, the differentiable per-dimension API (incl. separate RQS boxes,
rational_quadratic_spline.py:55-64 / reference test_spline.py:36-52) and the reverse=True sign convention."""
import pytest
import torch

import cases
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200 import _lib, _ops
from stribor_b200.spec import layers_from_spec

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _spline_flow(kind='quadratic', act='Tanh', rows=300, seed=5, dim=64):
    case = cases._mk_flow(kind, dim, [64], 2, 16, rows, seed, lower=-4., upper=4., scale=1.5, activation=act)()
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    return case, st.NormalizingFlow(st.UnitNormal(dim), layers)


def test_wrong_dtype_or_device_raises_python_errors():
    torch.manual_seed(0)
    f = st.ContinuousAffineCoupling(st.net.MLP(5, [16], 8), st.net.TimeLinear(8), 'ordered_0').to(DEV)
    x = torch.randn(7, 4, device=DEV)
    t = torch.rand(7, 1, device=DEV)
    with torch.no_grad():
        f(x, t=t)
        with pytest.raises(TypeError):
            f(x, t=t.double())                      # would be re-read as float32 pairs
        with pytest.raises(RuntimeError):
            f(x, t=t.cpu())
        with pytest.raises(TypeError):
            f(x.double(), t=t)
        g = st.Coupling(st.Affine(4, latent_net=st.net.MLP(4 + 3, [8], 8)), 'ordered_0').to(DEV)
        lat = torch.randn(7, 3, device=DEV)
        g(x, latent=lat)
        with pytest.raises(TypeError):
            g(x, latent=lat.double())
        cpu_module = st.Coupling(st.Affine(4, latent_net=st.net.MLP(4, [8], 8)), 'ordered_0')
        with pytest.raises(RuntimeError):
            cpu_module(x)                           # weights on the host: a foreign pointer for the kernel
        flow = st.NormalizingFlow(st.UnitNormal(4), [cpu_module])
        with pytest.raises(RuntimeError):
            flow.log_prob(x)


def test_packed_image_follows_the_weights():
    """in-place parameter updates repack automatically; writes through .data need invalidate_packed()"""
    case, flow = _spline_flow()
    x = case['inputs']['x'].to(DEV)
    with torch.no_grad():
        lp0 = flow.log_prob(x)
        lin = [m for m in flow.modules() if isinstance(m, torch.nn.Linear)]
        lin[1].weight.mul_(1.25)                    # bumps the version counter
        lp1 = flow.log_prob(x)
        spec = st.spec.spec_from_layers(list(flow.transforms))
        want1 = O.flow_log_prob(spec, x.cpu())
        assert (lp1 - lp0).abs().max() > 1e-3
        torch.testing.assert_close(lp1.cpu(), want1, rtol=1e-4, atol=2e-4)
        lin[1].weight.data.mul_(0.8)                # bypasses the counter ...
        st.invalidate_packed(flow)                  # ... so the owner is told
        lp2 = flow.log_prob(x)
        want2 = O.flow_log_prob(st.spec.spec_from_layers(list(flow.transforms)), x.cpu())
        torch.testing.assert_close(lp2.cpu(), want2, rtol=1e-4, atol=2e-4)
        assert (lp2 - lp1).abs().max() > 1e-3
        # mode switches and state-dict loads invalidate by themselves
        sd = {k: v.clone() for k, v in flow.state_dict().items()}
        lin[3].weight.data.mul_(1.5)
        flow.eval()
        lp3 = flow.log_prob(x)
        assert (lp3 - lp2).abs().max() > 1e-3
        flow.load_state_dict(sd)
        torch.testing.assert_close(flow.log_prob(x), lp2, rtol=0, atol=0)


def test_call_plan_is_reused_and_never_stale(monkeypatch):
    """flow.run_chain keeps the prepared layer array between calls (host time per small-batch call) and rebuilds it
    whenever a weight, an attribute or the module structure changes."""
    from stribor_b200 import flow as F
    from stribor_b200.flows.coupling import Coupling
    case, flow = _spline_flow()
    x = case['inputs']['x'].to(DEV)
    calls = [0]
    orig = Coupling.describe

    def counting(self, *a, **k):
        calls[0] += 1
        return orig(self, *a, **k)

    monkeypatch.setattr(Coupling, 'describe', counting)

    def oracle_lp():
        return O.flow_log_prob(st.spec.spec_from_layers(list(flow.transforms)), x.cpu())

    with torch.no_grad():
        lp0 = flow.log_prob(x)
        n0 = calls[0]
        assert n0 >= len(flow.transforms)
        for _ in range(3):
            torch.testing.assert_close(flow.log_prob(x), lp0, rtol=0, atol=0)
        assert calls[0] == n0, 'the plan was rebuilt although nothing changed'
        y0 = flow.inverse(x)                                      # same plan serves the other entry points
        assert calls[0] == n0
        # another stream: ordered after the pack event, same numbers
        s2 = torch.cuda.Stream()
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s2):
            lp_s = flow.log_prob(x)
        s2.synchronize()
        torch.testing.assert_close(lp_s, lp0, rtol=0, atol=0)
        # an attribute of a layer (the spline box) is honoured at the next call
        sp = flow.transforms[0].transform
        sp.lower, sp.upper = sp.lower * 1.5, sp.upper * 1.5
        lp1 = flow.log_prob(x)
        assert calls[0] > n0 and (lp1 - lp0).abs().max() > 1e-4
        torch.testing.assert_close(lp1.cpu(), oracle_lp(), rtol=1e-4, atol=2e-4)
        # a parameter OBJECT replaced (not just written)
        lin = [m for m in flow.modules() if isinstance(m, torch.nn.Linear)]
        lin[0].weight = torch.nn.Parameter(lin[0].weight.detach() * 0.5)
        lp2 = flow.log_prob(x)
        assert (lp2 - lp1).abs().max() > 1e-4
        torch.testing.assert_close(lp2.cpu(), oracle_lp(), rtol=1e-4, atol=2e-4)
        # a layer appended to the ModuleList
        extra = layers_from_spec(case['spec'])[0].to(DEV)
        flow.transforms.append(extra)
        lp3 = flow.log_prob(x)
        torch.testing.assert_close(lp3.cpu(), oracle_lp(), rtol=1e-4, atol=2e-4)
        torch.testing.assert_close(flow.inverse(x).cpu(), O.flow_inverse(st.spec.spec_from_layers(list(flow.transforms)), x.cpu()),
                                   rtol=1e-4, atol=2e-4)
    assert y0.shape == x.shape
    # with gradients on, the chain path steps aside (the plan is for inference)
    xg = x.clone().requires_grad_(True)
    assert flow.log_prob(xg).grad_fn is not None
    assert F._record(flow.transforms).chainable


@pytest.mark.parametrize('act', ['ReLU', 'ELU', 'SiLU'])
def test_unbounded_activation_stays_off_the_tensor_path_and_finite(act):
    """a hidden value above 65504 would split into (+inf, -inf) on the fp16 hi|lo tensor-core path"""
    case, flow = _spline_flow(act=act, rows=64)
    layer = flow.transforms[0]
    d = layer.describe(64, 0, torch.device(DEV))
    assert d['packed'] is None                      # no tensor-core image: the CUDA-core kernel runs
    x = case['inputs']['x'].clone()
    x[:8, 32:] *= 3e5                                # huge pass-through values feed the conditioner unchanged
    with torch.no_grad():
        lp = flow.log_prob(x.to(DEV)).cpu()
    want = O.flow_log_prob(O.spec_to(case['spec'], torch.float64), x.double())
    assert torch.isfinite(lp).all()
    assert ((lp.double() - want).abs() <= 1e-4 + 1e-5 * want.abs()).all()


def test_tanh_layers_keep_the_tensor_path():
    case, flow = _spline_flow(rows=32)
    d = flow.transforms[0].describe(64, 0, torch.device(DEV))
    L = _ops.make_struct(d['meta'], d['fmeta'], d['mask'], [p.detach() for p in d['params']], d['packed'])
    assert d['packed'] is not None and _lib.lib().stb_layer_uses_tensor_path(L) == 1


@pytest.mark.parametrize('shape', [(1, 1), (10, 2), (7, 4, 5)])
@pytest.mark.parametrize('inverse', [False, True])
def test_rqs_separate_boxes_values_and_gradients(shape, inverse):
    """unconstrained_rational_quadratic_spline(left, right, bottom, top): values and autograd against the
    oracle in fp64 (the configuration of the reference's UnboundedSpline test)."""
    torch.manual_seed(11)
    K = 5
    dim = shape[-1]
    x = torch.rand(*shape) * (5 if inverse else 2) - (3 if inverse else 1)        # codomain [-3, 2] / domain [-1, 1]
    w, h, dv = torch.randn(dim, K), torch.randn(dim, K), torch.randn(dim, K - 1)
    box = dict(left=-1, right=1, bottom=-3, top=2)
    leaves64 = [v.double().requires_grad_(True) for v in (x, w, h, dv)]
    o64, ld64 = O.rqs(leaves64[0], *leaves64[1:], inverse, -1., 1., **box)
    (o64.sin().sum() + (ld64 * ld64).sum()).backward()
    leaves = [v.to(DEV).requires_grad_(True) for v in (x, w, h, dv)]
    o, ld = st.util.unconstrained_rational_quadratic_spline(leaves[0], *leaves[1:], inverse=inverse, **box)
    torch.testing.assert_close(o.detach().cpu().double(), o64.detach(), rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(ld.detach().cpu().double(), ld64.detach(), rtol=1e-5, atol=5e-5)
    (o.sin().sum() + (ld * ld).sum()).backward()
    for a, b, nm in zip(leaves, leaves64, 'xwhd'):
        scale = b.grad.abs().max().clamp_min(1e-6)
        err = (a.grad.cpu().double() - b.grad).abs().max() / scale
        assert err < 2e-3, (nm, err.item())


@pytest.mark.parametrize('kind', ['quadratic', 'cubic'])
def test_log_diag_api_is_differentiable(kind):
    """Spline.forward_and_log_diag_jacobian / log_diag_jacobian carry gradients, learned and
    conditioned parameters"""
    torch.manual_seed(3)
    dim, K = 4, 6
    P = 3 * K - 1 if kind == 'quadratic' else 2 * K + 2
    for latent_dim in (0, 3):
        net = st.net.MLP(latent_dim, [12], dim * P) if latent_dim else None
        f = st.Spline(dim, K, latent_net=net, lower=-2, upper=2, spline_type=kind).to(DEV)
        x = (torch.rand(9, dim, device=DEV) * 5 - 2.5).requires_grad_(True)
        lat = torch.randn(9, latent_dim, device=DEV) if latent_dim else None
        y, ld = f.forward_and_log_diag_jacobian(x, lat)
        (y.sum() + ld.sum()).backward()
        assert x.grad is not None and torch.isfinite(x.grad).all()
        for p in f.parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all()
        # d y_i / d x_i == exp(log-diag) inside the box, 1 in the tails
        x2 = x.detach().clone().requires_grad_(True)
        y2, ld2 = f.forward_and_log_diag_jacobian(x2, lat)
        y2.sum().backward()
        torch.testing.assert_close(x2.grad, ld2.detach().exp(), rtol=2e-4, atol=2e-4)


def test_reverse_flag_returns_forward_log_det():
    """affine.py:97-109 / coupling.py:188-209: forward_and_log_det_jacobian(..., reverse=True) is the inverse
    map with the FORWARD log-det; inverse_and_log_det_jacobian negates it"""
    torch.manual_seed(4)
    x = torch.randn(6, 4, device=DEV)
    with torch.no_grad():
        f = st.Affine(4).to(DEV)
        y, ldj = f.forward_and_log_det_jacobian(x)
        xr, ldr = f.forward_and_log_det_jacobian(y, reverse=True)
        xi, ldi = f.inverse_and_log_det_jacobian(y)
        torch.testing.assert_close(xr, x, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ldr, ldj)
        torch.testing.assert_close(ldi, -ldj)
        t = torch.rand(6, 1, device=DEV)
        g = st.ContinuousAffineCoupling(st.net.MLP(5, [16], 8), st.net.TimeLinear(8), 'ordered_0').to(DEV)
        y, ldj = g.forward_and_log_det_jacobian(x, t)
        xr, ldr = g.forward_and_log_det_jacobian(y, t, reverse=True)
        torch.testing.assert_close(xr, x, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ldr, ldj, rtol=1e-5, atol=1e-6)


def test_spline_subclass_hooks_reach_every_method():
    """reference test_spline.py:43-46: overriding forward_and_log_diag_jacobian changes forward, inverse and
    the log-dets; a Coupling around such a subclass composes it the reference way"""
    class Boxed(st.Spline):
        def forward_and_log_diag_jacobian(self, x, latent=None, *, reverse=False, **kwargs):
            w, h, d = self._get_params(latent)
            return self.spline(x, w, h, d, inverse=reverse, left=-1, bottom=-3, top=2, right=1)

    torch.manual_seed(6)
    f = Boxed(dim=3, n_bins=5, spline_type='quadratic').to(DEV)
    assert not f.plain()
    x = torch.rand(8, 3, device=DEV)
    with torch.no_grad():
        y = f(x)
        assert y.min() >= -3 and y.max() <= 2 and (y - x).abs().max() > 1e-2
        torch.testing.assert_close(f.inverse(y), x, rtol=1e-4, atol=1e-4)
        _, ldj = f.forward_and_log_det_jacobian(x)
        _, ldi = f.inverse_and_log_det_jacobian(y)
        torch.testing.assert_close(ldj, -ldi, rtol=1e-4, atol=1e-4)
        want, wld = O.rqs(x.cpu(), f.width.cpu(), f.height.cpu(), f.derivative.cpu(), False, -1., 1.,
                          left=-1, right=1, bottom=-3, top=2)
        torch.testing.assert_close(y.cpu(), want, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ldj.cpu(), wld.sum(-1, keepdim=True), rtol=1e-4, atol=1e-4)
        # a coupling around a hooked transform
        net = st.net.MLP(4, [8], 4 * 14)
        c = st.Coupling(Boxed(dim=4, n_bins=5, latent_net=net, spline_type='quadratic'), 'ordered_0').to(DEV)
        x4 = torch.rand(8, 4, device=DEV)
        y4, l4 = c.forward_and_log_det_jacobian(x4)
        x4r, l4r = c.inverse_and_log_det_jacobian(y4)
        torch.testing.assert_close(x4r, x4, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(l4r, -l4, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(y4[:, 2:], x4[:, 2:], rtol=0, atol=0)


@pytest.mark.parametrize('kind,dim,hidden', [('quadratic', 64, [64]), ('cubic', 128, [64]), ('affine', 64, [128, 128]),
                                             ('affine', 16, [64])])
def test_sigmoid_hidden_activation_on_the_tensor_path(kind, dim, hidden):
    """Sigmoid is the second bounded activation the tensor-core kernels evaluate themselves (clamped ex2 form);
    large |pre-activation| must neither overflow nor lose parity"""
    rows = 500
    case = cases._mk_flow(kind, dim, hidden, 3, 16 if kind != 'affine' else 0, rows, 8100 + dim, lower=-4., upper=4.,
                          scale=1.5, activation='Sigmoid')()
    layers = [l.to(DEV) for l in layers_from_spec(case['spec'])]
    flow = st.NormalizingFlow(st.UnitNormal(dim), layers)
    d = layers[0].describe(dim, 0, torch.device(DEV))
    L = _ops.make_struct(d['meta'], d['fmeta'], d['mask'], [p.detach() for p in d['params']], d['packed'])
    assert _lib.lib().stb_layer_uses_tensor_path(L) == 1
    x = case['inputs']['x'].clone()
    x[:4] *= 200.0                                      # far outside the box: conditioner inputs of +-600
    with torch.no_grad():
        lp = flow.log_prob(x.to(DEV)).cpu()
        xr = flow.inverse(x.to(DEV))
        rt = (flow.forward(xr).cpu() - x).abs() / (1 + x.abs())
    want = O.flow_log_prob(O.spec_to(case['spec'], torch.float64), x.double())
    assert torch.isfinite(lp).all()
    err = (lp.double() - want).abs() / (1e-5 * want.abs() + 1e-5)
    frac = (err > 20).double().mean().item()
    assert frac <= (0.01 if kind == 'cubic' else 0.0), (frac, err.max().item())
    assert rt.max() < (5e-2 if kind == 'cubic' else 1e-4)
