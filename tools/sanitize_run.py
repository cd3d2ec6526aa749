"""Small run of every kernel (ragged tails included) for compute-sanitizer.

    compute-sanitizer --tool memcheck|synccheck|racecheck python tools/sanitize_run.py [only]

`only` (optional): a substring -- run just the sections whose name contains it (one process per kernel family, so a
tool that stops the process at its first report still covers the others)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import cases
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
dev = 'cuda'
ONLY = sys.argv[1] if len(sys.argv) > 1 else ''
def run(name, case, rows_list=(1, 255, 257, 700), grad=False):
    if ONLY and ONLY not in name:
        return
    layers = [l.to(dev) for l in layers_from_spec(case['spec'])]
    d = case['inputs']['x'].shape[-1]
    flow = st.NormalizingFlow(st.UnitNormal(d), layers)
    for r in rows_list:
        x = torch.randn(r, d, device=dev) * 1.5
        with torch.no_grad():
            lp = flow.log_prob(x); y, l = flow.forward_and_log_det_jacobian(x); xi = flow.inverse(y)
        assert torch.isfinite(lp).all()
    if grad:
        x = torch.randn(77, d, device=dev, requires_grad=True)
        (-flow.log_prob(x).mean()).backward()
    torch.cuda.synchronize()
    print('ok', name, flush=True)
run('tc quadratic', cases._mk_flow('quadratic', 64, [64], 2, 16, 8, 1)(), grad=True)
run('tc cubic d63', cases._mk_flow('cubic', 63, [64], 2, 16, 8, 2, masks=('parity_even', 'ordered_left_half'))(), grad=True)
run('wide quadratic d128 (128-row kernel, fused backward)', cases._mk_flow('quadratic', 128, [64], 2, 16, 8, 5)(), rows_list=(1, 127, 300), grad=True)
run('wide cubic d66 parity (odd chunk count)', cases._mk_flow('cubic', 66, [64], 2, 16, 8, 6, masks=('parity_even', 'parity_odd'))(), rows_list=(1, 130), grad=True)
run('chain 12 layers d30', cases._mk_flow('quadratic', 30, [64], 12, 16, 8, 7)(), rows_list=(1, 255, 600))
run('tc affine 256x256', cases._mk_flow('affine', 64, [256, 256], 2, 0, 8, 3)(), grad=True)
run('tc affine d30 h128', cases._mk_flow('affine', 30, [128], 2, 0, 8, 4)())
run('generic quadratic d5', cases.build_case('quadratic_d5_parity'), rows_list=(1, 33))
run('generic cubic d7', cases.build_case('cubic_d7_ordered'), rows_list=(1, 33), grad=True)
run('pointwise', cases.build_case('permute_quadratic_d16'), rows_list=(1, 9, 300))
import numpy as np
rs = np.random.RandomState(1)
if not ONLY or ONLY in 'neural tc':
    spec = [cases.cont_affine_spec(rs, 16, [64], 'ordered_0') for _ in range(2)]
    nf = st.NeuralFlow([l.to(dev) for l in layers_from_spec(spec)])
    with torch.no_grad():
        for r in (1, 129, 300):
            x = torch.randn(r, 16, device=dev); t = torch.rand(r, 1, device=dev)
            nf(x, t=t, t0=t * 0.5)
    torch.cuda.synchronize(); print('ok neural tc')
# round 2: chain kernels with folded permutations, the affine / continuous-affine chain kernel, sigmoid hidden units,
# the bin-index instrument
rs = np.random.RandomState(2)
spec = [{'type': 'flip'}]
for i in range(3):
    spec.append(cases.coupling_spec(rs, 'quadratic', 64, [64], 'ordered_right_half', n_bins=16, lower=-4., upper=4.,
                                    activation='Sigmoid' if i == 1 else 'Tanh'))
    spec.append({'type': 'permute', 'perm': rs.permutation(64).tolist()})
run('spline chain with folded permutations', {'spec': spec, 'inputs': {'x': torch.zeros(1, 64)}}, rows_list=(1, 257, 600))
spec = []
for i in range(5):
    spec.append(cases.cont_affine_spec(rs, 16, [64], ('ordered_0', 'ordered_1')[i % 2]))
    spec.append({'type': 'flip'})
if not ONLY or ONLY in 'affine chain kernel with folded permutations':
    nf = st.NeuralFlow([l.to(dev) for l in layers_from_spec(spec)])
    with torch.no_grad():
        for r in (1, 129, 300, 1000):
            x = torch.randn(r, 16, device=dev); t = torch.rand(r, 1, device=dev)
            nf(x, t=t, t0=t * 0.5)
    torch.cuda.synchronize(); print('ok affine chain kernel with folded permutations', flush=True)
from stribor_b200 import _ops, _lib
for kind, d in ((('quadratic', 64), ('cubic', 128)) if (not ONLY or ONLY in 'bins instrument') else ()):
    case = cases._mk_flow(kind, d, [64], 1, 16, 8, 9)()
    layer = layers_from_spec(case['spec'])[0].to(dev)
    ds = layer.describe(d, 0, torch.device(dev))
    for r in (1, 300):
        x = torch.randn(r, d, device=dev)
        for direction in (_lib.FORWARD, _lib.INVERSE):
            _ops.layer_apply_bins(x, None, None, ds['mask'], [p.detach() for p in ds['params']], ds['packed'], ds['meta'],
                                  ds['fmeta'], direction)
torch.cuda.synchronize()
if not ONLY or ONLY in 'bins instrument':
    print('ok bins instrument', flush=True)
# round 2, later: the wide-conditioner spline kernel (TMEM-resident activations, two issuers) and the chain4 kernel
run('hwide quadratic 256x256', cases._mk_flow('quadratic', 64, [256, 256], 2, 16, 8, 11)(), rows_list=(1, 129, 300))
run('hwide cubic 128 d30', cases._mk_flow('cubic', 30, [128], 2, 16, 8, 12)(), rows_list=(1, 257))
if not ONLY or ONLY in 'chain4 neural':
    spec = [cases.cont_affine_spec(rs, 16, [64], ('ordered_0', 'ordered_1')[i % 2]) for i in range(4)]
    nf = st.NeuralFlow([l.to(dev) for l in layers_from_spec(spec)])
    with torch.no_grad():
        for r in (1, 129, 513, 2000):
            x = torch.randn(r, 16, device=dev); t = torch.rand(r, 1, device=dev)
            nf(x, t=t, t0=t * 0.5)
    torch.cuda.synchronize(); print('ok chain4 neural', flush=True)
run('tc affine d63 h192 (pipe kernel, one hidden layer, odd dim)', cases._mk_flow('affine', 63, [192], 2, 0, 8, 13, masks=('ordered_left_half', 'parity_odd'))(), rows_list=(1, 129, 300))
# round 2, last: `latent=` columns on the tensor-core kernels (whole-flow kernel in its two-CTA form at these row counts,
# wide-conditioner spline kernel, affine pipelined kernel)
def run_latent(name, kind, d, hidden, latent_dim, n_layers, rows_list):
    if ONLY and ONLY not in name:
        return
    rs = cases._rs(77 + d)
    spec = [cases.coupling_spec(rs, kind, d, hidden, cases.ALT[i % 2], n_bins=16 if kind != 'affine' else 0,
                                lower=-4., upper=4., latent_dim=latent_dim) for i in range(n_layers)]
    flow = st.NormalizingFlow(st.UnitNormal(d), [l.to(dev) for l in layers_from_spec(spec)])
    for r in rows_list:
        x = torch.randn(r, d, device=dev) * 1.5
        lat = torch.randn(r, latent_dim, device=dev)
        with torch.no_grad():
            lp = flow.log_prob(x, latent=lat); y, l = flow.forward_and_log_det_jacobian(x, latent=lat)
        assert torch.isfinite(lp).all()
    torch.cuda.synchronize()
    print('ok', name, flush=True)
run_latent('latent spline chain d32', 'quadratic', 32, [64], 8, 3, (1, 129, 300))
run_latent('latent cubic single layer d40', 'cubic', 40, [64], 12, 1, (1, 257))
run_latent('latent hwide d32', 'quadratic', 32, [128, 128], 8, 2, (1, 130))
run_latent('latent affine pipe d32', 'affine', 32, [128, 128], 8, 2, (1, 129, 300))
run_latent('latent wide d64 (48 K columns)', 'quadratic', 64, [64], 16, 2, (1, 129, 300))
# fewer than 16 bins on the tensor-core kernels (padded 16-bin layout)
run('few bins pair kernel d64 k8', cases._mk_flow('quadratic', 64, [64], 3, 8, 8, 21)(), rows_list=(1, 129, 300))
run('few bins pair kernel cubic d33 k3', cases._mk_flow('cubic', 33, [64], 2, 3, 8, 22, masks=('ordered_left_half', 'parity_odd'))(), rows_list=(1, 257))
run('few bins wide d100 k8', cases._mk_flow('quadratic', 100, [64], 2, 8, 8, 23)(), rows_list=(1, 130))
run('few bins hwide d48 k10', cases._mk_flow('quadratic', 48, [128, 128], 2, 10, 8, 24)(), rows_list=(1, 130))
