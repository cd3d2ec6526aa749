"""Debug: per-layer comparison tensor path vs generic path vs fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import cases
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec

DEV = 'cuda'
kind = sys.argv[1] if len(sys.argv) > 1 else 'quadratic'
d = 64
case = cases._mk_flow(kind, d, [64], 1, 16, 4096, 964, masks=cases.ALT, lower=-4., upper=4., scale=1.7)()
spec = case['spec']
x = case['inputs']['x']
tl = [l.to(DEV) for l in layers_from_spec(spec)]
os.environ['STRIBOR_B200_FORCE_GENERIC'] = '1'
gl = [l.to(DEV) for l in layers_from_spec(spec)]
with torch.no_grad():
    xg, lg = gl[0].inverse_and_log_det_jacobian(x.to(DEV))
os.environ['STRIBOR_B200_FORCE_GENERIC'] = '0'
with torch.no_grad():
    xt, lt = tl[0].inverse_and_log_det_jacobian(x.to(DEV))
s64 = O.spec_to(spec, torch.float64)
x64, l64 = O.layer_apply(s64[0], x.double(), inverse=True)
x32, l32 = O.layer_apply(spec[0], x, inverse=True)
for tag, xx, ll in (('tensor', xt, lt), ('generic', xg, lg), ('oracle32', x32, l32)):
    ex = (xx.cpu().double() - x64).abs()
    el = (ll.cpu().double() - l64).abs()
    print(f'{tag:9s}: x err max {ex.max():.3e} mean {ex.mean():.3e} | ldj err max {el.max():.3e} mean {el.mean():.3e} '
          f'| rows ldj err>1e-4: {(el > 1e-4).sum().item()}  >1e-3: {(el > 1e-3).sum().item()}')
# per-element ld via diag for the worst rows: use oracle elementwise
el = (lt.cpu().double() - l64).abs().view(-1)
worst = torch.topk(el, 5).indices
tr = s64[0]['transform']
mask = O.make_mask(s64[0]['mask'], d).double()
z = x.double() * mask
p = O.transform_params(tr, z, torch.float64)
_, ld_inv_el, bins = O.rqs(x.double(), p[0], p[1], p[2], True, -4., 4., return_bins=True) if kind == 'quadratic' else O.cubic(x.double(), p[0], p[1], p[2], True, -4., 4., return_bins=True)
for r in worst.tolist():
    print('row', r, 'ldj tensor', lt[r].item(), 'generic', lg[r].item(), 'o64', l64[r].item(), 'o32', l32[r].item())
    dx = (xt[r].cpu().double() - x64[r]).abs()
    j = int(dx.argmax())
    print('   worst x dim', j, 'err', dx[j].item(), 'x in', x[r, j].item(), 'bin', bins[r, j].item(),
          'x64', x64[r, j].item(), 'xt', xt[r, j].item(), 'xg', xg[r, j].item())
