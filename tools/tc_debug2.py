import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from golden_util import close_or_arbitrated
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
DEV = 'cuda'
kind, d, masks = 'quadratic', 64, cases.ALT
case = cases._mk_flow(kind, d, [64], 3, 16, 700, 900 + d, masks=masks, lower=-4., upper=4., scale=1.7)()
spec = case['spec']
x = case['inputs']['x'].to(DEV)
if len(sys.argv) > 1:
    x[0, 0], x[1, 0], x[2, 1] = 4.0, -4.0, 5.5
tl = [l.to(DEV) for l in layers_from_spec(spec)]
os.environ['STRIBOR_B200_FORCE_GENERIC'] = '1'
gl = [l.to(DEV) for l in layers_from_spec(spec)]
os.environ['STRIBOR_B200_FORCE_GENERIC'] = '0'
s64 = O.spec_to(spec, torch.float64)
xc = x.cpu()
cur_t, cur_g, cur_o, cur_64 = x, x, xc, xc.double()
with torch.no_grad():
    for li in (2, 1, 0):
        # feed every implementation the SAME input (the fp32 oracle's) to isolate the layer
        inp = cur_o
        xt, lt = tl[li].inverse_and_log_det_jacobian(inp.to(DEV))
        xg, lg = gl[li].inverse_and_log_det_jacobian(inp.to(DEV))
        xo, lo = O.layer_apply(spec[li], inp, inverse=True)
        x6, l6 = O.layer_apply(s64[li], inp.double(), inverse=True)
        for tag, ll, xx in (('tensor', lt, xt), ('generic', lg, xg), ('oracle32', lo, xo)):
            el = (ll.cpu().double() - l6).abs().view(-1)
            ex = (xx.cpu().double() - x6).abs().max(-1).values
            bad = (el > 1e-3).nonzero().view(-1).tolist()
            print(f'layer {li} {tag:9s} ldj err max {el.max():.3e} mean {el.mean():.3e} x err max {ex.max():.3e} bad rows {bad[:12]}')
            for r in bad[:4]:
                print('      row', r, 'ldj', ll.view(-1)[r].item(), 'o64', l6.view(-1)[r].item(), 'xmax in', inp[r].abs().max().item(),
                      'x err', ex[r].item())
        cur_o = xo
