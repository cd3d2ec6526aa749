import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
DEV = 'cuda'
kind, d = sys.argv[1], int(sys.argv[2])
masks = tuple(sys.argv[3].split(',')) if len(sys.argv) > 3 else cases.ALT
case = cases._mk_flow(kind, d, [64], 3, 16, 700, 900 + d, masks=masks, lower=-4., upper=4., scale=1.7)()
spec = case['spec']
x = case['inputs']['x'].to(DEV)
x[2, 1] = 5.5
s64 = O.spec_to(spec, torch.float64)
xc = x.cpu()
def run(force):
    os.environ['STRIBOR_B200_FORCE_GENERIC'] = '1' if force else '0'
    layers = [l.to(DEV) for l in layers_from_spec(spec)]
    f = st.NormalizingFlow(st.UnitNormal(d), layers)
    with torch.no_grad():
        lp = f.log_prob(x)
        outs = []
        cur = x
        for l in reversed(layers):
            cur, ld = l.inverse_and_log_det_jacobian(cur)
            outs.append((cur.cpu(), ld.cpu()))
    return lp.cpu(), outs
lp_t, o_t = run(False)
lp_g, o_g = run(True)
lp32 = O.flow_log_prob(spec, xc); lp64 = O.flow_log_prob(s64, xc.double())
et = (lp_t.double() - lp64).abs().view(-1); eg = (lp_g.double() - lp64).abs().view(-1); e32 = (lp32.double() - lp64).abs().view(-1)
print('lp err  tensor max %.3e  generic max %.3e  oracle32 max %.3e' % (et.max(), eg.max(), e32.max()))
w = int(et.argmax())
print('worst row', w, 'x', xc[w].tolist(), 'lp t/g/32/64', lp_t[w].item(), lp_g[w].item(), lp32[w].item(), lp64[w].item())
cur32, cur64 = xc, xc.double()
for i, li in enumerate((2, 1, 0)):
    a32, l32 = O.layer_apply(spec[li], cur32, inverse=True)
    a64, l64 = O.layer_apply(s64[li], cur64, inverse=True)
    print(' layer', li, 'x_out t/g/32/64', o_t[i][0][w].tolist(), o_g[i][0][w].tolist(), a32[w].tolist(), a64[w].tolist())
    print('          ldj  t/g/32/64', o_t[i][1][w].item(), o_g[i][1][w].item(), l32[w].item(), l64[w].item())
    dt = (o_t[i][0][w].double() - a64[w]).abs(); j = int(dt.argmax())
    print('          worst dim', j, 'in t/64', (cur64[w, j].item()), 'out t/g/32/64', o_t[i][0][w, j].item(), o_g[i][0][w, j].item(), a32[w, j].item(), a64[w, j].item())
    cur32, cur64 = a32, a64
