"""One fused-backward call on a small case (debugging aid): prints gradient errors vs the hybrid path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
DEV = 'cuda'
d, rows, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 1
case = cases._mk_flow('quadratic', d, [64], L, 16, rows, 4100 + d, masks=cases.ALT, lower=-4., upper=4., scale=1.5)()
spec, x = case['spec'], case['inputs']['x']


def run(hybrid, gnet):
    os.environ['STRIBOR_B200_TRAIN_HYBRID'] = '1' if hybrid else '0'
    os.environ['STRIBOR_B200_TRAIN_GNET'] = '1' if gnet else '0'
    layers = [l.to(DEV) for l in layers_from_spec(spec)]
    flow = st.NormalizingFlow(st.UnitNormal(d), layers)
    xg = x.to(DEV).clone().requires_grad_(True)
    (-flow.log_prob(xg).mean()).backward()
    torch.cuda.synchronize()
    return [xg.grad.cpu()] + [p.grad.cpu() for p in flow.parameters()]


ref = run(True, False)
for name, gnet in (('gnet', True), ('fused', False)):
    got = run(False, gnet)
    for i, (g, r) in enumerate(zip(got, ref)):
        err = (g - r).abs().max().item()
        print(f'{name} tensor {i} shape {tuple(g.shape)} max|ref| {r.abs().max().item():.3e} max err {err:.3e} rel {err / max(r.abs().max().item(), 1e-30):.2e}')
    if name == 'gnet':
        e = (got[0] - ref[0]).abs()
        rows_bad = (e.max(1).values > 1e-5 + 1e-3 * ref[0].abs().max(1).values).nonzero().view(-1)
        print('bad rows', rows_bad.numel(), rows_bad[:40].tolist())
        if rows_bad.numel():
            r = int(rows_bad[0])
            cols = (e[r] > 1e-6).nonzero().view(-1)
            print(' row', r, 'bad cols', cols.tolist()[:64])
            print(' got', got[0][r, cols[:8]].tolist(), 'ref', ref[0][r, cols[:8]].tolist())
            print(' x  ', x[r, cols[:8]].tolist())
