"""Accuracy attribution: error of the tensor path vs the fp64 oracle under experiment flags."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import cases
from oracle import coupling_flow_oracle as O
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
DEV = 'cuda'
for d, rows in ((64, 8192), (2, 16384)):
    case = cases._mk_flow('quadratic', d, [64], 3, 16, rows, 900 + d, masks=cases.ALT, lower=-4., upper=4., scale=1.7)()
    spec = case['spec']; xc = case['inputs']['x']; x = xc.to(DEV)
    s64 = O.spec_to(spec, torch.float64)
    with torch.no_grad():
        x64, l64 = O.flow_inverse(s64, xc.double(), with_ldj=True)
        x32, l32 = O.flow_inverse(spec, xc, with_ldj=True)
    def stats(tag, xx, ll):
        el = (ll.cpu().double() - l64).abs().view(-1); ex = (xx.cpu().double() - x64).abs()
        print(f'd={d:3d} {tag:22s} ldj err mean {el.mean():.3e} p99 {el.kthvalue(int(0.99*el.numel())).values:.3e} max {el.max():.3e} | x err mean {ex.mean():.3e} max {ex.max():.3e}', flush=True)
    stats('oracle32', x32, l32)
    for flags, force in ((0, True), (0, False)):
        os.environ['STRIBOR_B200_FORCE_GENERIC'] = '1' if force else '0'
        os.environ['STRIBOR_B200_TC_FLAGS'] = str(flags)
        layers = [l.to(DEV) for l in layers_from_spec(spec)]
        f = st.NormalizingFlow(st.UnitNormal(d), layers)
        with torch.no_grad():
            xx, ll = f.inverse_and_log_det_jacobian(x)
        torch.cuda.synchronize()
        stats('generic' if force else 'tensor', xx, ll)
