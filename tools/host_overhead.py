"""Host-side cost of one flow.log_prob call (small batch: the kernel takes ~200 us, everything else is Python / ctypes)."""
import os, sys, cProfile, pstats, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch, cases
import stribor_b200 as st
from stribor_b200.spec import layers_from_spec
case = cases._mk_flow('quadratic', 64, [64], 8, 16, 16, 7, masks=cases.ALT, lower=-4., upper=4., scale=1.0)()
flow = st.NormalizingFlow(st.UnitNormal(64), [l.to('cuda') for l in layers_from_spec(case['spec'])])
x = torch.randn(4096, 64, device='cuda')
with torch.no_grad():
    for _ in range(10):
        flow.log_prob(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        flow.log_prob(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'host time per call {1e6 * (t1 - t0) / 200:.1f} us; drained after {1e6 * (t2 - t1):.1f} us more')
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        flow.log_prob(x)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
