"""Whole-flow kernel at small batch: one 256-row CTA per SM vs two 128-row CTAs per SM (STRIBOR_B200_PAIR)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    import torch, cases
    import stribor_b200 as st
    from stribor_b200.spec import layers_from_spec
    case = cases._mk_flow('quadratic', 64, [64], 8, 16, 16, 7, masks=cases.ALT, lower=-4., upper=4., scale=1.0)()
    flow = st.NormalizingFlow(st.UnitNormal(64), [l.to('cuda') for l in layers_from_spec(case['spec'])])
    for rows in (4096, 16384, 37888, 75776, 1 << 20):
        x = torch.randn(rows, 64, device='cuda')
        with torch.no_grad():
            for _ in range(5):
                flow.log_prob(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                flow.log_prob(x)
            e1.record(); torch.cuda.synchronize()
        print(f'  rows {rows:8d}: {e0.elapsed_time(e1) / 20 * 1e3:9.1f} us per log_prob')
else:
    for v in ('0', '1'):
        print('STRIBOR_B200_PAIR=' + v, flush=True)
        subprocess.run([sys.executable, __file__, 'child'], env=dict(os.environ, STRIBOR_B200_PAIR=v))
