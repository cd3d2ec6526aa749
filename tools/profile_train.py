"""Kernel-time split of one hybrid NLL training step (torch.profiler, CUDA activity)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import stribor_b200 as st
from stribor_b200.parallel import DataParallelNLL

d, K, L = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 16, 8
rows = 1 << 18
dev = torch.device('cuda')
torch.manual_seed(0)
layers = [st.Coupling(st.Spline(d, K, latent_net=st.net.MLP(d, [64], d * 47), lower=-4, upper=4, spline_type='quadratic'),
                      mask=('ordered_right_half', 'ordered_left_half')[i % 2]) for i in range(L)]
flow = st.NormalizingFlow(st.UnitNormal(d), layers).to(dev)
y = torch.randn(rows, d, device=dev)
dp = DataParallelNLL(flow, micro_rows=1 << 18)
dp.step(y, rows)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    dp.step(y, rows)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=70))
