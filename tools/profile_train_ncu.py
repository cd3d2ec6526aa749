"""One NLL training step at d = 128 (for `ncu -k regex:tc_wide`): python tools/profile_train_ncu.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stribor_b200 as st
from stribor_b200.parallel import DataParallelNLL

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
d, K, L = 128, 16, 2
dev = torch.device('cuda')
torch.manual_seed(0)
layers = [st.Coupling(st.Spline(d, K, latent_net=st.net.MLP(d, [64], d * 47), lower=-4, upper=4, spline_type='quadratic'),
                      mask=('ordered_right_half', 'ordered_left_half')[i % 2]) for i in range(L)]
flow = st.NormalizingFlow(st.UnitNormal(d), layers).to(dev)
y = torch.randn(rows, d, device=dev)
dp = DataParallelNLL(flow, micro_rows=rows)
for _ in range(2):
    dp.step(y, rows)
torch.cuda.synchronize()
