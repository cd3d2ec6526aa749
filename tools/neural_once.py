"""One NeuralFlow forward at the configs[3] shape (ncu target): python tools/neural_once.py [rows_B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import stribor_b200 as st
d, B, T, nl = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 64, 4
torch.manual_seed(123)
layers = [st.ContinuousAffineCoupling(st.net.MLP(d + 1, [64], 2 * d), st.net.TimeLinear(2 * d),
                                      ('ordered_0', 'ordered_1')[i % 2]) for i in range(nl)]
flow = st.NeuralFlow(layers).to('cuda').requires_grad_(False)
x = torch.randn(B, T, d, device='cuda')
t = torch.rand(B, T, 1, device='cuda')
with torch.no_grad():
    for _ in range(3):
        flow(x, t=t)
torch.cuda.synchronize()
