"""Per-warp phase clocks of the affine CHAIN kernel (tc_mlp.cu, profiling build with -DSTB_TCM_PROF).

    python tools/build_variants.py tcmprof:STB_TCM_PROF
    STRIBOR_B200_LIB=$PWD/variants/lib_tcmprof.so python tools/tcm_phase_prof.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import torch
import stribor_b200 as st
from stribor_b200 import _lib

d, B, T, nl = 16, 65536, 64, 4
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    flow(x, t=t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'NeuralFlow forward, {B * T} rows x {nl} layers: {ms:.3f} ms')
n = 160 * 32 * 8
buf = (ctypes.c_uint32 * n)()
lib = _lib.lib()
lib.stb_tcm_prof_read(buf, n)
a = np.frombuffer(buf, dtype=np.uint32).reshape(160, 32, 8).astype(np.float64)[:148]
V = int(os.environ.get('V', '2'))
names_e = ['stage x (+barrier)', 'A1 split + arrive', 'wait acc (GEMM1)', 'activation + arrive', 'wait acc (GEMM3)',
           'affine output', 'layer barrier', 'tile tail (ldj, store y)']
names_i = ['wait a_ready (A1)', 'issue GEMM1 + commit', 'wait a_ready (h)', 'issue GEMM3 + commit']
epi = np.concatenate([a[:, v * 12 + 2: v * 12 + 10, :] for v in range(V)], 1)
tot = epi.sum(-1).mean()
print(f'epilogue warps: mean total clocks {tot:.0f}  (kernel {ms * 1e-3 * 1.92e9:.0f} clocks at 1.92 GHz)')
for i, nm in enumerate(names_e):
    v = epi[:, :, i]
    print(f'  {nm:28s} mean {v.mean():12.0f} ({100 * v.mean() / tot:5.1f} %)  min {v.min():10.0f} max {v.max():10.0f}')
iss = np.stack([a[:, v * 12 + 1, :4] for v in range(V)], 1)
tot = iss.sum(-1).mean()
print(f'issuer warps: mean total clocks {tot:.0f}')
for i, nm in enumerate(names_i):
    v = iss[:, :, i]
    print(f'  {nm:28s} mean {v.mean():12.0f} ({100 * v.mean() / tot:5.1f} %)')
tiles_per_vc = B * T / 128 / (148 * V)
print(f'tile-layers per virtual CTA: {tiles_per_vc * nl:.0f}; clocks per tile-layer: {epi.sum(-1).mean() / (tiles_per_vc * nl):.0f}')
