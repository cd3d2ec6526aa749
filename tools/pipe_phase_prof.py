"""Per-warp phase clocks of tc_mlp_pipe_kernel (tc_mlp.cu, profiling build with -DSTB_TCM_PROF).

    python tools/build_variants.py tcmprof:STB_TCM_PROF
    STRIBOR_B200_LIB=$PWD/variants/lib_tcmprof.so python tools/pipe_phase_prof.py [H] [n_hidden]
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import torch
import stribor_b200 as st
from stribor_b200 import _lib

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d, B = 64, 1 << 20
torch.manual_seed(5)
layer = st.Coupling(st.Affine(d, latent_net=st.net.MLP(d, [H] * nh, 2 * d)), mask='ordered_right_half').to('cuda').requires_grad_(False)
flow = st.NormalizingFlow(st.UnitNormal(d), [layer])
x = torch.randn(B, d, device='cuda')
with torch.no_grad():
    for _ in range(3):
        flow.forward_and_log_det_jacobian(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    flow.forward_and_log_det_jacobian(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
tiles = B / 128 / 148
print(f'affine coupling MLP{[H] * nh}, {B} rows: {ms:.3f} ms = {ms * 1e-3 * 1.965e9 / tiles:.0f} clocks per 128-row tile at 1.965 GHz')
n = 160 * 32 * 8
buf = (ctypes.c_uint32 * n)()
_lib.lib().stb_tcm_prof_read(buf, n)
a = np.frombuffer(buf, dtype=np.uint32).reshape(160, 32, 8).astype(np.float64)[:148]
names_e = ['stage x (+barrier)', 'A1 split + arrive', 'wait acc1 (GEMM1)', 'activation 1 (waves)', 'wait acc2 (GEMM2 tail)',
           'activation 2 (waves)', 'wait acc3 (GEMM3 tail)', 'affine output, ldj, store y']
names_i = ['wait a1_ready', 'issue GEMM1', 'GEMM2: wait wave', 'GEMM2: wait weights', 'GEMM2: issue + commit',
           'GEMM3: wait wave', 'GEMM3: wait weights', 'GEMM3: issue + commit']
epi = a[:, 2:18, :]
print('epilogue warps, clocks per tile:')
for i, nm in enumerate(names_e):
    v = epi[:, :, i] / tiles
    print(f'  {nm:30s} mean {v.mean():8.0f}  min {v.min():8.0f} max {v.max():8.0f}')
print(f'  total {epi.sum(-1).mean() / tiles:.0f}')
iss = a[:, 1, :]
print('issuer thread, clocks per tile:')
for i, nm in enumerate(names_i):
    print(f'  {nm:30s} mean {iss[:, i].mean() / tiles:8.0f}')
print(f'  total {iss.sum(-1).mean() / tiles:.0f}')
