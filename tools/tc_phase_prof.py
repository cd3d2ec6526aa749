"""Per-warp phase clocks of the tensor-core kernel (profiling build with STB_TC_EXP bit 128).

    python tools/build_variants.py prof:STB_TC_EPI_PER_SUB=2,STB_TC_EXP=128
    STRIBOR_B200_LIB=$PWD/variants/lib_prof.so python tools/tc_phase_prof.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import torch
import cases
import stribor_b200 as st
from stribor_b200 import _lib
from stribor_b200.spec import layers_from_spec

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
NL = int(os.environ.get('L', '1'))
case = cases._mk_flow('quadratic', 64, [64], NL, 16, 16, 7, masks=cases.ALT, lower=-4., upper=4., scale=1.0)()
layers = [l.to('cuda') for l in layers_from_spec(case['spec'])]
flow = st.NormalizingFlow(st.UnitNormal(64), layers)
x = torch.randn(rows, 64, device='cuda')
with torch.no_grad():
    for _ in range(3):
        lp = flow.log_prob(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    lp = flow.log_prob(x)
e1.record(); torch.cuda.synchronize()
print(f'{NL} layer(s), {rows} rows: {e0.elapsed_time(e1):.3f} ms')
n = 160 * 32 * 8
buf = (ctypes.c_uint32 * n)()
lib = _lib.lib()
rc = lib.stb_tc_prof_read(buf, n)
a = np.frombuffer(buf, dtype=np.uint32).reshape(160, 32, 8).astype(np.float64)
nw = int(os.environ.get('NW', '19'))
if os.environ.get('HEAD'):
    names_e = ['stage x tile (+barrier)', 'A1 split', 'wait acc1_full (GEMM1)', 'tanh + h', 'tile tail', 'chunk loop']
else:
  names_e = ['tile head', 'wait acc_full', 'locate(ld..release)', 'finish', 'tile tail', 'chunk-loop overhead']
names_i = ['wait a1_ready', 'issue GEMM1', 'wait b_full', 'wait h_ready', 'wait acc_empty', 'issue GEMM2+commit']
blocks = a[:148]
tl = rows / 256 / 148 * NL
print(f'tile-layers per CTA: {tl:.1f}')
tot = blocks[:, 3:nw, :6].sum(-1).mean()
print('epilogue warps: mean total clocks', tot)
for i, nm in enumerate(names_e):
    v = blocks[:, 3:nw, i]
    print(f'  {nm:24s} mean {v.mean():12.0f} ({100 * v.mean() / tot:5.1f} %)  min {v.min():10.0f} max {v.max():10.0f}  per tile-layer {v.mean() / tl:8.0f}')
for w in (1, 2):
    tot = blocks[:, w, :6].sum(-1).mean()
    print(f'issuer warp {w}: total {tot:.0f}')
    for i, nm in enumerate(names_i):
        v = blocks[:, w, i]
        print(f'  {nm:24s} mean {v.mean():12.0f} ({100 * v.mean() / tot:5.1f} %)')
