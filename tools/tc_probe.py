"""GPU bring-up probe for the tcgen05 building blocks: prints the error of stb_tc_selftest against
torch for every (mode, variant, K, N).  Run under gpurun."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stribor_b200 import _lib

lib = _lib.lib()
dev = torch.device('cuda:0')
torch.manual_seed(0)
for mode in (0, 1, 2, 3):
    tf32 = mode & 1
    for variant in (0, 2):
        for K, N in ((64, 96), (64, 96), (64, 256), (64, 48)):
            if N % 16:
                continue
            A = (torch.rand(128, K, device=dev) * 2 - 1)
            B = (torch.rand(N, K, device=dev) * 2 - 1)
            D = torch.full((128, N), float('nan'), device=dev)
            rc = lib.stb_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, mode, variant,
                                     torch.cuda.current_stream().cuda_stream)
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print('mode', mode, 'variant', variant, K, N, 'CUDA ERROR', e)
                sys.exit(1)
            if rc:
                print('mode', mode, 'variant', variant, K, N, 'rc', rc, lib.stb_last_error())
                continue
            ref = A.double() @ B.double().t()
            err = (D.double() - ref).abs().max().item()
            e32 = ((A @ B.t()).double() - ref).abs()
            ed = (D.double() - ref)
            print(f'    fp32 matmul max err {e32.max().item():.3e} mean {e32.mean().item():.3e} | ours mean abs {ed.abs().mean().item():.3e} mean signed*sign(ref) {(ed * ref.sign()).mean().item():.3e}')
            print(f'mode {mode} ({"tf32" if tf32 else "fp16"}{" split3" if mode >> 1 else ""}) variant {variant} '
                  f'K={K} N={N}: max abs err {err:.3e}  (ref max {ref.abs().max().item():.2f})', flush=True)
