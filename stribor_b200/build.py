"""Build libstribor_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m stribor_b200.build [--force]
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libstribor_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(os.path.dirname(HERE), 'include', 'stribor_b200.h')]


def up_to_date() -> bool:
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(p) <= t for p in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str = OUT) -> str:
    """`defines` / `out`: profiling variants only (tools/build_variants.py); the product is the default."""
    if not force and out == OUT and up_to_date():
        return OUT
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + [f'-D{d}' for d in defines] + ['-o', out] + sources() + ['-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, 'build.log')
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + r.stdout)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed ({r.returncode}); see {log}')
    return out


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
