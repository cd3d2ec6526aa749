"""Profiling variants of the library: python tools/build_variants.py name:DEF1,DEF2 ...  -> gpurun_out/var/lib_<name>.so
Run one with STRIBOR_B200_LIB=<path> python bench.py ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
from stribor_b200 import build as B

os.makedirs(os.path.join(ROOT, 'variants'), exist_ok=True)


def one(spec):
    name, _, defs = spec.partition(':')
    out = os.path.join(ROOT, 'variants', f'lib_{name}.so')
    B.build(force=True, defines=[d for d in defs.split(',') if d], out=out)
    return out


with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print(o)
