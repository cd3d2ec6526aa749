"""Print the SASS of the chunk loop found by sass_loop.py:  python tools/sass_dump.py <file> <kernel-substring>"""
import sys
from sass_loop import kernels, loop_of

for name, insts in kernels(sys.argv[1]).items():
    if sys.argv[2] in name:
        lp = loop_of(insts)
        for a, t in insts:
            if lp[0] <= a <= lp[1]:
                print(f'{a:05x}  {t}')
