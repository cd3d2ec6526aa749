"""Static instruction count of the tensor-core kernel's per-chunk epilogue loop (one spline element
per thread), read from the SASS of an object file / shared library -- the offline figure of merit
for an issue-bound epilogue.

    python tools/sass_loop.py <file.o|.so> [kernel-substring ...]

The loop is located as the innermost backward branch that encloses the LAST `LDTM` (tcgen05.ld)
of the kernel."""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(['cuobjdump', '-sass', path], stdout=subprocess.PIPE, text=True).stdout
    cur, name = None, None
    res = {}
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            cur = res.setdefault(name, [])
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    return res


def loop_of(insts):
    ldtm = [a for a, t in insts if 'LDTM' in t]
    if not ldtm:
        return None
    last = max(ldtm)
    best = None
    for a, t in insts:
        m = re.search(r'BRA(?:\.\w+)*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)', t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= last < a and (best is None or a - tgt < best[1] - best[0]):
                best = (tgt, a)
    return best


def main():
    path = sys.argv[1]
    subs = sys.argv[2:] or ['tc_spline_layer_kernel']
    for name, insts in kernels(path).items():
        if not any(s in name for s in subs):
            continue
        lp = loop_of(insts)
        if not lp:
            continue
        body = [t for a, t in insts if lp[0] <= a <= lp[1]]
        ops = collections.Counter()
        for t in body:
            t = re.sub(r'^@!?U?P\d+\s+', '', t)
            ops[t.split()[0].split('.')[0]] += 1
        top = ', '.join(f'{k} {v}' for k, v in ops.most_common(14))
        print(f'{name}: loop {lp[0]:#x}..{lp[1]:#x} = {len(body)} instructions\n    {top}')


if __name__ == '__main__':
    main()
