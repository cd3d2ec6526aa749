"""Top SASS instructions of an ncu source-page CSV by stall samples, with their dominant stall reasons.

    ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_sass.py src.csv [top]
"""
import csv
import sys


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr, sass = None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[2].strip():
        sass.append(r)
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(num(r[si]) for r in sass)
agg = {}
for r in sass:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + num(r[i])
print('stall totals:', sorted(((v, k) for k, v in agg.items()), reverse=True)[:12])
for r in sorted(sass, key=lambda r: -num(r[si]))[:top]:
    st = sorted(((num(r[i]), hdr[i][6:]) for i in stall_cols), reverse=True)[:3]
    print(f"{100 * num(r[si]) / tot:5.2f}%  {r[2]:>6s} {r[3][:72]:72s} {st}")
