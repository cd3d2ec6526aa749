"""Summarise `ncu --page source --print-source cuda,sass --csv` per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
        ci = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
    elif hdr and r[0].isdigit() and len(r) > ci and r[ci].isdigit():
        out.append((int(r[ci]), int(r[si]) if r[si].isdigit() else 0, cur_file, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print('total warp-instructions', tot, 'samples', ts)
out.sort(reverse=True)
for o in out[:top]:
    print(f'{100*o[0]/tot:5.1f}% inst {100*o[1]/max(ts,1):5.1f}% smpl  {o[2]}:{o[3]:>4s}: {o[4]}')
