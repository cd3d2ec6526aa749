"""Source lines of an .ncu-rep sorted by STALL SAMPLES (where warps wait), not by instructions executed:
    python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
out, cur, hd = [], None, None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hd = r
        ci, si = hd.index('Instructions Executed'), hd.index('# Samples')
    elif hd and r[0].isdigit() and len(r) > ci and r[ci].isdigit():
        out.append((int(r[si]) if r[si].isdigit() else 0, int(r[ci]), cur, r[0], r[1].strip()[:105]))
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f'{ts} stall samples, {ti} warp-instructions')
for o in sorted(out, reverse=True)[:top]:
    print(f'{100 * o[0] / ts:5.1f}% smpl {100 * o[1] / ti:5.1f}% inst  {o[2]}:{o[3]:>4s}  {o[4]}')
