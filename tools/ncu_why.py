"""Where do the WORKING warps of a kernel wait, and why?  Per SASS instruction ncu gives stall samples by reason;
this groups them by CUDA source line and prints the dominant reasons, plus the reason totals.

    python tools/ncu_why.py gpurun_out/x.ncu-rep [top]
"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
lines = collections.OrderedDict()
cur_file, hd = None, None
tot = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hd = r
        cols = {n: i for i, n in enumerate(hd)}
        stalls = [n for n in hd if n.startswith('stall_') and 'Not Issued' not in n]
        continue
    if hd and r[0].isdigit() and len(r) >= len(hd):
        try:
            d = {n: int(r[cols[n]] or 0) for n in stalls}
            ie = int(r[cols['Instructions Executed']] or 0)
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        e = lines.setdefault(key, {'src': r[1].strip()[:95], 'ie': 0, 'st': collections.Counter()})
        e['ie'] += ie
        e['st'].update(d)
        tot.update(d)
T = sum(tot.values()) or 1
IE = sum(e['ie'] for e in lines.values()) or 1
print(f'{T} stall samples, {IE} warp-instructions')
print('reasons: ' + ', '.join(f'{n[6:]} {100 * v / T:.1f}%' for n, v in tot.most_common(9)))
for (f, ln), e in sorted(lines.items(), key=lambda kv: -sum(kv[1]['st'].values()))[:top]:
    s = sum(e['st'].values())
    why = ' '.join(f'{n[6:]}:{100 * v / max(s, 1):.0f}%' for n, v in e['st'].most_common(3) if v)
    print(f'{100 * s / T:5.1f}% smpl {100 * e["ie"] / IE:5.1f}% inst  {f}:{ln:<5d} {e["src"]:95s} [{why}]')
