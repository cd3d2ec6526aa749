"""Static SASS mnemonic counts per kernel (proof of tcgen05 / TMA / packed-fp32 use), no GPU needed:
    cuobjdump -sass stribor_b200/libstribor_b200.so | python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import re
import subprocess
import sys

KEYS = ('UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UBLKPF', 'REDG', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'MUFU.EX2',
        'MUFU.RCP', 'MUFU.LG2', 'MUFU.RSQ', 'F2FP', 'STG.E.ENL2.256', 'UCGABAR')
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    for k in KEYS:
        if k in line:
            cnt[cur][k] += 1
    if 'UTCHMMA tmem' in line:
        cnt[cur]['UTCHMMA(A in TMEM)'] += 1
print('SASS mnemonic census of libstribor_b200.so (cuobjdump -sass, static instruction counts per kernel)')
print('UTCHMMA = tcgen05.mma (operands `tmem[..]` first: A read from TMEM), LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, UBLKCP / UBLKPF = cp.async.bulk copy / L2 prefetch '
      '(TMA engine), SYNCS = mbarrier, REDG = red.global, FFMA2 / FADD2 / FMUL2 = packed fp32 pairs, '
      'STG.E.ENL2.256 = 32-byte global stores\n')
for f in sorted(cnt):
    if not any(k in cnt[f] for k in ('UTCHMMA', 'UBLKCP')):
        continue
    name = subprocess.run(['c++filt', f], stdout=subprocess.PIPE, text=True).stdout.strip()
    print(name[:120])
    print('   ' + ', '.join(f'{k} {v}' for k, v in sorted(cnt[f].items())))
