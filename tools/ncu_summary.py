"""Text summary of an .ncu-rep (run here, no GPU needed): key raw metrics + hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio']


def run(args):
    return subprocess.run(['ncu', '-i', rep] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


raw = list(csv.reader(io.StringIO(run(['--page', 'raw', '--csv']))))
hdr, units = raw[0], raw[1]
for k, row in enumerate(raw[2:]):
    name = row[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f'== launch {k}: {name}')
    for h, u, v in zip(hdr, units, row):
        if h in KEYS:
            print(f'   {h:75s} {v} {u}')
src = list(csv.reader(io.StringIO(run(['--page', 'source', '--print-source', 'cuda,sass', '--csv']))))
out, cur, hd = [], None, None
for r in src:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hd = r
        ci, si = hd.index('Instructions Executed'), hd.index('# Samples')
    elif hd and r[0].isdigit() and len(r) > ci and r[ci].isdigit():
        out.append((int(r[ci]), int(r[si]) if r[si].isdigit() else 0, cur, r[0], r[1].strip()[:95]))
ti, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f'\n== hottest source lines (first profiled launch): {ti} warp-instructions, {ts} stall samples')
for o in sorted(out, reverse=True)[:30]:
    print(f'  {100*o[0]/ti:5.1f}% inst {100*o[1]/ts:5.1f}% smpl  {o[2]}:{o[3]:>4s}  {o[4]}')
