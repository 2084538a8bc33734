"""Summarise an .ncu-rep: headline metrics per kernel + top stall locations (needs ncu on PATH; runs without a GPU)."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__average_warps_issue_stalled',
        'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_alu', 'sm__pipe_fma_cycles_active', 'sm__pipe_alu_cycles_active',
        'l1tex__data_bank_conflicts', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    print('=' * 20, r[hdr.index('Kernel Name')][:90])
    for h, v in zip(hdr, r):
        if any(h.startswith(w) for w in want) and 'pcsamp' not in h and v not in ('0', ''):
            print(f'  {h} = {v}')
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # first kernel only
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hi]
    data = []
    for r in rows[hi + 1:]:
        if not r or r[0] == 'Kernel Name':
            break
        data.append(r)
    si, ei = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    tot = sum(int(r[si]) for r in data)
    print('total samples', tot, 'warp-instr', sum(int(r[ei]) for r in data))
    for r in sorted(data, key=lambda r: -int(r[si]))[:int(sys.argv[2])]:
        print(r[si].rjust(7), r[ei].rjust(10), r[1].strip()[:110])
    c, s = Counter(), Counter()
    for r in data:
        op = [o for o in r[1].strip().split() if not o.startswith('@')]
        op = op[0] if op else ''
        c[op] += int(r[ei]); s[op] += int(r[si])
    print('--- opcode: executed, samples')
    for op, n in c.most_common(22):
        print(' ', op.ljust(30), n, s[op])
