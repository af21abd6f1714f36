#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full): per kernel launch the duration, DRAM traffic, tensor / issue utilisation, occupancy.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('launch__registers_per_thread', 'regs/thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts')]
lines = ['| kernel | ' + ' | '.join(n for _, n in want) + ' |', '|---|' + '---|' * len(want)]
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col['Kernel Name']].split('(')[0].replace('void ', '')
    vals = []
    for m, _ in want:
        if m in col:
            vals.append(f'{r[col[m]]} {units[col[m]]}'.strip())
        else:
            vals.append('-')
    lines.append(f'| `{name}` | ' + ' | '.join(vals) + ' |')
out = '\n'.join(lines) + '\n'
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(out)
print(out)
