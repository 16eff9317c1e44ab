#!/usr/bin/env python
"""Brief per-kernel summary of an .ncu-rep: tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('###', d.get('Kernel Name', '')[:100])
    for k in keys:
        if k in d: print(f'  {k:75s} {d[k]}')
    st = sorted(((float(d[k] or 0), k) for k in stall), reverse=True)[:7]
    for v, k in st: print(f'  stall {k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:28s} {v:.2f}')
