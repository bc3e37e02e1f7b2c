#!/usr/bin/env python
"""ncu_brief.py <report.ncu-rep> [kernel-substring]: duration, occupancy, issue, pipes and stall reasons per kernel."""
import csv, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for row in r[2:]:
    d = dict(zip(h, row))
    if pat not in d['Kernel Name']:
        continue
    print(d['Kernel Name'][:110])
    print('   ', {k.split('.')[0]: d.get(k) for k in keys})
    st = {}
    for n, v in d.items():
        if 'issue_stalled' in n and n.endswith('per_issue_active.ratio'):
            try:
                if float(v) > 0.4: st[n.split('issue_stalled_')[1].split('_per')[0]] = round(float(v), 2)
            except ValueError:
                pass
    print('    stalls/issue', st)
    pp = {}
    for n, v in d.items():
        if n.startswith('sm__inst_executed_pipe_') and n.endswith('.avg.pct_of_peak_sustained_active'):
            try:
                if float(v) > 4: pp[n.split('pipe_')[1].split('.')[0]] = round(float(v), 1)
            except ValueError:
                pass
    print('    pipes %', pp)
