#!/usr/bin/env python
"""Key metrics per launch of an .ncu-rep (ncu --set full): python tools/ncu_summary.py file.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct']
stall = [k for k in h if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio')]
for row in r[2:]:
    name = row[h.index('Kernel Name')]
    if rx and not rx.search(name): continue
    print(name[:90])
    for k in keys:
        if k in h: print("   %-62s %s" % (k, row[h.index(k)]))
    st = sorted(((float(row[h.index(k)] or 0), k) for k in stall), reverse=True)[:7]
    print("   stalls/issue: " + ", ".join("%s %.2f" % (k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v) for v, k in st))
