#!/usr/bin/env python
"""One CSV line per launch of an .ncu-rep captured with `ncu --set full`: duration, instructions, occupancy, issue-slot use,
DRAM bytes and the stall breakdown (cycles per issued instruction).  usage: tools/ncu_table.py a.ncu-rep [b.ncu-rep ...] > table.csv"""
import csv, subprocess, sys, re
cols = [("us", "gpu__time_duration.sum"), ("warp_inst", "smsp__inst_executed.sum"), ("regs", "launch__registers_per_thread"),
        ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"), ("l2_hit%", "lts__t_sector_hit_rate.pct")]
stalls = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "no_instruction", "membar", "sleeping",
          "branch_resolving", "lg_throttle", "mio_throttle", "barrier"]
print("kernel,grid,blk," + ",".join(c for c, _ in cols) + "," + ",".join("st_" + s for s in stalls) + ",dram_units")
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, units = r[0], r[1]
    for row in r[2:]:
        name = row[h.index("Kernel Name")]
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("h264r::", "").replace("(bool)", "").replace(", ", ";")
        v = [name, row[h.index("launch__grid_size")], row[h.index("launch__block_size")]]
        for _, k in cols:
            x = row[h.index(k)] if k in h else ""
            u = units[h.index(k)] if k in h else ""
            if x and k == "gpu__time_duration.sum": x = "%.3f" % (float(x) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0))
            if x and k.startswith("dram__bytes"): x = "%.0f" % (float(x) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0))
            v.append(x)
        for s in stalls:
            k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            v.append("%.2f" % float(row[h.index(k)]) if k in h and row[h.index(k)] else "")
        v.append("byte")
        print(",".join(v))
