#!/usr/bin/env python
"""Copies what tools/collect_profiles.sh left in gpurun_out/ (tag) into profiles/ under the round's names and prints the
launch shares.  usage: tools/publish_profiles.py TAG ROUNDPREFIX (e.g. r1h r1)"""
import collections, csv, json, os, shutil, subprocess, sys
tag, rp = sys.argv[1], sys.argv[2]
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
g = lambda n: os.path.join(R, "gpurun_out", n)
p = lambda n: os.path.join(R, "profiles", n)
shutil.copy(g("bench_%s_n1.json" % tag), p("%s_bench_n1.json" % rp))
shutil.copy(g("bench_%s_spec.json" % tag), p("%s_bench_n1_spec_mode.json" % rp))
shutil.copy(g("bench_%s_ref.json" % tag), p("%s_bench_reference_arm.json" % rp))
shutil.copy(g("launches_%s.csv" % tag), p("%s_launches.csv" % rp))
tool = os.path.join(R, "tools", "ncu_table.py")
open(p("%s_ncu_kernels.csv" % rp), "w").write(subprocess.run([sys.executable, tool, g("prof_step_%s.ncu-rep" % tag), g("prof_scan_%s.ncu-rep" % tag)], capture_output=True, text=True).stdout)
open(p("%s_ncu_kernels_spec_mode.csv" % rp), "w").write(subprocess.run([sys.executable, tool, g("prof_spec_%s.ncu-rep" % tag)], capture_output=True, text=True).stdout)
out = {}
for mode, f in (("compat", "%s_ncu_kernels.csv" % rp), ("spec", "%s_ncu_kernels_spec_mode.csv" % rp)):
    t = {}
    for r in csv.DictReader(open(p(f))):
        t.setdefault(r["kernel"].split("<")[0], []).append(float(r["dram_rd"]) + float(r["dram_wr"]))
    out[mode] = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in t.items()}
out["_source"] = "profiles/%s_ncu_kernels*.csv: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, timed resident step of bench.py (16 GOPs: 16 pictures per launch)" % rp
out["_pictures_per_launch"] = 16
json.dump(out, open(p("%s_ncu_traffic.json" % rp), "w"), indent=1)
rows = [r for r in csv.reader(open(p("%s_launches.csv" % rp))) if len(r) > 14 and r[0] != "ID"]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("h264r::", "")
    tot[name] += float(r[14]); cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(): print("%-28s n=%3d  %8.3f ms  %5.1f%%  avg %.1f us" % (k, cnt[k], v / 1e6, 100 * v / s, v / cnt[k] / 1e3))
print("launches", len(rows), "sum ms", s / 1e6)
for f in ("%s_bench_n1.json" % rp, "%s_bench_n1_spec_mode.json" % rp, "%s_bench_reference_arm.json" % rp):
    d = json.loads(open(p(f)).read().strip().split("\n")[-1])
    print(f, "value %.0f" % d["value"], "ms %.2f" % d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "roofline", d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("traffic"), d.get("cpu_baseline", {}).get("value"))
