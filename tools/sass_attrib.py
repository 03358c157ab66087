#!/usr/bin/env python
"""Joins an ncu source-page CSV (SASS view: executed instructions per address) with nvdisasm -g line info of the same
kernel and prints executed warp instructions per source region / line.  usage: sass_attrib.py src.csv kernel.sass n_units"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ie = hdr.index('Instructions Executed'); ss = hdr.index('# Samples')
cur = None; tags = []
for l in open(sys.argv[2]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s*/\*[0-9a-f]{4,}\*/', l): tags.append((cur, l.split('*/')[1].strip()))
data = rows[2:2 + len(tags)]
units = float(sys.argv[3])
for i, ((t, op), r) in enumerate(zip(tags, data)):
    if op.split(';')[0].strip().split()[0:1] != r[1].strip().split()[0:1]:
        print("mismatch at", i, op, r[1]); break
tot = sum(int(r[ie]) for r in data); print('total inst', tot, 'per unit', tot / units)
byline = collections.Counter(); bys = collections.Counter(); byop = collections.Counter()
for (t, op), r in zip(tags, data):
    byline[t] += int(r[ie]); bys[t] += int(r[ss]); byop[op.split()[0].split('.')[0]] += int(r[ie])
src = {}
import os
for f in ['h264r_core.cuh', 'h264r_kernels.cuh', 'h264r_intra.cuh', 'h264r_mc.cuh', 'h264r_mcgrp.cuh', 'h264r_seq.cuh']:
    src[f] = open(os.path.join(os.path.dirname(__file__), '../videodecoder_b200/csrc/cuda', f)).read().split('\n')
ts = sum(bys.values())
byfile = collections.Counter()
for (f, l), v in byline.items(): byfile[f] += v
for f, v in byfile.most_common(): print('%-28s %7.1f' % (f, v / units))
n = int(sys.argv[4]) if len(sys.argv) > 4 else 60
for (f, l), v in byline.most_common(n):
    s = src[f][l - 1].strip()[:105] if f in src else ''
    print('%-14s%5d %6.1f s%4.1f%%  %s' % (f[6:], l, v / units, 100 * bys[(f, l)] / ts, s))
print(byop.most_common(30))
