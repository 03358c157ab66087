#!/usr/bin/env python
"""SASS opcode summary of lib/libh264r.so per kernel (cuobjdump -sass): instruction count and the opcodes that show what
a kernel is made of -- IDP.4A (dp4a interpolation), PRMT / SHF (byte shuffling), UTMALDG / SYNCS (tensor copies and their
mbarrier), LDGSTS, REDUX, ATOM.  usage: tools/sass_summary.py [lib] > profiles/rN_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "videodecoder_b200", "lib", "libh264r.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, ops = None, {}
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("h264r::", "").replace("(bool)", "")
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and kern:
        ops[kern][m.group(1)] += 1
keys = ["IDP.4A", "PRMT", "SHF", "LOP3", "IMAD", "LDG", "LDS", "STS", "STG", "UTMALDG", "SYNCS", "LDGSTS", "REDUX", "ATOM", "RED", "STL", "LDL"]
print("%-34s %6s  %s" % ("kernel", "instrs", "  ".join("%s" % k for k in keys)))
for k, c in ops.items():
    tot = sum(c.values())
    def n(key):
        return sum(v for o, v in c.items() if o == key or o.startswith(key + "."))
    print("%-34s %6d  %s" % (k[:34], tot, "  ".join("%*d" % (len(key), n(key)) for key in keys)))
