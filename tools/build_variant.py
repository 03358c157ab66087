#!/usr/bin/env python
"""Builds lib/variants/libh264r_<tag>.so with extra nvcc flags (tuning runs; load with H264R_LIB=...).
usage: tools/build_variant.py tag -DH264R_INTER_MIN_CTAS=6 ..."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from videodecoder_b200 import build
tag, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(build.LIB, "variants", "libh264r_%s.so" % tag)
os.makedirs(os.path.dirname(out), exist_ok=True)
srcs = build._glob(os.path.join(build.CSRC, "cuda"), (".cu",))
build._run(["nvcc"] + build.NVCC_FLAGS + extra + ["-I", os.path.join(build.ROOT, "include")] + srcs + ["-o", out, "-lcudart"],
           log=os.path.join(build.LIB, "variants", "nvcc_%s.log" % tag))
print(out)
