"""Development tool: phase timing of the row kernels (needs a -DH264R_TRACE build of libh264r, see
h264r_kernels.cuh:WarpExecTrace).  Prints, for the warp of row 0 of the first picture of each launch, the average
clock cycles between consecutive stamps of a macroblock."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("H264R_LIB", os.path.join(ROOT, "videodecoder_b200", "lib", "variants", "libh264r_trace.so"))
from videodecoder_b200 import abi, engine, workload  # noqa: E402


def dump(L, title):
    buf = (ctypes.c_longlong * 4096)()
    n = L.h264r_debug_trace(buf, 4096)
    v = np.array(buf[:n], dtype=np.int64)
    # split per macroblock: a macroblock starts at marker -1 and ends after marker -2
    mbs, cur = [], None
    i = 0
    while i < n:
        if v[i] == -1:
            cur = [v[i + 1]]; i += 2
        elif v[i] == -2:
            if cur is not None:
                cur.append(v[i + 1]); mbs.append(cur)
            cur = None; i += 2
        else:
            if cur is not None:
                cur.append(v[i])
            i += 1
    if not mbs:
        print(title, "no samples"); return
    by_len = {}
    for m in mbs:
        by_len.setdefault(len(m), []).append(np.diff(np.array(m)))
    print(title, "macroblocks traced:", len(mbs))
    starts = np.array([m[0] for m in mbs])
    if len(starts) > 2:
        print("   start-to-start cycles (median):", int(np.median(np.diff(starts))))
    for k, d in sorted(by_len.items()):
        d = np.array(d)
        print("   %d stamps x %d MBs: median cycles per interval %s  sum %d" % (k, len(d), np.median(d, axis=0).astype(int).tolist(), int(np.median(d.sum(axis=1)))))


def main():
    compat = "--spec" not in sys.argv
    W, H = 120, 68
    cfg = workload.WorkloadConfig(compat=compat, p_t8x8=0.0, p_i8x8=0.0)
    rng = np.random.default_rng(3)
    gop = workload.make_gop(rng, W, H, 2, cfg)
    flags = abi.REF_COMPAT if compat else 0
    eng = engine.Engine(W, H, max_frames=2, flags=flags)
    L = eng.L
    L.h264r_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.h264r_debug_trace.restype = ctypes.c_int
    for rep in range(2):
        eng.submit_picture(0, gop[0]["pp"], gop[0]["hdr"], gop[0]["mv"], gop[0]["coeff"])
        eng.sync()
        dump(L, "I picture (run %d):" % rep)
        eng.submit_picture(1, gop[1]["pp"], gop[1]["hdr"], gop[1]["mv"], gop[1]["coeff"])
        eng.sync()
        dump(L, "P picture (run %d):" % rep)
    eng.close()


if __name__ == "__main__":
    main()
