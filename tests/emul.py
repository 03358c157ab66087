"""TEST INFRASTRUCTURE: builds and binds the host-side SIMT emulator of the kernel code
(tests/emul/emul.cpp).  Lets the CPU suite check the kernels' arithmetic against the oracle."""
import ctypes
import os
import subprocess

import numpy as np

from videodecoder_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "emul.cpp")
OUT = os.path.join(HERE, "emul", "_build", "libh264emul.so")
DEPS = [SRC] + [os.path.join(ROOT, "videodecoder_b200", "csrc", "cuda", f) for f in ("h264r_core.cuh", "h264r_seq.cuh", "h264r_mc.cuh", "h264r_mcgrp.cuh", "h264r_intra.cuh", "h264r_deblock.cuh")] + [
    os.path.join(ROOT, "include", "h264r.h")]

_lib = None
# test switches of the emulator (upper bits of `flags`, never seen by the engine)
PACKED = 0x40000000          # store coefficients packed (mask + offset + non-zero blocks) as h264r_submit_mbs_packed does
COMPACT = 0x20000000         # with PACKED: H264R_LEVELS_COMPACT storage (8-byte units, significance map + levels)
GROUP_RAW = 0x04000000       # with GROUP: windows staged as aligned words and shifted by the consumer (what the tensor-copy variant of k_inter_grp does)
PARTS = 0x08000000           # spec mode: inter macroblocks through the three parts of k_inter (luma 4x4 / luma 8x8 transform / chroma)
GROUP = 0x10000000           # with PACKED | COMPACT under REF_COMPAT: inter macroblocks through the group sequence of k_inter_grp
GENERIC_INTRA = 0x80000000   # run the generic intra phases of h264r_core.cuh instead of h264r_intra.cuh


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                            "-fsanitize=undefined", "-fno-sanitize-recover=undefined", SRC, "-o", OUT], check=True)
        _lib = ctypes.CDLL(OUT)
        _lib.h264emul_recon_picture.restype = ctypes.c_ulonglong
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def stream_expand(hdr8, mb_words, stream, max_units, max_mvs, strict=True, canary=None):
    """the stream form of a picture through the functions k_blk_scan / k_stream_expand call (h264r_core.cuh): 16-byte headers,
    dense masks, 8-byte units, partition-order vectors, the two offset scans, the intra list, words walked"""
    n = hdr8.shape[0]
    hdr8 = np.ascontiguousarray(hdr8, dtype=np.uint32)
    mb_words = np.ascontiguousarray(mb_words, dtype=np.uint8)
    stream = np.ascontiguousarray(stream, dtype=np.uint32)
    hdr16 = np.zeros(n, dtype=abi.MB_HDR_DTYPE)
    mask = np.zeros(n, dtype=np.uint32)
    pad = 0 if canary is None else 64
    mvbuf = np.zeros(max_mvs + 1 + pad, dtype=np.uint32)
    units = np.zeros((max_units + 2 + pad, 2), dtype=np.uint32)
    if canary is not None:          # guard words around what the expansion may write: they must come back untouched
        mvbuf[:pad] = canary
        units[max_units + 2:] = canary
    blk_off = np.zeros(n, dtype=np.uint32); mv_off = np.zeros(n, dtype=np.uint32); ilist = np.zeros(n, dtype=np.uint32)
    nu, nm, ni = ctypes.c_long(0), ctypes.c_long(0), ctypes.c_long(0)
    f = lib().h264emul_stream_expand
    f.restype = ctypes.c_long
    used = f(n, _ptr(hdr8), _ptr(mb_words), _ptr(stream), _ptr(hdr16), _ptr(mask),
             ctypes.c_void_p(mvbuf.ctypes.data + 4 * mvbuf.size), _ptr(units), _ptr(blk_off), _ptr(mv_off), _ptr(ilist),
             ctypes.byref(nu), ctypes.byref(nm), ctypes.byref(ni), 1 if strict else 0)
    if canary is not None:
        assert (mvbuf[:pad] == canary).all() and (units[max_units + 2:] == canary).all(), "the expansion wrote outside its areas"
        assert nu.value <= max_units and nm.value <= max_mvs
    mvs = mvbuf[::-1][:nm.value].copy().view(np.int16).reshape(-1, 2)
    return dict(hdr=hdr16, mask=mask, units=units[:nu.value].copy().view(np.uint8).reshape(-1, 8), mvs=mvs, blk_off=blk_off, mv_off=mv_off,
                intra_list=ilist[:ni.value].copy(), words=int(used))


def residual(w_mbs, h_mbs, flags, hdr, coeff):
    n = w_mbs * h_mbs
    hdr = np.ascontiguousarray(hdr)
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    out = np.zeros((n, 384), dtype=np.int16)
    lib().h264emul_residual(w_mbs, h_mbs, ctypes.c_uint(flags), _ptr(hdr), _ptr(coeff), _ptr(out))
    return (out[:, :256].reshape(n, 16, 16), out[:, 256:320].reshape(n, 8, 8), out[:, 320:].reshape(n, 8, 8))


def recon_picture(w_mbs, h_mbs, flags, pp, hdr, mv, coeff, refs=()):
    n = w_mbs * h_mbs
    W, H = w_mbs * 16, h_mbs * 16
    hdr = np.ascontiguousarray(hdr)
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    mv = np.ascontiguousarray(mv, dtype=np.int16) if mv is not None else None
    y = np.zeros((H, W), dtype=np.uint8)
    u = np.zeros((H // 2, W // 2), dtype=np.uint8)
    v = np.zeros((H // 2, W // 2), dtype=np.uint8)
    emap = np.zeros(n, dtype=np.uint8)
    arr = ctypes.c_void_p * abi.MAX_REFS
    ry, ru, rv = arr(), arr(), arr()
    keep = []
    for i, planes in enumerate(refs):
        planes = tuple(np.ascontiguousarray(x) for x in planes)
        keep.append(planes)
        ry[i], ru[i], rv[i] = (_ptr(x) for x in planes)
    exc = lib().h264emul_recon_picture(w_mbs, h_mbs, ctypes.c_uint(flags), ctypes.byref(pp), _ptr(hdr), _ptr(mv), _ptr(coeff),
                                       _ptr(y), _ptr(u), _ptr(v), ry, ru, rv, _ptr(emap))
    return {"y": y, "u": u, "v": v, "excursions": int(exc), "excursion_map": emap}


def digest(y, u, v):
    out = (ctypes.c_uint32 * 4)()
    y, u, v = (np.ascontiguousarray(x, dtype=np.uint8) for x in (y, u, v))
    lib().h264emul_digest(_ptr(y), _ptr(u), _ptr(v), ctypes.c_int(y.shape[1]), ctypes.c_int(y.shape[0]), out)
    return bytes(out)
