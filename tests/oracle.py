"""TEST INFRASTRUCTURE: ctypes binding of the CPU restatement (oracle/_ref/libh264restate.so,
source oracle/restate/h264_restate.c).  The product never imports this."""
import ctypes
import os

import numpy as np

from videodecoder_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBPATH = os.path.join(ROOT, "oracle", "_ref", "libh264restate.so")

_u8p = ctypes.POINTER(ctypes.c_uint8)


class OPic(ctypes.Structure):
    _fields_ = [
        ("w_mbs", ctypes.c_int32), ("h_mbs", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("y", ctypes.c_void_p), ("u", ctypes.c_void_p), ("v", ctypes.c_void_p),
        ("raw_y", ctypes.c_void_p),
        ("hdr", ctypes.c_void_p), ("mv", ctypes.c_void_p), ("coeff", ctypes.c_void_p),
        ("pp", ctypes.POINTER(abi.PicParams)),
        ("ref_y", ctypes.c_void_p * abi.MAX_REFS), ("ref_u", ctypes.c_void_p * abi.MAX_REFS),
        ("ref_v", ctypes.c_void_p * abi.MAX_REFS),
        ("ref_id", ctypes.c_int32 * abi.MAX_REFS),
        ("excursions", ctypes.c_uint64),
        ("excursion_map", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            from videodecoder_b200 import build
            build.build_oracle()
        _lib = ctypes.CDLL(LIBPATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def residual(w_mbs, h_mbs, flags, hdr, coeff):
    """residual::Decode for every MB: (res_y (n,16,16), res_u (n,8,8), res_v (n,8,8)) int32."""
    n = w_mbs * h_mbs
    hdr = np.ascontiguousarray(hdr)
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    pp = abi.default_pic_params()
    pic = OPic(w_mbs=w_mbs, h_mbs=h_mbs, flags=flags, hdr=_ptr(hdr), coeff=_ptr(coeff), pp=ctypes.pointer(pp))
    ry = np.zeros((n, 16, 16), dtype=np.int32)
    ru = np.zeros((n, 8, 8), dtype=np.int32)
    rv = np.zeros((n, 8, 8), dtype=np.int32)
    f = lib().h264o_residual_mb
    for a in range(n):
        f(ctypes.byref(pic), ctypes.c_int(a), _ptr(ry[a]), _ptr(ru[a]), _ptr(rv[a]))
    return ry, ru, rv


def recon_picture(w_mbs, h_mbs, flags, pp, hdr, mv, coeff, refs=(), ref_ids=None, want_raw=False):
    """Reconstructs one picture.  refs: list of (y,u,v) planes for ref_list0 entries.
    Returns dict(y,u,v, excursions, excursion_map[, raw])."""
    n = w_mbs * h_mbs
    W, H = w_mbs * 16, h_mbs * 16
    hdr = np.ascontiguousarray(hdr)
    coeff = np.ascontiguousarray(coeff, dtype=np.int16)
    mv = np.ascontiguousarray(mv, dtype=np.int16) if mv is not None else None
    y = np.zeros((H, W), dtype=np.uint8)
    u = np.zeros((H // 2, W // 2), dtype=np.uint8)
    v = np.zeros((H // 2, W // 2), dtype=np.uint8)
    raw = np.zeros((H, W), dtype=np.int32) if (flags & abi.REF_COMPAT) else None
    emap = np.zeros(n, dtype=np.uint8)
    pic = OPic(w_mbs=w_mbs, h_mbs=h_mbs, flags=flags, y=_ptr(y), u=_ptr(u), v=_ptr(v), raw_y=_ptr(raw),
               hdr=_ptr(hdr), mv=_ptr(mv), coeff=_ptr(coeff), pp=ctypes.pointer(pp), excursion_map=_ptr(emap))
    keep = []
    for i, (ry, ru, rv) in enumerate(refs):
        ry, ru, rv = (np.ascontiguousarray(x) for x in (ry, ru, rv))
        keep.append((ry, ru, rv))
        pic.ref_y[i], pic.ref_u[i], pic.ref_v[i] = _ptr(ry), _ptr(ru), _ptr(rv)
        pic.ref_id[i] = ref_ids[i] if ref_ids is not None else pp.ref_list0[i]
    lib().h264o_recon_picture(ctypes.byref(pic))
    out = {"y": y, "u": u, "v": v, "excursions": int(pic.excursions), "excursion_map": emap}
    if want_raw:
        out["raw"] = raw
    return out


def digest(y, u, v):
    out = (ctypes.c_uint8 * 16)()
    y, u, v = (np.ascontiguousarray(x, dtype=np.uint8) for x in (y, u, v))
    lib().h264o_digest(_ptr(y), _ptr(u), _ptr(v), ctypes.c_int(y.shape[1]), ctypes.c_int(y.shape[0]), out)
    return bytes(out)
