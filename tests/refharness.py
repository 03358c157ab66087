"""TEST INFRASTRUCTURE: runs the compiled reference decoder (oracle/_ref/h264ref_harness, built by
oracle/Makefile from the sources under /root/reference/h264 plus three shims) on a synthesized
stream and parses what it dumps.  Only tests/, smoke() and bench.py's reference legs use this."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "h264ref_harness")
HARNESS_CABAC = os.path.join(ROOT, "oracle", "_ref", "h264ref_cabac")   # the same, with the reference's own h264/cabac.cpp
ADAPTER = os.path.join(ROOT, "oracle", "_ref", "h264ref_adapter")      # reference host side + INTEGRATION.md's binding + libh264r.so

HDR_INTS = 25
MB_REC_INTS = 34 + 32 + 256 + 64 + 64 + 256


def have_harness():
    return os.path.exists(HARNESS)


def run_reference(stream, queue, level=2, keep=None):
    """Returns (pictures, stats).  pictures[i] = dict(y,u,v uint8 planes, raw int32 luma,
    mb = dict of per-MB arrays (level 2), ref_list = decode indices of list_Ref0)."""
    d = keep or tempfile.mkdtemp(prefix="h264ref_")
    sp, qp, dp = (os.path.join(d, n) for n in ("s.264", "q.bin", "dump.bin"))
    with open(sp, "wb") as f:
        f.write(stream)
    np.asarray(queue, dtype=np.int32).tofile(qp)
    r = subprocess.run([HARNESS, sp, qp, dp if level >= 0 else "-", str(max(level, 0))],
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference harness failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
    stats = json.loads(r.stderr.strip().splitlines()[-1])
    pics = parse_dump(dp) if level >= 0 else []
    if keep is None:
        for p in (sp, qp, dp):
            if os.path.exists(p):
                os.remove(p)
        os.rmdir(d)
    return pics, stats


def have_cabac_harness():
    return os.path.exists(HARNESS_CABAC)


def run_reference_cabac(stream, level=2, timeout=None):
    """The unmodified reference (bit reader, CABAC, macroblock layer, reconstruction) on a genuine CABAC stream.
    Returns (pictures, stats) like run_reference."""
    d = tempfile.mkdtemp(prefix="h264refc_")
    sp, dp = os.path.join(d, "s.264"), os.path.join(d, "dump.bin")
    try:
        with open(sp, "wb") as f:
            f.write(stream)
        r = subprocess.run([HARNESS_CABAC, sp, "-", dp if level >= 0 else "-", str(max(level, 0))],
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("reference (own CABAC) failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        stats = json.loads(r.stderr.strip().splitlines()[-1])
        return (parse_dump(dp) if level >= 0 else []), stats
    finally:
        for p in (sp, dp):
            if os.path.exists(p):
                os.remove(p)
        os.rmdir(d)


def have_adapter():
    return os.path.exists(ADAPTER)


def run_adapter(stream, queue, n_mbs=None, pack_only=False, transport=None):
    """Runs the reference-side binding (oracle/shims/adapter_main.cpp).  pack_only: returns the packed buffers per picture
    [(hdr (n,16) u8, mv (n,32) i16, coeff (n,384) i16)...] without touching the GPU -- with transport="buf" what
    include/h264r_writer.h wrote into a picture buffer: [dict(level_bytes, n_intra, hdr, blk_mask, units (n_units, 8|32) u8,
    mvs (n_mvs, 2) i16 in partition order)...]; else the adapter's verdict (dict)."""
    d = tempfile.mkdtemp(prefix="h264ada_")
    sp, qp, op = (os.path.join(d, n) for n in ("s.264", "q.bin", "pack.bin"))
    try:
        with open(sp, "wb") as f:
            f.write(stream)
        np.asarray(queue, dtype=np.int32).tofile(qp)
        r = subprocess.run([ADAPTER, sp, qp] + (["--pack-only", op] if pack_only else []) + (["--transport", transport] if transport else []),
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if r.returncode not in (0, 1):
            raise RuntimeError("adapter failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        verdict = json.loads(r.stdout.strip().splitlines()[-1])
        verdict["returncode"] = r.returncode
        if not pack_only:
            return verdict
        raw = np.fromfile(op, dtype=np.uint8)
        if transport == "buf":
            out, off = [], 0
            for _ in range(verdict["pictures"]):
                fmt, n_units, n_mvs, n_intra = (int(x) for x in raw[off:off + 16].view(np.int32)); off += 16
                usz = 32 if fmt == 2 else 8
                rec = {"level_bytes": fmt, "n_intra": n_intra}
                rec["hdr"] = raw[off:off + n_mbs * 16].reshape(n_mbs, 16); off += n_mbs * 16
                rec["blk_mask"] = raw[off:off + n_mbs * 4].view(np.uint32).copy(); off += n_mbs * 4
                rec["units"] = raw[off:off + n_units * usz].reshape(n_units, usz); off += n_units * usz
                rec["mvs"] = raw[off:off + n_mvs * 4].view(np.int16).reshape(n_mvs, 2)[::-1].copy(); off += n_mvs * 4
                out.append(rec)
            assert off == raw.size
            return out
        per = n_mbs * (16 + 64 + 768)
        assert raw.size == per * verdict["pictures"]
        out = []
        for i in range(verdict["pictures"]):
            b = raw[i * per:(i + 1) * per]
            out.append((b[:n_mbs * 16].reshape(n_mbs, 16), b[n_mbs * 16:n_mbs * 80].view(np.int16).reshape(n_mbs, 32),
                        b[n_mbs * 80:].view(np.int16).reshape(n_mbs, 384)))
        return out
    finally:
        for p in (sp, qp, op):
            if os.path.exists(p):
                os.remove(p)
        os.rmdir(d)


def parse_dump(path):
    data = np.fromfile(path, dtype=np.uint8)
    off = 0
    pics = []
    while off < data.size:
        hdr = data[off:off + 4 * HDR_INTS].view(np.int32)
        off += 4 * HDR_INTS
        assert hdr[0] == 0x44363248, "bad dump magic"
        wm, hm, level = int(hdr[2]), int(hdr[3]), int(hdr[6])
        W, H = wm * 16, hm * 16
        p = {"index": int(hdr[1]), "w_mbs": wm, "h_mbs": hm, "nal_type": int(hdr[4]), "poc": int(hdr[5]),
             "dec_num": int(hdr[7]), "ref_list": [int(x) for x in hdr[9:9 + int(hdr[8])]]}
        p["y"] = data[off:off + W * H].reshape(H, W).copy(); off += W * H
        p["u"] = data[off:off + W * H // 4].reshape(H // 2, W // 2).copy(); off += W * H // 4
        p["v"] = data[off:off + W * H // 4].reshape(H // 2, W // 2).copy(); off += W * H // 4
        if level >= 1:
            p["raw"] = data[off:off + 4 * W * H].view(np.int32).reshape(H, W).copy(); off += 4 * W * H
        if level >= 2:
            n = wm * hm
            rec = data[off:off + 4 * n * MB_REC_INTS].view(np.int32).reshape(n, MB_REC_INTS); off += 4 * n * MB_REC_INTS
            p["mb"] = {
                "type": rec[:, 0].copy(), "premode": rec[:, 1].copy(), "qpy": rec[:, 2].copy(), "qpc": rec[:, 3].copy(),
                "cbp_luma": rec[:, 4].copy(), "cbp_chroma": rec[:, 5].copy(), "chroma_mode": rec[:, 6].copy(),
                "t8x8": rec[:, 7].copy(), "num_part": rec[:, 8].copy(), "skip": rec[:, 9].copy(),
                "i4_modes": rec[:, 10:26].copy(), "sub_type": rec[:, 26:30].copy(), "ref_idx": rec[:, 30:34].copy(),
                "mv": rec[:, 34:66].reshape(n, 4, 4, 2).copy(),
                "res_y": rec[:, 66:322].reshape(n, 16, 16).copy(),
                "res_u": rec[:, 322:386].reshape(n, 8, 8).copy(),
                "res_v": rec[:, 386:450].reshape(n, 8, 8).copy(),
                "pred_y": rec[:, 450:706].reshape(n, 16, 16).copy(),
            }
        pics.append(p)
    return pics
