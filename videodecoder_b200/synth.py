"""Synthetic H.264 content: abstract macroblock syntax -> Annex-B stream + syntax-element queue.

There is no H.264 encoder (and no sample stream) in this image, so test and benchmark inputs are
synthesized.  A picture is an array of ``MB_DTYPE`` records (mirror of ``h264s_mb`` in
include/h264_syntax.h) holding the syntax-element values macroblock::Parse would read
(reference: h264/macroblock.cpp:143-475); libh264host.so serialises them.

The random content model honours the constraints under which the reference decoder is usable
(SURVEY.md 8(c)): CABAC/Main syntax, one slice per picture, slice_type 5/7, 4x4 transform only,
no I_PCM, no 8x4 sub-partitions, only intra modes whose neighbours exist.
"""
import ctypes
import os

import numpy as np

from . import build

I4x4, I16x16, P16x16, P16x8, P8x16, P8x8, PSKIP, I8x8 = range(8)
SLICE_P, SLICE_I = 0, 2

MB_DTYPE = np.dtype([
    ("kind", "u1"), ("i16_mode", "u1"), ("cbp_luma", "u1"), ("cbp_chroma", "u1"),
    ("chroma_mode", "u1"), ("qp_delta", "i1"), ("t8x8", "u1"), ("pad_", "u1"),
    ("prev_flag", "u1", (16,)), ("rem_mode", "u1", (16,)),
    ("sub_type", "u1", (4,)), ("ref_idx", "u1", (4,)),
    ("mvd", "<i2", (16, 2)),
    ("luma_dc", "<i2", (16,)), ("luma", "<i2", (16, 16)),
    ("chroma_dc", "<i2", (2, 4)), ("chroma_ac", "<i2", (2, 4, 16)),
])
assert MB_DTYPE.itemsize == 928


class Seq(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "w_mbs", "h_mbs", "max_num_ref_frames", "pic_init_qp", "chroma_qp_index_offset",
        "num_ref_idx_l0_default_active_minus1", "weighted_pred_flag",
        "deblocking_filter_control_present_flag", "log2_max_frame_num_minus4",
        "log2_max_poc_lsb_minus4", "profile_idc",
        "transform_8x8_mode_flag", "constrained_intra_pred_flag", "second_chroma_qp_index_offset", "pic_order_cnt_type")]


class Pic(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "is_idr", "slice_type", "nal_ref_idc", "frame_num", "idr_pic_id", "poc_lsb", "slice_qp_delta",
        "num_ref_idx_l0_active_minus1", "luma_log2_weight_denom", "chroma_log2_weight_denom")] + [
        ("luma_weight_flag", ctypes.c_int32 * 16), ("luma_weight", ctypes.c_int32 * 16),
        ("luma_offset", ctypes.c_int32 * 16)] + [(n, ctypes.c_int32) for n in (
            "disable_deblocking_filter_idc", "slice_alpha_c0_offset_div2", "slice_beta_offset_div2",
            "cabac_init_idc", "long_term_reference_flag")] + [
        ("chroma_weight_flag", ctypes.c_int32 * 16), ("chroma_weight", (ctypes.c_int32 * 2) * 16),
        ("chroma_offset", (ctypes.c_int32 * 2) * 16),
        ("n_mods", ctypes.c_int32), ("mods", (ctypes.c_int32 * 2) * 16),
        ("n_mmco", ctypes.c_int32), ("mmco", (ctypes.c_int32 * 3) * 16)]


_lib = None


def host_lib():
    global _lib
    if _lib is None:
        path = build.build_host()
        _lib = ctypes.CDLL(path)
        _lib.h264s_write_sps_pps.restype = ctypes.c_long
        _lib.h264s_write_slice_nal.restype = ctypes.c_long
        _lib.h264s_serialize_mbs.restype = ctypes.c_long
    return _lib


def make_seq(w_mbs, h_mbs, max_num_ref_frames=1, pic_init_qp=28, num_ref_idx_default=0,
             weighted_pred=0, deblocking_control=0, chroma_qp_index_offset=0, profile_idc=77, transform_8x8_mode=0,
             constrained_intra_pred=0, second_chroma_qp_index_offset=None, pic_order_cnt_type=0):
    """profile_idc 77 (Main) is what the reference can read; 100 (High) adds the 8x8 transform and the second
    chroma QP offset and is written by the h264e_* writer only."""
    if second_chroma_qp_index_offset is None:
        second_chroma_qp_index_offset = chroma_qp_index_offset
    return Seq(w_mbs, h_mbs, max_num_ref_frames, pic_init_qp, chroma_qp_index_offset, num_ref_idx_default,
               weighted_pred, deblocking_control, 4, 4, profile_idc, transform_8x8_mode, constrained_intra_pred,
               second_chroma_qp_index_offset, pic_order_cnt_type)


def make_pic(seq, index_in_gop, slice_type, slice_qp_delta=0, num_ref_active=None, weights=None,
             idr_pic_id=0, deblock=(1, 0, 0)):
    """Picture parameters for picture `index_in_gop` of an IDR-led GOP in which every picture is a
    reference (frame_num = index, POC = 2*index: the reference's output logic expects POC steps of
    2, h264/Decoder.cpp:474-497)."""
    p = Pic()
    p.is_idr = 1 if index_in_gop == 0 else 0
    p.slice_type = slice_type
    p.nal_ref_idc = 3 if p.is_idr else 2
    p.frame_num = index_in_gop
    p.idr_pic_id = idr_pic_id
    p.poc_lsb = 2 * index_in_gop
    p.slice_qp_delta = slice_qp_delta
    if num_ref_active is None:
        num_ref_active = seq.num_ref_idx_l0_default_active_minus1 + 1
    p.num_ref_idx_l0_active_minus1 = num_ref_active - 1
    if weights is not None:
        p.luma_log2_weight_denom = weights["log2_denom"]
        for i, (w, o) in enumerate(weights["wo"]):
            p.luma_weight_flag[i] = 1
            p.luma_weight[i] = w
            p.luma_offset[i] = o
    p.disable_deblocking_filter_idc, p.slice_alpha_c0_offset_div2, p.slice_beta_offset_div2 = deblock
    return p


def serialize_stream(seq, pictures):
    """pictures: list of (Pic, mbs ndarray[MB_DTYPE]).  Returns (annexb bytes, queue int32 ndarray)."""
    lib = host_lib()
    buf = (ctypes.c_uint8 * 4096)()
    n = lib.h264s_write_sps_pps(ctypes.byref(seq), buf, 4096)
    assert n > 0
    stream = [bytes(buf[:n])]
    queues = []
    for pic, mbs in pictures:
        n = lib.h264s_write_slice_nal(ctypes.byref(seq), ctypes.byref(pic), buf, 4096)
        assert n > 0
        stream.append(bytes(buf[:n]))
        mbs = np.ascontiguousarray(mbs)
        assert mbs.dtype == MB_DTYPE and mbs.size == seq.w_mbs * seq.h_mbs
        cap = int(mbs.size) * 2 * 1800 + 64
        q = np.empty(cap, dtype=np.int32)
        m = lib.h264s_serialize_mbs(ctypes.byref(seq), ctypes.byref(pic), mbs.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_int(mbs.size), q.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(cap))
        assert m > 0, "queue overflow"
        queues.append(q[:m].copy())
    # trailing start code so that the last NAL is terminated the same way as the others
    stream.append(b"\x00\x00\x00\x01\x0c\xff\x80")
    return b"".join(stream), np.concatenate(queues)


# ------------------------------------------------------------------------------------------------
# random content model
# ------------------------------------------------------------------------------------------------
BLK_RASTER = np.array([0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15])  # luma4x4BlkIdx -> raster (involution)
_V0 = np.array([10, 11, 13, 14, 16, 18])


def _amp_per_level(qp):
    return _V0[qp % 6] * (2.0 ** (qp // 6)) / 64.0


class ContentConfig:
    def __init__(self, **kw):
        self.p_i16 = 0.35            # I pictures: share of Intra16x16 MBs
        self.p_intra_in_p = 0.05     # P pictures: share of intra MBs
        self.p_skip = 0.35
        self.p_part = (0.35, 0.15, 0.15, 0.35)   # 16x16, 16x8, 8x16, 8x8 among coded inter MBs
        self.sub_types = (0, 2, 3)   # 8x8, 4x8, 4x4 (8x4 crashes the reference: SURVEY D9)
        self.mvd_range = 6           # |mvd| in quarter samples
        self.frac_ok = None          # None: any; else set of allowed (xFrac, yFrac) for the FINAL mv: enforced only by direct generators
        self.p_block = 0.45          # share of coded 4x4 blocks inside a coded 8x8
        self.p_cbp8 = 0.5            # share of coded luma 8x8 quadrants
        self.p_chroma = (0.4, 0.25, 0.35)        # cbp chroma 0 / 1 / 2
        self.max_amp = 14.0          # target residual amplitude (samples) of one coefficient
        self.nnz_max = 4
        self.qp_delta_range = 2
        self.p_qp_delta = 0.2
        self.num_ref = 1
        # the three below need absolute=True in random_picture (streams written with hoststream.E_ABSOLUTE) and, for the 8x8
        # tools, a High-profile sequence (make_seq(profile_idc=100, transform_8x8_mode=1)); the reference reads neither
        self.p_i8x8 = 0.0            # share of I_NxN macroblocks that are Intra8x8
        self.p_t8x8 = 0.0            # share of eligible inter macroblocks with transform_size_8x8_flag
        self.mv_range = 64           # absolute mode: final vectors uniform in +-mv_range quarter samples
        self.constrained_intra = False   # choose intra modes as constrained_intra_pred_flag = 1 allows (inter neighbours do not exist)
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)


def _fill_block(rng, levels, n, qp, cfg, first=0):
    """a few low-frequency non-zero levels in scan positions [first, n)."""
    k = int(rng.integers(1, cfg.nnz_max + 1))
    pos = first + np.minimum(rng.geometric(0.35, size=k) - 1, n - first - 1)
    amp = rng.uniform(1.0, cfg.max_amp, size=k) / _amp_per_level(qp)
    mag = np.maximum(1, np.rint(amp)).astype(np.int16)
    sign = rng.integers(0, 2, size=k) * 2 - 1
    levels[pos] = mag * sign


def random_picture(rng, seq, slice_type, cfg, slice_qp, absolute=False):
    """Random macroblock syntax for one picture.  Returns ndarray[MB_DTYPE] (raster order).

    absolute=True: rem_mode[] holds the FINAL intra prediction modes and mvd[] the FINAL vectors, for
    hoststream.encode_stream(..., flags=E_ABSOLUTE), which derives prev / rem / mvd against the standard's predictors.

    Intra 4x4 modes are chosen among those whose neighbouring samples exist and are then coded
    against the predicted mode exactly as the reference derives it (ParseIntra4x4PredMode,
    h264/macroblock.cpp:992-1035, including its treatment of inter neighbours as "DC predicted").
    """
    W, H = seq.w_mbs, seq.h_mbs
    n = W * H
    mbs = np.zeros(n, dtype=MB_DTYPE)
    # final Intra4x4PredMode per 4x4 block of the picture; -1 = MB is not Intra4x4 (-2: inter)
    modes = np.full((H * 4, W * 4), -1, dtype=np.int8)
    qp = slice_qp
    def _ok(x, y):       # macroblock (x, y) may be used for intra prediction
        if x < 0 or y < 0:
            return False
        return not (cfg.constrained_intra and mbs[y * W + x]["kind"] in (P16x16, P16x8, P8x16, P8x8, PSKIP))

    for a in range(n):
        mby, mbx = divmod(a, W)
        m = mbs[a]
        if slice_type == SLICE_I:
            kind = I16x16 if rng.random() < cfg.p_i16 else I4x4
        else:
            r = rng.random()
            if r < cfg.p_intra_in_p:
                kind = I16x16 if rng.random() < cfg.p_i16 else I4x4
            elif r < cfg.p_intra_in_p + cfg.p_skip:
                kind = PSKIP
            else:
                kind = (P16x16, P16x8, P8x16, P8x8)[rng.choice(4, p=cfg.p_part)]
        if kind == I4x4 and cfg.p_i8x8 > 0 and rng.random() < cfg.p_i8x8:
            kind = I8x8
        m["kind"] = kind
        if kind == PSKIP:
            modes[mby * 4:mby * 4 + 4, mbx * 4:mbx * 4 + 4] = -2
            continue
        # ---- residual shape ----
        if kind == I16x16:
            m["cbp_luma"] = 15 if rng.random() < 0.5 else 0
        else:
            m["cbp_luma"] = int(np.packbits((rng.random(4) < cfg.p_cbp8)[::-1], bitorder="little")[0]) & 15
        m["cbp_chroma"] = int(rng.choice(3, p=cfg.p_chroma))
        has_res = kind == I16x16 or m["cbp_luma"] or m["cbp_chroma"]
        if has_res and rng.random() < cfg.p_qp_delta:
            d = int(rng.integers(-cfg.qp_delta_range, cfg.qp_delta_range + 1))
            if 12 <= qp + d <= 44:
                m["qp_delta"] = d
                qp += d
        qpc = max(qp - 2, 0)    # the reference's chroma QP (h264/macroblock.cpp:487-488)
        if kind == I16x16:
            if rng.random() < 0.8:
                _fill_block(rng, m["luma_dc"], 16, qp, cfg)
                m["luma_dc"] //= 2      # the DC path has twice the gain of a 4x4 DC
            if m["cbp_luma"]:
                for b in range(16):
                    if rng.random() < cfg.p_block:
                        _fill_block(rng, m["luma"][b], 15, qp, cfg)
        else:
            t8 = kind == I8x8 or (cfg.p_t8x8 > 0 and kind in (P16x16, P16x8, P8x16, P8x8) and rng.random() < cfg.p_t8x8)
            m["t8x8"] = 1 if t8 else 0
            if t8:
                for q in range(4):
                    if (m["cbp_luma"] >> q) & 1 and rng.random() < 0.8:
                        _fill_block(rng, m["luma"][4 * q:4 * q + 4].reshape(-1), 64, qp, cfg)
            else:
                for b in range(16):
                    if (m["cbp_luma"] >> (b >> 2)) & 1 and rng.random() < cfg.p_block:
                        _fill_block(rng, m["luma"][b], 16, qp, cfg)
        if m["cbp_chroma"]:
            for c in range(2):
                if rng.random() < 0.7:
                    _fill_block(rng, m["chroma_dc"][c], 4, qpc, cfg)
            if m["cbp_chroma"] == 2:
                for c in range(2):
                    for b in range(4):
                        if rng.random() < cfg.p_block:
                            _fill_block(rng, m["chroma_ac"][c][b], 15, qpc, cfg)
        # ---- prediction syntax ----
        okA, okB, okD = _ok(mbx - 1, mby), _ok(mbx, mby - 1), _ok(mbx - 1, mby - 1)
        if kind in (I4x4, I16x16, I8x8):
            m["chroma_mode"] = _legal_chroma_mode(rng, okA, okB, okD)
        if kind == I16x16:
            legal = [2]
            if okB:
                legal.append(0)
            if okA:
                legal.append(1)
            if okA and okB and okD:
                legal.append(3)
            # the reference's top-only DC sums the wrong line -- still deterministic, keep it
            m["i16_mode"] = int(rng.choice(legal))
        elif kind == I8x8:
            for q in range(4):
                left, top = okA or (q & 1), okB or (q >> 1)
                corner = (okD, okB, okA, True)[q]
                legal = [2]
                if top:
                    legal += [0, 3, 7]
                if left:
                    legal += [1, 8]
                if left and top and corner:
                    legal += [4, 5, 6]
                m["rem_mode"][q] = int(rng.choice(legal))
            modes[mby * 4:mby * 4 + 4, mbx * 4:mbx * 4 + 4] = -1
        elif kind == I4x4:
            for blk in range(16):
                ras = BLK_RASTER[blk]
                by, bx = mby * 4 + ras // 4, mbx * 4 + ras % 4
                left, top = okA or ras % 4 > 0, okB or ras // 4 > 0
                corner = okD if ras == 0 else (okB if ras < 4 else (okA if ras % 4 == 0 else True))
                legal = [2]
                if top:
                    legal += [0, 3, 7]
                if left:
                    legal += [1, 8]
                if left and top and corner:
                    legal += [4, 5, 6]
                want = int(rng.choice(legal))
                if absolute:
                    m["rem_mode"][blk] = want
                    continue
                ma = modes[by, bx - 1] if left else -2
                mb_ = modes[by - 1, bx] if top else -2
                if ma == -2 or mb_ == -2:
                    pred = 2
                else:
                    pred = min(2 if ma < 0 else ma, 2 if mb_ < 0 else mb_)
                if want == pred:
                    m["prev_flag"][blk] = 1
                else:
                    m["prev_flag"][blk] = 0
                    m["rem_mode"][blk] = want if want < pred else want - 1
                modes[by, bx] = want
        else:
            modes[mby * 4:mby * 4 + 4, mbx * 4:mbx * 4 + 4] = -2
            if kind == P8x8:
                m["sub_type"][:] = rng.choice(cfg.sub_types, size=4)
            m["ref_idx"][:] = rng.integers(0, cfg.num_ref, size=4)
            r = cfg.mv_range if absolute else cfg.mvd_range
            m["mvd"][:] = rng.integers(-r, r + 1, size=(16, 2))
    return mbs


def _legal_chroma_mode(rng, ok_a, ok_b, ok_d):
    legal = [0]
    if ok_a:
        legal.append(1)
    if ok_b:
        legal.append(2)
    if ok_a and ok_b and ok_d:
        legal.append(3)
    return int(rng.choice(legal))


def zero_residual(mbs, addrs):
    """Removes all residual syntax of the given macroblocks (keeps prediction syntax)."""
    for a in addrs:
        m = mbs[a]
        m["cbp_luma"] = 0
        m["cbp_chroma"] = 0
        m["luma_dc"][:] = 0
        m["luma"][:] = 0
        m["chroma_dc"][:] = 0
        m["chroma_ac"][:] = 0
        # qp_delta is only sent with residual syntax; an Intra16x16 MB always sends it, keep its value
        if m["kind"] != I16x16:
            m["qp_delta"] = 0


def random_gop(rng, seq, n_pictures, cfg, slice_qp_delta=0, intra_only=False, absolute=False):
    """[(Pic, mbs)] for one IDR-led GOP: I then P pictures, every picture a reference."""
    pics = []
    for i in range(n_pictures):
        st = SLICE_I if (i == 0 or intra_only) else SLICE_P
        pic = make_pic(seq, i, st, slice_qp_delta, num_ref_active=min(cfg.num_ref, max(i, 1)) if st == SLICE_P else None)
        if intra_only and i > 0:
            # all-intra sequences are written as consecutive IDR pictures
            pic.is_idr, pic.nal_ref_idc, pic.frame_num, pic.poc_lsb, pic.idr_pic_id = 1, 3, 0, 0, i
        if st == SLICE_P:
            cfg_i = ContentConfig(**{**cfg.__dict__, "num_ref": pic.num_ref_idx_l0_active_minus1 + 1})
        else:
            cfg_i = cfg
        mbs = random_picture(rng, seq, st, cfg_i, seq.pic_init_qp + slice_qp_delta, absolute=absolute)
        pics.append((pic, mbs))
    return pics
