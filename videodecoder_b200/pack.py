"""Host-side packing of macroblock syntax into the flat buffers of include/h264r.h.

This is the Python twin of what the C++ host parser emits at the end of
Slice::PraseSliceDataer (reference h264/slice.cpp:278-389): coefficient levels leave scan order
(Inverse4x4Scan, reference h264/residual.cpp:136-168) and land block-major; motion vectors are
expanded to one per 4x4 block; prediction modes and QPs go into the 16-byte header records.
"""
import numpy as np

from . import abi
from . import synth

# scan index -> raster position inside a 4x4 block (reference h264/residual.cpp:136-141)
ZZ = np.array([0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15])
BLK_RASTER = synth.BLK_RASTER


def pack_coeff(mbs):
    """(n, 384) int16 from abstract syntax records (synth.MB_DTYPE)."""
    n = mbs.shape[0]
    out = np.zeros((n, 24, 16), dtype=np.int16)
    i16 = mbs["kind"] == synth.I16x16
    # 4x4 luma blocks: all 16 levels
    tmp = np.zeros((n, 16, 16), dtype=np.int16)
    tmp[:, :, ZZ] = mbs["luma"]
    out[~i16, :16] = tmp[~i16]
    # Intra16x16: 15 AC levels behind the DC slot, DC matrix element (r,c) -> block at raster (r,c)
    tmp = np.zeros((n, 16, 16), dtype=np.int16)
    tmp[:, :, ZZ[1:]] = mbs["luma"][:, :, :15]
    tmp[:, BLK_RASTER[ZZ], 0] = mbs["luma_dc"]
    out[i16, :16] = tmp[i16]
    ch = np.zeros((n, 2, 4, 16), dtype=np.int16)
    ch[:, :, :, ZZ[1:]] = mbs["chroma_ac"][:, :, :, :15]
    ch[:, :, :, 0] = mbs["chroma_dc"]
    out[:, 16:] = ch.reshape(n, 8, 16)
    return out.reshape(n, 384)


_CLASS = {synth.I4x4: abi.MB_I4x4, synth.I16x16: abi.MB_I16x16, synth.P16x16: abi.MB_P16x16,
          synth.P16x8: abi.MB_P16x8, synth.P8x16: abi.MB_P8x16, synth.P8x8: abi.MB_P8x8,
          synth.PSKIP: abi.MB_P16x16}


def expand_mv(kind, sub_type, part_mv):
    """part_mv[part][sub] = (x, y) -> 16 raster 4x4-block vectors (16, 2)."""
    out = np.zeros((16, 2), dtype=np.int16)
    for ras in range(16):
        bx, by = ras % 4, ras // 4
        q = (by >> 1) * 2 + (bx >> 1)
        if kind in (synth.P16x16, synth.PSKIP):
            part, sub = 0, 0
        elif kind == synth.P16x8:
            part, sub = by >> 1, 0
        elif kind == synth.P8x16:
            part, sub = bx >> 1, 0
        else:
            part = q
            st = sub_type[q]
            sub = {abi.SUB_8x8: 0, abi.SUB_8x4: by & 1, abi.SUB_4x8: bx & 1, abi.SUB_4x4: (by & 1) * 2 + (bx & 1)}[int(st)]
        out[ras] = part_mv[part][sub]
    return out


def expand_ref(kind, part_ref):
    """ref_idx per 8x8 quadrant."""
    if kind in (synth.P16x16, synth.PSKIP):
        return [part_ref[0]] * 4
    if kind == synth.P16x8:
        return [part_ref[0], part_ref[0], part_ref[1], part_ref[1]]
    if kind == synth.P8x16:
        return [part_ref[0], part_ref[1], part_ref[0], part_ref[1]]
    return list(part_ref[:4])


def pack_headers(w_mbs, h_mbs, mbs, qp_y, qp_c, i4_modes=None, part_ref=None):
    """Header records.  qp_y / qp_c: per-MB QP'Y and QP'C (both chroma planes use qp_c);
    i4_modes: (n,16) final Intra4x4PredMode per luma4x4BlkIdx; part_ref: (n,4) ref_idx_l0 per partition."""
    n = mbs.shape[0]
    hdr = np.zeros(n, dtype=abi.MB_HDR_DTYPE)
    hdr["flags"] = abi.avail_flags(w_mbs, h_mbs)
    hdr["qp_y"] = qp_y
    hdr["qp_cb"] = qp_c
    hdr["qp_cr"] = qp_c
    for a in range(n):
        k = int(mbs["kind"][a])
        hdr["mb_class"][a] = _CLASS[k]
        hdr["cbp"][a] = int(mbs["cbp_luma"][a]) | (int(mbs["cbp_chroma"][a]) << 4)
        if k == synth.PSKIP:
            hdr["flags"][a] |= abi.SKIP
        if k in (synth.I4x4, synth.I16x16):
            hdr["intra_modes"][a] = (int(mbs["i16_mode"][a]) & 3) | ((int(mbs["chroma_mode"][a]) & 3) << 2)
            if k == synth.I4x4:
                m = i4_modes[a]
                for j in range(8):
                    hdr["u"][a][j] = (int(m[2 * j]) & 15) | ((int(m[2 * j + 1]) & 15) << 4)
        else:
            st = mbs["sub_type"][a]
            if k == synth.P8x8:
                hdr["sub_mb_type"][a] = int(st[0]) | (int(st[1]) << 2) | (int(st[2]) << 4) | (int(st[3]) << 6)
            hdr["u"][a][:4] = expand_ref(k, [int(x) for x in part_ref[a]])
    return hdr


def pack_mvs(mbs, part_mv):
    """(n, 32) int16; part_mv: (n, 4, 4, 2) final vectors per partition / sub-partition."""
    n = mbs.shape[0]
    mv = np.zeros((n, 16, 2), dtype=np.int16)
    for a in range(n):
        k = int(mbs["kind"][a])
        if k >= synth.P16x16:
            mv[a] = expand_mv(k, mbs["sub_type"][a], part_mv[a])
    return mv.reshape(n, 32)


def pack_blocks(coeff):
    """Dense coefficient buffer (n, 384) int16 -> what h264r_submit_mbs_packed takes:
    (blk_mask uint32 (n,), blocks int16 (n_blocks, 16)); a block is kept iff it holds a non-zero level."""
    c = np.ascontiguousarray(coeff, dtype=np.int16).reshape(-1, 24, 16)
    present = (c != 0).any(axis=2)
    mask = (present.astype(np.uint32) << np.arange(24, dtype=np.uint32)).sum(axis=1, dtype=np.uint32)
    blocks = np.ascontiguousarray(c[present])
    return mask, blocks


def unpack_blocks(mask, blocks):
    """inverse of pack_blocks (tests)"""
    n = mask.shape[0]
    present = ((mask[:, None] >> np.arange(24, dtype=np.uint32)) & 1).astype(bool)
    c = np.zeros((n, 24, 16), dtype=np.int16)
    c[present] = blocks.reshape(-1, 16)
    return c.reshape(n, 384)


_SUB_COUNT = np.array([1, 2, 2, 4])
# first 4x4 block (raster, relative to the quadrant's first block) of each sub-macroblock partition
_SUB_FIRST = {0: [0], 1: [0, 4], 2: [0, 1], 3: [0, 1, 4, 5]}


def pack_mvs_partition(hdr, mv):
    """Per-4x4 vectors (n, 32) int16 -> partition order, as h264r_submit_picture(..., n_mvs) takes them:
    one vector per partition / sub-macroblock partition in bitstream order (mvd loops of macroblock::Parse,
    h264/macroblock.cpp:360-398).  Returns (n_mvs, 2) int16."""
    hdr = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
    mv = np.ascontiguousarray(mv, dtype=np.int16).reshape(-1, 16, 2)
    cls = hdr["mb_class"]
    out = []
    # vectorised per class; P_8x8 macroblocks individually ordered afterwards
    order = np.zeros(hdr.shape[0], dtype=np.int64)
    counts = np.zeros(hdr.shape[0], dtype=np.int64)
    counts[cls == abi.MB_P16x16] = 1
    counts[(cls == abi.MB_P16x8) | (cls == abi.MB_P8x16)] = 2
    p8 = np.flatnonzero(cls == abi.MB_P8x8)
    st = hdr["sub_mb_type"]
    for a in p8:
        counts[a] = sum(_SUB_COUNT[(int(st[a]) >> (2 * q)) & 3] for q in range(4))
    off = np.concatenate(([0], np.cumsum(counts)))
    res = np.zeros((int(off[-1]), 2), dtype=np.int16)
    m = cls == abi.MB_P16x16
    res[off[:-1][m]] = mv[m, 0]
    m = cls == abi.MB_P16x8
    res[off[:-1][m]] = mv[m, 0]
    res[off[:-1][m] + 1] = mv[m, 8]
    m = cls == abi.MB_P8x16
    res[off[:-1][m]] = mv[m, 0]
    res[off[:-1][m] + 1] = mv[m, 2]
    for a in p8:
        k = int(off[a])
        for q in range(4):
            b0 = (q >> 1) * 8 + (q & 1) * 2
            for rel in _SUB_FIRST[(int(st[a]) >> (2 * q)) & 3]:
                res[k] = mv[a, b0 + rel]
                k += 1
    return res


def pack_compact(blk_mask, blocks):
    """H264R_LEVELS_COMPACT storage of include/h264r.h from the packed form (blk_mask [n_mbs] uint32, blocks
    [n_blocks, 16] int16 with every level in [-128, 127]): 8-byte units.  A macroblock whose blocks all have at most six
    levels gets H264R_BLK_COMPACT in its mask word and one unit per block (significance map, then the levels in
    position order); any other macroblock two units per block (int8 raster).  Returns (mask, units [n_units, 8] uint8)."""
    blk_mask = np.asarray(blk_mask, dtype=np.uint32)
    blocks = np.asarray(blocks, dtype=np.int16).reshape(-1, 16)
    n_blocks = blocks.shape[0]
    per_mb = np.array([bin(int(m) & 0xFFFFFF).count("1") for m in blk_mask], dtype=np.int64) if blk_mask.size < 64 else \
        np.unpackbits((blk_mask & 0xFFFFFF).view(np.uint8).reshape(-1, 4), axis=1).sum(axis=1).astype(np.int64)
    assert int(per_mb.sum()) == n_blocks
    nz = blocks != 0
    nnz = nz.sum(axis=1)
    mb_of_block = np.repeat(np.arange(blk_mask.size), per_mb)
    too_many = np.zeros(blk_mask.size, dtype=np.int64)
    np.add.at(too_many, mb_of_block, (nnz > 6).astype(np.int64))
    mb_compact = (too_many == 0) & (per_mb > 0)
    mask = blk_mask.copy()
    mask[mb_compact] |= np.uint32(1 << 24)
    blk_compact = mb_compact[mb_of_block]
    units_per_block = np.where(blk_compact, 1, 2)
    first_unit = np.concatenate(([0], np.cumsum(units_per_block)))[:-1]
    n_units = int(units_per_block.sum())
    units = np.zeros((n_units, 8), dtype=np.uint8)
    # compact blocks: significance map + levels in position order
    cb = np.nonzero(blk_compact)[0]
    if cb.size:
        nzc = nz[cb]
        sig = (nzc.astype(np.uint32) << np.arange(16, dtype=np.uint32)).sum(axis=1).astype(np.uint16)
        order = np.argsort(~nzc, axis=1, kind="stable")[:, :6]             # positions of the levels first, ascending
        lv = np.take_along_axis(blocks[cb], order, axis=1).astype(np.int8)
        lv[np.arange(6)[None, :] >= nnz[cb][:, None]] = 0
        u = units[first_unit[cb]]
        u[:, 0] = sig & 0xFF
        u[:, 1] = sig >> 8
        u[:, 2:8] = lv.view(np.uint8)
        units[first_unit[cb]] = u
    pb = np.nonzero(~blk_compact)[0]
    if pb.size:
        raw = blocks[pb].astype(np.int8).view(np.uint8)
        units[first_unit[pb]] = raw[:, :8]
        units[first_unit[pb] + 1] = raw[:, 8:]
    return mask, units


def pack_hdr8(hdr):
    """8-byte headers (h264r_mb_hdr8 of include/h264r.h) from 16-byte ones: [n, 2] uint32."""
    h = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
    u32 = lambda a: a.astype(np.uint32)
    w0 = (u32(h["mb_class"]) & 7) | ((u32(h["flags"]) & 63) << 3) | ((u32(h["intra_modes"]) & 15) << 9) | \
         ((u32(h["qp_y"]) & 63) << 13) | ((u32(h["qp_cb"]) & 63) << 19) | ((u32(h["qp_cr"]) & 63) << 25)
    w1 = (u32(h["cbp"]) & 63) | (u32(h["sub_mb_type"]) << 6)
    u = h["u"].reshape(-1, 8)
    refs = ((u32(u[:, 0]) & 15) << 14) | ((u32(u[:, 1]) & 15) << 18) | ((u32(u[:, 2]) & 15) << 22) | ((u32(u[:, 3]) & 15) << 26)
    w1 = np.where(h["mb_class"] > abi.MB_I16x16, w1 | refs, w1)
    return np.stack([w0, w1], axis=1).astype(np.uint32)


def insert_intra_units(hdr, mask, units):
    """8-byte headers: every intra macroblock's u.pred4x4 bytes become the first unit of its blocks.  mask / units as
    returned by pack_compact; returns the unit stream with the extra units inserted."""
    h = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
    mask = np.asarray(mask, dtype=np.uint32)
    per_mb = np.unpackbits((mask & 0xFFFFFF).view(np.uint8).reshape(-1, 4), axis=1).sum(axis=1).astype(np.int64)
    per_mb = per_mb * np.where((mask >> 24) & 1, 1, 2)
    first = np.concatenate(([0], np.cumsum(per_mb)))[:-1]
    intra = np.flatnonzero(h["mb_class"] <= abi.MB_I16x16)
    if intra.size == 0:
        return units
    return np.insert(units, first[intra], h["u"].reshape(-1, 8)[intra].view(np.uint8), axis=0)


_ZZ8 = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


def syntax_from_buffers(hdr, mv, coeff, slice_qp):
    """Inverse of the host derivation (csrc/host/h264_derive.h): h264r buffers of one picture -> abstract syntax records
    (synth.MB_DTYPE) in ABSOLUTE form -- mvd[] holds the final vectors, rem_mode[] the final intra modes -- for
    hoststream.encode_stream(..., flags=E_ABSOLUTE), which turns them into a genuine CABAC bitstream.  Lets content that was
    generated directly in buffer form (workload.py: the benchmark's GOPs) be fed to the reference decoder and to libavcodec.

    What a bitstream cannot carry is normalised the way the syntax demands: QP changes only at macroblocks with residual syntax
    (the others inherit the running QP, which nothing reads in REF_COMPAT); a P_Skip flag is dropped (the macroblock is written as
    P_L0_16x16 without residual, so that it keeps its vector instead of the inferred one)."""
    hdr = np.ascontiguousarray(hdr).view(abi.MB_HDR_DTYPE).reshape(-1)
    n = hdr.shape[0]
    c = np.ascontiguousarray(coeff, dtype=np.int16).reshape(n, 24, 16)
    m = np.zeros(n, dtype=synth.MB_DTYPE)
    cls = hdr["mb_class"]
    kind = np.select([cls == abi.MB_I4x4, cls == abi.MB_I8x8, cls == abi.MB_I16x16, cls == abi.MB_P16x16, cls == abi.MB_P16x8,
                      cls == abi.MB_P8x16, cls == abi.MB_P8x8],
                     [synth.I4x4, synth.I8x8, synth.I16x16, synth.P16x16, synth.P16x8, synth.P8x16, synth.P8x8])
    m["kind"] = kind
    i16 = kind == synth.I16x16
    intra = cls <= abi.MB_I16x16
    m["cbp_luma"] = hdr["cbp"] & 15
    m["cbp_chroma"] = hdr["cbp"] >> 4
    m["i16_mode"] = np.where(i16, hdr["intra_modes"] & 3, 0)
    m["chroma_mode"] = np.where(intra, (hdr["intra_modes"] >> 2) & 3, 0)
    t8 = (hdr["flags"] & abi.T8x8) != 0
    m["t8x8"] = t8
    # ---- levels: raster inside the block -> scan order
    luma = c[:, :16]
    scan4 = luma[:, :, ZZ]                                   # [n, blk, k] = level at scan position k
    m["luma"] = np.where(i16[:, None, None], np.concatenate([luma[:, :, ZZ[1:]], np.zeros((n, 16, 1), np.int16)], axis=2), scan4)
    if t8.any():
        l8 = luma.reshape(n, 4, 64)[:, :, _ZZ8].reshape(n, 16, 16)
        m["luma"][t8] = l8[t8]
    m["luma_dc"] = np.where(i16[:, None], luma[:, BLK_RASTER[ZZ], 0], 0)
    ch = c[:, 16:].reshape(n, 2, 4, 16)
    m["chroma_dc"] = ch[:, :, :, 0]
    m["chroma_ac"][:, :, :, :15] = ch[:, :, :, ZZ[1:]]
    # ---- intra modes (final)
    u = hdr["u"]
    nib = np.stack([u & 15, u >> 4], axis=2).reshape(n, 16)    # mode of block 2j in the low nibble of byte j
    is4, is8 = kind == synth.I4x4, kind == synth.I8x8
    m["rem_mode"][is4] = nib[is4]
    m["rem_mode"][is8, :4] = nib[is8, :4]
    # ---- inter: sub types, reference indices per partition, final vector per partition
    st = hdr["sub_mb_type"]
    p88 = kind == synth.P8x8
    for q in range(4):
        m["sub_type"][:, q] = np.where(p88, (st >> (2 * q)) & 3, 0)
    ref = u[:, :4]
    m["ref_idx"][:, 0] = np.where(~intra, ref[:, 0], 0)
    m["ref_idx"][:, 1] = np.select([kind == synth.P16x8, kind == synth.P8x16, p88], [ref[:, 2], ref[:, 1], ref[:, 1]], 0)
    m["ref_idx"][:, 2] = np.where(p88, ref[:, 2], 0)
    m["ref_idx"][:, 3] = np.where(p88, ref[:, 3], 0)
    if mv is not None:
        v = np.ascontiguousarray(mv, dtype=np.int16).reshape(n, 16, 2)
        inter = ~intra
        m["mvd"][inter, 0] = v[inter, 0]
        k = kind == synth.P16x8
        m["mvd"][k, 4] = v[k, 8]
        k = kind == synth.P8x16
        m["mvd"][k, 4] = v[k, 2]
        for a in np.flatnonzero(p88):
            for q in range(4):
                b0 = (q >> 1) * 8 + (q & 1) * 2
                for j, rel in enumerate(_SUB_FIRST[int(m["sub_type"][a, q])]):
                    m["mvd"][a, q * 4 + j] = v[a, b0 + rel]
    # ---- QP chain: mb_qp_delta exists only with residual syntax
    coded = (m["cbp_luma"] != 0) | (m["cbp_chroma"] != 0) | i16
    qp = hdr["qp_y"].astype(np.int64)
    prev = slice_qp
    dq = np.zeros(n, dtype=np.int64)
    for a in np.flatnonzero(coded):
        d = int(qp[a]) - prev
        assert -26 <= d <= 25, "QP step outside mb_qp_delta"
        dq[a] = d
        prev = int(qp[a])
    m["qp_delta"] = dq
    return m
