"""Synthetic syntax dumps in the h264r buffer format, generated directly (no bitstream, no
reference decoder in the loop): the workloads of BASELINE.json configs 3-5 and of bench.py.

Everything is vectorised numpy so that 1080p GOPs are produced in seconds.  Content statistics
follow SURVEY.md 8(d): P pictures with skip / 16x16 / 16x8 / 8x16 / 8x8 (4x8 and 4x4
sub-partitions) macroblocks, quarter-sample vectors within +-16 samples, a few intra macroblocks,
sparse low-frequency coefficient levels; optionally transform_size_8x8_flag / Intra8x8 (spec mode
only -- the reference has neither).

`sanitize()` uses the engine's per-macroblock excursion map to zero the residual of macroblocks
whose luma leaves [0,255], so that REF_COMPAT runs of this content are comparable with the
reference (which has no Clip1).
"""
import numpy as np

from . import abi

_V0 = np.array([10, 11, 13, 14, 16, 18])
# raster index of the 4x4 block for (luma4x4BlkIdx)
_BLK_RASTER = np.array([0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15])
_RASTER_BLK = _BLK_RASTER  # involution


class WorkloadConfig:
    def __init__(self, **kw):
        self.qp = 28
        self.p_i16 = 0.35
        self.p_intra_in_p = 0.03
        self.p_skip = 0.35
        self.p_part = (0.40, 0.12, 0.12, 0.36)
        self.p_sub = (0.5, 0.0, 0.25, 0.25)       # 8x8, 8x4, 4x8, 4x4 (8x4 excluded by default: reference D9)
        self.mv_range = 64                         # quarter samples
        self.num_ref = 1
        self.p_cbp8 = 0.5
        self.p_block = 0.45
        self.p_chroma = (0.4, 0.25, 0.35)
        self.max_amp = 10.0
        self.p_t8x8 = 0.0                          # spec mode only
        self.p_i8x8 = 0.0                          # spec mode only (share of I_NxN MBs that are Intra8x8)
        self.compat = True
        self.chroma_qp_offset = 0
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)


def _levels(rng, shape, qp, cfg, p_nonzero):
    """sparse levels; shape = (..., 16) raster 4x4; low frequencies favoured."""
    amp_per_level = _V0[qp % 6] * (2.0 ** (qp // 6)) / 64.0
    maxlev = max(1, int(round(cfg.max_amp / amp_per_level)))
    # probability of a non-zero level decays with frequency index (row + col)
    freq = (np.arange(16) // 4 + np.arange(16) % 4).astype(np.float64)
    p = p_nonzero * np.exp(-0.9 * freq)
    nz = rng.random(shape) < p
    mag = rng.integers(1, maxlev + 1, size=shape)
    sign = rng.integers(0, 2, size=shape) * 2 - 1
    return (nz * mag * sign).astype(np.int16)


def make_picture(rng, w_mbs, h_mbs, is_p, cfg, ref_list=()):
    """One picture: dict(pp, hdr, mv, coeff)."""
    n = w_mbs * h_mbs
    hdr = np.zeros(n, dtype=abi.MB_HDR_DTYPE)
    hdr["flags"] = abi.avail_flags(w_mbs, h_mbs)
    qp = np.full(n, cfg.qp, dtype=np.int32) + rng.integers(-2, 3, size=n) * (rng.random(n) < 0.2)
    hdr["qp_y"] = qp
    qpc = abi.chroma_qp(qp, cfg.chroma_qp_offset, compat=cfg.compat)
    hdr["qp_cb"] = qpc
    hdr["qp_cr"] = qpc
    r = rng.random(n)
    if is_p:
        intra = r < cfg.p_intra_in_p
        skip = (~intra) & (r < cfg.p_intra_in_p + cfg.p_skip)
    else:
        intra = np.ones(n, dtype=bool)
        skip = np.zeros(n, dtype=bool)
    i16 = intra & (rng.random(n) < cfg.p_i16)
    inxn = intra & ~i16
    i8 = inxn & (rng.random(n) < cfg.p_i8x8)
    i4 = inxn & ~i8
    part = rng.choice(4, size=n, p=cfg.p_part)
    cls = np.where(i16, abi.MB_I16x16, np.where(i8, abi.MB_I8x8, np.where(i4, abi.MB_I4x4, abi.MB_P16x16 + part)))
    cls = np.where(skip, abi.MB_P16x16, cls)
    hdr["mb_class"] = cls
    hdr["flags"] |= np.where(skip, abi.SKIP, 0).astype(np.uint8)
    inter = ~intra
    t8 = (inter & ~skip & (cls != abi.MB_P8x8) & (rng.random(n) < cfg.p_t8x8)) | i8
    hdr["flags"] |= np.where(t8, abi.T8x8, 0).astype(np.uint8)

    # ---- coefficients -------------------------------------------------------------------------
    cbp8 = rng.random((n, 4)) < cfg.p_cbp8
    cbp8[i16] = (rng.random(int(i16.sum())) < 0.5)[:, None]
    cbp8[skip] = False
    cbpc = rng.choice(3, size=n, p=cfg.p_chroma)
    cbpc[skip] = 0
    luma = _levels(rng, (n, 16, 16), cfg.qp, cfg, 0.9)
    blk_coded = rng.random((n, 16)) < cfg.p_block
    quad_of_blk = np.arange(16) // 4
    luma *= (blk_coded & cbp8[:, quad_of_blk])[:, :, None]
    # Intra16x16: slot 0 carries the DC level (always allowed), AC only with cbp = 15
    dc16 = (_levels(rng, (n, 16, 16), cfg.qp, cfg, 0.9)[:, :, 0] // 2) * (rng.random((n, 16)) < 0.5)
    luma[i16, :, 0] = dc16[i16]
    # 8x8 transform blocks: the four 4x4 slots of a quadrant hold one raster 8x8; keep it low-frequency
    if t8.any():
        l8 = np.zeros((n, 4, 8, 8), dtype=np.int16)
        l8[:, :, :4, :4] = _levels(rng, (n, 4, 16), cfg.qp, cfg, 0.9).reshape(n, 4, 4, 4)
        l8 *= (cbp8 & (rng.random((n, 4)) < 0.8))[:, :, None, None]
        luma[t8] = l8.reshape(n, 16, 16)[t8]
    chroma = _levels(rng, (n, 8, 16), max(cfg.qp - 2, 0), cfg, 0.9)
    chroma[:, :, 1:] *= ((cbpc == 2)[:, None, None] & (rng.random((n, 8, 1)) < cfg.p_block))
    chroma[:, :, 0] *= (cbpc > 0)[:, None] & (rng.random((n, 8)) < 0.7)
    coeff = np.concatenate([luma, chroma], axis=1).reshape(n, 384)
    cbp_l = (cbp8 * np.array([1, 2, 4, 8])).sum(axis=1)
    hdr["cbp"] = (cbp_l | (cbpc << 4)).astype(np.uint8)

    # ---- intra syntax -------------------------------------------------------------------------
    mbx = np.tile(np.arange(w_mbs), h_mbs)
    mby = np.repeat(np.arange(h_mbs), w_mbs)
    A, B = mbx > 0, mby > 0
    m16 = rng.integers(0, 4, size=n)
    m16 = np.where((m16 == 0) & ~B, 2, m16)
    m16 = np.where((m16 == 1) & ~A, 2, m16)
    m16 = np.where((m16 == 3) & ~(A & B), 2, m16)
    mch = rng.integers(0, 4, size=n)
    mch = np.where((mch == 1) & ~A, 0, mch)
    mch = np.where((mch == 2) & ~B, 0, mch)
    mch = np.where((mch == 3) & ~(A & B), 0, mch)
    hdr["intra_modes"] = np.where(intra, (m16 & 3) | (mch << 2), 0).astype(np.uint8)
    # Intra4x4 modes per luma4x4BlkIdx, legal for the block's neighbours
    modes = rng.integers(0, 9, size=(n, 16))
    ras = _BLK_RASTER[np.arange(16)]
    left_ok = (ras % 4 > 0)[None, :] | A[:, None]
    top_ok = (ras // 4 > 0)[None, :] | B[:, None]
    need_top = np.isin(modes, (0, 3, 7)) | np.isin(modes, (4, 5, 6))
    need_left = np.isin(modes, (1, 8)) | np.isin(modes, (4, 5, 6))
    modes = np.where((need_top & ~top_ok) | (need_left & ~left_ok), 2, modes)
    # Intra8x8 modes per 8x8 quadrant
    modes8 = rng.integers(0, 9, size=(n, 4))
    q = np.arange(4)
    left8 = (q % 2 > 0)[None, :] | A[:, None]
    top8 = (q // 2 > 0)[None, :] | B[:, None]
    need_top8 = np.isin(modes8, (0, 3, 7, 4, 5, 6))
    need_left8 = np.isin(modes8, (1, 8, 4, 5, 6))
    modes8 = np.where((need_top8 & ~top8) | (need_left8 & ~left8), 2, modes8)
    packed = (modes[:, 0::2] | (modes[:, 1::2] << 4)).astype(np.uint8)
    packed8 = np.zeros((n, 8), dtype=np.uint8)
    packed8[:, 0] = modes8[:, 0] | (modes8[:, 1] << 4)
    packed8[:, 1] = modes8[:, 2] | (modes8[:, 3] << 4)
    u = np.zeros((n, 8), dtype=np.uint8)
    u[i4] = packed[i4]
    u[i8] = packed8[i8]

    # ---- inter syntax -------------------------------------------------------------------------
    mv = None
    if is_p:
        nref = max(1, min(cfg.num_ref, len(ref_list)))
        sub = rng.choice(4, size=(n, 4), p=cfg.p_sub)
        is88 = cls == abi.MB_P8x8
        hdr["sub_mb_type"] = np.where(is88, sub[:, 0] | (sub[:, 1] << 2) | (sub[:, 2] << 4) | (sub[:, 3] << 6), 0).astype(np.uint8)
        # one candidate vector per 4x4 block, then collapsed according to the partition shape
        cand = rng.integers(-cfg.mv_range, cfg.mv_range + 1, size=(n, 16, 2)).astype(np.int16)
        bx, by = np.arange(16) % 4, np.arange(16) // 4
        quad = (by // 2) * 2 + bx // 2
        src16 = np.zeros(16, dtype=np.int64)
        src168 = (by // 2) * 8
        src816 = (bx // 2) * 2
        src88 = (by // 2) * 8 + (bx // 2) * 2                     # top-left block of the quadrant
        src84 = by * 4 + (bx // 2) * 2                            # 8x4: row inside quadrant
        src48 = (by // 2) * 8 + bx                                # 4x8: column inside quadrant
        src44 = np.arange(16)
        sub_blk = sub[:, quad]                                    # (n,16)
        src_p88 = np.where(sub_blk == abi.SUB_8x8, src88[None, :], np.where(sub_blk == abi.SUB_8x4, src84[None, :],
                           np.where(sub_blk == abi.SUB_4x8, src48[None, :], src44[None, :])))
        src = np.where((cls == abi.MB_P16x16)[:, None], src16[None, :],
                       np.where((cls == abi.MB_P16x8)[:, None], src168[None, :],
                                np.where((cls == abi.MB_P8x16)[:, None], src816[None, :], src_p88)))
        mv = np.take_along_axis(cand, src[:, :, None].repeat(2, axis=2), axis=1)
        mv[intra] = 0
        refq = rng.integers(0, nref, size=(n, 4))
        # partitions share one ref_idx
        refq = np.where((cls == abi.MB_P16x16)[:, None], refq[:, :1],
                        np.where((cls == abi.MB_P16x8)[:, None], refq[:, [0, 0, 2, 2]],
                                 np.where((cls == abi.MB_P8x16)[:, None], refq[:, [0, 1, 0, 1]], refq)))
        u[inter, :4] = refq[inter].astype(np.uint8)
        mv = mv.reshape(n, 32)
    hdr["u"] = u
    pp = abi.default_pic_params(abi.SLICE_P if is_p else abi.SLICE_I, ref_list=list(ref_list)[:cfg.num_ref] if is_p else (),
                                deblock=(1, 0, 0) if cfg.compat else (0, 0, 0))
    return {"pp": pp, "hdr": hdr, "mv": mv, "coeff": np.ascontiguousarray(coeff)}


def make_gop(rng, w_mbs, h_mbs, n_frames, cfg):
    """IDR-led GOP: I then P pictures; ref_list0 entries are indices inside the GOP (most recent first)."""
    pics = []
    for i in range(n_frames):
        refs = [i - 1 - k for k in range(min(cfg.num_ref, i))]
        pics.append(make_picture(rng, w_mbs, h_mbs, i > 0, cfg, refs))
    return pics


def zero_mbs(pic, addrs):
    pic["coeff"][addrs] = 0
    pic["hdr"]["cbp"][addrs] = 0


def sanitize(engine_factory, w_mbs, h_mbs, gop, max_iter=24):
    """Zeroes the residual of macroblocks flagged by the engine's excursion map until a REF_COMPAT
    run of the GOP keeps every luma sample in [0,255].  Pictures are settled one at a time in decode
    order (a picture only depends on earlier ones), each with a few submit / inspect / zero rounds.
    Returns the total number of rounds."""
    eng = engine_factory(len(gop))
    rounds = 0
    try:
        for i, p in enumerate(gop):
            for it in range(max_iter):
                eng.submit_picture(i, p["pp"], p["hdr"], p["mv"], p["coeff"])
                eng.sync()
                rounds += 1
                if eng.range_excursions() == 0:
                    break
                bad = np.flatnonzero(eng.excursion_map(i))
                zero_mbs(p, bad)
            else:
                raise RuntimeError("picture %d did not converge to the [0,255] range" % i)
        return rounds
    finally:
        eng.close()
