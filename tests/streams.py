"""TEST INFRASTRUCTURE: recipes of the genuine H.264 bitstreams the tests use -- written by the repository's own CABAC writer
(libh264host.so, h264e_*), since the image holds no encoder and the reference ships no stream.

Three families:
  base_*   Baseline profile (CAVLC), standard-conformant: BASELINE config 1 ("baseline-profile stream") and config 2 geometry, decoded by
           libavcodec like the spec_* family;
  spec_*   High profile, standard-conformant: decoded by libavcodec (tests/lavc.py) to pin the spec-mode half of the engine and of
           the oracle -- Intra8x8, intra chroma prediction, chroma motion compensation, 8x8 transform, deblocking (SURVEY a13, a14,
           a19, a20, a21), which the reference decoder does not implement;
  ref_*    Main profile within what the reference can read (SURVEY 8(c) constraints + no sub-8x8 partitions, see D15 in DESIGN.md):
           decoded by the unmodified reference with its own arithmetic decoder (oracle/_ref/h264ref_cabac).
Fixtures made from them: tests/golden/make_spec_golden.py.
"""
import numpy as np

from videodecoder_b200 import hoststream, synth


def _high_seq(w, h, num_ref=1, **kw):
    return synth.make_seq(w, h, deblocking_control=1, max_num_ref_frames=max(num_ref, 1), num_ref_idx_default=num_ref - 1,
                          profile_idc=100, transform_8x8_mode=1, **kw)


def _encode(seq, pics, flags=hoststream.E_ABSOLUTE):
    stream, aus = hoststream.encode_stream(seq, pics, flags=flags, per_picture=True)
    return {"seq": seq, "pics": pics, "stream": stream, "aus": aus}


def spec_qcif_i():
    """BASELINE config 1 geometry, spec tools: 3 IDR pictures, Intra4x4 / Intra8x8 / Intra16x16 with every mode, deblocking on."""
    rng = np.random.default_rng(101)
    seq = _high_seq(11, 9, chroma_qp_index_offset=2, second_chroma_qp_index_offset=-3)
    pics = synth.random_gop(rng, seq, 3, synth.ContentConfig(p_i8x8=0.4), intra_only=True, absolute=True)
    for p, _ in pics:
        p.disable_deblocking_filter_idc = 0
    return _encode(seq, pics)


def spec_cif_ip():
    """BASELINE config 2 with the spec's tools: CIF, 1 I + 7 P, two references, every partition shape incl. 8x4, 8x8 transform on half
    of the eligible macroblocks, intra macroblocks (all three kinds) inside P pictures, deblocking with non-zero alpha / beta offsets."""
    rng = np.random.default_rng(102)
    seq = _high_seq(22, 18, num_ref=2, chroma_qp_index_offset=-2, second_chroma_qp_index_offset=3)
    cfg = synth.ContentConfig(num_ref=2, sub_types=(0, 1, 2, 3), p_intra_in_p=0.1, p_i8x8=0.4, p_t8x8=0.5, p_qp_delta=0.4, qp_delta_range=4)
    pics = synth.random_gop(rng, seq, 8, cfg, absolute=True)
    for i, (p, _) in enumerate(pics):
        p.disable_deblocking_filter_idc = 0
        p.slice_alpha_c0_offset_div2 = (0, 2, -2, 3)[i % 4]
        p.slice_beta_offset_div2 = (0, -1, 2, -3)[i % 4]
    return _encode(seq, pics)


def spec_qcif_wp():
    """explicit weighted prediction, luma and chroma, three references, vectors far outside the picture, deblocking on."""
    rng = np.random.default_rng(103)
    seq = _high_seq(11, 9, num_ref=3, weighted_pred=1)
    cfg = synth.ContentConfig(num_ref=3, sub_types=(0, 1, 2, 3), p_intra_in_p=0.1, p_t8x8=0.3, p_i8x8=0.3, mv_range=400, p_skip=0.2)
    pics = synth.random_gop(rng, seq, 6, cfg, absolute=True)
    for i, (p, _) in enumerate(pics):
        p.disable_deblocking_filter_idc = 0
        if p.slice_type == synth.SLICE_P:
            p.luma_log2_weight_denom = 5 if i % 2 else 0
            p.chroma_log2_weight_denom = 4 if i % 2 else 1
            for r in range(p.num_ref_idx_l0_active_minus1 + 1):
                p.luma_weight_flag[r] = 1 if (r + i) % 2 == 0 else 0
                p.luma_weight[r] = int(rng.integers(20, 44)) if i % 2 else 1
                p.luma_offset[r] = int(rng.integers(-6, 7))
                p.chroma_weight_flag[r] = 1 if (r + i) % 3 != 0 else 0
                for c in range(2):
                    p.chroma_weight[r][c] = int(rng.integers(10, 22)) if i % 2 else int(rng.integers(1, 4))
                    p.chroma_offset[r][c] = int(rng.integers(-5, 6))
    return _encode(seq, pics)


def spec_hd1080_ip(n_pictures=3):
    """BASELINE config 3: 1080p High (120x68 macroblocks), I + P, 8x8 transform and Intra8x8 on about half of the macroblocks,
    chroma prediction / MC, deblocking."""
    rng = np.random.default_rng(104)
    seq = _high_seq(120, 68)
    cfg = synth.ContentConfig(sub_types=(0, 1, 2, 3), p_intra_in_p=0.05, p_i8x8=0.5, p_t8x8=0.5)
    pics = synth.random_gop(rng, seq, n_pictures, cfg, absolute=True)
    for p, _ in pics:
        p.disable_deblocking_filter_idc = 0
    return _encode(seq, pics)


def spec_dpb():
    """reference-picture management: 4 reference frames, non-reference pictures, ref_pic_list_modification, memory management
    control operations 1 (drop a short-term picture) and 3 / 2 (long-term), a second IDR -- every P macroblock names a reference,
    so a wrong list shows in the samples."""
    rng = np.random.default_rng(105)
    seq = _high_seq(11, 9, num_ref=4)
    seq.log2_max_frame_num_minus4 = 0           # MaxFrameNum 16: frame_num wraps inside the sequence
    cfg = synth.ContentConfig(num_ref=4, sub_types=(0, 1, 2, 3), p_skip=0.1, p_intra_in_p=0.02, p_t8x8=0.3)
    pics = []
    frame_num = 0
    for i in range(30):
        idr = i in (0, 24)
        st = synth.SLICE_I if idr else synth.SLICE_P
        if idr:
            frame_num = 0
        pic = synth.make_pic(seq, 0, st)
        pic.is_idr = 1 if idr else 0
        is_ref = idr or (i % 5 != 3)                   # every fifth picture is a non-reference picture
        pic.nal_ref_idc = (3 if idr else 2) if is_ref else 0
        pic.frame_num = frame_num % 16
        pic.poc_lsb = 2 * (i if i < 24 else i - 24)
        pic.idr_pic_id = 1 if i == 24 else 0
        pic.disable_deblocking_filter_idc = 0
        n_avail = min(4, sum(1 for j, (q, _) in enumerate(pics) if q.nal_ref_idc and (i < 24 or j >= 24)))
        if st == synth.SLICE_P:
            n_active = max(1, min(n_avail, 1 + i % 4))
            pic.num_ref_idx_l0_active_minus1 = n_active - 1
            if i % 6 == 2 and n_avail >= 2:
                pic.n_mods = min(2, n_active)           # bring an older picture to the front, then the newest one behind it
                pic.mods[0][0], pic.mods[0][1] = 0, 1
                pic.mods[1][0], pic.mods[1][1] = 1, 0
            if is_ref and i == 7:
                pic.n_mmco = 1
                pic.mmco[0][0], pic.mmco[0][1] = 1, 1   # drop the picture with PicNum = CurrPicNum - 2
            if is_ref and i == 10:
                pic.n_mmco = 2                          # one short-term picture leaves, the previous reference picture becomes long-term 0
                pic.mmco[0][0], pic.mmco[0][1] = 1, 1
                pic.mmco[1][0], pic.mmco[1][1], pic.mmco[1][2] = 3, 0, 0
            if is_ref and i == 16:
                pic.n_mmco = 1
                pic.mmco[0][0], pic.mmco[0][1] = 2, 0   # long-term picture 0 leaves
            cfg_i = synth.ContentConfig(**{**cfg.__dict__, "num_ref": n_active})
        else:
            cfg_i = cfg
        mbs = synth.random_picture(rng, seq, st, cfg_i, seq.pic_init_qp, absolute=True)
        pics.append((pic, mbs))
        if is_ref:
            frame_num += 1
    return _encode(seq, pics)


def ref_cif_ip(n_pictures=6, w=22, h=18, seed=201, sanitize=None):
    """What the unmodified reference can read with its own CABAC: Main profile, I + P, 16x16 / 16x8 / 8x16 / 8x8 partitions (no
    sub-8x8 partitions: the reference's mvd context selection diverges from the standard there), two references.
    sanitize: callable(seq, pics) that removes residual where the reference's unclipped samples leave [0,255] (D3)."""
    rng = np.random.default_rng(seed)
    seq = synth.make_seq(w, h, max_num_ref_frames=2, num_ref_idx_default=1)
    cfg = synth.ContentConfig(num_ref=2, sub_types=(0,), p_intra_in_p=0.08)
    pics = synth.random_gop(rng, seq, n_pictures, cfg)
    if sanitize is not None:
        sanitize(seq, pics)
    out = _encode(seq, pics, flags=0)
    out["stream"] += b"\x00\x00\x00\x01\x0c\xff\x80"          # the reference's NAL scanner needs a start code after the last unit
    return out


def base_qcif_i():
    """BASELINE config 1 as written: QCIF 176x144, I slices only, BASELINE profile (profile_idc 66: CAVLC entropy coding, 4x4 transform):
    3 IDR pictures, Intra4x4 with every mode and Intra16x16, deblocking on."""
    rng = np.random.default_rng(111)
    seq = synth.make_seq(11, 9, deblocking_control=1, profile_idc=66, chroma_qp_index_offset=2)
    pics = synth.random_gop(rng, seq, 3, synth.ContentConfig(p_qp_delta=0.3, qp_delta_range=4), intra_only=True, absolute=True)
    for p, _ in pics:
        p.disable_deblocking_filter_idc = 0
    return _encode(seq, pics)


def base_cif_ip():
    """BASELINE config 2 geometry as a Baseline-profile (CAVLC) stream: CIF, 1 I + 5 P, two references, every partition shape, P_Skip
    runs, intra macroblocks inside P pictures, QP deltas, deblocking on."""
    rng = np.random.default_rng(112)
    seq = synth.make_seq(22, 18, deblocking_control=1, profile_idc=66, max_num_ref_frames=2, num_ref_idx_default=1, chroma_qp_index_offset=-2)
    cfg = synth.ContentConfig(num_ref=2, sub_types=(0, 1, 2, 3), p_intra_in_p=0.1, p_qp_delta=0.4, qp_delta_range=4)
    pics = synth.random_gop(rng, seq, 6, cfg, absolute=True)
    for i, (p, _) in enumerate(pics):
        p.disable_deblocking_filter_idc = 0
        p.slice_alpha_c0_offset_div2 = (0, 2, -2)[i % 3]
    return _encode(seq, pics)


SPEC_RECIPES = {"spec_qcif_i": spec_qcif_i, "spec_cif_ip": spec_cif_ip, "spec_qcif_wp": spec_qcif_wp, "spec_dpb": spec_dpb,
                "spec_hd1080_ip": spec_hd1080_ip, "base_qcif_i": base_qcif_i, "base_cif_ip": base_cif_ip}


# ---- known-answer streams: one tool at a time (each intra mode, each boundary-strength class) ------------------------
def _levels(rng, m, qp, p=0.5):
    """a few low-frequency levels in every coded block of record m (cbp and transform flags already set)"""
    cfg = synth.ContentConfig(nnz_max=3, max_amp=12.0)
    if m["kind"] == synth.I16x16:
        synth._fill_block(rng, m["luma_dc"], 16, qp, cfg)
        m["luma_dc"] //= 2
    for q in range(4):
        if not (int(m["cbp_luma"]) >> q) & 1:
            continue
        if m["t8x8"]:
            synth._fill_block(rng, m["luma"][4 * q:4 * q + 4].reshape(-1), 64, qp, cfg)
        else:
            for b in range(4 * q, 4 * q + 4):
                if rng.random() < p:
                    synth._fill_block(rng, m["luma"][b], 15 if m["kind"] == synth.I16x16 else 16, qp, cfg)
    if m["cbp_chroma"]:
        for c in range(2):
            synth._fill_block(rng, m["chroma_dc"][c], 4, qp, cfg)
            if m["cbp_chroma"] == 2:
                for b in range(4):
                    if rng.random() < p:
                        synth._fill_block(rng, m["chroma_ac"][c][b], 15, qp, cfg)


def kat_intra_modes():
    """One IDR picture per mode: Intra4x4 modes 0..8, Intra8x8 modes 0..8, Intra16x16 modes 0..3 -- every block of every macroblock
    uses the picture's mode wherever its neighbours allow it (DC elsewhere); intra_chroma_pred_mode cycles 0..3 over the pictures."""
    rng = np.random.default_rng(301)
    W, H = 6, 5
    seq = _high_seq(W, H)
    pics = []
    plan = [(synth.I4x4, m) for m in range(9)] + [(synth.I8x8, m) for m in range(9)] + [(synth.I16x16, m) for m in range(4)]
    for i, (kind, mode) in enumerate(plan):
        pic = synth.make_pic(seq, 0, synth.SLICE_I)
        pic.idr_pic_id = i
        pic.disable_deblocking_filter_idc = 1 if i % 2 else 0
        mbs = np.zeros(W * H, dtype=synth.MB_DTYPE)
        for a in range(W * H):
            mby, mbx = divmod(a, W)
            m = mbs[a]
            m["kind"] = kind
            cm = i % 4
            if (cm == 1 and mbx == 0) or (cm == 2 and mby == 0) or (cm == 3 and (mbx == 0 or mby == 0)):
                cm = 0
            m["chroma_mode"] = cm
            m["cbp_chroma"] = int(rng.integers(0, 3))
            if kind == synth.I16x16:
                ok = {0: mby > 0, 1: mbx > 0, 2: True, 3: mbx > 0 and mby > 0}[mode]
                m["i16_mode"] = mode if ok else 2
                m["cbp_luma"] = 15 if rng.random() < 0.5 else 0
            else:
                m["cbp_luma"] = int(rng.integers(0, 16))
                m["t8x8"] = 1 if kind == synth.I8x8 else 0
                nb, step = (16, 1) if kind == synth.I4x4 else (4, 2)
                for b in range(nb):
                    ras = synth.BLK_RASTER[b] if kind == synth.I4x4 else (b >> 1) * 8 + (b & 1) * 2
                    bx, by = (ras % 4), (ras // 4)
                    left, top = mbx > 0 or bx > 0, mby > 0 or by > 0
                    need_top, need_left = mode in (0, 3, 7, 4, 5, 6), mode in (1, 8, 4, 5, 6)
                    m["rem_mode"][b] = mode if (not need_top or top) and (not need_left or left) else 2
            _levels(rng, m, seq.pic_init_qp)
        pics.append((pic, mbs))
    return _encode(seq, pics)


def kat_deblock_strengths():
    """P pictures built so that each boundary-strength class of 8.7.2.1 dominates one picture: bS 0 (same vector, same reference, no
    levels), bS 1 by vector difference >= 4 quarter samples, bS 1 by different reference pictures, bS 2 (levels on one side of the edge,
    4x4 and 8x8 transform), bS 3 / 4 (intra macroblocks inside the P picture); filterOffsets -6 .. +6 and a QP step across macroblocks."""
    rng = np.random.default_rng(302)
    W, H = 8, 5
    seq = _high_seq(W, H, num_ref=2)
    n = W * H

    def p_picture(i, vec, ref, coded, intra=(), t8=False, offs=(0, 0)):
        pic = synth.make_pic(seq, i, synth.SLICE_P, num_ref_active=2 if i >= 2 else 1)
        pic.disable_deblocking_filter_idc = 0
        pic.slice_alpha_c0_offset_div2, pic.slice_beta_offset_div2 = offs
        mbs = np.zeros(n, dtype=synth.MB_DTYPE)
        for a in range(n):
            m = mbs[a]
            mby, mbx = divmod(a, W)
            if a in intra:
                m["kind"] = synth.I16x16 if a % 2 else synth.I4x4
                m["i16_mode"] = 2
                m["rem_mode"][:] = 2
                m["cbp_luma"] = 15 if a % 2 else 5
                m["cbp_chroma"] = 1
            else:
                m["kind"] = (synth.P16x16, synth.P16x8, synth.P8x16, synth.P8x8)[a % 4 if vec == "mixed" else 0]
                m["sub_type"][:] = (a // 4) % 4 if m["kind"] == synth.P8x8 else 0
                m["ref_idx"][:] = ref(a) if i >= 2 else 0
                v = {"same": (8, -4), "steps": (8 + 5 * (mbx % 3), -4 + 6 * (mby % 2)), "mixed": None}[vec]
                m["mvd"][:] = v if v is not None else rng.integers(-20, 21, size=(16, 2))
                if coded(a):
                    m["cbp_luma"] = int(rng.integers(1, 16))
                    m["cbp_chroma"] = int(rng.integers(0, 3))
                    m["t8x8"] = 1 if t8 and (m["kind"] != synth.P8x8 or not m["sub_type"].any()) else 0
                    m["qp_delta"] = int(rng.integers(-3, 4))
            _levels(rng, m, seq.pic_init_qp, p=0.7)
        return pic, mbs

    pic0 = synth.make_pic(seq, 0, synth.SLICE_I)
    pic0.disable_deblocking_filter_idc = 0
    mbs0 = synth.random_picture(rng, seq, synth.SLICE_I, synth.ContentConfig(p_i8x8=0.3), seq.pic_init_qp, absolute=True)
    pics = [(pic0, mbs0),
            p_picture(1, "same", lambda a: 0, lambda a: False),                                  # bS 0 everywhere
            p_picture(2, "steps", lambda a: 0, lambda a: False, offs=(3, -2)),                   # bS 1: vector differences
            p_picture(3, "same", lambda a: a % 2, lambda a: False, offs=(-6, 6)),                # bS 1: different reference pictures
            p_picture(4, "same", lambda a: 0, lambda a: (a + a // W) % 2 == 0, offs=(6, -6)),    # bS 2, 4x4 transform, QP steps
            p_picture(5, "same", lambda a: 0, lambda a: a % 3 == 0, t8=True),                    # bS 2, 8x8 transform
            p_picture(6, "mixed", lambda a: a % 2, lambda a: a % 2 == 0, intra={5, 6, 13, 14, 22, 27, 33}, t8=True, offs=(2, 2))]   # bS 3 / 4 and everything
    return _encode(seq, pics)


KAT_RECIPES = {"kat_intra_modes": kat_intra_modes, "kat_deblock_strengths": kat_deblock_strengths}
