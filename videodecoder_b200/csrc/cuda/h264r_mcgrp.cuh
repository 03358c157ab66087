// h264r_mcgrp.cuh -- K1+K3 for inter macroblocks, GROUP version (REF_COMPAT, lean pictures: picture buffers with
// partition-order vectors and compact levels, no weighting).
//
// k_inter (h264r_mc.cuh) spends a warp per macroblock on everything, and ncu's source page shows where that hurts
// (profiles/r2_ncu_kernels.csv, 908 warp instructions per macroblock): the residual phase runs with lane = block while
// 3.8 of a macroblock's 24 blocks are coded (about 210 instructions per macroblock for 16 % useful lanes), and the
// warp-uniform part -- header fields, offsets, coordinates, shape -- costs every lane the same instructions again.
// Here a warp takes GRP_MBS consecutive macroblocks at a time:
//   G0  lane i reads header + offsets of macroblock i and works out which of its blocks carry a residual
//       (ACTIVE blocks: stored luma blocks whose coded_block_pattern bit is set; the four blocks of a chroma plane
//       that has a stored block -- the chroma DC transform spreads over the plane), its coordinates, its record;
//   G1  the active blocks of the group become a dense item list (macroblock, block), slot = position in the list;
//   G2  lane = ITEM: dequantisation + inverse transform of 32 coded blocks at a time, every lane busy, into the
//       group's residual slots (reference residual::Decode and callees, h264/residual.cpp:184-434) -- none of this
//       reads a sample, so under programmatic stream serialisation it runs before the previous kernel has drained;
//   then per macroblock the window staging and the interpolation of h264r_mc.cuh (mc_patch), the residual rows
//       coming from the slots (a lane finds its block's slot with one popcount).
// Same arithmetic as seq_inter_mb_fast; the CPU test-suite runs this sequence in the host emulator against the oracle.
#pragma once
#include "h264r_mc.cuh"

namespace h264r {

#ifndef H264R_GRP_MBS
#define H264R_GRP_MBS 8
#endif
#ifndef H264R_GRP_SLOTS
#define H264R_GRP_SLOTS 64
#endif
constexpr int GRP_MBS = H264R_GRP_MBS;          // macroblocks a warp takes at a time
constexpr int GRP_SLOTS = H264R_GRP_SLOTS;      // residual blocks the working set holds; a group with more active blocks is taken in parts

struct alignas(16) GrpMb {          // record of one macroblock of the group, written by its lane in G0
    h264r_mb_hdr h;
    uint32_t mv_off, blk_mask, blk_off, active;     // active: bit b = block b has a residual slot
    int32_t yoff, coff;                             // byte offset of the macroblock's first luma / chroma sample from the plane origin
    uint32_t xy;                                    // mbx | mby << 16
    uint32_t meta;                                  // GRP_* fields below
};
// GrpMb::meta: slot of the first active block (G1, bits 0-7), number of active blocks (8-15), window shape (16-17), on the picture's
// border (18), inter macroblock inside the picture (19)
constexpr uint32_t GRP_BORDER = 1u << 18, GRP_INTER = 1u << 19;
HD int grp_slot0(uint32_t meta) { return (int)(meta & 255u); }
HD int grp_n_items(uint32_t meta) { return (int)((meta >> 8) & 255u); }
HD int grp_shape(uint32_t meta) { return (int)((meta >> 16) & 3u); }
// Windows are staged as the ALIGNED words that contain them: a window that starts `s` bytes into its first word is normalised by
// the consumer with one funnel shift per word (mc_groups<true>), so a staged word is one load and one store -- and a 16x16
// macroblock's window can equally well arrive as a tensor copy (TMA), whose boxes must start on a 16-byte column anyway.
// Tensor-copy boxes of 16x16 macroblocks (48 x 22 bytes, pitch 12 words) alternate between the first 1056 bytes of `win` and
// `win_b`, so that the box of the NEXT macroblock is in flight while the current one is interpolated.
constexpr int GRP_TMA_WORDS = TMA0_W * TMA0_H / 4;       // 264
struct alignas(128) GroupWork {
    uint32_t win[512];              // staged windows, layouts of h264r_mc.cuh (phase_inter_stage)
    uint4 slots[GRP_SLOTS * 2];     // residual blocks: 16 int16 each, raster 4x4
    GrpMb mb[GRP_MBS];
    uint8_t items[GRP_SLOTS];       // item -> macroblock of the group << 5 | block
    uint8_t need[32];               // host emulation of the warp reduction over the lanes' NEED bits
};

// what the tensor-copy variant keeps next to the working set (kept out of GroupWork: shared memory the default kernel does not
// use is L1 cache it loses)
struct alignas(128) GroupTma {
    uint32_t win_b[GRP_TMA_WORDS + 24];   // second tensor-copy buffer (+ slack: the consumer's fourth word of a row may lie behind the last row)
    unsigned long long mbar[2];           // completion barriers of the two tensor-copy buffers
};

// bit q of a 4-bit coded_block_pattern -> nibble q (luma4x4BlkIdx 4q..4q+3 belong to 8x8 block q)
HD uint32_t cbp_luma_blocks(uint32_t cbp_l)
{
    const uint32_t x = (cbp_l | (cbp_l << 3) | (cbp_l << 6) | (cbp_l << 9)) & 0x1111u;
    return x * 15u;
}

// G0: lane i < GRP_MBS prepares macroblock a0 + i.  A macroblock beyond the picture or an intra one gets no items and is
// skipped by the sequence (the intra kernels reconstruct it afterwards).
HD void phase_grp_load(GroupWork& w, const PicDesc& pd, const Geometry& g, int a0, int n_mbs, int lane)
{
    if (lane >= GRP_MBS) return;
    GrpMb m;
    const int a = a0 + lane;
    if (a >= n_mbs) {
        const int4 z = int4{0, 0, 0, 0};
        m.h = *reinterpret_cast<const h264r_mb_hdr*>(&z);
        m.mv_off = m.blk_mask = m.blk_off = m.active = 0u;
        m.yoff = m.coff = 0;
        m.xy = m.meta = 0u;
        w.mb[lane] = m;
        return;
    }
    const int4 raw = H264R_LDG(reinterpret_cast<const int4*>(pd.hdr + a));
    m.h = *reinterpret_cast<const h264r_mb_hdr*>(&raw);
    m.mv_off = H264R_LDG(pd.mv_off + a);
    m.blk_mask = H264R_LDG(pd.blk_mask + a);
    m.blk_off = H264R_LDG(pd.blk_off + a);
    uint32_t active = 0u, meta = 0u;
    if (!is_intra(m.h.mb_class)) {
        const uint32_t cbp_l = m.h.cbp & 15u, cbp_c = (m.h.cbp >> 4) & 3u;
        active = m.blk_mask & cbp_luma_blocks(cbp_l);
        if (cbp_c) {
            if (m.blk_mask & 0x0F0000u) active |= 0x0F0000u;
            if (m.blk_mask & 0xF00000u) active |= 0xF00000u;
        }
        meta = GRP_INTER | ((uint32_t)inter_shape_by_type(m.h) << 16);
    }
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    if (is_border_mb(g, mbx, mby)) meta |= GRP_BORDER;
    m.active = active;
    m.meta = meta | ((uint32_t)popc32(active) << 8);
    m.xy = (uint32_t)mbx | ((uint32_t)mby << 16);
    m.yoff = mby * 16 * g.pitch_y + mbx * 16;
    m.coff = mby * 8 * g.pitch_c + mbx * 8;
    w.mb[lane] = m;
}

// G1: the part of the group taken now = the longest run of macroblocks from `first` whose active blocks fit the slots
// (one macroblock has at most 24); lane i of the part numbers the items of macroblock i.
HD void phase_grp_items(GroupWork& w, int first, int lane, int& take_end_out, int& n_items_out)
{
    int cum = 0, take_end = first, slot0 = 0;
#pragma unroll
    for (int k = 0; k < GRP_MBS; k++) {
        const int c = grp_n_items(w.mb[k].meta);
        if (k >= first && take_end == k && cum + c <= GRP_SLOTS) {
            if (k == lane) slot0 = cum;
            cum += c;
            take_end = k + 1;
        }
    }
    if (lane >= first && lane < take_end) {
        w.mb[lane].meta |= (uint32_t)slot0;
        uint32_t act = w.mb[lane].active;
        int s = slot0;
        while (act) {
#if defined(__CUDA_ARCH__)
            const int b = __ffs((int)act) - 1;
#else
            const int b = __builtin_ctz(act);
#endif
            act &= act - 1;
            w.items[s++] = (uint8_t)((lane << 5) | b);
        }
    }
    take_end_out = take_end;       // the same in every lane
    n_items_out = cum;
}

// G2: lane = item.  Residual of one active block into its slot; arithmetic of phase_residual_fast (h264r_intra.cuh) for
// an inter macroblock with compact storage.
template <bool COMPAT>
HD void phase_grp_residual(GroupWork& w, const PicDesc& pd, int item, int n_items)
{
    if (item >= n_items) return;
    const int code = w.items[item], i = code >> 5, b = code & 31;
    const GrpMb& m = w.mb[i];
    const BlkRef br{m.blk_mask, m.blk_off};
    const int fmt = H264R_LEVELS_COMPACT;
    const bool luma = b < 16;
    const int cbp_c = (m.h.cbp >> 4) & 3;
    int16_t* dst = reinterpret_cast<int16_t*>(&w.slots[item * 2]);       // 4 rows of 4 int16, private to this lane
    const uint8_t* src = blk_ptr(pd, br, b, fmt);
    int4 v0 = int4{0, 0, 0, 0}, v1 = int4{0, 0, 0, 0};
    if (src && blk_compact(fmt, br)) {
        // compact block: the levels are scattered into the (zeroed) slot, which has the raster shape of a block, and read back as rows
        const uint2 v = H264R_LDG(reinterpret_cast<const uint2*>(src));
        uint32_t map = v.x & 0xffffu;
        unsigned long long lv = ((unsigned long long)v.y << 16) | (v.x >> 16);
        w.slots[item * 2] = uint4{0u, 0u, 0u, 0u};
        w.slots[item * 2 + 1] = uint4{0u, 0u, 0u, 0u};
        while (map) {
#if defined(__CUDA_ARCH__)
            const int q = __ffs((int)map) - 1;
#else
            const int q = __builtin_ctz(map);
#endif
            map &= map - 1;
            dst[q] = (int16_t)(int8_t)(lv & 255u);
            lv >>= 8;
        }
        const uint4 r0 = w.slots[item * 2], r1 = w.slots[item * 2 + 1];
        v0 = int4{(int)r0.x, (int)r0.y, (int)r0.z, (int)r0.w};
        if (luma || cbp_c == 2) v1 = int4{(int)r1.x, (int)r1.y, (int)r1.z, (int)r1.w};
    } else if (src) {
        v0 = blk_levels8(pd, br, src, 0, fmt);
        if (luma || cbp_c == 2) v1 = blk_levels8(pd, br, src, 8, fmt);
    }
    if (!luma && cbp_c != 2) { v0.x &= 0xffff; v0.y = v0.z = v0.w = 0; }        // DC only: AC levels are not coded
    int c[16];
    const int wv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; k++) { c[2 * k] = (int)(int16_t)(wv[k] & 0xffff); c[2 * k + 1] = wv[k] >> 16; }
    int qp = m.h.qp_y;
    bool bypass = false;
    if (!luma) {
        const int plane = (b - 16) >> 2, cb = (b - 16) & 3;
        qp = plane ? m.h.qp_cr : m.h.qp_cb;
        int f;
        const uint8_t* d3 = blk_ptr(pd, br, 19 + plane * 4, fmt);
        const int c3 = d3 ? blk_level0(pd, br, d3, fmt) : 0;
        if (COMPAT) f = (cb == 0 || cb == 3) ? c3 : -c3;      // reference's chroma DC product (h264/matrix.cpp:150-171)
        else {
            const uint8_t* d0 = blk_ptr(pd, br, 16 + plane * 4, fmt), * d1 = blk_ptr(pd, br, 17 + plane * 4, fmt), * d2 = blk_ptr(pd, br, 18 + plane * 4, fmt);
            const int c0 = d0 ? blk_level0(pd, br, d0, fmt) : 0, c1 = d1 ? blk_level0(pd, br, d1, fmt) : 0, c2 = d2 ? blk_level0(pd, br, d2, fmt) : 0;
            f = c0 + ((cb & 1) ? -c1 : c1) + ((cb & 2) ? -c2 : c2) + ((cb == 1 || cb == 2) ? -c3 : c3);
        }
        const int s = (qp * 43) >> 8, mq = qp - 6 * s;
        c[0] = ((f * 16 * norm_adjust_dc(mq)) * (1 << s)) >> 5;
        bypass = true;
    }
    int nz = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) nz |= c[k];
    uint4 o0 = uint4{0u, 0u, 0u, 0u}, o1 = o0;
    if (nz) {
        int r[16];
        const Dequant q = make_dequant(qp);
        idct4_fast(c, q, bypass, r);
        o0 = uint4{pack_sat_s16x2(r[0], r[1]), pack_sat_s16x2(r[2], r[3]), pack_sat_s16x2(r[4], r[5]), pack_sat_s16x2(r[6], r[7])};
        o1 = uint4{pack_sat_s16x2(r[8], r[9]), pack_sat_s16x2(r[10], r[11]), pack_sat_s16x2(r[12], r[13]), pack_sat_s16x2(r[14], r[15])};
    }
    w.slots[item * 2] = o0;
    w.slots[item * 2 + 1] = o1;
}

// the lane's own vector (block lane >> 1) of macroblock m
HD uint32_t grp_load_mv(const PicDesc& pd, const GrpMb& m, int lane)
{
    const uint32_t* mv32 = reinterpret_cast<const uint32_t*>(pd.mv);
    return H264R_LDG(mv32 - 1 - (ptrdiff_t)(m.mv_off + (uint32_t)mv_part_index(m.h.mb_class, m.h.sub_mb_type, lane >> 1)));
}

// Items first, first + STRIDE, ... of a window of NROWS x NWORDS ALIGNED words: word (row, j) is the aligned word at byte
// (sx & ~3) + 4 j of row sy + row of the padded reference plane (the plane origin is 16-byte aligned, so sx & 3 is the misalignment).
template <int NROWS, int NWORDS, int STRIDE, int DPITCH>
HD void stage_window_raw(uint32_t* dst, const uint8_t* ref, int gpitch, int sx, int sy, int first)
{
    constexpr int DR = STRIDE / NWORDS, DJ = STRIDE % NWORDS;       // one step = DR rows + DJ words
    int row = first / NWORDS, j = first - row * NWORDS;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(ref + (ptrdiff_t)(sy + row) * gpitch + (sx & ~3)) + j;
    const int wpitch = gpitch >> 2;
#pragma unroll
    for (int i = 0; i < (NROWS * NWORDS + STRIDE - 1) / STRIDE; i++) {
        const int it = first + i * STRIDE;
        if (it < NROWS * NWORDS) dst[DPITCH == NWORDS ? it : row * DPITCH + j] = H264R_LDG(q);
        q += DR * wpitch + DJ;
        if (DPITCH != NWORDS) row += DR;
        if (DJ) {
            j += DJ;
            if (j >= NWORDS) { j -= NWORDS; if (DPITCH != NWORDS) row++; q += wpitch - NWORDS; }
        }
    }
}

// integer position of the lane's prediction block (clamped as phase_inter_stage clamps it) -> x of the window's first byte
HD int grp_window_x(const Geometry& g, const GrpMb& m, int lane, uint32_t v)
{
    const int shape = grp_shape(m.meta), W = g.w_mbs * 16, x0 = (int)(m.xy & 0xffffu) * 16;
    const int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
    if (shape == 0) return clip3(-19, W + 2, x0 + (mv_x(v) >> 2)) - 2;
    if (shape == 1) return clip3(-11, W + 2, x0 + (q & 1) * 8 + (mv_x(v) >> 2)) - 2;
    return clip3(-7, W + 2, x0 + bx * 4 + (mv_x(v) >> 2)) - 2;
}

// M0: windows of macroblock m, load staging.  RAW: the aligned words that contain the window (the consumer shifts: the tensor-copy
// variant, where every shape shares the shifted consumer); else words already shifted to the window's first byte (stage_window).
template <bool COMPAT, bool RAW>
HD void phase_grp_stage(GroupWork& w, const PicDesc& pd, const Geometry& g, const GrpMb& m, int lane, uint32_t v)
{
    const h264r_mb_hdr& h = m.h;
    const int shape = grp_shape(m.meta);
    const int H = g.h_mbs * 16, y0 = (int)(m.xy >> 16) * 16;
    const int sx = grp_window_x(g, m, lane, v);
#if !defined(__CUDA_ARCH__)
    w.need[lane] = (uint8_t)mc_need(mv_x(v) & 3, mv_y(v) & 3);
#endif
    if (shape == 0) {
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, 0)];
        const int yb = clip3(-19, H + 3, y0 + (mv_y(v) >> 2));
        if (RAW) stage_window_raw<22, 6, 32, WIN0_PITCH>(w.win, ref, g.pitch_y, sx, yb - 3, lane);
        else stage_window<22, 6, 32, WIN0_PITCH>(w.win, ref, g.pitch_y, sx, yb - 3, lane);
    } else if (shape == 1) {
        const int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, q)];
        const int yb = clip3(-11, H + 3, y0 + (q >> 1) * 8 + (mv_y(v) >> 2));
        if (RAW) stage_window_raw<14, 4, 8, WIN1_PITCH>(w.win + q * WIN1_STRIDE, ref, g.pitch_y, sx, yb - 3, ((by & 1) * 2 + (bx & 1)) * 2 + (lane & 1));
        else stage_window<14, 4, 8, WIN1_PITCH>(w.win + q * WIN1_STRIDE, ref, g.pitch_y, sx, yb - 3, ((by & 1) * 2 + (bx & 1)) * 2 + (lane & 1));
    } else {
        const int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, q)];
        int ys = y0 + by * 4;
        // reference's source-row offset for 4x4 sub-partitions 2 and 3 (D8, h264/macroblock.cpp:786-789)
        if (COMPAT && h.mb_class == H264R_MB_P8x8 && ((h.sub_mb_type >> (2 * q)) & 3) == H264R_SUB_4x4 && (by & 1)) ys -= 3;
        const int yb = clip3(-7, H + 3, ys + (mv_y(v) >> 2));
        if (RAW) stage_window_raw<10, 3, 2, 3>(w.win + b * 30, ref, g.pitch_y, sx, yb - 3, lane & 1);
        else stage_window<10, 3, 2, 3>(w.win + b * 30, ref, g.pitch_y, sx, yb - 3, lane & 1);
    }
}

#if defined(__CUDACC__)
// M0 by tensor copy (16x16 macroblocks): lane 0 posts the 48 x 22 box of macroblock m -- the 16-byte columns that contain its
// window -- into `dst` and arms `bar`; mv0 = the macroblock's vector.  Tensor coordinates are relative to the padded plane.
__device__ __forceinline__ void grp_tma_post(uint32_t* dst, unsigned long long* bar, const PicDesc& pd, const Geometry& g, const GrpMb& m, uint32_t mv0)
{
    const int W = g.w_mbs * 16, H = g.h_mbs * 16, x0 = (int)(m.xy & 0xffffu) * 16, y0 = (int)(m.xy >> 16) * 16;
    const int xb = clip3(-19, W + 2, x0 + (mv_x(mv0) >> 2)), yb = clip3(-19, H + 3, y0 + (mv_y(mv0) >> 2));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was last read through the generic proxy
    mbar_expect_tx(bar, (uint32_t)(TMA0_W * TMA0_H));
    tma_load_2d(dst, static_cast<const char*>(pd.tmaps) + ((size_t)tma_ref_slot(pd, hdr_ref_idx(m.h, 0)) * 2) * 128, (PAD_X + xb - 2) & ~15, PAD_Y + yb - 3, bar);
}
#endif

// packed prediction + two residual pairs -> packed samples; returns the number of samples that left [0,255]
template <bool COMPAT>
HD uint32_t add_residual_words(uint32_t pred, uint32_t rx, uint32_t ry, uint32_t& packed)
{
    if (!COMPAT) { rx = clamp_s16x2(rx); ry = clamp_s16x2(ry); }
    uint32_t s0 = add_s16x2(prmt(pred, 0u, 0x4140), rx), s1 = add_s16x2(prmt(pred, 0u, 0x4342), ry);
    uint32_t n = 0;
    if ((s0 | s1) & 0xff00ff00u)
        n = ((s0 & 0xff00u) != 0) + ((s0 & 0xff000000u) != 0) + ((s1 & 0xff00u) != 0) + ((s1 & 0xff000000u) != 0);
    if (!COMPAT) { s0 = clip255_s16x2(s0); s1 = clip255_s16x2(s1); }
    packed = prmt(s0, s1, 0x6420);          // COMPAT: the plane keeps the low 8 bits (D3)
    return n;
}

// M1: interpolation, residual from the slots, stores (REF_COMPAT: chroma = residual, D4)
// tbuf: null = the windows were load-staged into w.win; else the tensor-copy buffer that holds the macroblock's 48 x 22 box
// SH: the windows hold aligned words (phase_grp_stage<.., RAW> / tensor copies) and the consumer shifts
template <bool COMPAT, bool SH>
HD void phase_grp_compute(GroupWork& w, const PicDesc& pd, const Geometry& g, const GrpMb& m, int a, int lane, uint32_t v, const uint32_t* tbuf = nullptr)
{
    static_assert(COMPAT, "the group path covers REF_COMPAT");
    int wn;
#if defined(__CUDA_ARCH__)
    wn = (int)__reduce_or_sync(0xffffffffu, (unsigned)mc_need(mv_x(v) & 3, mv_y(v) & 3));
#else
    {
        const uint4* np = reinterpret_cast<const uint4*>(w.need);
        const uint4 n0 = np[0], n1 = np[1];
        uint32_t o = n0.x | n0.y | n0.z | n0.w | n1.x | n1.y | n1.z | n1.w;
        o |= o >> 16; o |= o >> 8;
        wn = (int)(o & 15u);
    }
#endif
    const int b = lane >> 1, half = lane & 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
    const int shape = grp_shape(m.meta);
    const uint32_t* wp;
    int pitch;
    // the window starts sx & 3 bytes into its first staged word (load staging) / sx & 15 bytes into its box (tensor copy)
    const int sx = SH ? grp_window_x(g, m, lane, v) : 0;
    const int sh = (sx & 3) * 8;
    if (SH && tbuf) { wp = tbuf + (by * 4 + half * 2) * TMA0_PITCH + (((PAD_X + sx) & 15) >> 2) + bx; pitch = TMA0_PITCH; }
    else if (shape == 0) { wp = w.win + (by * 4 + half * 2) * WIN0_PITCH + bx; pitch = WIN0_PITCH; }
    else if (shape == 1) { wp = w.win + q * WIN1_STRIDE + ((by & 1) * 4 + half * 2) * WIN1_PITCH + (bx & 1); pitch = WIN1_PITCH; }
    else { wp = w.win + b * 30 + half * 2 * 3; pitch = 3; }
    const McSel sel = mc_select(mv_x(v) & 3, mv_y(v) & 3);
    uint32_t pred[2];
    if (wn == 0) mc_patch<COMPAT, 0, SH>(wp, pitch, sh, sel, pred, wn);
    else if (!(wn & NEED_J)) mc_patch<COMPAT, 1, SH>(wp, pitch, sh, sel, pred, wn);
    else mc_patch<COMPAT, 2, SH>(wp, pitch, sh, sel, pred, wn);
    const int row = by * 4 + half * 2, col = bx * 4;
    uint8_t* yp = pd.y + (m.yoff + row * g.pitch_y + col);
    uint32_t exc = 0;
    uint32_t chroma = 0;
    if (m.active) {          // warp-uniform
        const int slot0 = grp_slot0(m.meta);
        // luma: block b (raster) = luma4x4BlkIdx blk_raster(b); its slot holds rows 0..3, this lane takes rows 2 half, 2 half + 1
        const int bi = blk_raster(b);
        if ((m.active >> bi) & 1u) {
            const uint4 r = w.slots[(slot0 + popc32(m.active & ((1u << bi) - 1u))) * 2 + half];
            exc += add_residual_words<COMPAT>(pred[0], r.x, r.y, pred[0]);
            exc += add_residual_words<COMPAT>(pred[1], r.z, r.w, pred[1]);
        }
        // chroma: lane -> plane (lane >> 4), row cy, 4 samples at cx0: chroma block (cy >> 2) * 2 + (cx0 >> 2), its row cy & 3
        const int plane = lane >> 4, l = lane & 15, cy = l >> 1, ci = 16 + plane * 4 + (cy >> 2) * 2 + (l & 1);
        if ((m.active >> ci) & 1u) {
            const uint2* rows = reinterpret_cast<const uint2*>(&w.slots[(slot0 + popc32(m.active & ((1u << ci) - 1u))) * 2]);
            const uint2 r = rows[cy & 3];
            chroma = prmt(r.x, r.y, 0x6420);
        }
    }
    *reinterpret_cast<uint32_t*>(yp) = pred[0];
    *reinterpret_cast<uint32_t*>(yp + g.pitch_y) = pred[1];
    {
        const int plane = lane >> 4, l = lane & 15, cy = l >> 1, cx0 = (l & 1) * 4;
        uint8_t* cbase = (plane ? pd.v : pd.u) + (m.coff + cy * g.pitch_c + cx0);
        *reinterpret_cast<uint32_t*>(cbase) = chroma;
    }
    report_excursions(pd, a, exc);
}

// K1+K3 for the inter macroblocks among a0 .. a0 + GRP_MBS - 1.  `before_samples` runs once, after the residual of the
// first part and before the first reference sample is read (the kernel awaits its predecessor there).
// TMA: the windows of 16x16 macroblocks arrive as tensor copies, two deep -- while macroblock i is interpolated the box of the
// next macroblock of the part (if it is a 16x16 one) is already on its way into the other buffer -- and every other window as
// aligned words; the consumer shifts.  (On the host, where there are no tensor copies, TMA selects the aligned-word staging for
// every shape: the emulator covers the shifted consumer that way.)  par: phase parities of the two barriers, carried across
// groups by the kernel; tw: the second buffer and the barriers.
template <bool COMPAT, bool TMA, class Exec, class Before>
HD void seq_inter_group(Exec& ex, GroupWork& w, const PicDesc& pd, const Geometry& g, int a0, int n_mbs, Before before_samples, uint32_t* par = nullptr,
                        GroupTma* tw = nullptr)
{
    ex([&](int lane) { phase_grp_load(w, pd, g, a0, n_mbs, lane); });
    int first = 0;
    while (first < GRP_MBS) {
        PerLane<int> te, ni;
        ex([&](int lane) { phase_grp_items(w, first, lane, te(lane), ni(lane)); });
        const int take_end = te(0), n_items = ni(0);      // macroblocks [first, take_end) of the group are in this part, with n_items active blocks
        for (int base = 0; base < n_items; base += 32) ex([&](int lane) { phase_grp_residual<COMPAT>(w, pd, base + lane, n_items); });
        if (first == 0) before_samples();
        // tensor staging state: `have` = the box of the macroblock at hand was posted during the previous one, into buffer hb
        bool have = false;
        int hb = 1;
        for (int i = first; i < take_end; i++) {
            const GrpMb& m = w.mb[i];
            if (!(m.meta & GRP_INTER)) continue;       // warp-uniform: intra (the intra kernels reconstruct it afterwards) or beyond the picture
            const int a = a0 + i;
            PerLane<uint32_t> mv_own;
            const uint32_t* tbuf = nullptr;
#if defined(__CUDA_ARCH__)
            if (TMA) {
                const bool t0 = grp_shape(m.meta) == 0;                  // this macroblock's window is a tensor copy
                const int cb = t0 ? (have ? hb : 1) : -1;                // ... in this buffer
                int j = i + 1;                                           // the next inter macroblock of the part
                while (j < take_end && !(w.mb[j].meta & GRP_INTER)) j++;
                const bool tj = j < take_end && grp_shape(w.mb[j].meta) == 0;
                // its box goes to the buffer this macroblock leaves alone: the other tensor buffer, or win_b when this one is
                // load-staged (into w.win, whose head is tensor buffer 0)
                const int nb = t0 ? cb ^ 1 : 1;
                ex([&](int lane) {
                    mv_own(lane) = grp_load_mv(pd, m, lane);
                    if (!t0) phase_grp_stage<COMPAT, TMA>(w, pd, g, m, lane, mv_own(lane));
                    if (lane == 0) {
                        if (t0 && !have) grp_tma_post(tw->win_b, &tw->mbar[1], pd, g, m, mv_own(lane));
                        if (tj) grp_tma_post(nb ? tw->win_b : w.win, &tw->mbar[nb], pd, g, w.mb[j], grp_load_mv(pd, w.mb[j], 0));
                    }
                });
                if (t0) {
                    mbar_wait(&tw->mbar[cb], par[cb]);
                    par[cb] ^= 1u;
                    tbuf = cb ? tw->win_b : w.win;
                }
                have = tj;
                hb = nb;
            } else
#endif
            ex([&](int lane) {
                mv_own(lane) = grp_load_mv(pd, m, lane);
                phase_grp_stage<COMPAT, TMA>(w, pd, g, m, lane, mv_own(lane));
            });
            ex([&](int lane) { phase_grp_compute<COMPAT, TMA>(w, pd, g, m, a, lane, mv_own(lane), tbuf); });
            if (COMPAT && (m.meta & GRP_BORDER))
                ex([&](int lane) { pad_mb_luma_cold(pd.y, g.pitch_y, g.w_mbs, g.h_mbs, (int)(m.xy & 0xffffu), (int)(m.xy >> 16), lane); });
        }
        first = take_end;
    }
}

}  // namespace h264r
