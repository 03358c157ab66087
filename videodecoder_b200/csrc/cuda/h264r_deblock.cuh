// h264r_deblock.cuh -- K4, the in-loop deblocking filter of ITU-T H.264 8.7 (the reference has none: its
// h264/picture.cpp stops at reconstruction; only the slice-header fields are parsed, h264/slice.cpp:247-256),
// latency version for the row wavefront of h264r_kernels.cuh.
//
// Same results as phase_deblock_load / _bs / _dir / _store of h264r_core.cuh (the plain statement of the
// clause, kept as the second implementation the tests compare with), restructured so that ONE warp gets through
// ONE macroblock quickly, because a picture is a wavefront (W + 2H) macroblocks long:
//   * boundary strengths need, per 4x4 block, "has a non-zero level": the inter kernel leaves that as a 16-bit
//     mask per macroblock (PicDesc::nzmask), so nothing here touches coefficients;
//   * edge-ordered per macroblock row with vectorised pixel I/O: a lane owns one line of pixels across the
//     macroblock -- a row of 4 + 16 luma samples for the vertical edges, a column of 4 + 16 for the horizontal
//     ones (8 + 8 chroma lines on the other half-warp) -- loads it as 128-/32-bit words, runs its four (two)
//     edges in registers in edge order, and stores words; the only shared-memory traffic is the transpose
//     between the two directions;
//   * a direction whose 16 strengths are all zero is skipped by the whole warp.
// Per-lane phases only (cross-lane traffic through shared memory): the host emulator runs this code too.
#pragma once
#include "h264r_intra.cuh"

namespace h264r {

// Per-macroblock record written by the flat kernel k_deblock_prep (nothing in it depends on sample values):
//   bytes 0..31                              bs[dir*16 + edge*4 + k]                                   (8.7.2.1)
//   bytes 32 + 8*((plane*2 + dir)*2 + kind)  alpha, beta, tc0 for bS 1, 2, 3 of that plane / direction  (8.7.2.2)
//                                            kind 0 = the macroblock edge (QP averaged with the neighbour), 1 = inner edges
constexpr int DBREC_BYTES = 128;

// luma tile: sample (x, y), x = -4..15, y = -4..15; chroma tiles: x = -4..7 (only -2.. used), y = -2..7
constexpr int DB_P = 32, DB_X0 = 16, DBC_P = 16, DBC_X0 = 8;
struct alignas(16) DeblockWork {
    uint8_t tile[20 * DB_P];
    uint8_t ctile[2][10 * DBC_P];
    uint8_t rec[DBREC_BYTES];
    uint32_t stage[24];          // hand-over payload being assembled (see MBOX below)
};

// ---- hand-over between the rows of the wavefront ------------------------------------------------------------------------------
// What the row below needs of macroblock (x, y) -- luma rows 12..15 (p3..p0 of its top edge) and chroma rows 6, 7 of both planes,
// 96 bytes -- travels through a per-macroblock record in global memory, NOT through the frame: 24 pairs {4 payload bytes, stamp}
// written with 8-byte stores (8-byte accesses are single-copy atomic), stamp = the launch's number.  The consumer polls the
// pairs themselves until every stamp is this launch's: data and flag arrive together, so neither side needs a fence, and no
// L1 invalidation (fence.acq_rel.gpu) hits the other warps of the SM.  Payload bytes 0..63: luma rows 12..15, 16 bytes each;
// 64..79: Cb rows 6, 7; 80..95: Cr rows 6, 7.  Consequence for the frame: rows 13..15 (chroma row 7) of a macroblock are
// WRITTEN by the row below (after its top edge), never by the macroblock's own row -- except in the last row.
constexpr int MBOX_WORDS = 48;
HD uint2 mbox_load(const uint32_t* p)
{
#if defined(__CUDA_ARCH__)
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
#else
    return uint2{p[0], p[1]};
#endif
}
HD void mbox_store(uint32_t* p, uint32_t payload, uint32_t stamp)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(payload), "r"(stamp) : "memory");
#else
    p[0] = payload; p[1] = stamp;
#endif
}
#define DBT(x, y) w.tile[((y) + 4) * DB_P + (x) + DB_X0]
#define DBC(p, x, y) w.ctile[p][((y) + 2) * DBC_P + (x) + DBC_X0]

HD bool nz_bit(uint32_t nz, int ras) { return (nz >> ras) & 1u; }

// K4 preparation, one warp per macroblock: boundary strengths (lane = dir*16 + edge*4 + k, ITU-T H.264 8.7.2.1) and the
// filter parameters (lanes 0..11) into the macroblock's record.  h / hl / ht: headers of the macroblock, its left and
// its upper neighbour; nz*: their non-zero masks (PicDesc::nzmask).
HD void phase_deblock_prep(uint8_t* rec, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, const h264r_mb_hdr& hl,
                           const h264r_mb_hdr& ht, uint32_t nz, uint32_t nzl, uint32_t nzt, int a, int mbx, int mby, int lane)
{
    const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
    const bool t8 = (h.flags & H264R_T8x8) != 0;
    const bool mb_edge = e == 0;
    int bs = 0;
    bool active;
    int rasq, rasp, an = a;
    h264r_mb_hdr hn = h;          // by value: selecting between headers through pointers would put them in local memory
    uint32_t nzn = nz;
    if (mb_edge) {
        active = dir == 0 ? mbx > 0 : mby > 0;
        an = dir == 0 ? a - 1 : a - g.w_mbs;
        hn = dir == 0 ? hl : ht;
        nzn = dir == 0 ? nzl : nzt;
        rasq = dir == 0 ? k * 4 : k;
        rasp = dir == 0 ? k * 4 + 3 : 12 + k;
    } else {
        active = !(t8 && (e & 1));
        rasq = dir == 0 ? k * 4 + e : e * 4 + k;
        rasp = dir == 0 ? rasq - 1 : rasq - 4;
    }
    if (active) {
        if (is_intra(h.mb_class) || is_intra(hn.mb_class)) bs = mb_edge ? 4 : 3;
        else if (nz_bit(nz, rasq) || nz_bit(nzn, rasp)) bs = 2;
        else {
            const int qq = (rasq >> 3) * 2 + ((rasq & 3) >> 1), qp_ = (rasp >> 3) * 2 + ((rasp & 3) >> 1);
            const int refq = pd.ref_id[hdr_ref_idx(h, qq)], refp = pd.ref_id[hdr_ref_idx(hn, qp_)];
            if (refq != refp) bs = 1;
            else {
                const uint32_t mq = load_mv(pd, h, a, rasq), mp = load_mv(pd, hn, an, rasp);
                if (iabs(mv_x(mq) - mv_x(mp)) >= 4 || iabs(mv_y(mq) - mv_y(mp)) >= 4) bs = 1;
            }
        }
    }
    rec[lane] = (uint8_t)bs;
    if (lane < 12) {
        const int pl = lane >> 2, d = (lane >> 1) & 1, kind = lane & 1;
        const h264r_mb_hdr hq = d == 0 ? hl : ht;
        const int qp_q = pl == 0 ? h.qp_y : (pl == 1 ? h.qp_cb : h.qp_cr), qp_n = pl == 0 ? hq.qp_y : (pl == 1 ? hq.qp_cb : hq.qp_cr);
        const int qpav = kind == 0 ? (qp_q + qp_n + 1) >> 1 : qp_q;
        const DeblockTables& T = deblock_tables();
        const int ia = clip3(0, 51, qpav + pd.alpha_off), ib = clip3(0, 51, qpav + pd.beta_off);
        uint8_t* o = rec + 32 + 8 * lane;
        o[0] = T.alpha[ia]; o[1] = T.beta[ib]; o[2] = T.tc0[ia][0]; o[3] = T.tc0[ia][1]; o[4] = T.tc0[ia][2]; o[5] = 0; o[6] = 0; o[7] = 0;
    }
}

// One line across one edge, luma or chroma (chromaEdgeFlag): ITU-T H.264 8.7.2.3 (bS < 4) and 8.7.2.4 (bS = 4).
// p[0..7] = p3 p2 p1 p0 q0 q1 q2 q3.  Same arithmetic as deblock_luma_line / deblock_chroma_line of h264r_core.cuh.
HD void deblock_line(int* p, int bs, int alpha, int beta, int tc0, bool chroma)
{
    // Branch-free: both filters are evaluated and the results selected, so that the two sides and the two filter kinds form
    // independent instruction streams (one warp works on a macroblock alone: latency, not issue slots, is what counts) and
    // lanes with different strengths do not serialise.  bs = 0: nothing changes.
    const int P3 = p[0], P2 = p[1], P1 = p[2], P0 = p[3], Q0 = p[4], Q1 = p[5], Q2 = p[6], Q3 = p[7];
    const int dpq = iabs(P0 - Q0);
    const bool on = (bs != 0) & (dpq < alpha) & (iabs(P1 - P0) < beta) & (iabs(Q1 - Q0) < beta);
    const bool apb = !chroma & (iabs(P2 - P0) < beta), aqb = !chroma & (iabs(Q2 - Q0) < beta);
    // 8.7.2.3 (bS < 4)
    const int tc = tc0 + (chroma ? 1 : (int)apb + (int)aqb);
    const int delta = clip3(-tc, tc, (((Q0 - P0) * 4) + (P1 - Q1) + 4) >> 3);
    const int avg = (P0 + Q0 + 1) >> 1;
    const int n_p0 = clip1(P0 + delta), n_q0 = clip1(Q0 - delta);
    const int n_p1 = P1 + clip3(-tc0, tc0, (P2 + avg - (P1 * 2)) >> 1), n_q1 = Q1 + clip3(-tc0, tc0, (Q2 + avg - (Q1 * 2)) >> 1);
    // 8.7.2.4 (bS = 4)
    const bool small = dpq < ((alpha >> 2) + 2);
    const bool sp = apb & small, sq = aqb & small;
    const int s_p0 = sp ? (P2 + 2 * P1 + 2 * P0 + 2 * Q0 + Q1 + 4) >> 3 : (2 * P1 + P0 + Q1 + 2) >> 2;
    const int s_q0 = sq ? (P1 + 2 * P0 + 2 * Q0 + 2 * Q1 + Q2 + 4) >> 3 : (2 * Q1 + Q0 + P1 + 2) >> 2;
    const int s_p1 = (P2 + P1 + P0 + Q0 + 2) >> 2, s_q1 = (P0 + Q0 + Q1 + Q2 + 2) >> 2;
    const int s_p2 = (2 * P3 + 3 * P2 + P1 + P0 + Q0 + 4) >> 3, s_q2 = (2 * Q3 + 3 * Q2 + Q1 + Q0 + P0 + 4) >> 3;
    const bool strong = bs == 4;
    p[3] = on ? (strong ? s_p0 : n_p0) : P0;
    p[4] = on ? (strong ? s_q0 : n_q0) : Q0;
    p[2] = (on & (strong ? sp : apb)) ? (strong ? s_p1 : n_p1) : P1;
    p[5] = (on & (strong ? sq : aqb)) ? (strong ? s_q1 : n_q1) : Q1;
    p[1] = (on & strong & sp) ? s_p2 : P2;
    p[6] = (on & strong & sq) ? s_q2 : Q2;
}
// The edges of one line in edge order.  s[0..19] = positions -4..15 along the filtering direction; a chroma line sits
// in the same array with its two edges where luma edges 0 and 2 are (positions -2..1 at s[2..5], 2..5 at s[10..13]), so
// that both kinds of lines run the same instruction stream.  k = which of the four strength columns the line uses.
HD void deblock_line_edges(int s[20], uint4 bsd, uint4 par, int k, bool chroma)
{
    // bsd: the 16 strengths of the direction (byte e*4 + k = edge e, column k); par: {macroblock edge, inner edges} x {alpha | beta << 8
    // | tc0(bS 1) << 16 | tc0(bS 2) << 24, tc0(bS 3)} of the lane's plane and direction
    const uint32_t bw[4] = {bsd.x, bsd.y, bsd.z, bsd.w};
    int bs[4];
#pragma unroll
    for (int e = 0; e < 4; e++) bs[e] = (chroma && (e & 1)) ? 0 : (int)((bw[e] >> (8 * k)) & 255u);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (!bw[e]) continue;            // no line of the macroblock crosses this edge with a non-zero strength: warp-uniform
        const uint32_t px = e == 0 ? par.x : par.z, py = e == 0 ? par.y : par.w;
        const int alpha = (int)(px & 255u), beta = (int)((px >> 8) & 255u);
        const int tc0 = bs[e] == 1 ? (int)((px >> 16) & 255u) : (bs[e] == 2 ? (int)(px >> 24) : (bs[e] == 3 ? (int)(py & 255u) : 0));
        deblock_line(s + e * 4, bs[e], alpha, beta, tc0, chroma);
    }
}
HD void unpack4(uint32_t v, int* d) { d[0] = (int)(v & 255u); d[1] = (int)((v >> 8) & 255u); d[2] = (int)((v >> 16) & 255u); d[3] = (int)(v >> 24); }
HD uint32_t pack4(const int* s) { return (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24); }

// Phase D2: horizontal edges.  Lanes 0..15: luma column `lane` (y = -4..15) out of the tile, filtered in registers,
// rows -3..15 back.  Lanes 16..31: chroma column (y = -2..5; p0 / q0 of its two edges back).
HD void phase_deblock_horz(DeblockWork& w, uint4 bsh, uint4 par_h, int lane)
{
    const bool luma = lane < 16;
    const int p = luma ? 0 : (lane - 16) >> 3, c = luma ? lane : (lane - 16) & 7, k = luma ? c >> 2 : c >> 1;
    if (!(((bsh.x | bsh.y | bsh.z | bsh.w) >> (8 * k)) & 255u)) return;
    int s[20];
    if (luma) {
#pragma unroll
        for (int i = 0; i < 20; i++) s[i] = DBT(c, i - 4);
    } else {
#pragma unroll
        for (int i = 0; i < 20; i++) s[i] = 0;
        s[2] = DBC(p, c, -2); s[3] = DBC(p, c, -1); s[4] = DBC(p, c, 0); s[5] = DBC(p, c, 1);
        s[10] = DBC(p, c, 2); s[11] = DBC(p, c, 3); s[12] = DBC(p, c, 4); s[13] = DBC(p, c, 5);
    }
    deblock_line_edges(s, bsh, par_h, k, !luma);
    if (luma) {
#pragma unroll
        for (int i = 1; i < 20; i++) DBT(c, i - 4) = (uint8_t)s[i];
    } else {
        DBC(p, c, -1) = (uint8_t)s[3]; DBC(p, c, 0) = (uint8_t)s[4]; DBC(p, c, 3) = (uint8_t)s[11]; DBC(p, c, 4) = (uint8_t)s[12];
    }
}

#undef DBT
#undef DBC

// K4 for one macroblock of a row of the wavefront.  What a lane needs of the macroblock that does not depend on the neighbours'
// filtering is fetched by the caller one macroblock ahead, straight into registers: the strengths of both directions (the same 32
// bytes for every lane: one transaction), the filter parameters of the lane's plane, and the lane's own line of (unfiltered) samples.
struct DeblockIn { uint4 bsv, bsh, par_v, par_h, own_y; uint2 own_c; };
HD DeblockIn deblock_fetch(const PicDesc& pd, const Geometry& g, const uint8_t* dbrec, int a, int mbx, int mby, int lane)
{
    DeblockIn in;
    const uint4* rec = reinterpret_cast<const uint4*>(dbrec + (size_t)a * DBREC_BYTES);
    const int plane = lane < 16 ? 0 : 1 + ((lane - 16) >> 3);
    in.bsv = H264R_LDG(rec);
    in.bsh = H264R_LDG(rec + 1);
    in.par_v = H264R_LDG(rec + 2 + plane * 2);          // bytes 32 + 8 * ((plane * 2 + dir) * 2): 16 bytes per (plane, dir)
    in.par_h = H264R_LDG(rec + 2 + plane * 2 + 1);
    in.own_y = uint4{0, 0, 0, 0}; in.own_c = uint2{0, 0};
    if (lane < 16) in.own_y = H264R_LDG(reinterpret_cast<const uint4*>(pd.y + (ptrdiff_t)(mby * 16 + lane) * g.pitch_y + mbx * 16));
    else {
        const int p = (lane - 16) >> 3, r = (lane - 16) & 7;
        in.own_c = H264R_LDG(reinterpret_cast<const uint2*>((p ? pd.v : pd.u) + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8));
    }
    return in;
}

#define DBT(x, y) w.tile[((y) + 4) * DB_P + (x) + DB_X0]
#define DBC(p, x, y) w.ctile[p][((y) + 2) * DBC_P + (x) + DBC_X0]
// Phase D1: vertical edges into the tiles.  Lanes 0..15: luma row `lane`, lanes 16..31: chroma plane (lane-16)>>3, row (lane-16)&7.
// The four samples to the left are the last columns of the macroblock the warp filtered just before (carry), which the lane
// itself wrote: no hand-over inside the warp.  The rows above arrive through the hand-over record (phase_mbox_top).
HD void phase_deblock_vert2(DeblockWork& w, const DeblockIn& in, bool carry, bool any, int lane)
{
    const bool luma = lane < 16;
    const int p = luma ? 0 : (lane - 16) >> 3, r = luma ? lane : (lane - 16) & 7;
    int s[20], t[4], c[8];
    if (luma) {
        unpack4(carry ? *reinterpret_cast<const uint32_t*>(&DBT(12, r)) : 0u, s);
        unpack4(in.own_y.x, s + 4); unpack4(in.own_y.y, s + 8); unpack4(in.own_y.z, s + 12); unpack4(in.own_y.w, s + 16);
    } else {
        unpack4(carry ? *reinterpret_cast<const uint32_t*>(&DBC(p, 4, r)) : 0u, t);
        unpack4(in.own_c.x, c); unpack4(in.own_c.y, c + 4);
#pragma unroll
        for (int i = 0; i < 20; i++) s[i] = 0;
        s[2] = t[2]; s[3] = t[3]; s[4] = c[0]; s[5] = c[1]; s[10] = c[2]; s[11] = c[3]; s[12] = c[4]; s[13] = c[5];
    }
    if (any) deblock_line_edges(s, in.bsv, in.par_v, luma ? r >> 2 : r >> 1, !luma);
    if (luma) {
        *reinterpret_cast<uint32_t*>(&DBT(-4, r)) = pack4(s);
        *reinterpret_cast<uint4*>(&DBT(0, r)) = uint4{pack4(s + 4), pack4(s + 8), pack4(s + 12), pack4(s + 16)};
    } else {
        t[2] = s[2]; t[3] = s[3]; c[0] = s[4]; c[1] = s[5]; c[2] = s[10]; c[3] = s[11]; c[4] = s[12]; c[5] = s[13];
        *reinterpret_cast<uint32_t*>(&DBC(p, -4, r)) = pack4(t);
        *reinterpret_cast<uint2*>(&DBC(p, 0, r)) = uint2{pack4(c), pack4(c + 4)};
    }
}
// The hand-over payload of the macroblock in the tiles (its own columns): luma rows 12..15, chroma rows 6, 7 -> w.stage.
HD void phase_mbox_stage(DeblockWork& w, int lane)
{
    if (lane < 4) *reinterpret_cast<uint4*>(&w.stage[lane * 4]) = *reinterpret_cast<const uint4*>(&DBT(0, 12 + lane));
    else if (lane < 8) { const int p = (lane - 4) >> 1, r = 6 + ((lane - 4) & 1); *reinterpret_cast<uint2*>(&w.stage[16 + (lane - 4) * 2]) = *reinterpret_cast<const uint2*>(&DBC(p, 0, r)); }
}
// Publishes the record of the macroblock to the LEFT (mbox_mb), final now that this macroblock's left edge is through: the payload
// staged at the end of its own turn, except its last four columns, which are the carry columns of the tiles (luma columns 12..15
// = tile columns -4..-1; chroma 4..7 likewise).
HD void phase_mbox_publish_left(const DeblockWork& w, uint32_t* mbox_mb, uint32_t stamp, int lane)
{
    if (lane >= 24) return;
    uint32_t v = w.stage[lane];
    if (lane < 16) { if ((lane & 3) == 3) v = *reinterpret_cast<const uint32_t*>(&DBT(-4, 12 + (lane >> 2))); }
    else if (lane & 1) { const int j = (lane - 16) >> 1; v = *reinterpret_cast<const uint32_t*>(&DBC(j >> 1, -4, 6 + (j & 1))); }
    mbox_store(mbox_mb + 2 * lane, v, stamp);
}
HD void phase_mbox_publish(const DeblockWork& w, uint32_t* mbox_mb, uint32_t stamp, int lane)
{
    if (lane < 24) mbox_store(mbox_mb + 2 * lane, w.stage[lane], stamp);
}
// The rows above (payload word `lane` < 24 in `top`) into the tiles: luma rows -4..-1, chroma rows -2, -1.
HD void phase_mbox_unpack(DeblockWork& w, uint32_t top, int lane)
{
    if (lane < 16) *reinterpret_cast<uint32_t*>(&DBT((lane & 3) * 4, (lane >> 2) - 4)) = top;
    else if (lane < 24) { const int k = lane - 16, p = k >> 2, r = ((k >> 1) & 1) - 2; *reinterpret_cast<uint32_t*>(&DBC(p, (k & 1) * 4, r)) = top; }
}
// Phase D3 under the hand-over rule: this row writes its own rows 0..12 (chroma 0..6) -- all of them in the picture's last row
// -- with the four columns to the left, and ALWAYS the rows -3..-1 (chroma -1) of the macroblock above, which that row left to it.
// own: the macroblock itself changed (some strength non-zero); else only the rows above are written.  Lanes 0..15: luma row
// `lane`, 16..31: chroma rows; then lanes 0..2 / 16, 17 the rows above.
HD void phase_deblock_store2(const DeblockWork& w, const PicDesc& pd, const Geometry& g, int mbx, int mby, bool own, bool last_row, int lane)
{
    if (lane < 16) {
        uint8_t* row = pd.y + (ptrdiff_t)(mby * 16 + lane) * g.pitch_y + mbx * 16;
        if (own && (lane <= 12 || last_row)) {
            *reinterpret_cast<uint4*>(row) = *reinterpret_cast<const uint4*>(&DBT(0, lane));
            if (mbx > 0) *reinterpret_cast<uint32_t*>(row - 4) = *reinterpret_cast<const uint32_t*>(&DBT(-4, lane));
        }
        if (mby > 0 && lane < 3) {
            const int r = lane - 3;
            *reinterpret_cast<uint4*>(row - (ptrdiff_t)3 * g.pitch_y) = *reinterpret_cast<const uint4*>(&DBT(0, r));      // row (lane - 3): three rows up from row `lane`
        }
    } else {
        const int p = (lane - 16) >> 3, r = (lane - 16) & 7;
        uint8_t* row = (p ? pd.v : pd.u) + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8;
        if (own && (r <= 6 || last_row)) {
            *reinterpret_cast<uint2*>(row) = *reinterpret_cast<const uint2*>(&DBC(p, 0, r));
            if (mbx > 0) *reinterpret_cast<uint32_t*>(row - 4) = *reinterpret_cast<const uint32_t*>(&DBC(p, -4, r));
        }
        if (mby > 0 && r == 0) *reinterpret_cast<uint2*>(row - g.pitch_c) = *reinterpret_cast<const uint2*>(&DBC(p, 0, -1));
    }
}
#undef DBT
#undef DBC

// The sequence for macroblock (mbx, mby): four phases.  top_poll(lane): the caller's way of obtaining payload word `lane` of the
// macroblock above once its stamp is this launch's (the kernel polls global memory; the emulator, which runs rows in order, reads).
template <class Exec, class Poll>
HD void seq_deblock_mb_fast(Exec& ex, DeblockWork& w, const PicDesc& pd, const Geometry& g, PerLane<DeblockIn>& in, bool carry, int mbx, int mby,
                            uint32_t stamp, Poll&& top_poll)
{
    const int a = mby * g.w_mbs + mbx;
    const bool last_row = mby == g.h_mbs - 1, last_col = mbx == g.w_mbs - 1;
    // every lane holds the same strengths: warp-uniform
    const DeblockIn& i0 = in(0);
    const bool any = (i0.bsv.x | i0.bsv.y | i0.bsv.z | i0.bsv.w | i0.bsh.x | i0.bsh.y | i0.bsh.z | i0.bsh.w) != 0;
    ex([&](int lane) { phase_deblock_vert2(w, in(lane), carry, any, lane); });
    // the macroblock to the left is final now: its record goes out BEFORE this one waits for the row above
    if ((carry && !last_row) || mby > 0)
        ex([&](int lane) {
            if (carry && !last_row) phase_mbox_publish_left(w, pd.mbox + (size_t)(a - 1) * MBOX_WORDS, stamp, lane);
            if (mby > 0) phase_mbox_unpack(w, top_poll(lane), lane);
        });
    if (any) ex([&](int lane) { phase_deblock_horz(w, in(lane).bsh, in(lane).par_h, lane); });
    ex([&](int lane) {
        phase_deblock_store2(w, pd, g, mbx, mby, any, last_row, lane);
        if (!last_row) phase_mbox_stage(w, lane);
    });
    if (last_col && !last_row) ex([&](int lane) { phase_mbox_publish(w, pd.mbox + (size_t)a * MBOX_WORDS, stamp, lane); });
}

}  // namespace h264r
