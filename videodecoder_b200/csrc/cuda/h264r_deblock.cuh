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
#include "h264r_mc.cuh"

namespace h264r {

// luma tile: sample (x, y), x = -4..15, y = -4..15; chroma tiles: x = -4..7 (only -2.. used), y = -2..7
constexpr int DB_P = 32, DB_X0 = 16, DBC_P = 16, DBC_X0 = 8;
struct alignas(16) DeblockWork {
    uint8_t tile[20 * DB_P];
    uint8_t ctile[2][10 * DBC_P];
    uint8_t bs[32];             // [dir][edge][k]
};
#define DBT(x, y) w.tile[((y) + 4) * DB_P + (x) + DB_X0]
#define DBC(p, x, y) w.ctile[p][((y) + 2) * DBC_P + (x) + DBC_X0]

HD bool nz_bit(uint32_t nz, int ras) { return (nz >> ras) & 1u; }

// Phase D0: boundary strengths (lane = dir*16 + edge*4 + k, ITU-T H.264 8.7.2.1) and the rows above the
// macroblock into the tiles (lanes 0..3 luma rows -4..-1, lanes 4..7 chroma rows -2..-1).
// h / hl / ht: headers of the macroblock, its left and its upper neighbour; nz*: their non-zero masks.
HD void phase_deblock_bs_fast(DeblockWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, const h264r_mb_hdr& hl,
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
    w.bs[lane] = (uint8_t)bs;
    if (mby > 0) {
        if (lane < 4) {
            const int r = lane - 4;
            *reinterpret_cast<uint4*>(&DBT(0, r)) = H264R_LDCG(reinterpret_cast<const uint4*>(pd.y + (ptrdiff_t)(mby * 16 + r) * g.pitch_y + mbx * 16));
        } else if (lane < 8) {
            const int p = (lane - 4) >> 1, r = ((lane - 4) & 1) - 2;
            const uint8_t* c = p ? pd.v : pd.u;
            *reinterpret_cast<uint2*>(&DBC(p, 0, r)) = H264R_LDCG(reinterpret_cast<const uint2*>(c + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8));
        }
    }
}

// alpha, beta, tc0 of one edge (8.7.2.2): qp = average QP of the two sides
struct DbParams { int alpha, beta, tc0; };
HD DbParams db_params(const PicDesc& pd, int qpav, int bs)
{
    const DeblockTables& T = deblock_tables();
    const int ia = clip3(0, 51, qpav + pd.alpha_off), ib = clip3(0, 51, qpav + pd.beta_off);
    return DbParams{T.alpha[ia], T.beta[ib], bs < 4 ? T.tc0[ia][bs > 0 ? bs - 1 : 0] : 0};
}

// the edges of one luma line held in s[0..19] (positions -4..15 along the filtering direction); bs4[e] = strength
// of edge e for this line; qp_n = QP'Y of the neighbouring macroblock across edge 0
HD void deblock_luma_line4(int s[20], const uint8_t bs4[4], const PicDesc& pd, int qp, int qp_n)
{
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int bs = bs4[e];
        if (!bs) continue;
        const DbParams P = db_params(pd, e == 0 ? (qp + qp_n + 1) >> 1 : qp, bs);
        int* p = s + e * 4;         // p[0..3] = p3..p0, p[4..7] = q0..q3
        deblock_luma_line(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], bs, P.alpha, P.beta, P.tc0);
    }
}
// chroma: positions -2..7 in s[0..9], edges at 0 (s[2]) and 4 (s[6])
HD void deblock_chroma_line2(int s[10], int bs0, int bs2, const PicDesc& pd, int qp, int qp_n)
{
    if (bs0) {
        const DbParams P = db_params(pd, (qp + qp_n + 1) >> 1, bs0);
        deblock_chroma_line(s[0], s[1], s[2], s[3], bs0, P.alpha, P.beta, P.tc0);
    }
    if (bs2) {
        const DbParams P = db_params(pd, qp, bs2);
        deblock_chroma_line(s[4], s[5], s[6], s[7], bs2, P.alpha, P.beta, P.tc0);
    }
}
HD void unpack4(uint32_t v, int* d) { d[0] = (int)(v & 255u); d[1] = (int)((v >> 8) & 255u); d[2] = (int)((v >> 16) & 255u); d[3] = (int)(v >> 24); }
HD uint32_t pack4(const int* s) { return (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16) | ((uint32_t)s[3] << 24); }

// Phase D1: vertical edges.  Lanes 0..15: luma row `lane` straight from the frame (x = -4..15), filtered in
// registers, into the tile.  Lanes 16..31: chroma plane (lane-16)>>3, row (lane-16)&7 (x = -2..7).
HD void phase_deblock_vert(DeblockWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, const h264r_mb_hdr& hl, int mbx, int mby,
                           int lane)
{
    if (lane < 16) {
        const uint8_t* row = pd.y + (ptrdiff_t)(mby * 16 + lane) * g.pitch_y + mbx * 16;
        const uint4 v = H264R_LDCG(reinterpret_cast<const uint4*>(row));
        const uint32_t vl = mbx > 0 ? H264R_LDCG(reinterpret_cast<const uint32_t*>(row - 4)) : 0u;
        int s[20];
        unpack4(vl, s); unpack4(v.x, s + 4); unpack4(v.y, s + 8); unpack4(v.z, s + 12); unpack4(v.w, s + 16);
        uint8_t bs4[4];
#pragma unroll
        for (int e = 0; e < 4; e++) bs4[e] = w.bs[e * 4 + (lane >> 2)];
        deblock_luma_line4(s, bs4, pd, h.qp_y, hl.qp_y);
        *reinterpret_cast<uint32_t*>(&DBT(-4, lane)) = pack4(s);
        *reinterpret_cast<uint4*>(&DBT(0, lane)) = uint4{pack4(s + 4), pack4(s + 8), pack4(s + 12), pack4(s + 16)};
    } else {
        const int p = (lane - 16) >> 3, r = (lane - 16) & 7;
        const uint8_t* row = (p ? pd.v : pd.u) + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8;
        const uint2 v = H264R_LDCG(reinterpret_cast<const uint2*>(row));
        const uint32_t vl = mbx > 0 ? H264R_LDCG(reinterpret_cast<const uint32_t*>(row - 4)) : 0u;
        int t[4], s[10];
        unpack4(vl, t); s[0] = t[2]; s[1] = t[3];
        unpack4(v.x, s + 2); unpack4(v.y, s + 6);
        deblock_chroma_line2(s, w.bs[r >> 1], w.bs[8 + (r >> 1)], pd, p ? h.qp_cr : h.qp_cb, p ? hl.qp_cr : hl.qp_cb);
        t[2] = s[0]; t[3] = s[1];
        *reinterpret_cast<uint32_t*>(&DBC(p, -4, r)) = pack4(t);
        *reinterpret_cast<uint2*>(&DBC(p, 0, r)) = uint2{pack4(s + 2), pack4(s + 6)};
    }
}

// Phase D2: horizontal edges.  Lanes 0..15: luma column `lane` (y = -4..15) out of the tile, filtered in
// registers, rows -3..15 back.  Lanes 16..31: chroma column.
HD void phase_deblock_horz(DeblockWork& w, const PicDesc& pd, const h264r_mb_hdr& h, const h264r_mb_hdr& ht, int lane)
{
    if (lane < 16) {
        uint8_t bs4[4];
        int any = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) { bs4[e] = w.bs[16 + e * 4 + (lane >> 2)]; any |= bs4[e]; }
        if (!any) return;
        int s[20];
#pragma unroll
        for (int i = 0; i < 20; i++) s[i] = DBT(lane, i - 4);
        deblock_luma_line4(s, bs4, pd, h.qp_y, ht.qp_y);
#pragma unroll
        for (int i = 1; i < 20; i++) DBT(lane, i - 4) = (uint8_t)s[i];
    } else {
        const int p = (lane - 16) >> 3, c = (lane - 16) & 7;
        const int bs0 = w.bs[16 + (c >> 1)], bs2 = w.bs[16 + 8 + (c >> 1)];
        if (!(bs0 | bs2)) return;
        int s[10];
#pragma unroll
        for (int i = 0; i < 10; i++) s[i] = DBC(p, c, i - 2);
        deblock_chroma_line2(s, bs0, bs2, pd, p ? h.qp_cr : h.qp_cb, p ? ht.qp_cr : ht.qp_cb);
#pragma unroll
        for (int i = 1; i < 10; i++) DBC(p, c, i - 2) = (uint8_t)s[i];
    }
}

// Phase D3: write back what this macroblock may have changed: luma rows 0..15 x columns -4..15 and rows -3..-1 x
// columns 0..15; chroma rows 0..7 x columns -4..7 (of which -1 may have changed) and row -1 x columns 0..7.
HD void phase_deblock_store_fast(const DeblockWork& w, const PicDesc& pd, const Geometry& g, int mbx, int mby, int lane)
{
    for (int t = lane; t < 19 + 18; t += 32) {
        if (t < 19) {
            const int r = t < 16 ? t : t - 19;          // 0..15, then -3..-1
            if (r < 0 && mby == 0) continue;
            uint8_t* row = pd.y + (ptrdiff_t)(mby * 16 + r) * g.pitch_y + mbx * 16;
            *reinterpret_cast<uint4*>(row) = *reinterpret_cast<const uint4*>(&DBT(0, r));
            if (r >= 0 && mbx > 0) *reinterpret_cast<uint32_t*>(row - 4) = *reinterpret_cast<const uint32_t*>(&DBT(-4, r));
        } else {
            const int u = t - 19, p = u / 9, k = u - p * 9, r = k < 8 ? k : -1;
            if (r < 0 && mby == 0) continue;
            uint8_t* row = (p ? pd.v : pd.u) + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8;
            *reinterpret_cast<uint2*>(row) = *reinterpret_cast<const uint2*>(&DBC(p, 0, r));
            if (r >= 0 && mbx > 0) *reinterpret_cast<uint32_t*>(row - 4) = *reinterpret_cast<const uint32_t*>(&DBC(p, -4, r));
        }
    }
}

#undef DBT
#undef DBC

// K4 for one macroblock: its neighbours A, B, C are completely filtered.  hl / ht / nz*: see phase_deblock_bs_fast.
template <class Exec>
HD void seq_deblock_mb_fast(Exec& ex, DeblockWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, const h264r_mb_hdr& hl,
                            const h264r_mb_hdr& ht, uint32_t nz, uint32_t nzl, uint32_t nzt, int a, int mbx, int mby)
{
    ex([&](int lane) { phase_deblock_bs_fast(w, pd, g, h, hl, ht, nz, nzl, nzt, a, mbx, mby, lane); });
    ex([&](int lane) { phase_deblock_vert(w, pd, g, h, hl, mbx, mby, lane); });
    ex([&](int lane) { phase_deblock_horz(w, pd, h, ht, lane); });
    ex([&](int lane) { phase_deblock_store_fast(w, pd, g, mbx, mby, lane); });
}

}  // namespace h264r
