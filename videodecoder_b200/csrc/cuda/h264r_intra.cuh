// h264r_intra.cuh -- K1+K2 for Intra4x4 / Intra16x16 macroblocks, latency version.
//
// An I picture is a wavefront whose critical path is (W + 2H) macroblocks long (h264r_kernels.cuh), so what
// matters here is the time ONE warp needs for ONE macroblock.  Same results as the generic phases of
// h264r_core.cuh (phase_intra_edges / phase_residual / phase_intra16 / phase_intra4 / phase_store, which stay
// as the plain statement of the arithmetic and as the path for Intra8x8), restructured:
//   * residual with the register transform of h264r_mc.cuh (plus the Intra16x16 DC Hadamard);
//   * neighbouring samples arrive as 32-bit words (top row) and, when the warp has just finished the
//     macroblock to the left, from its own tile (no trip to L2 for the left column);
//   * Intra4x4: every directional mode is (w0 e[i] + w1 e[i+1] + w2 e[i+2] + round) >> shift over the 13
//     neighbouring samples e = L K J I M A..H; (i, w, shift) per (mode, sample) comes from a 144-entry
//     table, and the sixteen blocks run as a 10-step anti-diagonal wavefront INSIDE the macroblock, two
//     blocks per step on the two half-warps (block (bx,by) needs steps d-1, d-2, d-3 only, d = bx + 2by);
//   * Intra16x16: sums and plane gradients with dp4a on the word-aligned top row, 8 packed samples per
//     lane, stored straight to the frame.
// Every function is a per-lane phase, so the host emulator of the CPU test-suite runs this code too.
#pragma once
#include "h264r_mc.cuh"

namespace h264r {

// ---- K1 for any 4x4-transform macroblock ---------------------------------------------------------------
// Lane b < 24 owns block b (luma4x4BlkIdx order, then Cb, Cr) and writes its residual into res[384]
// (Y 16x16, U 8x8, V 8x8).  Inter / Intra4x4: luma blocks whose coded_block_pattern bit is clear are zero
// (the reference zero-fills them, h264/residual.cpp:44-46).  Intra16x16: DC levels go through the 4x4
// Hadamard (reference residual::Decode_Intra16x16, h264/residual.cpp:286-319) and bypass the scaling.
// nz_flags (optional, 16 bytes): lane b < 16 records whether luma block b has a non-zero level, at its raster index.
template <bool COMPAT, bool LEAN, bool INTER>
HD void phase_residual_fast(int16_t* res, const PicDesc& pd, const h264r_mb_hdr& h, int a, int lane, uint8_t* nz_flags, bool chroma_only,
                            const MbPre* pre, bool luma_only)
{
    if (lane >= 24 || (chroma_only && lane < 16) || (luma_only && lane >= 16)) return;
    const int fmt = LEAN ? H264R_LEVELS_COMPACT : pd.level_fmt;       // LEAN: see k_inter
    const BlkRef br = pre ? BlkRef{pre->blk_mask, pre->blk_off} : load_blkref(pd, a);
    const int cbp_l = h.cbp & 15, cbp_c = (h.cbp >> 4) & 3;
    const bool luma = lane < 16, i16 = INTER ? false : h.mb_class == H264R_MB_I16x16;      // INTER: the caller only has inter macroblocks
    bool active = luma ? (i16 || ((cbp_l >> (lane >> 2)) & 1) != 0) : cbp_c != 0;
    int r[16];
    bool zero = true;
    // where the block's residual goes: 4 rows of 4 int16 (8 bytes each), private to this lane during the phase
    int16_t* dst;
    int pitch;
    if (luma) { int ras = blk_raster(lane); dst = res + (ras >> 2) * 64 + (ras & 3) * 4; pitch = 16; }
    else { int plane = (lane - 16) >> 2, b = (lane - 16) & 3; dst = res + 256 + plane * 64 + (b >> 1) * 32 + (b & 1) * 4; pitch = 8; }
    if (active) {
        const uint8_t* src = blk_ptr(pd, br, lane, fmt);
        int4 v0 = int4{0, 0, 0, 0}, v1 = int4{0, 0, 0, 0};
        if (src && blk_compact(fmt, br)) {
            // compact block: the levels are scattered into the lane's own (zeroed) 4x4 field of res[], which has the
            // raster shape of a block, and read back as rows
            const uint2 v = H264R_LDG(reinterpret_cast<const uint2*>(src));
            uint32_t map = v.x & 0xffffu;
            unsigned long long lv = ((unsigned long long)v.y << 16) | (v.x >> 16);
#pragma unroll
            for (int i = 0; i < 4; i++) *reinterpret_cast<uint2*>(dst + i * pitch) = uint2{0u, 0u};
            while (map) {
#if defined(__CUDA_ARCH__)
                const int q = __ffs((int)map) - 1;
#else
                const int q = __builtin_ctz(map);
#endif
                map &= map - 1;
                dst[(q >> 2) * pitch + (q & 3)] = (int16_t)(int8_t)(lv & 255u);
                lv >>= 8;
            }
            const uint2 r0 = *reinterpret_cast<const uint2*>(dst), r1 = *reinterpret_cast<const uint2*>(dst + pitch);
            const uint2 r2 = *reinterpret_cast<const uint2*>(dst + 2 * pitch), r3 = *reinterpret_cast<const uint2*>(dst + 3 * pitch);
            v0 = int4{(int)r0.x, (int)r0.y, (int)r1.x, (int)r1.y};
            if (luma || cbp_c == 2) v1 = int4{(int)r2.x, (int)r2.y, (int)r3.x, (int)r3.y};
        } else if (src) { v0 = blk_levels8(pd, br, src, 0, fmt); if (luma || cbp_c == 2) v1 = blk_levels8(pd, br, src, 8, fmt); }
        if (!luma && cbp_c != 2) { v0.x &= 0xffff; v0.y = v0.z = v0.w = 0; }        // DC only: AC levels are not coded
        int c[16];
        const int w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) { c[2 * k] = (int)(int16_t)(w[k] & 0xffff); c[2 * k + 1] = w[k] >> 16; }
        int qp = h.qp_y;
        bool bypass = false;
        if (luma && i16) {
            // element (i, j) of T * DC * T for the block at raster (i, j); DC(k, l) sits in slot 0 of the block at raster (k, l)
            const int ras = blk_raster(lane), i = ras >> 2, j = ras & 3;
            int f = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int t = 0;
#pragma unroll
                for (int l = 0; l < 4; l++) {
                    const uint8_t* dp = blk_ptr(pd, br, blk_raster(k * 4 + l), fmt);
                    int dc = dp ? blk_level0(pd, br, dp, fmt) : 0;
                    t += ((0xA6C0 >> (l * 4 + j)) & 1) ? -dc : dc;       // T rows {++++},{++--},{+--+},{+-+-}: bit = minus
                }
                f += ((0xA6C0 >> (i * 4 + k)) & 1) ? -t : t;
            }
            int s = (qp * 43) >> 8, m = qp - 6 * s, ls = 16 * norm_adjust_dc(m);
            c[0] = (s >= 6) ? (f * ls) * (1 << (s - 6)) : ((f * ls + (1 << (5 - s))) >> (6 - s));
            bypass = true;
        }
        if (!luma) {
            int plane = (lane - 16) >> 2, b = (lane - 16) & 3;
            qp = plane ? h.qp_cr : h.qp_cb;
            const uint8_t* d0 = blk_ptr(pd, br, 16 + plane * 4, fmt), * d1 = blk_ptr(pd, br, 17 + plane * 4, fmt);
            const uint8_t* d2 = blk_ptr(pd, br, 18 + plane * 4, fmt), * d3 = blk_ptr(pd, br, 19 + plane * 4, fmt);
            int c0 = d0 ? blk_level0(pd, br, d0, fmt) : 0, c1 = d1 ? blk_level0(pd, br, d1, fmt) : 0, c2 = d2 ? blk_level0(pd, br, d2, fmt) : 0, c3 = d3 ? blk_level0(pd, br, d3, fmt) : 0;
            int f;
            if (COMPAT) f = (b == 0 || b == 3) ? c3 : -c3;      // reference's chroma DC product (h264/matrix.cpp:150-171)
            else f = c0 + ((b & 1) ? -c1 : c1) + ((b & 2) ? -c2 : c2) + ((b == 1 || b == 2) ? -c3 : c3);
            int s = (qp * 43) >> 8, m = qp - 6 * s;
            c[0] = ((f * 16 * norm_adjust_dc(m)) * (1 << s)) >> 5;
            bypass = true;
        }
        int nz = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) nz |= c[k];
        if (nz) {
            Dequant q = make_dequant(qp);
            idct4_fast(c, q, bypass, r);
            zero = false;
        }
    }
    if (nz_flags && luma) nz_flags[blk_raster(lane)] = zero ? 0 : 1;
    // store the block
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t lo = 0, hi = 0;
        if (!zero) {
            lo = pack_sat_s16x2(r[i * 4], r[i * 4 + 1]);
            hi = pack_sat_s16x2(r[i * 4 + 2], r[i * 4 + 3]);
        }
        *reinterpret_cast<uint2*>(dst + i * pitch) = uint2{lo, hi};
    }
}

// ---- neighbouring samples ----------------------------------------------------------------------------------
#define TL(x, y) w.tile[H264R_TI(x, y)]
#define CTP(p, x, y) w.ctile[p][H264R_CTI(x, y)]

// Tile borders of the macroblock: row y = -1 for x = -1..23 as seven words (lanes 0..6), column x = -1 for
// y = 0..15 (lanes 8..23), chroma borders when `chroma` (top: lanes 24..29, left: lanes 8..23).  Unavailable
// neighbours read as 0.  `carry`: the warp's tiles still hold the macroblock to the left, whose last column
// is the left neighbour.  Other neighbours were written by other warps of the same kernel: ld.global.cg.
// left_col: the right column of the macroblock to the left as its warp left it in MbWork::colx (this warp's own or the partner
// warp's of a row pair), or null: from the frame.
// parts: 1 = the row above (B, C, D), 2 = the left column, 3 = both.
HD void phase_intra_edges_fast(MbWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane, bool chroma, const uint8_t* left_col,
                               int parts = 3)
{
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    const bool A = h.flags & H264R_AVAIL_A, B = h.flags & H264R_AVAIL_B, C = h.flags & H264R_AVAIL_C, D = h.flags & H264R_AVAIL_D;
    if (lane < 7) {
        if (!(parts & 1)) return;
        const int x = lane * 4 - 4;
        const bool ok = lane == 0 ? D : (lane <= 4 ? B : C);
        uint32_t v = 0;
        if (ok) v = H264R_LDCG(reinterpret_cast<const uint32_t*>(pd.y + (ptrdiff_t)(mby * 16 - 1) * g.pitch_y + mbx * 16 + x));
        if (lane == 0) v &= 0xff000000u;
        *reinterpret_cast<uint32_t*>(&TL(x, -1)) = v;
    } else if (lane >= 8 && lane < 24) {
        if (!(parts & 2)) return;
        const int yy = lane - 8;
        uint8_t v = 0;
        if (A) v = left_col ? left_col[yy] : H264R_LDCG(pd.y + (ptrdiff_t)(mby * 16 + yy) * g.pitch_y + mbx * 16 - 1);
        TL(-1, yy) = v;
        if (chroma) {
            const int p = yy >> 3, r = yy & 7;
            const uint8_t* c = p ? pd.v : pd.u;
            uint8_t cv = 0;
            if (A) cv = left_col ? left_col[16 + yy] : H264R_LDCG(c + (ptrdiff_t)(mby * 8 + r) * g.pitch_c + mbx * 8 - 1);
            CTP(p, -1, r) = cv;
        }
    } else if (chroma && lane >= 24 && lane < 30) {
        if (!(parts & 1)) return;
        const int p = (lane - 24) / 3, k = (lane - 24) - p * 3, x = k * 4 - 4;
        const uint8_t* c = p ? pd.v : pd.u;
        const bool ok = k == 0 ? D : B;
        uint32_t v = 0;
        if (ok) v = H264R_LDCG(reinterpret_cast<const uint32_t*>(c + (ptrdiff_t)(mby * 8 - 1) * g.pitch_c + mbx * 8 + x));
        if (k == 0) v &= 0xff000000u;
        *reinterpret_cast<uint32_t*>(&CTP(p, x, -1)) = v;
    }
}

// ---- Intra16x16 --------------------------------------------------------------------------------------------
// Lane -> row (lane >> 1), columns 8*(lane & 1)..+7.  Prediction (reference macroblock::Prediction_Intra16x16_
// {V,H,DC,Plane}, h264/macroblock.cpp:1373-1441) + residual; writes the tile and the frame.
template <bool COMPAT>
HD void phase_intra16_fast(MbWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane)
{
    const bool A = h.flags & H264R_AVAIL_A, B = h.flags & H264R_AVAIL_B;
    const int mode = h.intra_modes & 3;
    const int row = lane >> 1, c0 = (lane & 1) * 8;
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    uint32_t p0, p1;
    if (mode == 0) {
        p0 = *reinterpret_cast<const uint32_t*>(&TL(c0, -1));
        p1 = *reinterpret_cast<const uint32_t*>(&TL(c0 + 4, -1));
    } else if (mode == 1) {
        p0 = p1 = (uint32_t)TL(-1, row) * 0x01010101u;
    } else if (mode == 2) {
        int st = 0, sl = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) st = dp4a_us(*reinterpret_cast<const uint32_t*>(&TL(4 * k, -1)), 0x01010101u, st);
#pragma unroll
        for (int k = 0; k < 16; k++) sl += TL(-1, k);
        int dc;
        if (A && B) dc = (sl + st + 16) >> 5;
        else if (A) dc = (sl + 8) >> 4;
        else if (B) {
            if (COMPAT) {
                // D5: top-only DC sums column 15 of the MB above (h264/macroblock.cpp:1410-1411)
                int sc = 0;
                for (int k = 0; k < 16; k++) sc += H264R_LDCG(pd.y + (ptrdiff_t)(mby * 16 - 16 + k) * g.pitch_y + mbx * 16 + 15);
                dc = (sc + 8) >> 4;
            } else dc = (st + 8) >> 4;
        } else dc = 128;
        p0 = p1 = (uint32_t)dc * 0x01010101u;
    } else {
        // plane: H = sum i*(top[7+i] - top[7-i]), V likewise on the left column, i = 1..8 (index -1 = corner)
        const uint32_t wm4 = *reinterpret_cast<const uint32_t*>(&TL(-4, -1)), w0 = *reinterpret_cast<const uint32_t*>(&TL(0, -1)),
                       w4 = *reinterpret_cast<const uint32_t*>(&TL(4, -1)), w8 = *reinterpret_cast<const uint32_t*>(&TL(8, -1)),
                       w12 = *reinterpret_cast<const uint32_t*>(&TL(12, -1));
        int Hs = dp4a_us(wm4, taps4(0, 0, 0, -8), 0);
        Hs = dp4a_us(w0, taps4(-7, -6, -5, -4), Hs);
        Hs = dp4a_us(w4, taps4(-3, -2, -1, 0), Hs);
        Hs = dp4a_us(w8, taps4(1, 2, 3, 4), Hs);
        Hs = dp4a_us(w12, taps4(5, 6, 7, 8), Hs);
        int Vs = 0;
#pragma unroll
        for (int i = 1; i <= 8; i++) Vs += i * ((int)TL(-1, 7 + i) - (int)TL(-1, 7 - i));
        const int pa = 16 * ((int)TL(15, -1) + (int)TL(-1, 15)), pb = (5 * Hs + 32) >> 6, pc = (5 * Vs + 32) >> 6;
        int v = pa + pb * (c0 - 7) + pc * (row - 7) + 16;
        int s[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { s[k] = v >> 5; v += pb; }
        p0 = pack_sat_u8(s[0], s[1], s[2], s[3]);
        p1 = pack_sat_u8(s[4], s[5], s[6], s[7]);
    }
    uint32_t exc = add_residual<COMPAT>(p0, &w.res[row * 16 + c0], p0);
    exc += add_residual<COMPAT>(p1, &w.res[row * 16 + c0 + 4], p1);
    *reinterpret_cast<uint2*>(&TL(c0, row)) = uint2{p0, p1};
    if (lane & 1) w.colx[row] = (uint8_t)(p1 >> 24);
    *reinterpret_cast<uint2*>(pd.y + (size_t)(mby * 16 + row) * g.pitch_y + mbx * 16 + c0) = uint2{p0, p1};
    report_excursions(pd, a, exc);
}

// ---- Intra4x4 ----------------------------------------------------------------------------------------------
// Table entry for (mode, sample x, y): bits 0-3 i0, 4-5 w0, 6-7 w1, 8-9 w2, 10-11 shift;
// pred = (w0 e[i0] + w1 e[i0+1] + w2 e[i0+2] + (1 << shift >> 1)) >> shift over e = L K J I M A B C D E F G H
// (reference Prediction_Intra4x4_* , h264/macroblock.cpp:1037-1365 = ITU-T H.264 8.3.1.2.1-8.3.1.2.9).  Mode 2 (DC)
// is not table driven.
HD uint16_t intra4_entry(int mode, int x, int y)
{
    int i0 = 0, w0 = 1, w1 = 0, w2 = 0, sh = 0;
    switch (mode) {
    case 0: i0 = 5 + x; break;
    case 1: i0 = 3 - y; break;
    case 3:
        if (x == 3 && y == 3) { i0 = 11; w1 = 3; sh = 2; }
        else { i0 = 5 + x + y; w1 = 2; w2 = 1; sh = 2; }
        break;
    case 4: i0 = 3 - y + x; w1 = 2; w2 = 1; sh = 2; break;
    case 5: {
        int z = 2 * x - y, i = x - (y >> 1);
        if (z >= 0 && !(z & 1)) { i0 = 4 + i; w1 = 1; sh = 1; }
        else if (z >= 0) { i0 = 3 + i; w1 = 2; w2 = 1; sh = 2; }
        else if (z == -1) { i0 = 3; w1 = 2; w2 = 1; sh = 2; }
        else { i0 = 4 - y; w1 = 2; w2 = 1; sh = 2; }
        break;
    }
    case 6: {
        int z = 2 * y - x, i = y - (x >> 1);
        if (z >= 0 && !(z & 1)) { i0 = 3 - i; w1 = 1; sh = 1; }
        else if (z >= 0) { i0 = 3 - i; w1 = 2; w2 = 1; sh = 2; }
        else if (z == -1) { i0 = 3; w1 = 2; w2 = 1; sh = 2; }
        else { i0 = 2 + x; w1 = 2; w2 = 1; sh = 2; }
        break;
    }
    case 7: {
        int i = x + (y >> 1);
        i0 = 5 + i;
        if (!(y & 1)) { w1 = 1; sh = 1; } else { w1 = 2; w2 = 1; sh = 2; }
        break;
    }
    case 8: {
        int z = x + 2 * y, i = y + (x >> 1);
        if (z > 5) { i0 = 0; }
        else if (z == 5) { i0 = 0; w0 = 3; w1 = 1; sh = 2; }
        else if (!(z & 1)) { i0 = 2 - i; w1 = 1; sh = 1; }
        else { i0 = 1 - i; w1 = 2; w2 = 1; sh = 2; }
        break;
    }
    default: break;
    }
    return (uint16_t)(i0 | (w0 << 4) | (w1 << 6) | (w2 << 8) | (sh << 10));
}
// The table the plan reads: for (mode, up-right samples usable, sample position) the two plan words of a block at the tile's
// origin -- a: the three tile offsets and the shift, b: output offset, weights and kind (see I4Plan).  A block at (X, Y) adds its
// origin to every offset field.  Built once per CTA in shared memory.
constexpr int INTRA4_LUT_SIZE = 9 * 2 * 16;
HD uint2 intra4_plan_entry(int mode, int tr_ok, int pos)
{
    const int x = pos & 3, y = pos >> 2;
    const uint32_t e = intra4_entry(mode, x, y);
    const int i0 = (int)(e & 15u), trmax = tr_ok ? 7 : 3;
    const int left = H264R_TI(-1, 3), top = H264R_TI(0, -1);
    uint32_t o[3];
    for (int t = 0; t < 3; t++) {
        const int i = i0 + t;
        const int k = i - 5 < trmax ? i - 5 : trmax;
        o[t] = (uint32_t)(i <= 4 ? left - i * TILE_P : top + k);
    }
    return uint2{o[0] | (o[1] << 10) | (o[2] << 20) | (((e >> 10) & 3u) << 30), (uint32_t)H264R_TI(x, y) | (((e >> 4) & 63u) << 10) | (1u << 16)};
}
HD void intra4_lut_fill(uint2* lut, int first, int stride)
{
    for (int k = first; k < INTRA4_LUT_SIZE; k += stride) lut[k] = intra4_plan_entry(k >> 5, (k >> 4) & 1, k & 15);
}

// The macroblock's inner wavefront: step d = 0..9 handles the blocks (bx, by) with bx + 2 by = d; half-warp
// `lane >> 4` takes the block with by = by_min + (lane >> 4), lane & 15 = sample.  Everything about a lane's sample
// of step d that does not depend on sample VALUES -- which block, its mode, the table entry, the three tile
// offsets with the up-right substitution applied, the residual -- is worked out for all ten steps at once before
// the first step (ten independent chains), so that a step itself is three byte loads, three multiply-adds, a
// shift, the residual add and one store.  Availability and the up-right substitution rule follow
// macroblock::Calc / get_4x4Neighbour (h264/macroblock.cpp:28-33, 1092-1110).
//   a[d] = o0 | o1 << 10 | o2 << 20 | shift << 30        (tile byte offsets; DC: o0 = top word, o1 = first left sample)
//   b[d] = out | w0 << 10 | w1 << 12 | w2 << 14 | kind << 16   (kind 0 idle, 1 directional, 2..5 DC with both / left /
//          top / no neighbours) | column-15 flag << 19 | row << 20
//   r[d >> 1] = residual of steps d (low half) and d + 1 (high half)
struct I4Plan { uint32_t a[10], b[10], r[5]; };

HD void intra4_plan(const MbWork& w, const h264r_mb_hdr& h, const uint2* lut, int lane, I4Plan& pl)
{
    const int pos = lane & 15, x = pos & 3, y = pos >> 2;
    const bool avA = (h.flags & H264R_AVAIL_A) != 0, avB = (h.flags & H264R_AVAIL_B) != 0, avC = (h.flags & H264R_AVAIL_C) != 0;
#pragma unroll
    for (int d = 0; d < 10; d++) {
        const int by_min = d > 3 ? (d - 2) >> 1 : 0, by_max = (d >> 1) < 3 ? (d >> 1) : 3;
        const int by = by_min + (lane >> 4);
        uint32_t A = 0, B = 0;
        int res = 0;
        if (by <= by_max) {
            const int bx = d - 2 * by, ras = by * 4 + bx, blk = blk_raster(ras);
            const int mode = hdr_pred_mode(h, blk);
            const int X = bx * 4, Y = by * 4;
            const uint32_t org = (uint32_t)(Y * TILE_P + X);
            res = w.res[(Y + y) * 16 + X + x];
            if (mode == 2) {
                const bool left_ok = bx > 0 || avA, top_ok = by > 0 || avB;
                A = (uint32_t)H264R_TI(X, Y - 1) | ((uint32_t)H264R_TI(X - 1, Y) << 10);
                B = (uint32_t)H264R_TI(X + x, Y + y) | ((uint32_t)(left_ok ? (top_ok ? 2 : 3) : (top_ok ? 4 : 5)) << 16);
            } else {
                bool tr_ok;
                if (blk == 3 || blk == 11) tr_ok = false;
                else if (bx == 3) tr_ok = (by == 0) && avC && avB;
                else if (by > 0) tr_ok = true;
                else tr_ok = avB;
                const uint2 e = lut[(mode * 2 + (tr_ok ? 1 : 0)) * 16 + pos];
                A = e.x + org * 0x00100401u;          // the origin into each of the three 10-bit offsets
                B = e.y + org;
            }
            if (X + x == 15) B |= (1u << 19) | ((uint32_t)(Y + y) << 20);      // column 15 also goes to MbWork::colx
        }
        pl.a[d] = A; pl.b[d] = B;
        if (d & 1) pl.r[d >> 1] |= (uint32_t)res << 16; else pl.r[d >> 1] = (uint32_t)res & 0xffffu;
    }
}

template <bool COMPAT, int D>
HD void phase_intra4_step(MbWork& w, const PicDesc& pd, int a, const I4Plan& pl, int lane)
{
    (void)lane;
    const uint32_t A = pl.a[D], B = pl.b[D];
    const int kind = (int)((B >> 16) & 7u);
    if (!kind) return;
    const uint8_t* t = w.tile;
    int v;
    if (kind == 1) {
        const int sh = (int)(A >> 30);
        v = ((int)((B >> 10) & 3u) * (int)t[A & 1023u] + (int)((B >> 12) & 3u) * (int)t[(A >> 10) & 1023u] +
             (int)((B >> 14) & 3u) * (int)t[(A >> 20) & 1023u] + ((1 << sh) >> 1)) >> sh;
    } else {
        const uint8_t* l = t + ((A >> 10) & 1023u);
        const int st = dp4a_us(*reinterpret_cast<const uint32_t*>(t + (A & 1023u)), 0x01010101u, 0);
        const int sl = (int)l[0] + (int)l[TILE_P] + (int)l[2 * TILE_P] + (int)l[3 * TILE_P];
        v = kind == 2 ? (st + sl + 4) >> 3 : (kind == 3 ? (sl + 2) >> 2 : (kind == 4 ? (st + 2) >> 2 : 128));
    }
    const int res = (int)(int16_t)(pl.r[D >> 1] >> (16 * (D & 1)));
    uint32_t exc = 0;
    const uint8_t s = recon_luma<COMPAT>(v + res, exc);
    w.tile[B & 1023u] = s;
    if (B & (1u << 19)) w.colx[(B >> 20) & 15u] = s;
    report_excursions(pd, a, exc);
}

// ---- stores --------------------------------------------------------------------------------------------------
// luma rows of the tile: lanes 0..15, one 128-bit row each
HD void store_luma_tile(const MbWork& w, const PicDesc& pd, const Geometry& g, int a, int lane)
{
    if (lane >= 16) return;
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    *reinterpret_cast<uint4*>(pd.y + (size_t)(mby * 16 + lane) * g.pitch_y + mbx * 16) = *reinterpret_cast<const uint4*>(&TL(0, lane));
}
// chroma rows: lanes 16..31, one 64-bit row each.  COMPAT: the reference has no chroma prediction, sample =
// residual (low 8 bits; D4, h264/macroblock.cpp:1477-1480); otherwise from the chroma tiles.
template <bool COMPAT>
HD void store_chroma(const MbWork& w, const PicDesc& pd, const Geometry& g, int a, int lane)
{
    if (lane < 16) return;
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    const int p = (lane - 16) >> 3, row = (lane - 16) & 7;
    uint2 v;
    if (COMPAT) {
        const uint4 r = *reinterpret_cast<const uint4*>(&w.res[256 + p * 64 + row * 8]);
        v = uint2{prmt(r.x, r.y, 0x6420), prmt(r.z, r.w, 0x6420)};
    } else v = *reinterpret_cast<const uint2*>(&CTP(p, 0, row));
    *reinterpret_cast<uint2*>((p ? pd.v : pd.u) + (size_t)(mby * 8 + row) * g.pitch_c + mbx * 8) = v;
}

// the left neighbours of the tiles from a column another warp left in its MbWork::colx (generic sequences: their edge phase reads
// the frame, where that macroblock may not have arrived yet)
HD void patch_left_col(MbWork& w, const uint8_t* left_col, bool chroma, int lane)
{
    if (lane < 16) TL(-1, lane) = left_col[lane];
    else if (chroma) { const int p = (lane - 16) >> 3, r = (lane - 16) & 7; CTP(p, -1, r) = left_col[lane]; }
}
// the right column of the macroblock in the tiles -> MbWork::colx (the generic sequences do not write it as they go)
HD void save_right_col(MbWork& w, int lane)
{
    if (lane < 16) w.colx[lane] = TL(15, lane);
    else { const int p = (lane - 16) >> 3, r = (lane - 16) & 7; w.colx[lane] = CTP(p, 7, r); }
}

#undef TL
#undef CTP

}  // namespace h264r
