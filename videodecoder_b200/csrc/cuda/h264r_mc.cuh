// h264r_mc.cuh -- fast path of K1+K3 (inter macroblocks).
//
// Same results as phase_residual / phase_inter_luma / phase_inter_chroma of h264r_core.cuh (which
// stay as the straightforward statement of the arithmetic and as the fallback for the 8x8
// transform), restructured for instruction count:
//   * residual: lanes 0..23 own one 4x4 block each and keep it in registers; blocks whose
//     coded_block_pattern bit is clear are neither loaded nor transformed (the reference zero-fills
//     them: h264/residual.cpp:44-46), a macroblock with cbp == 0 skips the phase altogether;
//   * motion compensation: the reference window of every prediction block is staged ONCE in shared
//     memory with the coordinate clamp applied while staging (reference h264/macroblock.cpp:828-830),
//     then every lane produces 8 samples (a row segment) with dp4a: each 6-tap sum is two dp4a on
//     byte groups obtained with funnel shifts, and every quantity of the 16 quarter-sample positions,
//     including the reference's own tap pattern for j and s (D6/D7), is a per-row dp4a accumulation;
//   * prediction block size is chosen per macroblock from the 16 per-4x4 vectors: 16x16 when they
//     are all equal, 8x8 when every quadrant is uniform, 4x4 otherwise.
// Every function is a per-lane phase (no shuffles), so the host emulator runs this code too.
#pragma once
#include "h264r_core.cuh"

namespace h264r {

// ---- intrinsics with host emulation ----------------------------------------------------------------
HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int shift_bits)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift_bits);
#else
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (shift_bits & 31));
#endif
}
// sum over the four bytes: (unsigned byte of a) * (signed byte of b) + c
HD int dp4a_us(uint32_t a, uint32_t b, int c)
{
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 255u) * (int)(int8_t)((b >> (8 * i)) & 255u);
    return c;
#endif
}
HD constexpr uint32_t taps4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(c & 255) << 16) | ((uint32_t)(d & 255) << 24);
}

// Working set of the fast inter path (one per warp)
struct alignas(16) InterWork {
    uint8_t win[2560];          // staged reference windows (see window geometry below)
    int16_t res[384];           // residual
    uint32_t excursions;
    int32_t bs;                 // prediction block size chosen for the macroblock (16 / 8 / 4)
    uint32_t pad_[2];
};

// ---- K1: residual in registers -----------------------------------------------------------------------
struct Dequant {                // per QP: d = (c * mul + add) >> shr for the three position classes
    int mul[3], add, shr;
};
HD Dequant make_dequant(int qp)
{
    const int v[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
    int s = (qp * 43) >> 8, m = qp - 6 * s;         // qp / 6, qp % 6 for 0 <= qp < 64
    Dequant q;
    if (s >= 4) { for (int k = 0; k < 3; k++) q.mul[k] = (16 * v[m][k]) << (s - 4); q.add = 0; q.shr = 0; }
    else { for (int k = 0; k < 3; k++) q.mul[k] = 16 * v[m][k]; q.add = 1 << (3 - s); q.shr = 4 - s; }
    return q;
}
// c: 16 levels (raster); returns residual in r.  Same arithmetic as scale_idct4.
HD void idct4_fast(const int c[16], const Dequant& q, bool dc_bypass, int r[16])
{
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int cls = ((k >> 2) & 1) == 0 && (k & 1) == 0 ? 0 : ((((k >> 2) & 1) && (k & 1)) ? 1 : 2);
        d[k] = (c[k] * q.mul[cls] + q.add) >> q.shr;
    }
    if (dc_bypass) d[0] = c[0];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int e0 = d[i * 4] + d[i * 4 + 2], e1 = d[i * 4] - d[i * 4 + 2];
        int e2 = (d[i * 4 + 1] >> 1) - d[i * 4 + 3], e3 = d[i * 4 + 1] + (d[i * 4 + 3] >> 1);
        d[i * 4] = e0 + e3; d[i * 4 + 1] = e1 + e2; d[i * 4 + 2] = e1 - e2; d[i * 4 + 3] = e0 - e3;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int g0 = d[j] + d[8 + j], g1 = d[j] - d[8 + j];
        int g2 = (d[4 + j] >> 1) - d[12 + j], g3 = d[4 + j] + (d[12 + j] >> 1);
        r[j] = (g0 + g3 + 32) >> 6; r[4 + j] = (g1 + g2 + 32) >> 6;
        r[8 + j] = (g1 - g2 + 32) >> 6; r[12 + j] = (g0 - g3 + 32) >> 6;
    }
}

// Phase: residual of an inter macroblock (4x4 transform) straight from global memory into res[].
// Lane b < 24 owns block b.  Uncoded blocks write zeros.
template <bool COMPAT>
HD void phase_residual_inter_fast(int16_t* res, const PicDesc& pd, const h264r_mb_hdr& h, int a, int lane)
{
    if (lane >= 24) return;
    const int16_t* cbase = pd.coeff + (size_t)a * H264R_COEFF_PER_MB;
    const int cbp_l = h.cbp & 15, cbp_c = (h.cbp >> 4) & 3;
    bool luma = lane < 16;
    bool active = luma ? ((cbp_l >> (lane >> 2)) & 1) != 0 : cbp_c != 0;
    int r[16];
    bool zero = true;
    if (active) {
        const int4* src = reinterpret_cast<const int4*>(cbase) + lane * 2;
        int4 v0 = H264R_LDG(src), v1 = (luma || cbp_c == 2) ? H264R_LDG(src + 1) : int4{0, 0, 0, 0};
        if (!luma && cbp_c != 2) { v0.x &= 0xffff; v0.y = v0.z = v0.w = 0; }        // DC only: AC levels are not coded
        int c[16];
        const int w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; k++) { c[2 * k] = (int)(int16_t)(w[k] & 0xffff); c[2 * k + 1] = w[k] >> 16; }
        int qp = h.qp_y;
        bool bypass = false;
        if (!luma) {
            int plane = (lane - 16) >> 2, b = (lane - 16) & 3;
            qp = plane ? h.qp_cr : h.qp_cb;
            const int16_t* dcp = cbase + (16 + plane * 4) * 16;
            int c0 = H264R_LDG(dcp), c1 = H264R_LDG(dcp + 16), c2 = H264R_LDG(dcp + 32), c3 = H264R_LDG(dcp + 48);
            int f;
            if (COMPAT) f = (b == 0 || b == 3) ? c3 : -c3;      // reference's chroma DC product (h264/matrix.cpp:150-171)
            else f = c0 + ((b & 1) ? -c1 : c1) + ((b & 2) ? -c2 : c2) + ((b == 1 || b == 2) ? -c3 : c3);
            const int v0t[6] = {10, 11, 13, 14, 16, 18};
            int s = (qp * 43) >> 8, m = qp - 6 * s;
            c[0] = ((f * 16 * v0t[m]) * (1 << s)) >> 5;
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
    // store the block: 4 rows of 4 int16 (8 bytes each)
    int16_t* dst;
    int pitch;
    if (luma) { int ras = blk_raster(lane); dst = res + (ras >> 2) * 64 + (ras & 3) * 4; pitch = 16; }
    else { int plane = (lane - 16) >> 2, b = (lane - 16) & 3; dst = res + 256 + plane * 64 + (b >> 1) * 32 + (b & 1) * 4; pitch = 8; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t lo = 0, hi = 0;
        if (!zero) {
            lo = (uint32_t)(sat16(r[i * 4]) & 0xffff) | ((uint32_t)sat16(r[i * 4 + 1]) << 16);
            hi = (uint32_t)(sat16(r[i * 4 + 2]) & 0xffff) | ((uint32_t)sat16(r[i * 4 + 3]) << 16);
        }
        *reinterpret_cast<uint2*>(dst + i * pitch) = uint2{lo, hi};
    }
}

// ---- K3: window staging --------------------------------------------------------------------------------
// Window geometry (bytes): sample (px, py) of the prediction block sits at win[(py + 3) * pitch + px + 4];
// rows -3..bh+2, columns -4..: 16x16 block: 22 rows x 24 bytes (pitch 32); 8x8: 14 x 16 (pitch 16);
// 4x4: 10 x 12 (pitch 16).
HD uint32_t ref_word(const uint8_t* ref, int gpitch, int W, int H, int x, int y)
{
    // 4 samples (x..x+3, y) with the coordinate clamp; word loads when the span is inside the picture
    if (y >= 0 && y < H && x >= 0 && x + 7 < W) {
        const uint8_t* p = ref + (size_t)y * gpitch + x;
        uintptr_t ad = reinterpret_cast<uintptr_t>(p);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
        int sh = (int)(ad & 3) * 8;
        uint32_t lo = H264R_LDG(q);
        if (sh == 0) return lo;
        uint32_t hi = H264R_LDG(q + 1);
        return funnel_r(lo, hi, sh);
    }
    int yy = clip3(0, H - 1, y);
    const uint8_t* row = ref + (size_t)yy * gpitch;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= (uint32_t)H264R_LDG(row + clip3(0, W - 1, x + i)) << (8 * i);
    return v;
}
// items first, first + STRIDE, ... of a window of NROWS x NWORDS words
template <int NROWS, int NWORDS, int PITCH, int STRIDE>
HD void stage_window(uint8_t* dst, const uint8_t* ref, int gpitch, int W, int H, int sx, int sy, int first)
{
#pragma unroll 1
    for (int it = first; it < NROWS * NWORDS; it += STRIDE) {
        int row = it / NWORDS, j = it - row * NWORDS;
        *reinterpret_cast<uint32_t*>(dst + row * PITCH + j * 4) = ref_word(ref, gpitch, W, H, sx + 4 * j, sy + row);
    }
}

// ---- K3: one row segment of N samples ---------------------------------------------------------------------
// Table-driven: for every source row dr = -3..3 a position needs, up to four raw quantities are
// accumulated with dp4a on the byte groups A = x[-2..1], B = x[2..5] of each sample:
//   QH  6-tap of row 0 (b1)                QS  the "s1" sum (row +1; the reference takes tap K from row -1, D6)
//   QV  vertical 6-tap (column 0 or +1)    QJ  the part of j1 that is not 20*QH + 20*QS
// so that in BOTH modes j1 = QJ + 20 QH + 20 QS.  Spec: QJ = H(-2) - 5 H(-1) - 5 H(2) + H(3).  Reference
// (D7, h264/macroblock.cpp:843-846; with L = x[-2]-5x[-1], M = 20x[0]+20x[1], R = -5x[2]+x[3]):
//   j1 = aa - 5bb + 20b1 + 20s1 - 5gg + hh,  aa = L(-2)+M(-2)+R(-2), bb = L(-2)+M(-1)+R(-2),
//   gg = L(2)+M(2)+R(-2), hh = L(3)+M(3)+R(-3)
//   => QJ = R(-3) + [-4L + M - 9R](-2) + [-5M](-1) + [-5L - 5M](2) + [L + M](3).
struct McRowTaps { uint32_t sa, sb, ja, jb; int jm, vt; };
template <bool COMPAT>
HD McRowTaps mc_row_taps(int dr)
{
    constexpr uint32_t TA = taps4(1, -5, 20, 20), TB = taps4(-5, 1, 0, 0);
    McRowTaps t{0, 0, 0, 0, 0, 0};
    t.vt = dr == -2 || dr == 3 ? 1 : (dr == -1 || dr == 2 ? -5 : (dr == 0 || dr == 1 ? 20 : 0));
    if (COMPAT) {
        if (dr == -1) t.sa = taps4(1, 0, 0, 0);
        if (dr == 1) { t.sa = taps4(0, -5, 20, 20); t.sb = TB; }
        t.jm = 1;
        if (dr == -3) t.jb = TB;
        if (dr == -2) { t.ja = taps4(-4, 20, 20, 20); t.jb = taps4(45, -9, 0, 0); }
        if (dr == -1) t.ja = taps4(0, 0, -100, -100);
        if (dr == 2) t.ja = taps4(-5, 25, -100, -100);
        if (dr == 3) t.ja = TA;
    } else {
        if (dr == 1) { t.sa = TA; t.sb = TB; }
        if (dr == -2 || dr == -1 || dr == 2 || dr == 3) { t.ja = TA; t.jb = TB; t.jm = t.vt; }
    }
    return t;
}
// bit 0: QH, bit 1: QS, bit 2: QV (column 0), bit 3: QV (column +1), bit 4: QJ
HD int mc_needs(int sel)
{
    // index = xFrac * 4 + yFrac:   G   d   h   n   a   e   i   p   b   f   j   q   c   g   k   r
    const uint8_t need[16] = {0, 4, 4, 4, 1, 5, 0x17, 6, 1, 0x13, 0x13, 0x13, 1, 9, 0x1b, 10};
    return need[sel];
}

template <bool COMPAT, int N>
HD void mc_row(const uint8_t* win, int pitch, int r, int c0, int xf, int yf, int out[N])
{
    constexpr int NW = N == 8 ? 4 : 3;
    constexpr uint32_t TA = taps4(1, -5, 20, 20), TB = taps4(-5, 1, 0, 0);
    const int sel = xf * 4 + yf;
    const int need = mc_needs(sel);
    const bool nH = need & 1, nS = (need & 2) != 0, nV = (need & 12) != 0, nJ = (need & 16) != 0;
    const int vshift = (need & 8) ? 24 : 16;
    uint32_t w[NW];
    int G[N], Hn[N], Mn[N], QH[N], QS[N], QV[N], QJ[N];
#pragma unroll
    for (int k = 0; k < N; k++) { G[k] = Hn[k] = Mn[k] = QH[k] = QS[k] = QV[k] = QJ[k] = 0; }
    const int dr_lo = nJ ? (COMPAT ? -3 : -2) : (nV ? -2 : ((COMPAT && nS) ? -1 : 0));
    const int dr_hi = (nJ || nV) ? 3 : ((nS || sel == 3) ? 1 : 0);
#pragma unroll 1
    for (int dr = dr_lo; dr <= dr_hi; dr++) {
        const McRowTaps t = mc_row_taps<COMPAT>(dr);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(win + (r + dr) * pitch + c0);
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = p[i];
        const bool doH = nH && dr == 0, doS = nS && (t.sa | t.sb) != 0, doJ = nJ && (t.ja | t.jb) != 0, doV = nV && t.vt != 0;
        const uint32_t tv = (uint32_t)(t.vt & 255) << vshift;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const int sA = 2 + k, sB = 6 + k;
            uint32_t A = (sA & 3) == 0 ? w[sA >> 2] : funnel_r(w[sA >> 2], w[(sA >> 2) + 1], (sA & 3) * 8);
            uint32_t B = (sB & 3) == 0 ? w[sB >> 2] : funnel_r(w[sB >> 2], w[(sB >> 2) + 1 < NW ? (sB >> 2) + 1 : (sB >> 2)], (sB & 3) * 8);
            if (dr == 0) { G[k] = (int)((A >> 16) & 255u); Hn[k] = (int)(A >> 24); }
            if (dr == 1) Mn[k] = (int)((A >> 16) & 255u);
            if (doH) QH[k] = dp4a_us(B, TB, dp4a_us(A, TA, 0));
            if (doS) QS[k] = dp4a_us(B, t.sb, dp4a_us(A, t.sa, QS[k]));
            if (doJ) QJ[k] += t.jm * dp4a_us(B, t.jb, dp4a_us(A, t.ja, 0));
            if (doV) QV[k] = dp4a_us(A, tv, QV[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
        int b = clip1((QH[k] + 16) >> 5), v = clip1((QV[k] + 16) >> 5), s = clip1((QS[k] + 16) >> 5);
        int j = clip1((QJ[k] + 20 * QH[k] + 20 * QS[k] + 512) >> 10);
        int o;
        switch (sel) {
        case 0: o = G[k]; break;
        case 1: o = (G[k] + v + 1) >> 1; break;
        case 2: o = v; break;
        case 3: o = (Mn[k] + v + 1) >> 1; break;
        case 4: o = (G[k] + b + 1) >> 1; break;
        case 5: o = (b + v + 1) >> 1; break;
        case 6: o = (v + j + 1) >> 1; break;
        case 7: o = (v + s + 1) >> 1; break;
        case 8: o = b; break;
        case 9: o = (b + j + 1) >> 1; break;
        case 10: o = j; break;
        case 11: o = (j + s + 1) >> 1; break;
        case 12: o = (Hn[k] + b + 1) >> 1; break;
        case 13: o = (b + v + 1) >> 1; break;
        case 14: o = (j + v + 1) >> 1; break;
        default: o = (v + s + 1) >> 1; break;
        }
        out[k] = o;
    }
}

// prediction block size of an inter macroblock: 16, 8 or 4 (see file header); mv[16] = packed (x | y << 16)
template <bool COMPAT>
HD int inter_block_size(const h264r_mb_hdr& h, const uint32_t mv[16])
{
    bool d8 = false;
    if (COMPAT && h.mb_class == H264R_MB_P8x8) { uint32_t s = h.sub_mb_type; d8 = ((s & (s >> 1)) & 0x55u) != 0; }
    if (d8) return 4;
    bool q_uniform = true, all = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int b0 = (q >> 1) * 8 + (q & 1) * 2;
        q_uniform = q_uniform && mv[b0] == mv[b0 + 1] && mv[b0] == mv[b0 + 4] && mv[b0] == mv[b0 + 5];
        all = all && mv[b0] == mv[0];
    }
    if (!q_uniform) return 4;
    bool ref_same = h.u.ref_idx[0] == h.u.ref_idx[1] && h.u.ref_idx[0] == h.u.ref_idx[2] && h.u.ref_idx[0] == h.u.ref_idx[3];
    return (all && ref_same) ? 16 : 8;
}
HD void load_mvs(const PicDesc& pd, int a, uint32_t mv[16])
{
    const uint4* p = reinterpret_cast<const uint4*>(pd.mv + (size_t)a * H264R_MV_PER_MB);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 v = H264R_LDG(p + i);
        mv[4 * i] = v.x; mv[4 * i + 1] = v.y; mv[4 * i + 2] = v.z; mv[4 * i + 3] = v.w;
    }
}
HD int mv_x(uint32_t v) { return (int)(int16_t)(v & 0xffff); }
HD int mv_y(uint32_t v) { return (int)v >> 16; }

HD uint32_t load_mv1(const PicDesc& pd, int a, int blk)
{
    return H264R_LDG(reinterpret_cast<const uint32_t*>(pd.mv + (size_t)a * H264R_MV_PER_MB) + blk);
}

// Phase M0: stage the windows of the macroblock.
template <bool COMPAT>
HD void phase_inter_stage(InterWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane)
{
    int bs;
    {
        uint32_t mv[16];
        load_mvs(pd, a, mv);
        bs = inter_block_size<COMPAT>(h, mv);
    }
    if (lane == 0) w.bs = bs;
    int mbx = a % g.w_mbs, mby = a / g.w_mbs;
    const int W = g.w_mbs * 16, H = g.h_mbs * 16;
    if (bs == 16) {
        uint32_t v = load_mv1(pd, a, 0);
        const uint8_t* ref = pd.ref_y[h.u.ref_idx[0] & 15];
        stage_window<22, 6, 32, 32>(w.win, ref, g.pitch_y, W, H, mbx * 16 + (mv_x(v) >> 2) - 4, mby * 16 + (mv_y(v) >> 2) - 3, lane);
    } else if (bs == 8) {
        int q = lane >> 3, b0 = (q >> 1) * 8 + (q & 1) * 2;
        uint32_t v = load_mv1(pd, a, b0);
        const uint8_t* ref = pd.ref_y[h.u.ref_idx[q] & 15];
        stage_window<14, 4, 16, 8>(w.win + q * (14 * 16), ref, g.pitch_y, W, H, mbx * 16 + (q & 1) * 8 + (mv_x(v) >> 2) - 4,
                                   mby * 16 + (q >> 1) * 8 + (mv_y(v) >> 2) - 3, lane & 7);
    } else {
        int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        uint32_t v = load_mv1(pd, a, b);
        const uint8_t* ref = pd.ref_y[h.u.ref_idx[q] & 15];
        int ys = mby * 16 + by * 4;
        // reference's source-row offset for 4x4 sub-partitions 2 and 3 (D8, h264/macroblock.cpp:786-789)
        if (COMPAT && h.mb_class == H264R_MB_P8x8 && ((h.sub_mb_type >> (2 * q)) & 3) == H264R_SUB_4x4 && (by & 1)) ys -= 3;
        stage_window<10, 3, 16, 2>(w.win + b * 160, ref, g.pitch_y, W, H, mbx * 16 + bx * 4 + (mv_x(v) >> 2) - 4, ys + (mv_y(v) >> 2) - 3, lane & 1);
    }
}

// luma: prediction (+ weighting) + residual -> packed bytes; `bad` collects the sign of out-of-range values
template <bool COMPAT, int N>
HD void finish_luma(int pred[N], const int16_t* res_row, bool coded, const PicDesc& pd, int ri, int& bad, uint32_t packed[N / 4])
{
    if (pd.weighted_pred) {
#pragma unroll
        for (int k = 0; k < N; k++) pred[k] = weight_pred(pred[k], pd.luma_log2_wd, pd.luma_weight[ri], pd.luma_offset[ri]);
    }
    if (coded) {
#pragma unroll
        for (int k = 0; k < N; k += 2) {
            uint32_t rr = *reinterpret_cast<const uint32_t*>(res_row + k);
            pred[k] += (int)(int16_t)(rr & 0xffff);
            pred[k + 1] += (int)rr >> 16;
        }
    }
#pragma unroll
    for (int i = 0; i < N / 4; i++) packed[i] = 0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        int v = pred[k];
        bad |= v | (255 - v);                       // sign bit set when v < 0 or v > 255
        packed[k >> 2] |= (uint32_t)(COMPAT ? (v & 255) : clip1(v)) << ((k & 3) * 8);
    }
}
template <int N>
HD uint32_t count_excursions(const int pred[N])
{
    uint32_t n = 0;
    for (int k = 0; k < N; k++) n += (pred[k] < 0 || pred[k] > 255);
    return n;
}

// Phase M1: compute and store the macroblock (luma from the staged windows; chroma: COMPAT = residual
// only, D4; spec = bilinear from the reference planes).
template <bool COMPAT>
HD void phase_inter_compute(InterWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane, bool coded)
{
    const int bs = w.bs;
    int mbx = a % g.w_mbs, mby = a / g.w_mbs;
    uint8_t* ybase = pd.y + (size_t)(mby * 16) * g.pitch_y + mbx * 16;
    uint32_t exc = 0;
    if (bs != 4) {
        int row, col, ri, pitch, wr, c0, blk;
        const uint8_t* win;
        if (bs == 16) { row = lane >> 1; col = (lane & 1) * 8; win = w.win; pitch = 32; wr = row + 3; c0 = col; blk = 0; ri = h.u.ref_idx[0] & 15; }
        else {
            int q = lane >> 3, rr = lane & 7;
            row = (q >> 1) * 8 + rr; col = (q & 1) * 8; win = w.win + q * (14 * 16); pitch = 16; wr = rr + 3; c0 = 0;
            blk = (q >> 1) * 8 + (q & 1) * 2; ri = h.u.ref_idx[q] & 15;
        }
        uint32_t v = load_mv1(pd, a, blk);
        int pred[8];
        mc_row<COMPAT, 8>(win, pitch, wr, c0, mv_x(v) & 3, mv_y(v) & 3, pred);
        uint32_t packed[2];
        int bad = 0;
        finish_luma<COMPAT, 8>(pred, &w.res[row * 16 + col], coded, pd, ri, bad, packed);
        if (bad < 0) exc += count_excursions<8>(pred);
        *reinterpret_cast<uint2*>(ybase + (size_t)row * g.pitch_y + col) = uint2{packed[0], packed[1]};
    } else {
        int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        int ri = h.u.ref_idx[q] & 15;
        uint32_t v = load_mv1(pd, a, b);
        const uint8_t* win = w.win + b * 160;
#pragma unroll 1
        for (int n = 0; n < 2; n++) {
            int rr = (lane & 1) * 2 + n, row = by * 4 + rr, col = bx * 4;
            int pred[4];
            mc_row<COMPAT, 4>(win, 16, rr + 3, 0, mv_x(v) & 3, mv_y(v) & 3, pred);
            uint32_t packed[1];
            int bad = 0;
            finish_luma<COMPAT, 4>(pred, &w.res[row * 16 + col], coded, pd, ri, bad, packed);
            if (bad < 0) exc += count_excursions<4>(pred);
            *reinterpret_cast<uint32_t*>(ybase + (size_t)row * g.pitch_y + col) = packed[0];
        }
    }
    // chroma: lane -> plane (lane >> 4), row (l >> 1), 4 samples
    {
        int plane = lane >> 4, l = lane & 15, cy = l >> 1, cx0 = (l & 1) * 4;
        uint8_t* cbase = (plane ? pd.v : pd.u) + (size_t)(mby * 8 + cy) * g.pitch_c + mbx * 8 + cx0;
        uint32_t packed = 0;
        if (COMPAT) {
            if (coded) {
                const uint32_t* rp = reinterpret_cast<const uint32_t*>(&w.res[256 + plane * 64 + cy * 8 + cx0]);
                uint32_t r0 = rp[0], r1 = rp[1];
                packed = (r0 & 255u) | ((r0 >> 8) & 0xff00u) | ((r1 & 255u) << 16) | ((r1 << 8) & 0xff000000u);
            }
        } else {
            const int cw = g.w_mbs * 8, ch = g.h_mbs * 8;
            for (int k = 0; k < 4; k++) {
                int cx = cx0 + k, blk = (cy >> 1) * 4 + (cx >> 1), q = (cy >> 2) * 2 + (cx >> 2);
                int ri = h.u.ref_idx[q] & 15;
                const uint8_t* ref = plane ? pd.ref_v[ri] : pd.ref_u[ri];
                uint32_t v2 = load_mv1(pd, a, blk);
                int mvx = mv_x(v2), mvy = mv_y(v2);
                int xf = mvx & 7, yf = mvy & 7;
                int xi = mbx * 8 + cx + (mvx >> 3), yi = mby * 8 + cy + (mvy >> 3);
                int xa = clip3(0, cw - 1, xi), xb = clip3(0, cw - 1, xi + 1), ya = clip3(0, ch - 1, yi), yb = clip3(0, ch - 1, yi + 1);
                int A = H264R_LDG(ref + (size_t)ya * g.pitch_c + xa), B = H264R_LDG(ref + (size_t)ya * g.pitch_c + xb);
                int C = H264R_LDG(ref + (size_t)yb * g.pitch_c + xa), D = H264R_LDG(ref + (size_t)yb * g.pitch_c + xb);
                int v = ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6;
                if (pd.weighted_pred) v = weight_pred(v, pd.chroma_log2_wd, pd.chroma_weight[ri][plane], pd.chroma_offset[ri][plane]);
                if (coded) v += w.res[256 + plane * 64 + cy * 8 + cx];
                packed |= (uint32_t)clip1(v) << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t*>(cbase) = packed;
    }
    if (exc) {
#if defined(__CUDA_ARCH__)
        atomicAdd(&w.excursions, exc);
#else
        w.excursions += exc;
#endif
    }
}

// K1+K3 fast sequence for one inter macroblock (4x4 transform).  8x8-transform macroblocks take
// seq_inter_mb of h264r_seq.cuh.
template <bool COMPAT, class Exec>
HD void seq_inter_mb_fast(Exec& ex, InterWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a)
{
    const bool coded = h.cbp != 0;
    ex([&](int lane) {
        if (lane == 0) w.excursions = 0;
        if (coded) phase_residual_inter_fast<COMPAT>(w.res, pd, h, a, lane);
        phase_inter_stage<COMPAT>(w, pd, g, h, a, lane);
    });
    ex([&](int lane) { phase_inter_compute<COMPAT>(w, pd, g, h, a, lane, coded); });
    ex([&](int lane) {
        if (lane == 0 && w.excursions && pd.excursion_map) pd.excursion_map[a] = 1;
    });
}

}  // namespace h264r
