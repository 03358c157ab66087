// h264r_mc.cuh -- K1+K3 for inter macroblocks (4x4 transform): residual in registers, reference
// windows staged once in shared memory, quarter-sample interpolation with dp4a on packed bytes.
//
// Same results as phase_residual / phase_inter_luma / phase_inter_chroma of h264r_core.cuh (kept as
// the plain statement of the arithmetic and as the path for the 8x8 transform); this file is the
// throughput version:
//   * reference pictures carry a replicated border (PAD_* below, written by k_pad), so the
//     coordinate clamp of the reference (h264/macroblock.cpp:828-830, = ITU-T H.264 eq. 8-239..8-240)
//     becomes a clamp of the BLOCK position and window staging is plain word loads;
//   * the window of every prediction block is staged once (16x16: one 22x24-byte window; uniform 8x8
//     quadrants: four 14x16; otherwise sixteen 10x12), already shifted so that consumers read
//     aligned words;
//   * every lane produces a 4x2 patch (block lane>>1, rows 2*(lane&1)..+1) from 8 window rows: per
//     row three LDS + six PRMT give the eight byte groups A_k = x[k-2..k+1], and every quantity of
//     the 16 quarter-sample positions is a dp4a accumulation over those groups -- including the
//     reference's own tap pattern for j and s (D6/D7), see the derivation at mc_accum;
//   * no per-lane branches: lanes that do not need a quantity compute it anyway and drop it through
//     select masks; what NO lane of the warp needs is skipped by choosing one of three instantiations
//     (full-sample only / no centre position j / everything) with a warp-uniform switch;
//   * rounding, Clip1, the (P+Q+1)>>1 averages, the residual add and the range check run on packed
//     bytes / 16-bit pairs.
// Every function is a per-lane phase (cross-lane traffic goes through shared memory), so the host
// emulator of the CPU test-suite runs this code too.
#pragma once
#include "h264r_core.cuh"

namespace h264r {

// ---- intrinsics with host emulation -------------------------------------------------------------------
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
// byte permute: result byte i = byte (sel nibble i) of the 8-byte pool {a (0..3), b (4..7)}; selectors 0..7 only
HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * i)) & 7))) & 255u) << (8 * i);
    return r;
#endif
}
// bytes {sat8(v0), sat8(v1), sat8(v2), sat8(v3)}, sat8 = clamp to [0,255]
HD uint32_t pack_sat_u8(int v0, int v1, int v2, int v3)
{
#if defined(__CUDA_ARCH__)
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(v1), "r"(v0), "r"(hi));
    return r;
#else
    return (uint32_t)clip1(v0) | ((uint32_t)clip1(v1) << 8) | ((uint32_t)clip1(v2) << 16) | ((uint32_t)clip1(v3) << 24);
#endif
}
// {sat16(lo), sat16(hi)} as two 16-bit halves, sat16 = clamp to [-32768, 32767]
HD uint32_t pack_sat_s16x2(int lo, int hi)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
#else
    return (uint32_t)(clip3(-32768, 32767, lo) & 0xffff) | ((uint32_t)clip3(-32768, 32767, hi) << 16);
#endif
}
// per byte (a + b + 1) >> 1
HD uint32_t avg_u8x4(uint32_t a, uint32_t b) { return (a | b) - (((a ^ b) & 0xfefefefeu) >> 1); }
// per 16-bit half a + b (wrapping)
HD uint32_t add_s16x2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
// per 16-bit half clamp to [0,255]
HD uint32_t clip255_s16x2(uint32_t a)
{
#if defined(__CUDA_ARCH__)
    return __vimin_s16x2_relu(a, 0x00ff00ffu);
#else
    int lo = (int16_t)(a & 0xffff), hi = (int16_t)(a >> 16);
    return (uint32_t)clip1(lo) | ((uint32_t)clip1(hi) << 16);
#endif
}
// per 16-bit half clamp to [-32000, 32000]
HD uint32_t clamp_s16x2(uint32_t a)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(__vmins2(a, 0x7d007d00u), 0x83008300u);
#else
    int lo = (int16_t)(a & 0xffff), hi = (int16_t)(a >> 16);
    return (uint32_t)(clip3(-32000, 32000, lo) & 0xffff) | ((uint32_t)(clip3(-32000, 32000, hi) & 0xffff) << 16);
#endif
}
HD constexpr uint32_t taps4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 255) | ((uint32_t)(b & 255) << 8) | ((uint32_t)(c & 255) << 16) | ((uint32_t)(d & 255) << 24);
}

// Replicated border around every plane of the frame pool (samples).  Luma: a 16x16 block clamped to
// [-(16+3), W+2] x [-(16+3), H+3] reads columns -21..W+20 (+ up to 7 bytes of word alignment) and rows
// -22..H+21.  Chroma (spec mode): coordinates are clamped to [-8, Wc+7] x [-8, Hc+7].
constexpr int PAD_X = 32, PAD_Y = 24, PAD_XC = 16, PAD_YC = 12;

// Border replication of one row r (-pady <= r < H + pady) of a plane whose sample (0,0) is at org.
// Rows inside the picture get their two side borders, rows above / below are copies of the (extended)
// first / last row.  Lanes stride over 4-byte words.
HD void pad_row_phase(uint8_t* org, int pitch, int W, int H, int padx, int pady, int r, int lane)
{
    (void)pady;
    const int rs = clip3(0, H - 1, r);
    const uint8_t* src = org + (ptrdiff_t)rs * pitch;
    uint8_t* dst = org + (ptrdiff_t)r * pitch;
    const uint32_t left = (uint32_t)H264R_LDCG(src) * 0x01010101u, right = (uint32_t)H264R_LDCG(src + W - 1) * 0x01010101u;
    const int pw = padx >> 2;
    if (r >= 0 && r < H) {
        for (int i = lane; i < 2 * pw; i += 32) {
            if (i < pw) *reinterpret_cast<uint32_t*>(dst - padx + 4 * i) = left;
            else *reinterpret_cast<uint32_t*>(dst + W + 4 * (i - pw)) = right;
        }
    } else {
        for (int i = lane; i < (W >> 2) + 2 * pw; i += 32) {
            int x = 4 * i - padx;
            uint32_t v = x < 0 ? left : (x >= W ? right : H264R_LDCG(reinterpret_cast<const uint32_t*>(src + x)));
            *reinterpret_cast<uint32_t*>(dst + x) = v;
        }
    }
}

// The same border written by the warp that reconstructed a BORDER macroblock (REF_COMPAT luma: nothing modifies a
// sample after its macroblock's store, so the replication need not wait for the whole picture and k_pad's launch
// disappears from the chain picture -> next picture).  Runs as its own phase after the macroblock's stores: the
// samples are read back through L2.  Left / right: the macroblock's 16 rows x PAD_X bytes; top / bottom: PAD_Y copies of
// its first / last row, and next to a corner macroblock the PAD_X x PAD_Y corner block.  All stores are 16 bytes
// (plane origin, pitch, PAD_X and macroblock columns are multiples of 16).
HD void phase_pad_mb_luma(uint8_t* org, int pitch, int w_mbs, int h_mbs, int mbx, int mby, int lane)
{
    const int W = w_mbs * 16, H = h_mbs * 16, x0 = mbx * 16, y0 = mby * 16;
    const bool left = mbx == 0, right = mbx == w_mbs - 1;
    auto splat = [&](int x, int y) -> uint4 {
        const uint32_t s = (uint32_t)H264R_LDCG(org + (ptrdiff_t)y * pitch + x) * 0x01010101u;
        return uint4{s, s, s, s};
    };
    if (left || right) {
        const int r = y0 + (lane >> 1), k = (lane & 1) * 16;
        if (left) *reinterpret_cast<uint4*>(org + (ptrdiff_t)r * pitch - PAD_X + k) = splat(0, r);
        if (right) *reinterpret_cast<uint4*>(org + (ptrdiff_t)r * pitch + W + k) = splat(W - 1, r);
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        if (side == 0 ? mby != 0 : mby != h_mbs - 1) continue;
        const int ys = side == 0 ? 0 : H - 1;
        if (lane < PAD_Y) {
            uint8_t* dst = org + (ptrdiff_t)(side == 0 ? -1 - lane : H + lane) * pitch;
            *reinterpret_cast<uint4*>(dst + x0) = H264R_LDCG(reinterpret_cast<const uint4*>(org + (ptrdiff_t)ys * pitch + x0));
            if (left) { const uint4 s = splat(0, ys); *reinterpret_cast<uint4*>(dst - PAD_X) = s; *reinterpret_cast<uint4*>(dst - PAD_X + 16) = s; }
            if (right) { const uint4 s = splat(W - 1, ys); *reinterpret_cast<uint4*>(dst + W) = s; *reinterpret_cast<uint4*>(dst + W + 16) = s; }
        }
    }
}
// out of line on the device: one macroblock in twenty takes it, and the inter kernel's hot code should stay small
#if defined(__CUDACC__)
__device__ __noinline__ void pad_mb_luma_dev(uint8_t* org, int pitch, int w_mbs, int h_mbs, int mbx, int mby, int lane)
{
    phase_pad_mb_luma(org, pitch, w_mbs, h_mbs, mbx, mby, lane);
}
#endif
HD void pad_mb_luma_cold(uint8_t* org, int pitch, int w_mbs, int h_mbs, int mbx, int mby, int lane)
{
#if defined(__CUDA_ARCH__)
    pad_mb_luma_dev(org, pitch, w_mbs, h_mbs, mbx, mby, lane);
#else
    phase_pad_mb_luma(org, pitch, w_mbs, h_mbs, mbx, mby, lane);
#endif
}
HD bool is_border_mb(const Geometry& g, int mbx, int mby) { return mbx == 0 || mby == 0 || mbx == g.w_mbs - 1 || mby == g.h_mbs - 1; }

// Working set of the inter path (one per warp)
struct alignas(128) InterWork {
    uint32_t win[512];          // staged windows, see phase_inter_stage (TMA destination: 128-byte aligned)
    int16_t res[384];           // residual
    uint8_t need[32];           // per lane: NEED_* of its block
    uint8_t nzf[32];            // spec mode: "has a non-zero level" per luma block (raster; 4x4 transform) or per (quadrant, row) of
                                // the 8x8 transform, for K4
    int32_t t8[4 * 72];         // 8x8 transform: the four blocks after the row pass, block q at 72 q
    int32_t shape;              // 0: one 16x16 window, 1: four 8x8 windows, 2: sixteen 4x4 windows
};
static_assert(sizeof(InterWork) % 128 == 0, "TMA destinations of consecutive warps must stay 128-byte aligned");
// mbarrier of a warp's TMA loads and the phase parity its current macroblock completes (lives next to, not inside,
// the working sets: the generic phases reuse their memory)
struct TmaState { unsigned long long mbar; uint32_t parity; uint32_t pad_; };
// window layouts (words), load-staged: shape 0: 22 rows, pitch 6 (conflict-free for the 4x2 patches); shape 1: window q
// at 56 q, 14 rows, pitch 4; shape 2: window b at 30 b, 10 rows, pitch 3
constexpr int WIN0_PITCH = 6, WIN1_STRIDE = 56, WIN1_PITCH = 4;
// TMA windows.  A tensor copy must start on a 16-byte boundary of the innermost dimension (measured on B200: any
// other start coordinate raises an illegal-instruction fault), while a prediction window starts at an arbitrary byte.
// The box is therefore widened to the enclosing 16-byte columns -- 48 x 22 bytes for a 16x16 block, 32 x 14 for an
// 8x8 quadrant -- and the consumer applies the remaining offset: whole words through its row pointer, 0..3 bytes
// with one funnel shift per word (mc_groups_shifted).
constexpr int TMA0_W = 48, TMA0_H = 22, TMA0_PITCH = 12, TMA1_W = 32, TMA1_H = 14, TMA1_PITCH = 8, TMA1_STRIDE = 128;
// 8x8 transform: 32 (quadrant, row) flags -> 16-bit mask (the four 4x4 blocks of a quadrant share its flag)
HD uint32_t nz_pack8(const uint8_t* f)
{
    const uint2* v = reinterpret_cast<const uint2*>(f);
    uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint2 w = v[q];
        if (w.x | w.y) m |= 0x33u << ((q >> 1) * 8 + (q & 1) * 2);     // raster blocks b0, b0+1, b0+4, b0+5
    }
    return m;
}
// 16 byte flags -> 16-bit mask
HD uint32_t nz_pack(const uint8_t* f)
{
    const uint4 v = *reinterpret_cast<const uint4*>(f);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) m |= (((w[i] & 1u) | ((w[i] >> 7) & 2u) | ((w[i] >> 14) & 4u) | ((w[i] >> 21) & 8u)) << (4 * i));
    return m;
}

// ---- K1: residual in registers -----------------------------------------------------------------------
struct Dequant {                // per QP: d = (c * mul + add) >> shr for the three position classes
    int mul[3], add, shr;
};
// normAdjust4x4 columns as byte tables indexed by qp % 6: (even,even) 10 11 13 14 16 18, (odd,odd) 16 18 20 23 25 29,
// other 13 14 16 18 20 23
HD constexpr uint64_t norm_adjust_ee() { return 0x12100E0D0B0Aull; }
HD constexpr uint64_t norm_adjust_oo() { return 0x1D1917141210ull; }
HD constexpr uint64_t norm_adjust_other() { return 0x171412100E0Dull; }
HD int norm_adjust_dc(int m) { return (int)((norm_adjust_ee() >> (8 * m)) & 255); }
HD Dequant make_dequant(int qp)
{
    // normAdjust4x4 (reference residual::LevelScale4x4, h264/residual.cpp:170-183): columns = (even,even) / (odd,odd) / other
    // one byte per qp % 6 = 0..5, no local-memory table
    int s = (qp * 43) >> 8, m = qp - 6 * s;         // qp / 6, qp % 6 for 0 <= qp < 64
    const int v[3] = {(int)((norm_adjust_ee() >> (8 * m)) & 255), (int)((norm_adjust_oo() >> (8 * m)) & 255),
                      (int)((norm_adjust_other() >> (8 * m)) & 255)};
    Dequant q;
    if (s >= 4) { for (int k = 0; k < 3; k++) q.mul[k] = (16 * v[k]) << (s - 4); q.add = 0; q.shr = 0; }
    else { for (int k = 0; k < 3; k++) q.mul[k] = 16 * v[k]; q.add = 1 << (3 - s); q.shr = 4 - s; }
    return q;
}
// c: 16 levels (raster); returns residual in r.  Same arithmetic as scale_idct4
// (reference residual::ScalingAndTransform_Residual4x4Blocks, h264/residual.cpp:367-434).
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

// K1 for any 4x4-transform macroblock: h264r_intra.cuh
template <bool COMPAT, bool LEAN = false, bool INTER = false>
HD void phase_residual_fast(int16_t* res, const PicDesc& pd, const h264r_mb_hdr& h, int a, int lane, uint8_t* nz_flags = nullptr,
                            bool chroma_only = false, const MbPre* pre = nullptr, bool luma_only = false);

// ---- K1 with the 8x8 transform (ITU-T H.264 8.5.13, flat scaling lists; the reference has no 8x8 transform) ------------
// LevelScale8x8 / 16 for position (i, j), qP % 6 = m: six classes of positions, one byte per class and m
HD int level_scale8_fast(int m, int i, int j)
{
    int k;
    if ((i & 3) == 0 && (j & 3) == 0) k = 0;
    else if ((i & 1) && (j & 1)) k = 1;
    else if ((i & 3) == 2 && (j & 3) == 2) k = 2;
    else if (((i & 3) == 0 && (j & 1)) || ((i & 1) && (j & 3) == 0)) k = 3;
    else if (((i & 3) == 0 && (j & 3) == 2) || ((i & 3) == 2 && (j & 3) == 0)) k = 4;
    else k = 5;
    // normAdjust8x8 columns {20,18,32,19,25,24} {22,19,35,21,28,26} {26,23,42,24,33,31} {28,25,45,26,35,33} {32,28,51,30,40,38}
    // {36,32,58,34,46,43} as one 64-bit word per m, class k in byte k
    const uint64_t t = m == 0 ? 0x181913201214ull : (m == 1 ? 0x1A1C15231316ull : (m == 2 ? 0x1F21182A171Aull :
                       (m == 3 ? 0x21231A2D191Cull : (m == 4 ? 0x26281E331C20ull : 0x2B2E223A2024ull))));
    return 16 * (int)((t >> (8 * k)) & 255u);
}
// Row pass: lane (q = lane >> 3, r = lane & 7) dequantises and transforms row r of 8x8 block q into t8.
HD void phase_residual8_rows(InterWork& w, const PicDesc& pd, const h264r_mb_hdr& h, int a, int lane)
{
    const int q = lane >> 3, r = lane & 7;
    const BlkRef br = load_blkref(pd, a);
    const uint8_t* src = blk_ptr(pd, br, q * 4 + (r >> 1));
    int x[8];
    int nz = 0;
    if (src && ((h.cbp >> q) & 1)) {
        const int4 v = blk_levels8(pd, br, src, (r & 1) * 8);
        const int wv[4] = {v.x, v.y, v.z, v.w};
        const int qp = h.qp_y, s = (qp * 43) >> 8, m = qp - 6 * s;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int c0 = (int)(int16_t)(wv[k] & 0xffff), c1 = wv[k] >> 16;
            nz |= c0 | c1;
            const int l0 = level_scale8_fast(m, r, 2 * k), l1 = level_scale8_fast(m, r, 2 * k + 1);
            x[2 * k] = s >= 6 ? (c0 * l0) * (1 << (s - 6)) : ((c0 * l0 + (1 << (5 - s))) >> (6 - s));
            x[2 * k + 1] = s >= 6 ? (c1 * l1) * (1 << (s - 6)) : ((c1 * l1 + (1 << (5 - s))) >> (6 - s));
        }
        idct8_1d(x);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = 0;
    }
    w.nzf[lane] = nz ? 1 : 0;
    int4* dst = reinterpret_cast<int4*>(&w.t8[q * 72 + r * 8]);
    dst[0] = int4{x[0], x[1], x[2], x[3]};
    dst[1] = int4{x[4], x[5], x[6], x[7]};
}
// Column pass: lane (q, c) transforms column c of block q and writes the residual.
HD void phase_residual8_cols(InterWork& w, int lane)
{
    const int q = lane >> 3, c = lane & 7;
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = w.t8[q * 72 + i * 8 + c];
    idct8_1d(x);
#pragma unroll
    for (int i = 0; i < 8; i++) w.res[((q >> 1) * 8 + i) * 16 + (q & 1) * 8 + c] = (int16_t)sat16((x[i] + 32) >> 6);
}

// ---- motion vectors, window shape --------------------------------------------------------------------
HD void load_mvs(const PicDesc& pd, int a, uint32_t mv[16])
{
    const uint4* p = reinterpret_cast<const uint4*>(pd.mv + (size_t)a * H264R_MV_PER_MB);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 v = H264R_LDG(p + i);
        mv[4 * i] = v.x; mv[4 * i + 1] = v.y; mv[4 * i + 2] = v.z; mv[4 * i + 3] = v.w;
    }
}
// 0: all sixteen vectors (and reference indices) equal; 1: every 8x8 quadrant uniform; 2: anything else.
// Under COMPAT a P_8x8 macroblock with a 4x4 sub-partition always takes shape 2: the reference's source-row
// offset of sub-blocks 2 and 3 (D8, h264/macroblock.cpp:786-789) makes equal vectors read different rows.
// partition layout of the vectors: the shape follows from the macroblock type alone
HD int inter_shape_by_type(const h264r_mb_hdr& h)
{
    if (h.mb_class == H264R_MB_P16x16) return 0;
    if (h.mb_class != H264R_MB_P8x8) return 1;
    return h.sub_mb_type == 0 ? 1 : 2;
}
template <bool COMPAT>
HD int inter_shape(const h264r_mb_hdr& h, const uint32_t mv[16])
{
    if (COMPAT && h.mb_class == H264R_MB_P8x8) { uint32_t s = h.sub_mb_type; if ((s & (s >> 1)) & 0x55u) return 2; }
    bool q_uniform = true, all = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int b0 = (q >> 1) * 8 + (q & 1) * 2;
        q_uniform = q_uniform && mv[b0] == mv[b0 + 1] && mv[b0] == mv[b0 + 4] && mv[b0] == mv[b0 + 5];
        all = all && mv[b0] == mv[0];
    }
    if (!q_uniform) return 2;
    const uint32_t rw = hdr_word(h, 0);
    bool ref_same = rw == (rw & 255u) * 0x01010101u;
    return (all && ref_same) ? 0 : 1;
}

// ---- what a quarter-sample position needs ---------------------------------------------------------------
// Position sel = xFrac*4 + yFrac produces (P + Q + 1) >> 1 with P, Q out of
//   F  a full sample: G (0,0), H (+1,0) or M (0,+1)          b  Clip1((b1+16)>>5), b1 = 6-tap of row 0
//   v  Clip1((v1+16)>>5), v1 = 6-tap of column 0 (h) or +1 (m)  s  Clip1((s1+16)>>5), s1 = 6-tap of row +1
//   j  Clip1((j1+512)>>10)
// (reference predPart table, h264/macroblock.cpp:893-901).
constexpr int NEED_H = 1, NEED_S = 2, NEED_V = 4, NEED_J = 8;
struct McSel {
    // select masks (0 / ~0): P <- G, H, M, b, v, j;  Q <- G, b, v, s, j
    uint32_t pG, pH, pM, pb, pv, pj, qG, qb, qv, qs, qj;
    uint32_t tv1, tv5, tv20;     // vertical taps 1, -5, 20 in the byte of column 0 (h) or +1 (m) of A_k
    int need;
};
HD McSel mc_select(int xf, int yf)
{
    const int sel = xf * 4 + yf;
    //                      sel:  0  1  2  3  4  5  6  7  8  9 10 11 12 13 14 15
    // P: 0 G, 1 H, 2 M, 3 b, 4 v, 5 j   0  0  4  2  0  3  4  4  3  3  5  5  1  3  5  4
    // Q: 0 G, 1 b, 2 v, 3 s, 4 j        0  2  2  2  1  2  4  3  1  4  4  3  1  2  2  3
    const uint32_t PTl = 0x44302400u, PTh = 0x45315533u, QTl = 0x34212220u, QTh = 0x32213441u;
    const int sh4 = 4 * (sel & 7);
    const int p = (int)(((sel & 8 ? PTh : PTl) >> sh4) & 15u), q = (int)(((sel & 8 ? QTh : QTl) >> sh4) & 15u);
    McSel m;
    m.pG = p == 0 ? ~0u : 0u; m.pH = p == 1 ? ~0u : 0u; m.pM = p == 2 ? ~0u : 0u;
    m.pb = p == 3 ? ~0u : 0u; m.pv = p == 4 ? ~0u : 0u; m.pj = p == 5 ? ~0u : 0u;
    m.qG = q == 0 ? ~0u : 0u; m.qb = q == 1 ? ~0u : 0u; m.qv = q == 2 ? ~0u : 0u;
    m.qs = q == 3 ? ~0u : 0u; m.qj = q == 4 ? ~0u : 0u;
    const bool nj = p == 5 || q == 4;
    m.need = ((p == 3 || q == 1 || nj) ? NEED_H : 0) | ((q == 3 || nj) ? NEED_S : 0) | ((p == 4 || q == 2) ? NEED_V : 0) | (nj ? NEED_J : 0);
    const int vshift = xf == 3 ? 24 : 16;
    m.tv1 = 1u << vshift; m.tv5 = (uint32_t)(-5 & 255) << vshift; m.tv20 = 20u << vshift;
    return m;
}
HD int mc_need(int xf, int yf)
{
    // mc_select(xf, yf).need as a table, 4 bits per position xf*4 + yf (checked against mc_select by the test emulator):
    // 0:- 1:V 2:V 3:V 4:H 5:HV 6:HSVJ 7:SV 8:H 9:HSJ 10:HSJ 11:HSJ 12:H 13:HV 14:HSVJ 15:SV
    const int sel = xf * 4 + yf;
    const uint32_t lo = 0x6F514440u, hi = 0x6F51BBB1u;
    return (int)(((sel & 8 ? hi : lo) >> (4 * (sel & 7))) & 15u);
}

// bytes (x .. x+3, y) of a padded reference plane, x of any alignment
HD uint32_t ref_word(const uint8_t* ref, int gpitch, int x, int y)
{
    const uintptr_t ad = reinterpret_cast<uintptr_t>(ref + (ptrdiff_t)y * gpitch + x);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(ad & ~(uintptr_t)3);
    return funnel_r(H264R_LDG(q), H264R_LDG(q + 1), (int)(ad & 3) * 8);
}

// ---- window staging ----------------------------------------------------------------------------------------
// Items first, first + STRIDE, ... of a window of NROWS x NWORDS words whose top-left byte is (sx, sy) of the padded
// reference plane.  A staged word = 4 bytes at any alignment = funnel shift of two aligned words; the alignment is
// the same for every word of a window (the pitch is a multiple of 4), and the source address advances incrementally.
template <int NROWS, int NWORDS, int STRIDE, int DPITCH>
HD void stage_window(uint32_t* dst, const uint8_t* ref, int gpitch, int sx, int sy, int first)
{
    constexpr int DR = STRIDE / NWORDS, DJ = STRIDE % NWORDS;       // one step = DR rows + DJ words
    int row = first / NWORDS, j = first - row * NWORDS;
    const uint8_t* p0 = ref + (ptrdiff_t)(sy + row) * gpitch + sx + 4 * j;
    const int sh = (int)(reinterpret_cast<uintptr_t>(p0) & 3) * 8;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p0) & ~(uintptr_t)3);
    const int wpitch = gpitch >> 2;
#pragma unroll
    for (int i = 0; i < (NROWS * NWORDS + STRIDE - 1) / STRIDE; i++) {
        const int it = first + i * STRIDE;
        if (it < NROWS * NWORDS) dst[DPITCH == NWORDS ? it : row * DPITCH + j] = funnel_r(H264R_LDG(q), H264R_LDG(q + 1), sh);
        q += DR * wpitch + DJ;
        if (DPITCH != NWORDS) row += DR;
        if (DJ) {
            j += DJ;
            if (j >= NWORDS) { j -= NWORDS; if (DPITCH != NWORDS) row++; q += wpitch - NWORDS; }
        }
    }
}

#if defined(__CUDACC__)
// ---- TMA: one cp.async.bulk.tensor per window, completion on the warp's mbarrier -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity)
{
    const uint32_t addr = smem_u32(b);
    uint32_t done = 0;
    for (int spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spin > (1 << 22)) __trap();       // a load that never lands must not hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, unsigned long long* b)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(b)) : "memory");
}
// frame slot whose tensor maps serve ref_list0[ri]; a stray index beyond the list (ref_id -1) falls back to entry 0,
// as the plane pointers of unused entries do (fill_desc)
__device__ __forceinline__ int tma_ref_slot(const PicDesc& pd, int ri)
{
    const int r = pd.ref_id[ri];
    return r >= 0 ? r : pd.ref_id[0];
}
#endif

// Phase M0: shape of the macroblock, this lane's NEED bits, and the windows.
// Window of a bw x bh block at integer position (xb, yb): bytes xb-2 .. and rows yb-3 .. yb+bh+2
// (the reference's j reaches row -3, D7).  Block positions are clamped to [-(bw+3), W+2] x [-(bh+3), H+3],
// which gives every sample the value of the reference's per-sample coordinate clamp.
template <bool COMPAT, bool TMA, bool LEAN>
HD void phase_inter_stage(InterWork& w, TmaState& ts, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane, uint32_t& mv_own,
                          const MbPre& pre)
{
    (void)ts;
    const bool partition_layout = LEAN ? true : pd.mv_off != nullptr;
    int shape;
    if (partition_layout) shape = inter_shape_by_type(h);
    else {
        uint32_t mv[16];
        load_mvs(pd, a, mv);
        shape = inter_shape<COMPAT>(h, mv);
    }
    if (!LEAN && lane == 0) w.shape = shape;
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    const int W = g.w_mbs * 16, H = g.h_mbs * 16;
    // the vector of the lane's own 4x4 block (lane >> 1): fetched once per macroblock, it serves the NEED bits, the window
    // position below (a window's lanes are lanes of its blocks, which share the vector) and the compute phase
    const uint32_t v = load_mv_pre(pd, h, a, lane >> 1, pre, partition_layout);
    mv_own = v;
#if defined(__CUDA_ARCH__)
    if (TMA)        // (the load-staged device path takes the union of the NEED bits with one warp reduction, see phase_inter_compute)
#endif
    w.need[lane] = (uint8_t)mc_need(mv_x(v) & 3, mv_y(v) & 3);
#if defined(__CUDA_ARCH__)
    if (TMA && shape < 2) {
        // reference-frame tiles by TMA: lane 0 posts the byte count and one tensor copy per window; every lane waits on
        // the warp's mbarrier at the start of the compute phase.  Tensor coordinates are relative to the padded plane.
        if (lane == 0) {
            ts.parity ^= 1u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the windows were last read through the generic proxy
            const char* maps = static_cast<const char*>(pd.tmaps);
            if (shape == 0) {
                const uint32_t v = load_mv(pd, h, a, 0);
                const int xb = clip3(-19, W + 2, mbx * 16 + (mv_x(v) >> 2)), yb = clip3(-19, H + 3, mby * 16 + (mv_y(v) >> 2));
                mbar_expect_tx(&ts.mbar, (uint32_t)(TMA0_W * TMA0_H));
                tma_load_2d(w.win, maps + ((size_t)tma_ref_slot(pd, hdr_ref_idx(h, 0)) * 2) * 128, (PAD_X + xb - 2) & ~15, PAD_Y + yb - 3, &ts.mbar);
            } else {
                mbar_expect_tx(&ts.mbar, (uint32_t)(4 * TMA1_W * TMA1_H));
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t v = load_mv(pd, h, a, (q >> 1) * 8 + (q & 1) * 2);
                    const int xb = clip3(-11, W + 2, mbx * 16 + (q & 1) * 8 + (mv_x(v) >> 2)), yb = clip3(-11, H + 3, mby * 16 + (q >> 1) * 8 + (mv_y(v) >> 2));
                    tma_load_2d(w.win + q * TMA1_STRIDE, maps + ((size_t)tma_ref_slot(pd, hdr_ref_idx(h, q)) * 2 + 1) * 128, (PAD_X + xb - 2) & ~15, PAD_Y + yb - 3, &ts.mbar);
                }
            }
        }
        return;
    }
#endif
    if (shape == 0) {
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, 0)];
        int xb = clip3(-19, W + 2, mbx * 16 + (mv_x(v) >> 2)), yb = clip3(-19, H + 3, mby * 16 + (mv_y(v) >> 2));
        stage_window<22, 6, 32, WIN0_PITCH>(w.win, ref, g.pitch_y, xb - 2, yb - 3, lane);
    } else if (shape == 1) {
        // quadrant q is staged by the eight lanes of its four blocks
        const int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, q)];
        int xb = clip3(-11, W + 2, mbx * 16 + (q & 1) * 8 + (mv_x(v) >> 2)), yb = clip3(-11, H + 3, mby * 16 + (q >> 1) * 8 + (mv_y(v) >> 2));
        stage_window<14, 4, 8, WIN1_PITCH>(w.win + q * WIN1_STRIDE, ref, g.pitch_y, xb - 2, yb - 3, ((by & 1) * 2 + (bx & 1)) * 2 + (lane & 1));
    } else {
        const int b = lane >> 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
        const uint8_t* ref = pd.ref_y[hdr_ref_idx(h, q)];
        int ys = mby * 16 + by * 4;
        // reference's source-row offset for 4x4 sub-partitions 2 and 3 (D8, h264/macroblock.cpp:786-789)
        if (COMPAT && h.mb_class == H264R_MB_P8x8 && ((h.sub_mb_type >> (2 * q)) & 3) == H264R_SUB_4x4 && (by & 1)) ys -= 3;
        int xb = clip3(-7, W + 2, mbx * 16 + bx * 4 + (mv_x(v) >> 2)), yb = clip3(-7, H + 3, ys + (mv_y(v) >> 2));
        stage_window<10, 3, 2, 3>(w.win + b * 30, ref, g.pitch_y, xb - 2, yb - 3, lane & 1);
    }
}

// ---- interpolation ----------------------------------------------------------------------------------------------
// One window row contributes to one output row at vertical offset DY = -3..3.  A[k] = x[k-2 .. k+1] for the
// lane's samples k = 0..3 and A[k+4] = x[k+2 .. k+5].  With TA = (1,-5,20,20) on A[k] and TB = (-5,1,0,0) on
// A[k+4] a dp4a pair is the 6-tap row filter; a vertical tap is a dp4a with one non-zero byte.
//
// j under COMPAT: the reference computes j1 = tap6(aa, bb, b1, s1, gg, hh) from row sums whose outer taps
// come from the wrong rows (D7, h264/macroblock.cpp:843-846) and s1 with tap K from row -1 (D6,
// h264/staticcharts.h:1505).  With L = x[-2]-5x[-1], M = 20x[0]+20x[1], R = -5x[2]+x[3] of a row:
//   aa = L(-2)+M(-2)+R(-2)   bb = L(-2)+M(-1)+R(-2)   gg = L(2)+M(2)+R(-2)   hh = L(3)+M(3)+R(-3)
//   j1 = aa - 5bb + 20 b1 + 20 s1 - 5gg + hh
//      = R(-3) + [-4L + M - 9R](-2) + [-5M](-1) + [-5L - 5M](2) + [L + M](3)  +  20 b1 + 20 s1
// so QJ accumulates the bracketed part with one or two dp4a per row.  Spec mode: j1 = sum over rows of
// vertical tap * row filter = (F(-2) + F(3)) - 5 (F(-1) + F(2)) + 20 b1 + 20 s1, accumulated in QJ / QJ5.
struct McAcc { int QH[4], QS[4], QV[4], QJ[4], QJ5[4]; uint32_t P, Q; };

// wn: union over the warp of the lanes' NEED bits (warp-uniform): a quantity no lane needs is neither accumulated nor finished.
// The rounding constant of (x + 16) >> 5 rides in the first tap of QH / QS / QV (j1 takes it out again, see mc_finish).
template <bool COMPAT, int CLS, int DY>
HD void mc_accum(const uint32_t A[8], const McSel& m, McAcc& c, int wn)
{
    constexpr uint32_t TA = taps4(1, -5, 20, 20), TB = taps4(-5, 1, 0, 0);
    if (DY < -3 || DY > 3) return;
    if (DY == 0) {       // full samples G (x[0] = byte 2 of A[k]: the four of them are A[2]) and H (A[3])
        c.P = (A[2] & m.pG) | (A[3] & m.pH);
        c.Q = A[2] & m.qG;
    }
    if (DY == 1) c.P |= A[2] & m.pM;
    if (CLS == 0) return;
    (void)wn;
    constexpr bool nh = true, ns = true, nv = true;      // (gating quantities by the warp's NEED bits at run time was measured: the extra
                                                         // branches and spills cost more than the skipped taps save -- 0.197 against 0.152 ms per launch)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (DY == 0 && nh) c.QH[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, 16));
        if (ns) {
            if (COMPAT) {
                if (DY == -1) c.QS[k] = dp4a_us(A[k], taps4(1, 0, 0, 0), 16);
                if (DY == 1) c.QS[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], taps4(0, -5, 20, 20), c.QS[k]));
            } else {
                if (DY == 1) c.QS[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, 16));
            }
        }
        if (nv) {
            if (DY == -2) c.QV[k] = dp4a_us(A[k], m.tv1, 16);
            if (DY == -1 || DY == 2) c.QV[k] = dp4a_us(A[k], m.tv5, c.QV[k]);
            if (DY == 0 || DY == 1) c.QV[k] = dp4a_us(A[k], m.tv20, c.QV[k]);
            if (DY == 3) c.QV[k] = dp4a_us(A[k], m.tv1, c.QV[k]);
        }
        if (CLS == 2) {
            if (COMPAT) {
                if (DY == -3) c.QJ[k] = dp4a_us(A[k + 4], TB, 0);
                if (DY == -2) c.QJ[k] = dp4a_us(A[k + 4], taps4(45, -9, 0, 0), dp4a_us(A[k], taps4(-4, 20, 20, 20), c.QJ[k]));
                if (DY == -1) c.QJ[k] = dp4a_us(A[k], taps4(0, 0, -100, -100), c.QJ[k]);
                if (DY == 2) c.QJ[k] = dp4a_us(A[k], taps4(-5, 25, -100, -100), c.QJ[k]);
                if (DY == 3) c.QJ[k] = dp4a_us(A[k], TA, c.QJ[k]);
            } else {
                if (DY == -2) c.QJ[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, 0));
                if (DY == 3) c.QJ[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, c.QJ[k]));
                if (DY == -1) c.QJ5[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, 0));
                if (DY == 2) c.QJ5[k] = dp4a_us(A[k + 4], TB, dp4a_us(A[k], TA, c.QJ5[k]));
            }
        }
    }
}

// the four prediction samples of one output row, packed
template <bool COMPAT, int CLS>
HD uint32_t mc_finish(const McSel& m, const McAcc& c, int wn)
{
    uint32_t P = c.P, Q = c.Q;
    if (CLS >= 1) {
        (void)wn;
        const uint32_t b = pack_sat_u8(c.QH[0] >> 5, c.QH[1] >> 5, c.QH[2] >> 5, c.QH[3] >> 5);
        const uint32_t s = pack_sat_u8(c.QS[0] >> 5, c.QS[1] >> 5, c.QS[2] >> 5, c.QS[3] >> 5);
        const uint32_t v = pack_sat_u8(c.QV[0] >> 5, c.QV[1] >> 5, c.QV[2] >> 5, c.QV[3] >> 5);
        P |= (b & m.pb) | (v & m.pv);
        Q |= (b & m.qb) | (v & m.qv) | (s & m.qs);
    }
    if (CLS == 2) {
        int j1[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            j1[k] = c.QJ[k] + 20 * (c.QH[k] + c.QS[k]) + (512 - 20 * 32);      // QH and QS carry +16 each
            if (!COMPAT) j1[k] -= 5 * c.QJ5[k];
        }
        uint32_t j = pack_sat_u8(j1[0] >> 10, j1[1] >> 10, j1[2] >> 10, j1[3] >> 10);
        P |= j & m.pj;
        Q |= j & m.qj;
    }
    return avg_u8x4(P, Q);
}

// A[0..7] of one window row from its three words; SH: the row starts `sh` bits into its first word (TMA windows)
template <bool SH>
HD void mc_groups(const uint32_t* row, int sh, uint32_t A[8])
{
    uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
    if (SH) {
        const uint32_t w3 = row[3];
        w0 = funnel_r(w0, w1, sh); w1 = funnel_r(w1, w2, sh); w2 = funnel_r(w2, w3, sh);
    }
    A[0] = w0; A[1] = prmt(w0, w1, 0x4321); A[2] = prmt(w0, w1, 0x5432); A[3] = prmt(w0, w1, 0x6543);
    A[4] = w1; A[5] = prmt(w1, w2, 0x4321); A[6] = prmt(w1, w2, 0x5432); A[7] = prmt(w1, w2, 0x6543);
}

// the lane's 4x2 patch: wp = first of its 8 window rows (row -3 of output row 0), pitch in words
template <bool COMPAT, int CLS, bool SH>
HD void mc_patch(const uint32_t* wp, int pitch, int sh, const McSel& m, uint32_t out[2], int wn)
{
    McAcc c0, c1;
    uint32_t A[8];
#define MC_ROW(RR)                                       \
    mc_groups<SH>(wp + (RR) * pitch, sh, A);             \
    mc_accum<COMPAT, CLS, (RR) - 3>(A, m, c0, wn);       \
    mc_accum<COMPAT, CLS, (RR) - 4>(A, m, c1, wn);
    if (CLS == 0) {
        // every lane of the warp sits on a full-sample position: row 0 of both output rows
        MC_ROW(3) MC_ROW(4)
    } else {
        if (CLS == 2 && COMPAT) { MC_ROW(0) }          // row -3 is reached only by the reference's j (D7)
        MC_ROW(1) MC_ROW(2) MC_ROW(3) MC_ROW(4) MC_ROW(5) MC_ROW(6) MC_ROW(7)
    }
#undef MC_ROW
    out[0] = mc_finish<COMPAT, CLS>(m, c0, wn);
    out[1] = mc_finish<COMPAT, CLS>(m, c1, wn);
}

// packed prediction + residual -> packed samples; returns the number of samples that left [0,255]
template <bool COMPAT>
HD uint32_t add_residual(uint32_t pred, const int16_t* res_row, uint32_t& packed)
{
    uint2 r = *reinterpret_cast<const uint2*>(res_row);
    if (!COMPAT) { r.x = clamp_s16x2(r.x); r.y = clamp_s16x2(r.y); }      // keeps pred + residual inside 16 bits; Clip1 gives the same sample
    uint32_t s0 = add_s16x2(prmt(pred, 0u, 0x4140), r.x), s1 = add_s16x2(prmt(pred, 0u, 0x4342), r.y);
    uint32_t n = 0;
    if ((s0 | s1) & 0xff00ff00u)
        n = ((s0 & 0xff00u) != 0) + ((s0 & 0xff000000u) != 0) + ((s1 & 0xff00u) != 0) + ((s1 & 0xff000000u) != 0);
    if (!COMPAT) { s0 = clip255_s16x2(s0); s1 = clip255_s16x2(s1); }
    packed = prmt(s0, s1, 0x6420);          // COMPAT: the plane keeps the low 8 bits (D3)
    return n;
}
HD uint32_t weight_packed(uint32_t pred, int logwd, int wgt, int off)
{
    return (uint32_t)weight_pred((int)(pred & 255u), logwd, wgt, off) | ((uint32_t)weight_pred((int)((pred >> 8) & 255u), logwd, wgt, off) << 8) |
           ((uint32_t)weight_pred((int)((pred >> 16) & 255u), logwd, wgt, off) << 16) | ((uint32_t)weight_pred((int)(pred >> 24), logwd, wgt, off) << 24);
}

// Phase M1: compute and store the macroblock (luma from the staged windows; chroma: COMPAT = residual
// only, D4; spec = bilinear from the padded reference planes, ITU-T H.264 8.4.2.2.2).
// PART: which share of the macroblock this instantiation reconstructs -- 0 everything; 1 / 2 its luma (instantiations for the
// macroblocks without / with the 8x8 transform differ only in the residual phases of seq_inter_mb_fast); 3 its chroma.  The parts
// exist for spec mode, whose single instantiation is 60 KB of hot code against a 32 KB instruction cache (ncu: 12 of 22 stall
// cycles per issue are "no instruction"): three launches of kernels that fit run faster than one that does not.
constexpr int PART_ALL = 0, PART_LUMA4 = 1, PART_LUMA8 = 2, PART_CHROMA = 3;
template <bool COMPAT, bool TMA, bool LEAN, int PART = PART_ALL>
HD void phase_inter_compute(InterWork& w, TmaState& ts, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, int lane, bool coded,
                            uint32_t mv_own, const MbPre* pre = nullptr)
{
    (void)ts;
    const bool weighted = LEAN ? false : pd.weighted_pred != 0;
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    uint32_t exc = 0;
    if (PART != PART_CHROMA) {
    // warp-wide union of the NEED bits
    int wn;
#if defined(__CUDA_ARCH__)
    if (!TMA) wn = (int)__reduce_or_sync(0xffffffffu, (unsigned)mc_need(mv_x(mv_own) & 3, mv_y(mv_own) & 3));      // every lane of the warp runs this phase
    else
#endif
    {
        const uint4* np = reinterpret_cast<const uint4*>(w.need);
        uint4 n0 = np[0], n1 = np[1];
        uint32_t o = n0.x | n0.y | n0.z | n0.w | n1.x | n1.y | n1.z | n1.w;
        o |= o >> 16; o |= o >> 8;
        wn = (int)(o & 15u);
    }
    const int b = lane >> 1, half = lane & 1, bx = b & 3, by = b >> 2, q = (by >> 1) * 2 + (bx >> 1);
    const int shape = LEAN ? inter_shape_by_type(h) : w.shape;       // (partition layout: the type says it)
    const uint32_t* wp;
    int pitch;
    bool tma = false;
    int sh = 0;
#if defined(__CUDA_ARCH__)
    tma = TMA && shape < 2;
    if (tma) {
        // where the window starts inside its 16-byte aligned box: same block position as in phase M0
        int mbx_, mby_;
        mb_xy(g, a, mbx_, mby_);
        const uint32_t vw = load_mv(pd, h, a, shape == 0 ? 0 : (q >> 1) * 8 + (q & 1) * 2);
        const int bw = shape == 0 ? 16 : 8;
        const int xb = clip3(-(bw + 3), g.w_mbs * 16 + 2, mbx_ * 16 + (shape == 0 ? 0 : (q & 1) * 8) + (mv_x(vw) >> 2));
        const int o = (PAD_X + xb - 2) & 15;
        sh = (o & 3) * 8;
        if (shape == 0) { wp = w.win + (by * 4 + half * 2) * TMA0_PITCH + (o >> 2) + bx; pitch = TMA0_PITCH; }
        else { wp = w.win + q * TMA1_STRIDE + ((by & 1) * 4 + half * 2) * TMA1_PITCH + (o >> 2) + (bx & 1); pitch = TMA1_PITCH; }
        mbar_wait(&ts.mbar, ts.parity);      // the windows posted in phase M0 have landed
    } else
#endif
    if (shape == 0) { wp = w.win + (by * 4 + half * 2) * WIN0_PITCH + bx; pitch = WIN0_PITCH; }
    else if (shape == 1) { wp = w.win + q * WIN1_STRIDE + ((by & 1) * 4 + half * 2) * WIN1_PITCH + (bx & 1); pitch = WIN1_PITCH; }
    else { wp = w.win + b * 30 + half * 2 * 3; pitch = 3; }
    const uint32_t v = mv_own;
    const McSel m = mc_select(mv_x(v) & 3, mv_y(v) & 3);
    uint32_t pred[2];
    if (TMA && tma) {
        if (wn == 0) mc_patch<COMPAT, 0, TMA>(wp, pitch, sh, m, pred, wn);
        else if (!(wn & NEED_J)) mc_patch<COMPAT, 1, TMA>(wp, pitch, sh, m, pred, wn);
        else mc_patch<COMPAT, 2, TMA>(wp, pitch, sh, m, pred, wn);
    } else {
        if (wn == 0) mc_patch<COMPAT, 0, false>(wp, pitch, 0, m, pred, wn);
        else if (!(wn & NEED_J)) mc_patch<COMPAT, 1, false>(wp, pitch, 0, m, pred, wn);
        else mc_patch<COMPAT, 2, false>(wp, pitch, 0, m, pred, wn);
    }
    const int ri = hdr_ref_idx(h, q);
    if (weighted) {
        pred[0] = weight_packed(pred[0], pd.luma_log2_wd, pd.luma_weight[ri], pd.luma_offset[ri]);
        pred[1] = weight_packed(pred[1], pd.luma_log2_wd, pd.luma_weight[ri], pd.luma_offset[ri]);
    }
    const int row = by * 4 + half * 2, col = bx * 4;
    uint8_t* yp = pd.y + (size_t)(mby * 16 + row) * g.pitch_y + mbx * 16 + col;
    if (coded) {
        exc += add_residual<COMPAT>(pred[0], &w.res[row * 16 + col], pred[0]);
        exc += add_residual<COMPAT>(pred[1], &w.res[(row + 1) * 16 + col], pred[1]);
    }
    *reinterpret_cast<uint32_t*>(yp) = pred[0];
    *reinterpret_cast<uint32_t*>(yp + g.pitch_y) = pred[1];
    }
    // chroma: lane -> plane (lane >> 4), row (l >> 1), 4 samples
    if (PART == PART_ALL || PART == PART_CHROMA) {
        int plane = lane >> 4, l = lane & 15, cy = l >> 1, cx0 = (l & 1) * 4;
        uint8_t* cbase = (plane ? pd.v : pd.u) + (size_t)(mby * 8 + cy) * g.pitch_c + mbx * 8 + cx0;
        uint32_t packed = 0;
        if (COMPAT) {
            if (coded) {
                const uint2 r = *reinterpret_cast<const uint2*>(&w.res[256 + plane * 64 + cy * 8 + cx0]);
                packed = prmt(r.x, r.y, 0x6420);
            }
        } else {
            // ITU-T H.264 8.4.2.2.2.  The lane's four samples are two pairs; a pair lies in one 4x4 luma block's 2x2 chroma
            // block and shares its vector, so it needs 3 reference samples from each of 2 rows: per row one unaligned word
            // (two aligned loads + funnel).  The pair's position is clamped as a whole, which the replicated border makes
            // equal to the per-sample clamp of the coordinates.
            const int cw = g.w_mbs * 8, ch = g.h_mbs * 8;
#pragma unroll
            for (int k2 = 0; k2 < 2; k2++) {
                const int cx = cx0 + 2 * k2, blk = (cy >> 1) * 4 + (cx >> 1), qq = (cy >> 2) * 2 + (cx >> 2);
                const int rc = hdr_ref_idx(h, qq);
                const uint8_t* ref = plane ? pd.ref_v[rc] : pd.ref_u[rc];
                // (the vector offset came with the macroblock's prefetch: one memory round trip less in front of the reference loads)
                const uint32_t v2 = pre ? load_mv_pre(pd, h, a, blk, *pre, LEAN ? true : pd.mv_off != nullptr) : load_mv(pd, h, a, blk);
                const int mvx = mv_x(v2), mvy = mv_y(v2);
                const int xf = mvx & 7, yf = mvy & 7;
                const int xi = clip3(-8, cw + 5, mbx * 8 + cx + (mvx >> 3)), yi = clip3(-8, ch + 6, mby * 8 + cy + (mvy >> 3));
                const uint32_t r0 = ref_word(ref, g.pitch_c, xi, yi), r1 = ref_word(ref, g.pitch_c, xi, yi + 1);
                const int w00 = (8 - xf) * (8 - yf), w01 = xf * (8 - yf), w10 = (8 - xf) * yf, w11 = xf * yf;
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int A = (int)((r0 >> (8 * k)) & 255u), B = (int)((r0 >> (8 * k + 8)) & 255u);
                    const int C = (int)((r1 >> (8 * k)) & 255u), D = (int)((r1 >> (8 * k + 8)) & 255u);
                    int pv = (w00 * A + w01 * B + w10 * C + w11 * D + 32) >> 6;
                    if (weighted) pv = weight_pred(pv, pd.chroma_log2_wd, pd.chroma_weight[rc][plane], pd.chroma_offset[rc][plane]);
                    if (coded) pv += w.res[256 + plane * 64 + cy * 8 + cx + k];
                    packed |= (uint32_t)clip1(pv) << (8 * (2 * k2 + k));
                }
            }
        }
        *reinterpret_cast<uint32_t*>(cbase) = packed;
    }
    report_excursions(pd, a, exc);
}

// K1+K3 sequence for one inter macroblock (4x4 transform).  8x8-transform macroblocks take seq_inter_mb
// of h264r_seq.cuh.
template <bool COMPAT, bool TMA, bool LEAN, int PART = PART_ALL, class Exec>
HD void seq_inter_mb_fast(Exec& ex, InterWork& w, TmaState& ts, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, const MbPre& pre)
{
    static_assert(PART == PART_ALL || !COMPAT, "the parts are spec-mode instantiations");
    const bool t8flag = !COMPAT && h.cbp != 0 && (h.flags & H264R_T8x8);        // REF_COMPAT has no 8x8 transform
    if (PART == PART_LUMA4 && t8flag) return;
    if (PART == PART_LUMA8 && !t8flag) return;
    // what "coded" means for the share at hand: any residual / luma residual / chroma residual
    const bool coded = PART == PART_ALL ? h.cbp != 0 : (PART == PART_CHROMA ? (h.cbp >> 4) != 0 : (h.cbp & 15) != 0);
    const bool t8 = PART == PART_CHROMA ? false : (PART == PART_LUMA8 ? coded : t8flag);       // (LUMA8 with cbp luma = 0: nothing to transform)
    PerLane<uint32_t> mv_own;
    if (PART == PART_CHROMA) {
        if (coded) ex([&](int lane) { phase_residual_fast<COMPAT, LEAN, true>(w.res, pd, h, a, lane, nullptr, true, &pre); });
        ex([&](int lane) { phase_inter_compute<COMPAT, TMA, LEAN, PART>(w, ts, pd, g, h, a, lane, coded, 0u, &pre); });
        return;
    }
    ex([&](int lane) {
        if (t8) {
            phase_residual8_rows(w, pd, h, a, lane);
            if (PART == PART_ALL) phase_residual_fast<COMPAT, LEAN, true>(w.res, pd, h, a, lane, nullptr, true, &pre);
        } else if (coded) phase_residual_fast<COMPAT, LEAN, true>(w.res, pd, h, a, lane, (!COMPAT && pd.nzmask) ? w.nzf : nullptr, false, &pre, PART != PART_ALL);
        phase_inter_stage<COMPAT, TMA, LEAN>(w, ts, pd, g, h, a, lane, mv_own(lane), pre);
    });
    if (t8) ex([&](int lane) { phase_residual8_cols(w, lane); });
    ex([&](int lane) {
        if (!COMPAT && pd.nzmask && lane == 0) pd.nzmask[a] = !coded ? 0u : (t8 ? nz_pack8(w.nzf) : nz_pack(w.nzf));
        phase_inter_compute<COMPAT, TMA, LEAN, PART>(w, ts, pd, g, h, a, lane, coded, mv_own(lane), &pre);
    });
    if (COMPAT) {        // border macroblocks replicate their samples into the frame's border (see phase_pad_mb_luma)
        int mbx, mby;
        mb_xy(g, a, mbx, mby);
        if (is_border_mb(g, mbx, mby)) ex([&](int lane) { pad_mb_luma_cold(pd.y, g.pitch_y, g.w_mbs, g.h_mbs, mbx, mby, lane); });
    }
}

}  // namespace h264r
