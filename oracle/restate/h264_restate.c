/*
 * TEST INFRASTRUCTURE -- CPU restatement of the H.264 macroblock-reconstruction path.
 * See h264_restate.h for who may use this and for the pinning status.
 *
 * Each function cites the reference code it restates (paths relative to /root/reference).
 * Plain scalar C; written for clarity, not speed.  Not derived from the CUDA kernels and not
 * used by them: the two implementations only share the buffer formats of include/h264r.h.
 */
#include "h264_restate.h"
#include <stdlib.h>
#include <string.h>

#define COMPAT(p) (((p)->flags & H264R_REF_COMPAT) != 0)

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip1(int v) { return clip3(0, 255, v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* luma4x4BlkIdx <-> raster index inside the MB (h264/staticcharts.h:1378-1383; an involution) */
static const uint8_t kBlkRaster[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};

/* ------------------------------------------------------------------------------------------
 * Residual: dequantisation and inverse transforms
 * ---------------------------------------------------------------------------------------- */

/* LevelScale4x4 with flat scaling lists (h264/residual.cpp:170-183) */
static int level_scale4(int m, int i, int j)
{
    static const uint8_t v[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
    if (i % 2 == 0 && j % 2 == 0) return 16 * v[m][0];
    if (i % 2 == 1 && j % 2 == 1) return 16 * v[m][1];
    return 16 * v[m][2];
}

/* residual::ScalingAndTransform_Residual4x4Blocks (h264/residual.cpp:367-434).
 * c: 4x4 levels in raster order (after the inverse zig-zag); dc_bypass: element (0,0) is already
 * scaled (Intra16x16 / chroma DC path, `mode == 1` at h264/residual.cpp:383). */
static void scale_idct4(const int c[16], int qp, int dc_bypass, int r[16])
{
    int d[16], e[16], f[16], g[16], h[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int ls = level_scale4(qp % 6, i, j);
            if (qp >= 24) d[i * 4 + j] = (c[i * 4 + j] * ls) * (1 << (qp / 6 - 4));
            else d[i * 4 + j] = (c[i * 4 + j] * ls + (1 << (3 - qp / 6))) >> (4 - qp / 6);
        }
    if (dc_bypass) d[0] = c[0];
    for (int i = 0; i < 4; i++) {
        e[i * 4 + 0] = d[i * 4 + 0] + d[i * 4 + 2];
        e[i * 4 + 1] = d[i * 4 + 0] - d[i * 4 + 2];
        e[i * 4 + 2] = (d[i * 4 + 1] >> 1) - d[i * 4 + 3];
        e[i * 4 + 3] = d[i * 4 + 1] + (d[i * 4 + 3] >> 1);
        f[i * 4 + 0] = e[i * 4 + 0] + e[i * 4 + 3];
        f[i * 4 + 1] = e[i * 4 + 1] + e[i * 4 + 2];
        f[i * 4 + 2] = e[i * 4 + 1] - e[i * 4 + 2];
        f[i * 4 + 3] = e[i * 4 + 0] - e[i * 4 + 3];
    }
    for (int j = 0; j < 4; j++) {
        g[0 * 4 + j] = f[0 * 4 + j] + f[2 * 4 + j];
        g[1 * 4 + j] = f[0 * 4 + j] - f[2 * 4 + j];
        g[2 * 4 + j] = (f[1 * 4 + j] >> 1) - f[3 * 4 + j];
        g[3 * 4 + j] = f[1 * 4 + j] + (f[3 * 4 + j] >> 1);
        h[0 * 4 + j] = g[0 * 4 + j] + g[3 * 4 + j];
        h[1 * 4 + j] = g[1 * 4 + j] + g[2 * 4 + j];
        h[2 * 4 + j] = g[1 * 4 + j] - g[2 * 4 + j];
        h[3 * 4 + j] = g[0 * 4 + j] - g[3 * 4 + j];
    }
    for (int k = 0; k < 16; k++) r[k] = (h[k] + 32) >> 6;
}

/* 8x8 scaling + inverse transform, ITU-T H.264 8.5.13 (flat scaling lists).
 * No reference counterpart (the reference cannot parse 8x8 blocks: h264/block.cpp:50-51). */
static int level_scale8(int m, int i, int j)
{
    static const uint8_t v[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
                                    {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
    int k;
    if (i % 4 == 0 && j % 4 == 0) k = 0;
    else if (i % 2 == 1 && j % 2 == 1) k = 1;
    else if (i % 4 == 2 && j % 4 == 2) k = 2;
    else if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) k = 3;
    else if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) k = 4;
    else k = 5;
    return 16 * v[m][k];
}
static void idct8_1d(const int* in, int stride, int* out, int ostride)
{
    int a0 = in[0 * stride] + in[4 * stride];
    int a4 = in[0 * stride] - in[4 * stride];
    int a2 = (in[2 * stride] >> 1) - in[6 * stride];
    int a6 = in[2 * stride] + (in[6 * stride] >> 1);
    int a1 = -in[3 * stride] + in[5 * stride] - in[7 * stride] - (in[7 * stride] >> 1);
    int a3 = in[1 * stride] + in[7 * stride] - in[3 * stride] - (in[3 * stride] >> 1);
    int a5 = -in[1 * stride] + in[7 * stride] + in[5 * stride] + (in[5 * stride] >> 1);
    int a7 = in[3 * stride] + in[5 * stride] + in[1 * stride] + (in[1 * stride] >> 1);
    int b0 = a0 + a6, b2 = a4 + a2, b4 = a4 - a2, b6 = a0 - a6;
    int b1 = a1 + (a7 >> 2), b7 = -(a1 >> 2) + a7, b3 = a3 + (a5 >> 2), b5 = (a3 >> 2) - a5;
    out[0 * ostride] = b0 + b7; out[1 * ostride] = b2 + b5; out[2 * ostride] = b4 + b3; out[3 * ostride] = b6 + b1;
    out[4 * ostride] = b6 - b1; out[5 * ostride] = b4 - b3; out[6 * ostride] = b2 - b5; out[7 * ostride] = b0 - b7;
}
static void scale_idct8(const int c[64], int qp, int r[64])
{
    int d[64], g[64], m[64];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            int ls = level_scale8(qp % 6, i, j);
            if (qp >= 36) d[i * 8 + j] = (c[i * 8 + j] * ls) * (1 << (qp / 6 - 6));
            else d[i * 8 + j] = (c[i * 8 + j] * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
        }
    for (int i = 0; i < 8; i++) idct8_1d(d + i * 8, 1, g + i * 8, 1);
    for (int j = 0; j < 8; j++) idct8_1d(g + j, 8, m + j, 8);
    for (int k = 0; k < 64; k++) r[k] = (m[k] + 32) >> 6;
}

/* coefficient block b (0..23) of macroblock a as ints */
static void load_block(const h264o_pic* p, int a, int b, int c[16])
{
    const int16_t* src = p->coeff + (size_t)a * H264R_COEFF_PER_MB + b * 16;
    for (int k = 0; k < 16; k++) c[k] = src[k];
}

/* residual::Decode_Intra16x16 DC part (h264/residual.cpp:286-319): the 16 DC levels sit in slot 0
 * of their blocks; DC matrix element (r,c) belongs to the block at raster position (r,c). */
static void luma_dc16(const h264o_pic* p, int a, int qp, int dc[16])
{
    static const int T[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
    int c[4][4], t[4][4], f[4][4];
    for (int ras = 0; ras < 16; ras++)
        c[ras / 4][ras % 4] = p->coeff[(size_t)a * H264R_COEFF_PER_MB + kBlkRaster[ras] * 16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int s = 0;
            for (int k = 0; k < 4; k++) s += T[i][k] * c[k][j];
            t[i][j] = s;
        }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            int s = 0;
            for (int k = 0; k < 4; k++) s += t[i][k] * T[k][j];
            f[i][j] = s;
        }
    int ls = level_scale4(qp % 6, 0, 0);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            if (qp >= 36) dc[i * 4 + j] = (f[i][j] * ls) * (1 << (qp / 6 - 6));
            else dc[i * 4 + j] = (f[i][j] * ls + (1 << (5 - qp / 6))) >> (6 - qp / 6);
        }
}

/* residual::Decode_Chroma DC part (h264/residual.cpp:202-219).  REF_COMPAT reproduces the
 * reference's non-accumulating matrix product (h264/matrix.cpp:150-171): T*c*T collapses to
 * c[1][1] * [[1,-1],[-1,1]]. */
static void chroma_dc(const h264o_pic* p, int a, int plane, int qp, int dc[4])
{
    int c[4], f[4];
    for (int i = 0; i < 4; i++) c[i] = p->coeff[(size_t)a * H264R_COEFF_PER_MB + (16 + plane * 4 + i) * 16];
    if (COMPAT(p)) {
        f[0] = c[3]; f[1] = -c[3]; f[2] = -c[3]; f[3] = c[3];
    } else {
        f[0] = c[0] + c[1] + c[2] + c[3];
        f[1] = c[0] - c[1] + c[2] - c[3];
        f[2] = c[0] + c[1] - c[2] - c[3];
        f[3] = c[0] - c[1] - c[2] + c[3];
    }
    int ls = level_scale4(qp % 6, 0, 0);
    for (int i = 0; i < 4; i++) dc[i] = ((f[i] * ls) * (1 << (qp / 6))) >> 5;
}

void h264o_residual_mb(const h264o_pic* p, int a, int32_t* out_y, int32_t* out_u, int32_t* out_v)
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int c[64], r[64];
    if (h->mb_class == H264R_MB_I16x16) {
        int dc[16];
        luma_dc16(p, a, h->qp_y, dc);
        for (int ras = 0; ras < 16; ras++) {
            load_block(p, a, kBlkRaster[ras], c);
            c[0] = dc[ras];
            scale_idct4(c, h->qp_y, 1, r);
            for (int k = 0; k < 16; k++) out_y[(ras / 4 * 4 + k / 4) * 16 + ras % 4 * 4 + k % 4] = r[k];
        }
    } else if (h->flags & H264R_T8x8) {
        for (int q = 0; q < 4; q++) {
            const int16_t* src = p->coeff + (size_t)a * H264R_COEFF_PER_MB + q * 64;
            for (int k = 0; k < 64; k++) c[k] = src[k];
            scale_idct8(c, h->qp_y, r);
            for (int k = 0; k < 64; k++) out_y[(q / 2 * 8 + k / 8) * 16 + q % 2 * 8 + k % 8] = r[k];
        }
    } else {
        /* residual::Decode_Intra4x4, also used for every inter MB (h264/residual.cpp:259-284, 197) */
        for (int b = 0; b < 16; b++) {
            int ras = kBlkRaster[b];
            load_block(p, a, b, c);
            scale_idct4(c, h->qp_y, 0, r);
            for (int k = 0; k < 16; k++) out_y[(ras / 4 * 4 + k / 4) * 16 + ras % 4 * 4 + k % 4] = r[k];
        }
    }
    for (int plane = 0; plane < 2; plane++) {
        int qp = plane ? h->qp_cr : h->qp_cb;
        int32_t* out = plane ? out_v : out_u;
        int dc[4];
        chroma_dc(p, a, plane, qp, dc);
        for (int b = 0; b < 4; b++) {
            load_block(p, a, 16 + plane * 4 + b, c);
            c[0] = dc[b];
            scale_idct4(c, qp, 1, r);
            for (int k = 0; k < 16; k++) out[(b / 2 * 4 + k / 4) * 8 + b % 2 * 4 + k % 4] = r[k];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Sample access
 * ---------------------------------------------------------------------------------------- */
static inline int W(const h264o_pic* p) { return p->w_mbs * 16; }
static inline int H(const h264o_pic* p) { return p->h_mbs * 16; }

/* neighbouring luma sample for intra prediction.  The reference reads the unclipped int samples
 * of neighbouring macroblocks (h264/macroblock.cpp:1066-1116); raw_y models that when present. */
static inline int luma_at(const h264o_pic* p, int x, int y)
{
    if (COMPAT(p) && p->raw_y) return p->raw_y[(size_t)y * W(p) + x];
    return p->y[(size_t)y * W(p) + x];
}
static void store_luma(h264o_pic* p, int a, int x, int y, int v)
{
    if (v < 0 || v > 255) {
        p->excursions++;
        if (p->excursion_map) p->excursion_map[a] = 1;
    }
    if (COMPAT(p)) {
        /* macroblock::ConstructPicture has no Clip1 (h264/macroblock.cpp:1467-1481); the plane is
         * what picture::get_SampleXY returns: the low 8 bits (h264/picture.cpp:396-409) */
        if (p->raw_y) p->raw_y[(size_t)y * W(p) + x] = v;
        p->y[(size_t)y * W(p) + x] = (uint8_t)v;
    } else {
        p->y[(size_t)y * W(p) + x] = (uint8_t)clip1(v);
    }
}

/* ------------------------------------------------------------------------------------------
 * Intra prediction
 * ---------------------------------------------------------------------------------------- */

/* Intra4x4 prediction of the block at raster index `ras` (blkIdx = kBlkRaster[ras]).
 * Reference: macroblock::Prediction_Intra4x4 + the nine mode functions and get_4x4Neighbour
 * (h264/macroblock.cpp:973-990, 1037-1365); neighbour availability from picture::get_BLneighbour
 * (h264/picture.cpp:102-183): inside the MB always, outside through the AVAIL_* flags; the
 * up-right samples are replaced by D for blkIdx 3 and 11, for the right column (no MB there yet)
 * and when MB C is unavailable (h264/macroblock.cpp:1092-1110). */
static void intra4x4_pred(const h264o_pic* p, int a, int ras, int mode, int pred[16])
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int mbx = a % p->w_mbs, mby = a / p->w_mbs;
    int bx = ras % 4, by = ras / 4, blk = kBlkRaster[ras];
    int x0 = mbx * 16 + bx * 4, y0 = mby * 16 + by * 4;
    int left_ok = bx > 0 || (h->flags & H264R_AVAIL_A);
    int top_ok = by > 0 || (h->flags & H264R_AVAIL_B);
    int tl_ok;
    if (bx > 0 && by > 0) tl_ok = 1;
    else if (bx > 0) tl_ok = (h->flags & H264R_AVAIL_B) != 0;
    else if (by > 0) tl_ok = (h->flags & H264R_AVAIL_A) != 0;
    else tl_ok = (h->flags & H264R_AVAIL_D) != 0;
    int tr_ok;
    if (blk == 3 || blk == 11) tr_ok = 0;
    else if (bx == 3) tr_ok = (by == 0) && (h->flags & H264R_AVAIL_C);
    else if (by > 0) tr_ok = 1;        /* inside the MB: raster order has already produced it */
    else tr_ok = (h->flags & H264R_AVAIL_B) != 0;
    /* e[0..12] = L K J I M A B C D E F G H */
    int e[13];
    for (int k = 0; k < 13; k++) e[k] = 0;
    if (left_ok) for (int k = 0; k < 4; k++) e[3 - k] = luma_at(p, x0 - 1, y0 + k);
    if (tl_ok) e[4] = luma_at(p, x0 - 1, y0 - 1);
    if (top_ok) {
        for (int k = 0; k < 4; k++) e[5 + k] = luma_at(p, x0 + k, y0 - 1);
        if (tr_ok) for (int k = 0; k < 4; k++) e[9 + k] = luma_at(p, x0 + 4 + k, y0 - 1);
        else for (int k = 0; k < 4; k++) e[9 + k] = e[8];
    }
    const int* top = e + 5;            /* top[-1] = M */
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) {
            int v = 0;
            switch (mode) {
            case 0: v = top[x]; break;
            case 1: v = e[3 - y]; break;
            case 2:
                if (left_ok && top_ok) v = (top[0] + top[1] + top[2] + top[3] + e[0] + e[1] + e[2] + e[3] + 4) >> 3;
                else if (left_ok) v = (e[0] + e[1] + e[2] + e[3] + 2) >> 2;
                else if (top_ok) v = (top[0] + top[1] + top[2] + top[3] + 2) >> 2;
                else v = 128;
                break;
            case 3:
                if (x == 3 && y == 3) v = (top[6] + 3 * top[7] + 2) >> 2;
                else v = (top[x + y] + 2 * top[x + y + 1] + top[x + y + 2] + 2) >> 2;
                break;
            case 4: v = (e[3 - y + x] + 2 * e[4 - y + x] + e[5 - y + x] + 2) >> 2; break;
            case 5: {
                int z = 2 * x - y, i = x - (y >> 1);
                if (z >= 0 && (z & 1) == 0) v = (top[i - 1] + top[i] + 1) >> 1;
                else if (z >= 0) v = (top[i - 2] + 2 * top[i - 1] + top[i] + 2) >> 2;
                else if (z == -1) v = (e[3] + 2 * e[4] + e[5] + 2) >> 2;
                else v = (e[4 - y] + 2 * e[5 - y] + e[6 - y] + 2) >> 2;   /* left[y-1], left[y-2], left[y-3]; left[-1] = M */
                break;
            }
            case 6: {
                int z = 2 * y - x, i = y - (x >> 1);                       /* left[k] = e[3-k], left[-1] = M = e[4] */
                if (z >= 0 && (z & 1) == 0) v = (e[4 - i] + e[3 - i] + 1) >> 1;
                else if (z >= 0) v = (e[5 - i] + 2 * e[4 - i] + e[3 - i] + 2) >> 2;
                else if (z == -1) v = (e[3] + 2 * e[4] + e[5] + 2) >> 2;
                else v = (top[x - 1] + 2 * top[x - 2] + top[x - 3] + 2) >> 2;
                break;
            }
            case 7: {
                int i = x + (y >> 1);
                if ((y & 1) == 0) v = (top[i] + top[i + 1] + 1) >> 1;
                else v = (top[i] + 2 * top[i + 1] + top[i + 2] + 2) >> 2;
                break;
            }
            case 8: {
                int z = x + 2 * y, i = y + (x >> 1);
                if (z > 5) v = e[0];
                else if (z == 5) v = (e[1] + 3 * e[0] + 2) >> 2;
                else if ((z & 1) == 0) v = (e[3 - i] + e[2 - i] + 1) >> 1;
                else v = (e[3 - i] + 2 * e[2 - i] + e[1 - i] + 2) >> 2;
                break;
            }
            default: break;
            }
            pred[y * 4 + x] = v;
        }
}

/* macroblock::Prediction_Intra16x16 and its four modes (h264/macroblock.cpp:1373-1441). */
static void intra16x16_pred(const h264o_pic* p, int a, int mode, int pred[256])
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int mbx = a % p->w_mbs, mby = a / p->w_mbs;
    int x0 = mbx * 16, y0 = mby * 16;
    int A = (h->flags & H264R_AVAIL_A) != 0, B = (h->flags & H264R_AVAIL_B) != 0, D = (h->flags & H264R_AVAIL_D) != 0;
    memset(pred, 0, 256 * sizeof(int));
    if (mode == 0) {
        if (B) for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) pred[y * 16 + x] = luma_at(p, x0 + x, y0 - 1);
    } else if (mode == 1) {
        if (A) for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) pred[y * 16 + x] = luma_at(p, x0 - 1, y0 + y);
    } else if (mode == 2) {
        int sl = 0, st = 0, v;
        if (A) for (int k = 0; k < 16; k++) sl += luma_at(p, x0 - 1, y0 + k);
        if (B) for (int k = 0; k < 16; k++) st += luma_at(p, x0 + k, y0 - 1);
        if (A && B) v = (sl + st + 16) >> 5;
        else if (A) v = (sl + 8) >> 4;
        else if (B) {
            if (COMPAT(p)) {
                /* the top-only case sums column 15 of the MB above (h264/macroblock.cpp:1410-1411) */
                int sc = 0;
                for (int k = 0; k < 16; k++) sc += luma_at(p, x0 + 15, y0 - 16 + k);
                v = (sc + 8) >> 4;
            } else v = (st + 8) >> 4;
        } else v = 128;
        for (int k = 0; k < 256; k++) pred[k] = v;
    } else {
        int hs[17], vs[17];
        memset(hs, 0, sizeof(hs)); memset(vs, 0, sizeof(vs));
        if (D) hs[0] = vs[0] = luma_at(p, x0 - 1, y0 - 1);
        if (A) for (int k = 0; k < 16; k++) vs[k + 1] = luma_at(p, x0 - 1, y0 + k);
        if (B) for (int k = 0; k < 16; k++) hs[k + 1] = luma_at(p, x0 + k, y0 - 1);
        int Hs = 0, Vs = 0;
        for (int i = 9; i < 17; i++) { Hs += (i - 8) * (hs[i] - hs[16 - i]); Vs += (i - 8) * (vs[i] - vs[16 - i]); }
        int aa = 16 * (hs[16] + vs[16]), bb = (5 * Hs + 32) >> 6, cc = (5 * Vs + 32) >> 6;
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++) pred[y * 16 + x] = clip1((aa + bb * (x - 7) + cc * (y - 7) + 16) >> 5);
    }
}

/* Intra8x8 (ITU-T H.264 8.3.2) with reference-sample filtering; no reference counterpart
 * (macroblock::Prediction_Intra8x8 is empty, h264/macroblock.cpp:1369-1372). */
static void intra8x8_pred(const h264o_pic* p, int a, int q, int mode, int pred[64])
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int mbx = a % p->w_mbs, mby = a / p->w_mbs;
    int bx = q % 2, by = q / 2;
    int x0 = mbx * 16 + bx * 8, y0 = mby * 16 + by * 8;
    int left_ok = bx > 0 || (h->flags & H264R_AVAIL_A);
    int top_ok = by > 0 || (h->flags & H264R_AVAIL_B);
    int tl_ok, tr_ok;
    if (bx > 0 && by > 0) tl_ok = 1;
    else if (bx > 0) tl_ok = (h->flags & H264R_AVAIL_B) != 0;
    else if (by > 0) tl_ok = (h->flags & H264R_AVAIL_A) != 0;
    else tl_ok = (h->flags & H264R_AVAIL_D) != 0;
    if (q == 0) tr_ok = (h->flags & H264R_AVAIL_B) != 0;
    else if (q == 1) tr_ok = (h->flags & H264R_AVAIL_C) != 0;
    else if (q == 2) tr_ok = 1;
    else tr_ok = 0;
    /* unfiltered: t[-1..15] top row (t[-1] = corner), l[0..7] left column */
    int tu[17], lu[8], tf[17], lf[8];
    int* t = tu + 1;
    int* f = tf + 1;
    for (int k = 0; k < 17; k++) tu[k] = tf[k] = 0;
    for (int k = 0; k < 8; k++) lu[k] = lf[k] = 0;
    if (tl_ok) t[-1] = p->y[(size_t)(y0 - 1) * W(p) + x0 - 1];
    if (top_ok) {
        for (int k = 0; k < 8; k++) t[k] = p->y[(size_t)(y0 - 1) * W(p) + x0 + k];
        if (tr_ok) for (int k = 8; k < 16; k++) t[k] = p->y[(size_t)(y0 - 1) * W(p) + x0 + k];
        else for (int k = 8; k < 16; k++) t[k] = t[7];
    }
    if (left_ok) for (int k = 0; k < 8; k++) lu[k] = p->y[(size_t)(y0 + k) * W(p) + x0 - 1];
    /* 8.3.2.2.1 reference sample filtering */
    if (top_ok) {
        f[0] = tl_ok ? (t[-1] + 2 * t[0] + t[1] + 2) >> 2 : (3 * t[0] + t[1] + 2) >> 2;
        for (int k = 1; k < 15; k++) f[k] = (t[k - 1] + 2 * t[k] + t[k + 1] + 2) >> 2;
        f[15] = (t[14] + 3 * t[15] + 2) >> 2;
    }
    if (tl_ok) {
        if (!top_ok && left_ok) f[-1] = (3 * t[-1] + lu[0] + 2) >> 2;
        else if (top_ok && !left_ok) f[-1] = (3 * t[-1] + t[0] + 2) >> 2;
        else if (top_ok && left_ok) f[-1] = (t[0] + 2 * t[-1] + lu[0] + 2) >> 2;
        else f[-1] = t[-1];
    }
    if (left_ok) {
        lf[0] = tl_ok ? (t[-1] + 2 * lu[0] + lu[1] + 2) >> 2 : (3 * lu[0] + lu[1] + 2) >> 2;
        for (int k = 1; k < 7; k++) lf[k] = (lu[k - 1] + 2 * lu[k] + lu[k + 1] + 2) >> 2;
        lf[7] = (lu[6] + 3 * lu[7] + 2) >> 2;
    }
    /* e[0..24]: e[7-k] = left[k], e[8] = corner, e[9+k] = top[k] */
    int e[25];
    for (int k = 0; k < 8; k++) e[7 - k] = lf[k];
    e[8] = f[-1];
    for (int k = 0; k < 16; k++) e[9 + k] = f[k];
    const int* T = e + 9;              /* T[-1] = corner */
#define L(k) e[7 - (k)]                 /* L(-1) = corner */
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            int v = 0;
            switch (mode) {
            case 0: v = T[x]; break;
            case 1: v = L(y); break;
            case 2: {
                int s = 0;
                if (left_ok && top_ok) { for (int k = 0; k < 8; k++) s += T[k] + L(k); v = (s + 8) >> 4; }
                else if (left_ok) { for (int k = 0; k < 8; k++) s += L(k); v = (s + 4) >> 3; }
                else if (top_ok) { for (int k = 0; k < 8; k++) s += T[k]; v = (s + 4) >> 3; }
                else v = 128;
                break;
            }
            case 3:
                if (x == 7 && y == 7) v = (T[14] + 3 * T[15] + 2) >> 2;
                else v = (T[x + y] + 2 * T[x + y + 1] + T[x + y + 2] + 2) >> 2;
                break;
            case 4: v = (e[7 - y + x] + 2 * e[8 - y + x] + e[9 - y + x] + 2) >> 2; break;
            case 5: {
                int z = 2 * x - y, i = x - (y >> 1);
                if (z >= 0 && (z & 1) == 0) v = (T[i - 1] + T[i] + 1) >> 1;
                else if (z >= 0) v = (T[i - 2] + 2 * T[i - 1] + T[i] + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T[-1] + T[0] + 2) >> 2;
                else v = (L(y - 2 * x - 1) + 2 * L(y - 2 * x - 2) + L(y - 2 * x - 3) + 2) >> 2;
                break;
            }
            case 6: {
                int z = 2 * y - x, i = y - (x >> 1);
                if (z >= 0 && (z & 1) == 0) v = (L(i - 1) + L(i) + 1) >> 1;
                else if (z >= 0) v = (L(i - 2) + 2 * L(i - 1) + L(i) + 2) >> 2;
                else if (z == -1) v = (L(0) + 2 * T[-1] + T[0] + 2) >> 2;
                else v = (T[x - 2 * y - 1] + 2 * T[x - 2 * y - 2] + T[x - 2 * y - 3] + 2) >> 2;
                break;
            }
            case 7: {
                int i = x + (y >> 1);
                if ((y & 1) == 0) v = (T[i] + T[i + 1] + 1) >> 1;
                else v = (T[i] + 2 * T[i + 1] + T[i + 2] + 2) >> 2;
                break;
            }
            case 8: {
                int z = x + 2 * y, i = y + (x >> 1);
                if (z > 13) v = L(7);
                else if (z == 13) v = (L(6) + 3 * L(7) + 2) >> 2;
                else if ((z & 1) == 0) v = (L(i) + L(i + 1) + 1) >> 1;
                else v = (L(i) + 2 * L(i + 1) + L(i + 2) + 2) >> 2;
                break;
            }
            default: break;
            }
            pred[y * 8 + x] = v;
        }
#undef L
}

/* Intra chroma prediction, ITU-T H.264 8.3.4 (4:2:0).  No reference counterpart: the reference
 * parses intra_chroma_pred_mode and never uses it (h264/macroblock.cpp:340-341). */
static void intra_chroma_pred(const h264o_pic* p, int a, const uint8_t* plane, int mode, int pred[64])
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int mbx = a % p->w_mbs, mby = a / p->w_mbs;
    int cw = p->w_mbs * 8;
    int x0 = mbx * 8, y0 = mby * 8;
    int A = (h->flags & H264R_AVAIL_A) != 0, B = (h->flags & H264R_AVAIL_B) != 0;
    int top[8], left[8], tl = 0;
    for (int k = 0; k < 8; k++) {
        top[k] = B ? plane[(size_t)(y0 - 1) * cw + x0 + k] : 0;
        left[k] = A ? plane[(size_t)(y0 + k) * cw + x0 - 1] : 0;
    }
    if (A && B) tl = plane[(size_t)(y0 - 1) * cw + x0 - 1];
    if (mode == 0) {
        for (int b = 0; b < 4; b++) {
            int xo = (b & 1) * 4, yo = (b >> 1) * 4, st = 0, sl = 0, v;
            for (int k = 0; k < 4; k++) { st += top[xo + k]; sl += left[yo + k]; }
            if (b == 0 || b == 3) {
                if (A && B) v = (st + sl + 4) >> 3;
                else if (A) v = (sl + 2) >> 2;
                else if (B) v = (st + 2) >> 2;
                else v = 128;
            } else if (b == 1) {
                if (B) v = (st + 2) >> 2;
                else if (A) v = (sl + 2) >> 2;
                else v = 128;
            } else {
                if (A) v = (sl + 2) >> 2;
                else if (B) v = (st + 2) >> 2;
                else v = 128;
            }
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) pred[(yo + y) * 8 + xo + x] = v;
        }
    } else if (mode == 1) {
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pred[y * 8 + x] = left[y];
    } else if (mode == 2) {
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) pred[y * 8 + x] = top[x];
    } else {
        int Hs = 0, Vs = 0;
        for (int k = 0; k < 4; k++) {
            Hs += (k + 1) * (top[4 + k] - (k == 3 ? tl : top[2 - k]));
            Vs += (k + 1) * (left[4 + k] - (k == 3 ? tl : left[2 - k]));
        }
        int aa = 16 * (left[7] + top[7]), bb = (34 * Hs + 32) >> 6, cc = (34 * Vs + 32) >> 6;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) pred[y * 8 + x] = clip1((aa + bb * (x - 3) + cc * (y - 3) + 16) >> 5);
    }
}

/* ------------------------------------------------------------------------------------------
 * Inter prediction
 * ---------------------------------------------------------------------------------------- */
static inline int ref_luma(const h264o_pic* p, const uint8_t* ref, int x, int y)
{
    /* coordinate clamp of macroblock::Prediction_Inter_LumaSampleInterpolation
     * (h264/macroblock.cpp:828-830); samples are what picture::get_SampleXY returns (uint8_t) */
    return ref[(size_t)clip3(0, H(p) - 1, y) * W(p) + clip3(0, W(p) - 1, x)];
}
static inline int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

/* one luma sample at integer position (xi, yi) + fraction (xf, yf).
 * REF_COMPAT: the reference's tap coordinates, including its deviations D6/D7
 * (h264/staticcharts.h:1492-1517, h264/macroblock.cpp:837-901).  Otherwise ITU-T H.264 8.4.2.2.1. */
static int luma_interp(const h264o_pic* p, const uint8_t* ref, int xi, int yi, int xf, int yf)
{
#define S(dx, dy) ref_luma(p, ref, xi + (dx), yi + (dy))
    int G = S(0, 0);
    if (xf == 0 && yf == 0) return G;
    int b1, h1, s1, m1, aa, bb, gg, hh;
    int Hh = S(1, 0), M = S(0, 1);
    b1 = tap6(S(-2, 0), S(-1, 0), G, Hh, S(2, 0), S(3, 0));
    h1 = tap6(S(0, -2), S(0, -1), G, M, S(0, 2), S(0, 3));
    m1 = tap6(S(1, -2), S(1, -1), Hh, S(1, 1), S(1, 2), S(1, 3));
    aa = tap6(S(-2, -2), S(-1, -2), S(0, -2), S(1, -2), S(2, -2), S(3, -2));
    if (COMPAT(p)) {
        s1 = tap6(S(-2, -1), S(-1, 1), M, S(1, 1), S(2, 1), S(3, 1));
        bb = tap6(S(-2, -2), S(-1, -2), S(0, -1), S(1, -1), S(2, -2), S(3, -2));
        gg = tap6(S(-2, 2), S(-1, 2), S(0, 2), S(1, 2), S(2, -2), S(3, -2));
        hh = tap6(S(-2, 3), S(-1, 3), S(0, 3), S(1, 3), S(2, -3), S(3, -3));
    } else {
        s1 = tap6(S(-2, 1), S(-1, 1), M, S(1, 1), S(2, 1), S(3, 1));
        bb = tap6(S(-2, -1), S(-1, -1), S(0, -1), S(1, -1), S(2, -1), S(3, -1));
        gg = tap6(S(-2, 2), S(-1, 2), S(0, 2), S(1, 2), S(2, 2), S(3, 2));
        hh = tap6(S(-2, 3), S(-1, 3), S(0, 3), S(1, 3), S(2, 3), S(3, 3));
    }
#undef S
    int b = clip1((b1 + 16) >> 5), h = clip1((h1 + 16) >> 5);
    int j = clip1((tap6(aa, bb, b1, s1, gg, hh) + 512) >> 10);
    int s = clip1((s1 + 16) >> 5), m = clip1((m1 + 16) >> 5);
    switch (xf * 4 + yf) {
    case 1: return (G + h + 1) >> 1;       /* d */
    case 2: return h;
    case 3: return (M + h + 1) >> 1;       /* n */
    case 4: return (G + b + 1) >> 1;       /* a */
    case 5: return (b + h + 1) >> 1;       /* e */
    case 6: return (h + j + 1) >> 1;       /* i */
    case 7: return (h + s + 1) >> 1;       /* p */
    case 8: return b;
    case 9: return (b + j + 1) >> 1;       /* f */
    case 10: return j;
    case 11: return (j + s + 1) >> 1;      /* q */
    case 12: return (Hh + b + 1) >> 1;     /* c */
    case 13: return (b + m + 1) >> 1;      /* g */
    case 14: return (j + m + 1) >> 1;      /* k */
    default: return (m + s + 1) >> 1;      /* r */
    }
}

/* chroma sample, ITU-T H.264 8.4.2.2.2; no reference counterpart */
static int chroma_interp(const h264o_pic* p, const uint8_t* ref, int xi, int yi, int xf, int yf)
{
    int cw = p->w_mbs * 8, ch = p->h_mbs * 8;
#define C(dx, dy) ref[(size_t)clip3(0, ch - 1, yi + (dy)) * cw + clip3(0, cw - 1, xi + (dx))]
    int A = C(0, 0), B = C(1, 0), Cc = C(0, 1), D = C(1, 1);
#undef C
    return ((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * Cc + xf * yf * D + 32) >> 6;
}

static int weight(int v, int logwd, int w, int o)
{
    /* macroblock::Weight_CoefficWeight, single-list case (h264/macroblock.cpp:544-553) = 8.4.2.3 */
    if (logwd >= 1) return clip1(((v * w + (1 << (logwd - 1))) >> logwd) + o);
    return clip1(v * w + o);
}

/* ------------------------------------------------------------------------------------------
 * Macroblock reconstruction
 * ---------------------------------------------------------------------------------------- */
static int is_intra(const h264r_mb_hdr* h) { return h->mb_class <= H264R_MB_I16x16; }

void h264o_recon_mb(h264o_pic* p, int a)
{
    const h264r_mb_hdr* h = &p->hdr[a];
    int mbx = a % p->w_mbs, mby = a / p->w_mbs;
    int res_y[256], res_u[64], res_v[64];
    int pred[256];
    int cw = p->w_mbs * 8;
    h264o_residual_mb(p, a, res_y, res_u, res_v);

    if (h->mb_class == H264R_MB_I16x16) {
        intra16x16_pred(p, a, h->intra_modes & 3, pred);
        for (int y = 0; y < 16; y++)
            for (int x = 0; x < 16; x++) store_luma(p, a, mbx * 16 + x, mby * 16 + y, pred[y * 16 + x] + res_y[y * 16 + x]);
    } else if (h->mb_class == H264R_MB_I4x4) {
        /* macroblock::Calc predicts the 16 blocks in raster order (h264/macroblock.cpp:28-33) */
        for (int ras = 0; ras < 16; ras++) {
            int blk = kBlkRaster[ras];
            int mode = (h->u.pred4x4[blk >> 1] >> ((blk & 1) * 4)) & 15;
            intra4x4_pred(p, a, ras, mode, pred);
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) {
                    int px = ras % 4 * 4 + x, py = ras / 4 * 4 + y;
                    store_luma(p, a, mbx * 16 + px, mby * 16 + py, pred[y * 4 + x] + res_y[py * 16 + px]);
                }
        }
    } else if (h->mb_class == H264R_MB_I8x8) {
        for (int q = 0; q < 4; q++) {
            int mode = (h->u.pred4x4[q >> 1] >> ((q & 1) * 4)) & 15;
            intra8x8_pred(p, a, q, mode, pred);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    int px = q % 2 * 8 + x, py = q / 2 * 8 + y;
                    store_luma(p, a, mbx * 16 + px, mby * 16 + py, pred[y * 8 + x] + res_y[py * 16 + px]);
                }
        }
    } else {
        /* inter: macroblock::Decode + Prediction_Inter (h264/macroblock.cpp:591-813), evaluated per
         * 4x4 block: every sample of a partition is interpolated independently from the partition's
         * vector, so the partition shape only matters through the reference's sub-partition source
         * offset (SURVEY D8, h264/macroblock.cpp:786-789): for 4x4 sub-partitions 2 and 3 the
         * source row offset is 1 instead of 4. */
        const int16_t* mv = p->mv + (size_t)a * H264R_MV_PER_MB;
        for (int ras = 0; ras < 16; ras++) {
            int bx = ras % 4, by = ras / 4, q = (by >> 1) * 2 + (bx >> 1);
            int ri = h->u.ref_idx[q];
            const uint8_t* ref = p->ref_y[ri];
            int mvx = mv[ras * 2], mvy = mv[ras * 2 + 1];
            int x0 = mbx * 16 + bx * 4, y0 = mby * 16 + by * 4;
            int ys = y0;
            if (COMPAT(p) && h->mb_class == H264R_MB_P8x8 && ((h->sub_mb_type >> (2 * q)) & 3) == H264R_SUB_4x4 && (by & 1))
                ys = y0 - 3;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) {
                    int v = luma_interp(p, ref, x0 + x + (mvx >> 2), ys + y + (mvy >> 2), mvx & 3, mvy & 3);
                    if (p->pp->weighted_pred) v = weight(v, p->pp->luma_log2_weight_denom, p->pp->luma_weight[ri], p->pp->luma_offset[ri]);
                    store_luma(p, a, x0 + x, y0 + y, v + res_y[(by * 4 + y) * 16 + bx * 4 + x]);
                }
        }
    }

    /* chroma */
    for (int plane = 0; plane < 2; plane++) {
        uint8_t* dst = plane ? p->v : p->u;
        const int* res = plane ? res_v : res_u;
        int cp[64];
        if (COMPAT(p)) {
            /* pred_U/pred_V are never written, sample = residual (h264/macroblock.cpp:1477-1480) */
            memset(cp, 0, sizeof(cp));
        } else if (is_intra(h)) {
            intra_chroma_pred(p, a, dst, (h->intra_modes >> 2) & 3, cp);
        } else {
            const int16_t* mv = p->mv + (size_t)a * H264R_MV_PER_MB;
            for (int ras = 0; ras < 16; ras++) {
                int bx = ras % 4, by = ras / 4, q = (by >> 1) * 2 + (bx >> 1);
                int ri = h->u.ref_idx[q];
                const uint8_t* ref = plane ? p->ref_v[ri] : p->ref_u[ri];
                int mvx = mv[ras * 2], mvy = mv[ras * 2 + 1];
                for (int y = 0; y < 2; y++)
                    for (int x = 0; x < 2; x++) {
                        int cx = bx * 2 + x, cy = by * 2 + y;
                        int v = chroma_interp(p, ref, mbx * 8 + cx + (mvx >> 3), mby * 8 + cy + (mvy >> 3), mvx & 7, mvy & 7);
                        if (p->pp->weighted_pred)
                            v = weight(v, p->pp->chroma_log2_weight_denom, p->pp->chroma_weight[ri][plane], p->pp->chroma_offset[ri][plane]);
                        cp[cy * 8 + cx] = v;
                    }
            }
        }
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++) {
                int v = cp[y * 8 + x] + res[y * 8 + x];
                dst[(size_t)(mby * 8 + y) * cw + mbx * 8 + x] = COMPAT(p) ? (uint8_t)v : (uint8_t)clip1(v);
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * Deblocking filter, ITU-T H.264 8.7 (frame pictures, one slice, 4:2:0).  No reference counterpart.
 * ---------------------------------------------------------------------------------------- */
static const uint8_t kAlpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13,
                                   15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144,
                                   162, 182, 203, 226, 255, 255};
static const uint8_t kBeta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                  6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16,
                                  17, 17, 18, 18};
static const uint8_t kTc0[52][3] = {
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1},
    {0, 0, 1}, {0, 1, 1}, {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2},
    {1, 1, 2}, {1, 2, 3}, {1, 2, 3}, {2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6},
    {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10}, {6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18},
    {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};

/* non-zero transform coefficients in the 4x4 luma block at raster position ras of MB a
 * (for 8x8-transform MBs: in the 8x8 block containing it) */
static int blk_coded(const h264o_pic* p, int a, int ras)
{
    const h264r_mb_hdr* h = &p->hdr[a];
    const int16_t* c = p->coeff + (size_t)a * H264R_COEFF_PER_MB;
    if (h->flags & H264R_T8x8) {
        int q = (ras / 8) * 2 + (ras % 4) / 2;
        for (int k = 0; k < 64; k++) if (c[q * 64 + k]) return 1;
        return 0;
    }
    int blk = kBlkRaster[ras];
    for (int k = 0; k < 16; k++) if (c[blk * 16 + k]) return 1;
    return 0;
}

static int bs_for(const h264o_pic* p, int ap, int rasp, int aq, int rasq, int mb_edge)
{
    const h264r_mb_hdr* hp = &p->hdr[ap];
    const h264r_mb_hdr* hq = &p->hdr[aq];
    if (is_intra(hp) || is_intra(hq)) return mb_edge ? 4 : 3;
    if (blk_coded(p, ap, rasp) || blk_coded(p, aq, rasq)) return 2;
    int qp_ = (rasp / 8) * 2 + (rasp % 4) / 2, qq = (rasq / 8) * 2 + (rasq % 4) / 2;
    int refp = p->ref_id[hp->u.ref_idx[qp_]], refq = p->ref_id[hq->u.ref_idx[qq]];
    if (refp != refq) return 1;
    const int16_t* mvp = p->mv + (size_t)ap * H264R_MV_PER_MB + rasp * 2;
    const int16_t* mvq = p->mv + (size_t)aq * H264R_MV_PER_MB + rasq * 2;
    if (iabs(mvp[0] - mvq[0]) >= 4 || iabs(mvp[1] - mvq[1]) >= 4) return 1;
    return 0;
}

/* filters one line of samples across an edge: pix points at q0, step is the distance between
 * p_i and p_{i+1} (and q_i, q_{i+1}) */
static void filter_line(uint8_t* pix, int step, int bs, int alpha, int beta, int index_a, int luma)
{
    int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;
    if (luma) {
        int p2 = pix[-3 * step], q2 = pix[2 * step];
        int ap = iabs(p2 - p0), aq = iabs(q2 - q0);
        if (bs < 4) {
            int tc0 = kTc0[index_a][bs - 1];
            int tc = tc0 + (ap < beta) + (aq < beta);
            int delta = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
            pix[-step] = (uint8_t)clip1(p0 + delta);
            pix[0] = (uint8_t)clip1(q0 - delta);
            if (ap < beta) pix[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 * 2)) >> 1));
            if (aq < beta) pix[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 * 2)) >> 1));
        } else {
            int small = iabs(p0 - q0) < ((alpha >> 2) + 2);
            if (ap < beta && small) {
                int p3 = pix[-4 * step];
                pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
                pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
                pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
            } else pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
            if (aq < beta && small) {
                int q3 = pix[3 * step];
                pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
                pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
                pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
            } else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    } else {
        if (bs < 4) {
            int tc = kTc0[index_a][bs - 1] + 1;
            int delta = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
            pix[-step] = (uint8_t)clip1(p0 + delta);
            pix[0] = (uint8_t)clip1(q0 - delta);
        } else {
            pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
            pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
        }
    }
}

void h264o_deblock_picture(h264o_pic* p)
{
    const h264r_pic_params* pp = p->pp;
    if (pp->disable_deblocking_filter_idc == 1) return;
    int wl = W(p), wc = p->w_mbs * 8;
    int off_a = pp->slice_alpha_c0_offset_div2 * 2, off_b = pp->slice_beta_offset_div2 * 2;
    for (int a = 0; a < p->w_mbs * p->h_mbs; a++) {
        const h264r_mb_hdr* h = &p->hdr[a];
        int mbx = a % p->w_mbs, mby = a / p->w_mbs;
        int t8 = (h->flags & H264R_T8x8) != 0;
        /* bS of the 32 luma edge segments: [dir][edge][k] */
        int bs[2][4][4];
        for (int dir = 0; dir < 2; dir++)
            for (int e = 0; e < 4; e++)
                for (int k = 0; k < 4; k++) {
                    int v = 0;
                    if (e == 0) {
                        int have = dir == 0 ? mbx > 0 : mby > 0;
                        if (have) {
                            int an = dir == 0 ? a - 1 : a - p->w_mbs;
                            int rasq = dir == 0 ? k * 4 : k;
                            int rasp = dir == 0 ? k * 4 + 3 : 12 + k;
                            v = bs_for(p, an, rasp, a, rasq, 1);
                        }
                    } else if (!(t8 && (e & 1))) {
                        int rasq = dir == 0 ? k * 4 + e : e * 4 + k;
                        int rasp = dir == 0 ? rasq - 1 : rasq - 4;
                        v = bs_for(p, a, rasp, a, rasq, 0);
                    }
                    bs[dir][e][k] = v;
                }
        for (int plane = 0; plane < 3; plane++) {
            uint8_t* base = plane == 0 ? p->y : (plane == 1 ? p->u : p->v);
            int pitch = plane == 0 ? wl : wc;
            int size = plane == 0 ? 16 : 8;
            for (int dir = 0; dir < 2; dir++)
                for (int e = 0; e < 4; e++) {
                    if (plane && (e & 1)) continue;            /* chroma: edges at 0 and 4 chroma samples */
                    int qp_q = plane == 0 ? h->qp_y : (plane == 1 ? h->qp_cb : h->qp_cr);
                    int qp_p = qp_q;
                    if (e == 0) {
                        int have = dir == 0 ? mbx > 0 : mby > 0;
                        if (!have) continue;
                        const h264r_mb_hdr* hn = &p->hdr[dir == 0 ? a - 1 : a - p->w_mbs];
                        qp_p = plane == 0 ? hn->qp_y : (plane == 1 ? hn->qp_cb : hn->qp_cr);
                    }
                    int qpav = (qp_p + qp_q + 1) >> 1;
                    int ia = clip3(0, 51, qpav + off_a), ib = clip3(0, 51, qpav + off_b);
                    int alpha = kAlpha[ia], beta = kBeta[ib];
                    int pos = plane == 0 ? e * 4 : e * 2;      /* sample offset of the edge inside the MB */
                    for (int k = 0; k < size; k++) {
                        int b = bs[dir][e][plane == 0 ? k / 4 : k / 2];
                        if (!b) continue;
                        uint8_t* pix;
                        int step;
                        if (dir == 0) { pix = base + (size_t)(mby * size + k) * pitch + mbx * size + pos; step = 1; }
                        else { pix = base + (size_t)(mby * size + pos) * pitch + mbx * size + k; step = pitch; }
                        filter_line(pix, step, b, alpha, beta, ia, plane == 0);
                    }
                }
        }
    }
}

void h264o_recon_picture(h264o_pic* p)
{
    int n = p->w_mbs * p->h_mbs;
    for (int a = 0; a < n; a++) h264o_recon_mb(p, a);
    if (!COMPAT(p)) h264o_deblock_picture(p);
}

/* position-dependent checksum of the three planes (DESIGN.md "K5 frame digest"): four sums modulo 2^32 over
 * the little-endian 32-bit words (4 samples) of Y, then U, then V, each plane row-major without padding (plane
 * widths are multiples of 8), so that it can be accumulated in any order. */
void h264o_digest(const uint8_t* y, const uint8_t* u, const uint8_t* v, int w, int h, uint8_t out[16])
{
    uint32_t d[4] = {0, 0, 0, 0};
    const uint8_t* planes[3] = {y, u, v};
    uint32_t sizes[3] = {(uint32_t)(w * h / 4), (uint32_t)(w * h / 16), (uint32_t)(w * h / 16)};
    uint32_t i = 0;
    for (int pl = 0; pl < 3; pl++)
        for (uint32_t k = 0; k < sizes[pl]; k++, i++) {
            const uint8_t* b = planes[pl] + 4u * k;
            uint32_t s = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
            uint32_t m = i * 0x9E3779B1u + 0x7F4A7C15u;
            d[0] += s;
            d[1] += s * (i + 1u);
            d[2] += s * m;
            d[3] += (s ^ 0xA5A5A5A5u) * (m ^ (m >> 15));
        }
    memcpy(out, d, 16);
}
