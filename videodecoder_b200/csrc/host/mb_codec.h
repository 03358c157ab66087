// mb_codec.h -- the macroblock layer of a CABAC slice (ITU-T H.264 7.3.4, 7.3.5, 7.3.5.1-3; binarisations 9.3.2, context
// index increments 9.3.3.1.1-3), written ONCE against the engine interface of cabac.h and instantiated for both directions:
// with CabacDecoder it fills h264s_mb records from the bitstream, with CabacEncoder it writes them.  Every bin goes through
//      bin = e.decision(ctx, E::kEncoding ? <the bit the record asks for> : 0)
// so what the writer emits is by construction what the reader expects.
//
// Replaces, on the reference's side: the per-syntax-element readers of cabac.cpp (h264/cabac.cpp:274-1145: binarisation tables
// searched linearly per bin, h264/cabac.cpp:240-252; neighbour lookups through picture::get_BLneighbour / get_PartNeighbour per
// bin), macroblock::Parse's syntax order (h264/macroblock.cpp:143-475), residual::Parse (h264/residual.cpp:21-129) and
// residual_block_cabac (h264/residual.cpp:561-614).  Here the neighbour information a context increment needs is one 48-byte
// record per finished macroblock (MbInfo); binarisations are straight-line code.
//
// Scope: I and P slices of progressive frames, 4:2:0; Intra4x4 / Intra8x8 / Intra16x16, P_L0_16x16 / 16x8 / 8x16 / P_8x8
// (P_8x8ref0 does not exist in CABAC), P_Skip, transform_size_8x8_flag.  I_PCM is reported as unsupported.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../../include/h264_syntax.h"
#include "cabac.h"

namespace h264host {

struct SliceParams {
    int w_mbs = 0, h_mbs = 0;
    bool p_slice = false;
    int num_ref_idx_l0 = 1;
    bool transform_8x8_mode = false;
};

// what context selection of later macroblocks reads from a finished one
struct MbInfo {
    uint16_t slice;        // 0 = not decoded yet in this picture; else 1 + index of the slice it belongs to
    uint8_t kind;          // H264S_*
    uint8_t t8x8;
    uint8_t cbp_luma, cbp_chroma;
    uint8_t chroma_mode;   // intra_chroma_pred_mode (0 for inter)
    uint8_t ref_gt0;       // bit q: ref_idx_l0 of 8x8 quadrant q is > 0
    uint16_t cbf_luma;     // bit r = raster index of the 4x4 block: coded_block_flag (all four of a coded 8x8 transform block)
    uint8_t cbf_dc;        // bit 0 Intra16x16 DC, bit 1 Cb DC, bit 2 Cr DC
    uint8_t cbf_chroma;    // bits 0-3 Cb AC blocks, bits 4-7 Cr AC blocks
    uint8_t mvd[16][2];    // |mvd_l0| per 4x4 block (raster), saturated at 127
};

static const uint8_t kBlkRaster[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};     // luma4x4BlkIdx <-> raster (involution)

// Table 9-43, frame coded blocks: ctxIdxInc of significant_coeff_flag / last_significant_coeff_flag for ctxBlockCat 5
static const uint8_t kSig8x8[63] = {
    0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
    7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
static const uint8_t kLast8x8[63] = {
    0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
    3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};

enum { kCatLumaDC = 0, kCatLumaAC = 1, kCatLuma4x4 = 2, kCatChromaDC = 3, kCatChromaAC = 4, kCatLuma8x8 = 5 };

template <class E>
class SliceCodec {
public:
    SliceCodec(E& e, const SliceParams& sp, MbInfo* info, int slice_no) : e_(e), sp_(sp), info_(info), slice_((uint16_t)(slice_no + 1)) {}

    // One macroblock (mb_skip_flag .. residual); end_of_slice_flag is the caller's.  Returns 0 or a negative error.
    int mb(int addr, h264s_mb& m)
    {
        const int mbx = addr % sp_.w_mbs;
        a_ = mbx > 0 && info_[addr - 1].slice == slice_ ? &info_[addr - 1] : nullptr;
        b_ = addr >= sp_.w_mbs && info_[addr - sp_.w_mbs].slice == slice_ ? &info_[addr - sp_.w_mbs] : nullptr;
        MbInfo& c = info_[addr];
        memset(&c, 0, sizeof(c));
        cur_ = &c;
        if (!E::kEncoding) clear_syntax(m);
        if (sp_.p_slice) {
            const int inc = (a_ && a_->kind != H264S_PSKIP) + (b_ && b_->kind != H264S_PSKIP);
            const int skip = e_.decision(11 + inc, m.kind == H264S_PSKIP);
            if (skip) {
                m.kind = H264S_PSKIP;
                c.kind = H264S_PSKIP;
                c.slice = slice_;
                last_dqp_nz_ = false;
                return 0;
            }
        }
        int rc = mb_type(m);
        if (rc) return rc;
        c.kind = m.kind;
        intra_ = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        bool t8 = false;
        if (m.kind == H264S_I4x4 || m.kind == H264S_I8x8) {                   // I_NxN
            if (sp_.transform_8x8_mode) {
                t8 = transform8x8_flag(m.kind == H264S_I8x8);
                m.kind = t8 ? H264S_I8x8 : H264S_I4x4;
            } else if (m.kind == H264S_I8x8) return -1;
            c.kind = m.kind;
            const int nb = t8 ? 4 : 16;
            for (int i = 0; i < nb; i++) {
                m.prev_flag[i] = (uint8_t)e_.decision(68, m.prev_flag[i] != 0);
                if (!m.prev_flag[i]) {
                    int v = 0;
                    for (int k = 0; k < 3; k++) v |= e_.decision(69, (m.rem_mode[i] >> k) & 1) << k;
                    m.rem_mode[i] = (uint8_t)v;
                }
            }
        }
        if (intra_) {
            m.chroma_mode = (uint8_t)chroma_pred_mode(m.chroma_mode);
            c.chroma_mode = m.chroma_mode;
        } else {
            inter_prediction(m);
        }
        if (m.kind != H264S_I16x16) {
            coded_block_pattern(m);
            if (m.cbp_luma && sp_.transform_8x8_mode && !intra_ &&
                (m.kind != H264S_P8x8 || (m.sub_type[0] | m.sub_type[1] | m.sub_type[2] | m.sub_type[3]) == 0))
                t8 = transform8x8_flag(m.t8x8 != 0);
        }
        m.t8x8 = t8;
        c.t8x8 = t8;
        c.cbp_luma = m.cbp_luma;
        c.cbp_chroma = m.cbp_chroma;
        c.slice = slice_;
        if (m.cbp_luma || m.cbp_chroma || m.kind == H264S_I16x16) {
            m.qp_delta = (int8_t)mb_qp_delta(m.qp_delta);
            last_dqp_nz_ = m.qp_delta != 0;
            rc = residual(m);
            if (rc) return rc;
        } else {
            if (E::kEncoding && m.qp_delta != 0) return -1;                   // cannot be expressed
            m.qp_delta = 0;
            last_dqp_nz_ = false;
        }
        return 0;
    }
    int end_of_slice(int last) { return e_.terminate(last); }
    void begin_slice() { last_dqp_nz_ = false; }

private:
    static void clear_syntax(h264s_mb& m)
    {
        memset(&m, 0, sizeof(m));
    }

    // ---- mb_type (Table 9-36 binarisations; 9.3.3.1.1.3, 9.3.3.1.2) ----------------------------------------------------------
    int mb_type(h264s_mb& m)
    {
        const bool want_intra = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        if (sp_.p_slice) {
            if (!e_.decision(14, want_intra)) {
                // P_L0_16x16 "0 0 0", P_8x8 "0 0 1", P_L0_L0_8x16 "0 1 0", P_L0_L0_16x8 "0 1 1"
                const int b1 = e_.decision(15, m.kind == H264S_P16x8 || m.kind == H264S_P8x16);
                const int b2 = e_.decision(b1 ? 17 : 16, b1 ? m.kind == H264S_P16x8 : m.kind == H264S_P8x8);
                m.kind = (uint8_t)(b1 ? (b2 ? H264S_P16x8 : H264S_P8x16) : (b2 ? H264S_P8x8 : H264S_P16x16));
                return 0;
            }
            return intra_mb_type(m, 17, false);
        }
        return intra_mb_type(m, 3, true);
    }
    int intra_mb_type(h264s_mb& m, int base, bool i_slice)
    {
        int ctx0 = base;
        if (i_slice) ctx0 += (a_ && a_->kind != H264S_I4x4 && a_->kind != H264S_I8x8) + (b_ && b_->kind != H264S_I4x4 && b_->kind != H264S_I8x8);
        const int is16 = e_.decision(ctx0, m.kind == H264S_I16x16);
        if (!is16) {
            if (!E::kEncoding) m.kind = H264S_I4x4;              // I_NxN: transform_size_8x8_flag tells 4x4 from 8x8
            return 0;
        }
        if (e_.terminate(0)) return -2;                          // I_PCM
        // ctxIdxInc of bins 2.. : I slices 3,4,(5|6),(6|7),7; suffix in P slices 1,2,(2|3),3,3
        const int c2 = base + (i_slice ? 3 : 1), c3 = base + (i_slice ? 4 : 2);
        const int c4a = base + (i_slice ? 5 : 2), c4b = base + (i_slice ? 6 : 3);
        const int c5a = base + (i_slice ? 6 : 3), c5b = base + (i_slice ? 7 : 3), c6 = base + (i_slice ? 7 : 3);
        const int luma = e_.decision(c2, m.cbp_luma != 0);
        int chroma = e_.decision(c3, m.cbp_chroma != 0);
        int mode;
        if (chroma) {
            chroma = 1 + e_.decision(c4a, m.cbp_chroma == 2);
            mode = e_.decision(c5a, (m.i16_mode >> 1) & 1) << 1;
            mode |= e_.decision(c6, m.i16_mode & 1);
        } else {
            mode = e_.decision(c4b, (m.i16_mode >> 1) & 1) << 1;
            mode |= e_.decision(c5b, m.i16_mode & 1);
        }
        m.kind = H264S_I16x16;
        m.cbp_luma = luma ? 15 : 0;
        m.cbp_chroma = (uint8_t)chroma;
        m.i16_mode = (uint8_t)mode;
        return 0;
    }

    bool transform8x8_flag(bool want)
    {
        const int inc = (a_ && a_->t8x8) + (b_ && b_->t8x8);
        return e_.decision(399 + inc, want) != 0;
    }

    int chroma_pred_mode(int want)
    {
        const int inc = (a_ && a_->chroma_mode != 0) + (b_ && b_->chroma_mode != 0);
        if (!e_.decision(64 + inc, want > 0)) return 0;
        if (!e_.decision(64 + 3, want > 1)) return 1;
        return 2 + e_.decision(64 + 3, want > 2);
    }

    // ---- inter prediction syntax: sub_mb_type, ref_idx_l0, mvd_l0 (7.3.5.1, 7.3.5.2) ---------------------------------------------
    // |mvd| of the 4x4 block left of / above block (bx, by) of the current macroblock, component comp
    int mvd_abs_at(int bx, int by, int comp) const
    {
        if (bx < 0) return a_ ? a_->mvd[by * 4 + 3][comp] : 0;
        if (by < 0) return b_ ? b_->mvd[12 + bx][comp] : 0;
        return cur_->mvd[by * 4 + bx][comp];
    }
    int ref_gt0_at(int qx, int qy) const          // 8x8 quadrant coordinates relative to the current macroblock
    {
        if (qx < 0) return a_ ? (a_->ref_gt0 >> (qy * 2 + 1)) & 1 : 0;
        if (qy < 0) return b_ ? (b_->ref_gt0 >> (2 + qx)) & 1 : 0;
        return (cur_->ref_gt0 >> (qy * 2 + qx)) & 1;
    }
    int ref_idx(int qx, int qy, int qw, int qh, int want)
    {
        const int inc = ref_gt0_at(qx - 1, qy) + 2 * ref_gt0_at(qx, qy - 1);
        int v = 0;
        if (e_.decision(54 + inc, want > 0)) {
            v = 1;
            if (e_.decision(54 + 4, want > 1)) {
                v = 2;
                while (v < 32 && e_.decision(54 + 5, want > v)) v++;
            }
        }
        if (v > 0)
            for (int y = qy; y < qy + qh; y++)
                for (int x = qx; x < qx + qw; x++) cur_->ref_gt0 |= (uint8_t)(1u << (y * 2 + x));
        return v;
    }
    int mvd_comp(int bx, int by, int comp, int want)
    {
        const int base = comp ? 47 : 40;
        const int sum = mvd_abs_at(bx - 1, by, comp) + mvd_abs_at(bx, by - 1, comp);
        const int inc = sum < 3 ? 0 : (sum > 32 ? 2 : 1);
        const int wabs = want < 0 ? -want : want;
        if (!e_.decision(base + inc, wabs > 0)) return 0;
        int v = 1;
        // truncated unary prefix up to 9 with ctxIdxInc 3, 4, 5, 6, 6, ...
        while (v < 9 && e_.decision(base + (v < 4 ? 2 + v : 6), wabs > v)) v++;
        if (v == 9) v += (int)egk_bypass(3, (uint32_t)(wabs - 9));
        return e_.bypass(want < 0) ? -v : v;
    }
    void mvd_part(h264s_mb& m, int idx, int bx, int by, int bw, int bh)
    {
        const int x = mvd_comp(bx, by, 0, m.mvd[idx][0]);
        const int y = mvd_comp(bx, by, 1, m.mvd[idx][1]);
        m.mvd[idx][0] = (int16_t)x;
        m.mvd[idx][1] = (int16_t)y;
        const uint8_t ax = (uint8_t)((x < 0 ? -x : x) > 127 ? 127 : (x < 0 ? -x : x));
        const uint8_t ay = (uint8_t)((y < 0 ? -y : y) > 127 ? 127 : (y < 0 ? -y : y));
        for (int j = by; j < by + bh; j++)
            for (int i = bx; i < bx + bw; i++) { cur_->mvd[j * 4 + i][0] = ax; cur_->mvd[j * 4 + i][1] = ay; }
    }
    void inter_prediction(h264s_mb& m)
    {
        const bool send_ref = sp_.num_ref_idx_l0 > 1;
        if (m.kind == H264S_P8x8) {
            for (int q = 0; q < 4; q++) {
                // P_L0_8x8 "1", 8x4 "0 0", 4x8 "0 1 1", 4x4 "0 1 0"
                const int want = m.sub_type[q];
                int st;
                if (e_.decision(21, want == 0)) st = 0;
                else if (!e_.decision(22, want != 1)) st = 1;
                else st = e_.decision(23, want == 2) ? 2 : 3;
                m.sub_type[q] = (uint8_t)st;
            }
            for (int q = 0; q < 4; q++) m.ref_idx[q] = send_ref ? (uint8_t)ref_idx(q & 1, q >> 1, 1, 1, m.ref_idx[q]) : 0;
            for (int q = 0; q < 4; q++) {
                const int bx = (q & 1) * 2, by = (q >> 1) * 2;
                switch (m.sub_type[q]) {
                case 0: mvd_part(m, q * 4, bx, by, 2, 2); break;
                case 1: mvd_part(m, q * 4, bx, by, 2, 1); mvd_part(m, q * 4 + 1, bx, by + 1, 2, 1); break;
                case 2: mvd_part(m, q * 4, bx, by, 1, 2); mvd_part(m, q * 4 + 1, bx + 1, by, 1, 2); break;
                default:
                    for (int s = 0; s < 4; s++) mvd_part(m, q * 4 + s, bx + (s & 1), by + (s >> 1), 1, 1);
                }
            }
        } else if (m.kind == H264S_P16x16) {
            m.ref_idx[0] = send_ref ? (uint8_t)ref_idx(0, 0, 2, 2, m.ref_idx[0]) : 0;
            mvd_part(m, 0, 0, 0, 4, 4);
        } else if (m.kind == H264S_P16x8) {
            for (int p = 0; p < 2; p++) m.ref_idx[p] = send_ref ? (uint8_t)ref_idx(0, p, 2, 1, m.ref_idx[p]) : 0;
            for (int p = 0; p < 2; p++) mvd_part(m, p * 4, 0, p * 2, 4, 2);
        } else {
            for (int p = 0; p < 2; p++) m.ref_idx[p] = send_ref ? (uint8_t)ref_idx(p, 0, 1, 2, m.ref_idx[p]) : 0;
            for (int p = 0; p < 2; p++) mvd_part(m, p * 4, p * 2, 0, 2, 4);
        }
    }

    // ---- coded_block_pattern (9.3.2.6, 9.3.3.1.1.4), mb_qp_delta (9.3.2.7, 9.3.3.1.1.5) ------------------------------------------
    void coded_block_pattern(h264s_mb& m)
    {
        int luma = 0;
        for (int b8 = 0; b8 < 4; b8++) {
            int ca, cb;       // condTermFlag: the neighbouring 8x8 block is available and has its bit CLEAR
            if (b8 & 1) ca = !((luma >> (b8 - 1)) & 1);
            else ca = a_ ? !((a_->cbp_luma >> (b8 + 1)) & 1) : 0;
            if (b8 & 2) cb = !((luma >> (b8 - 2)) & 1);
            else cb = b_ ? !((b_->cbp_luma >> (b8 + 2)) & 1) : 0;
            luma |= e_.decision(73 + ca + 2 * cb, (m.cbp_luma >> b8) & 1) << b8;
        }
        int chroma = 0;
        {
            const int ca = a_ && a_->cbp_chroma != 0, cb = b_ && b_->cbp_chroma != 0;
            if (e_.decision(77 + ca + 2 * cb, m.cbp_chroma != 0)) {
                const int ca2 = a_ && a_->cbp_chroma == 2, cb2 = b_ && b_->cbp_chroma == 2;
                chroma = 1 + e_.decision(77 + 4 + ca2 + 2 * cb2, m.cbp_chroma == 2);
            }
        }
        m.cbp_luma = (uint8_t)luma;
        m.cbp_chroma = (uint8_t)chroma;
    }
    int mb_qp_delta(int want)
    {
        const int wmap = want > 0 ? 2 * want - 1 : -2 * want;
        int v = 0;
        if (e_.decision(60 + (last_dqp_nz_ ? 1 : 0), wmap > 0)) {
            v = 1;
            if (e_.decision(62, wmap > 1)) {
                v = 2;
                while (v < 104 && e_.decision(63, wmap > v)) v++;
            }
        }
        return (v & 1) ? (v + 1) >> 1 : -(v >> 1);
    }

    // k-th order Exp-Golomb through the bypass engine (UEGk suffix, 9.3.2.5)
    uint32_t egk_bypass(int k, uint32_t want)
    {
        uint32_t v = 0, w = want;
        for (;;) {
            const bool more = w >= (1u << k);
            if (e_.bypass(more) && k < 30) {
                v += 1u << k;
                if (E::kEncoding) w -= 1u << k;
                k++;
            } else break;
        }
        while (k-- > 0) v += (uint32_t)e_.bypass((w >> k) & 1) << k;
        return v;
    }

    // ---- residual ---------------------------------------------------------------------------------------------------------------
    // One residual block (7.3.5.3.3): lv[0 .. n) in scan order.  cbf_inc < 0: no coded_block_flag in the bitstream (8x8 blocks).
    // Returns coded_block_flag.
    int block(int cat, int cbf_inc, int16_t* lv, int n)
    {
        static const int kCbfOff[5] = {0, 4, 8, 12, 16};
        static const int kSigOff[5] = {0, 15, 29, 44, 47};
        static const int kAbsOff[5] = {0, 10, 20, 30, 39};
        int last_nz = -1;
        if (E::kEncoding)
            for (int i = 0; i < n; i++)
                if (lv[i]) last_nz = i;
        if (cbf_inc >= 0) {
            if (!e_.decision(85 + kCbfOff[cat] + cbf_inc, last_nz >= 0)) {
                if (!E::kEncoding) memset(lv, 0, sizeof(int16_t) * (size_t)n);
                return 0;
            }
        }
        const int sig_base = cat == 5 ? 402 : 105 + kSigOff[cat];
        const int last_base = cat == 5 ? 417 : 166 + kSigOff[cat];
        const int abs_base = cat == 5 ? 426 : 227 + kAbsOff[cat];
        uint8_t pos[64];
        int count = 0, i = 0;
        for (; i < n - 1; i++) {
            const int si = cat == 5 ? kSig8x8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
            if (e_.decision(sig_base + si, E::kEncoding ? lv[i] != 0 : 0)) {
                pos[count++] = (uint8_t)i;
                const int li = cat == 5 ? kLast8x8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
                if (e_.decision(last_base + li, i == last_nz)) break;
            } else if (!E::kEncoding) lv[i] = 0;
        }
        if (i == n - 1) pos[count++] = (uint8_t)(n - 1);          // the last position is significant by inference
        else if (!E::kEncoding)
            for (int k = i + 1; k < n; k++) lv[k] = 0;
        int eq1 = 0, gt1 = 0;
        for (int k = count - 1; k >= 0; k--) {
            const int p = pos[k];
            int want = 0;
            if (E::kEncoding) want = (lv[p] < 0 ? -lv[p] : lv[p]) - 1;
            const int c0 = abs_base + (gt1 ? 0 : (eq1 < 3 ? 1 + eq1 : 4));
            int v = 0;
            if (e_.decision(c0, want > 0)) {
                const int c1 = abs_base + 5 + (gt1 < 4 - (cat == 3) ? gt1 : 4 - (cat == 3));
                v = 1;
                while (v < 14 && e_.decision(c1, want > v)) v++;
                if (v == 14) v += (int)egk_bypass(0, (uint32_t)(want - 14));
                gt1++;
            } else eq1++;
            const int neg = e_.bypass(E::kEncoding ? lv[p] < 0 : 0);
            if (!E::kEncoding) lv[p] = (int16_t)(neg ? -(v + 1) : v + 1);
        }
        return 1;
    }

    // coded_block_flag context increment (9.3.3.1.1.9) from the flags of the blocks to the left and above; an unavailable
    // macroblock counts as coded for intra macroblocks, as not coded for inter ones
    int cbf_inc(int fa, int fb) const { return fa + 2 * fb; }
    int na() const { return intra_ ? 1 : 0; }
    int luma_cbf_at(int bx, int by) const
    {
        if (bx < 0) return a_ ? (a_->cbf_luma >> (by * 4 + 3)) & 1 : na();
        if (by < 0) return b_ ? (b_->cbf_luma >> (12 + bx)) & 1 : na();
        return (cur_->cbf_luma >> (by * 4 + bx)) & 1;
    }
    int chroma_cbf_at(int c, int bx, int by) const
    {
        if (bx < 0) return a_ ? (a_->cbf_chroma >> (c * 4 + by * 2 + 1)) & 1 : na();
        if (by < 0) return b_ ? (b_->cbf_chroma >> (c * 4 + 2 + bx)) & 1 : na();
        return (cur_->cbf_chroma >> (c * 4 + by * 2 + bx)) & 1;
    }

    int residual(h264s_mb& m)
    {
        MbInfo& c = *cur_;
        if (m.kind == H264S_I16x16) {
            const int fa = a_ ? a_->cbf_dc & 1 : 1, fb = b_ ? b_->cbf_dc & 1 : 1;
            if (block(kCatLumaDC, cbf_inc(fa, fb), m.luma_dc, 16)) c.cbf_dc |= 1;
        }
        for (int q = 0; q < 4; q++) {
            if (!((m.cbp_luma >> q) & 1)) {
                if (!E::kEncoding) memset(m.luma[q * 4], 0, sizeof(int16_t) * 64);
                continue;
            }
            const int qx = (q & 1) * 2, qy = (q >> 1) * 2;
            if (m.t8x8) {
                block(kCatLuma8x8, -1, m.luma[q * 4], 64);
                c.cbf_luma |= (uint16_t)(0x33u << (qy * 4 + qx));
                continue;
            }
            for (int s = 0; s < 4; s++) {
                const int bx = qx + (s & 1), by = qy + (s >> 1);
                const int inc = cbf_inc(luma_cbf_at(bx - 1, by), luma_cbf_at(bx, by - 1));
                const int coded = m.kind == H264S_I16x16 ? block(kCatLumaAC, inc, m.luma[q * 4 + s], 15)
                                                         : block(kCatLuma4x4, inc, m.luma[q * 4 + s], 16);
                if (coded) c.cbf_luma |= (uint16_t)(1u << (by * 4 + bx));
            }
        }
        for (int ch = 0; ch < 2; ch++) {
            if (!(m.cbp_chroma & 3)) {
                if (!E::kEncoding) memset(m.chroma_dc[ch], 0, sizeof(m.chroma_dc[ch]));
                continue;
            }
            const int fa = a_ ? (a_->cbf_dc >> (1 + ch)) & 1 : na(), fb = b_ ? (b_->cbf_dc >> (1 + ch)) & 1 : na();
            if (block(kCatChromaDC, cbf_inc(fa, fb), m.chroma_dc[ch], 4)) c.cbf_dc |= (uint8_t)(2u << ch);
        }
        for (int ch = 0; ch < 2; ch++) {
            if (!(m.cbp_chroma & 2)) {
                if (!E::kEncoding) memset(m.chroma_ac[ch], 0, sizeof(m.chroma_ac[ch]));
                continue;
            }
            for (int s = 0; s < 4; s++) {
                const int bx = s & 1, by = s >> 1;
                const int inc = cbf_inc(chroma_cbf_at(ch, bx - 1, by), chroma_cbf_at(ch, bx, by - 1));
                if (block(kCatChromaAC, inc, m.chroma_ac[ch][s], 15)) c.cbf_chroma |= (uint8_t)(1u << (ch * 4 + s));
            }
        }
        return 0;
    }

    E& e_;
    const SliceParams& sp_;
    MbInfo* info_;
    uint16_t slice_;
    const MbInfo* a_ = nullptr;
    const MbInfo* b_ = nullptr;
    MbInfo* cur_ = nullptr;
    bool intra_ = false;
    bool last_dqp_nz_ = false;
};

}  // namespace h264host
