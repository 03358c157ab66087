// cavlc.h -- the macroblock layer of a CAVLC slice (entropy_coding_mode_flag = 0: ITU-T H.264 7.3.4, 7.3.5, 7.3.5.1-3 with the
// ue(v) / se(v) / te(v) / me(v) descriptors and residual_block_cavlc, 7.3.5.3.2 + 9.2), written once against a small bit-I/O
// interface and instantiated for both directions, like mb_codec.h for CABAC: with VlcReader it fills h264s_mb records from the
// bitstream, with VlcWriter it writes them.  Baseline-profile streams (BASELINE config 1, "baseline-profile stream") are CAVLC;
// High profile allows it as well (the 8x8 transform then travels as four interleaved 4x4 blocks).
//
// The reference has no CAVLC: residual_block_cavlc is an empty stub (h264/residual.cpp:488-560) and its macroblock layer reads
// every element through the arithmetic decoder.  No counterpart; pinned against libavcodec (tests/test_host_codec.py).
//
// Scope as mb_codec.h: I and P slices of progressive frames, 4:2:0; I_PCM is reported as unsupported.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../../include/h264_syntax.h"
#include "bits.h"
#include "cavlc_tables.h"
#include "mb_codec.h"

namespace h264host {

// total_coeff of the blocks of a finished macroblock: what nC of later blocks reads (9.2.1)
struct NnzInfo {
    uint8_t luma[16];          // raster index of the 4x4 block
    uint8_t chroma[2][4];
};

struct VlcReader {
    static const bool kEncoding = false;
    BitReader& r;
    uint32_t ue(uint32_t) { return r.ue(); }
    int32_t se(int32_t) { return r.se(); }
    uint32_t u(int n, uint32_t) { return r.u(n); }
    // index of the table entry whose code comes next, or -1
    int vlc(const uint8_t* len, const uint8_t* bits, int n, int)
    {
        uint32_t code = 0;
        for (int l = 1; l <= 16; l++) {
            code = (code << 1) | r.u1();
            for (int i = 0; i < n; i++)
                if (len[i] == l && bits[i] == code) return i;
        }
        return -1;
    }
    bool overrun() const { return r.overrun(); }
};
struct VlcWriter {
    static const bool kEncoding = true;
    BitWriter& w;
    uint32_t ue(uint32_t want) { w.ue(want); return want; }
    int32_t se(int32_t want) { w.se(want); return want; }
    uint32_t u(int n, uint32_t want) { w.put(want, n); return want; }
    int vlc(const uint8_t* len, const uint8_t* bits, int n, int want)
    {
        if (want < 0 || want >= n || !len[want]) return -1;
        w.put(bits[want], len[want]);
        return want;
    }
    bool overrun() const { return false; }
};

template <class IO>
class CavlcSliceCodec {
public:
    CavlcSliceCodec(IO& io, const SliceParams& sp, MbInfo* info, NnzInfo* nnz, int slice_no)
        : io_(io), sp_(sp), info_(info), nnz_(nnz), slice_((uint16_t)(slice_no + 1)) {}

    // a skipped macroblock (counted in mb_skip_run): no syntax, every total_coeff 0
    void skipped(int addr, h264s_mb& m)
    {
        if (!IO::kEncoding) memset(&m, 0, sizeof(m));
        m.kind = H264S_PSKIP;
        MbInfo& c = info_[addr];
        memset(&c, 0, sizeof(c));
        c.kind = H264S_PSKIP;
        c.slice = slice_;
        memset(&nnz_[addr], 0, sizeof(NnzInfo));
    }

    // macroblock_layer() of one macroblock that is not skipped.  Returns 0, -1 (malformed / not expressible), -2 (I_PCM).
    int mb(int addr, h264s_mb& m)
    {
        const int mbx = addr % sp_.w_mbs;
        a_ = mbx > 0 && info_[addr - 1].slice == slice_ ? addr - 1 : -1;
        b_ = addr >= sp_.w_mbs && info_[addr - sp_.w_mbs].slice == slice_ ? addr - sp_.w_mbs : -1;
        MbInfo& c = info_[addr];
        memset(&c, 0, sizeof(c));
        cur_ = addr;
        memset(&nnz_[addr], 0, sizeof(NnzInfo));
        if (!IO::kEncoding) memset(&m, 0, sizeof(m));
        // mb_type, ue(v): Tables 7-11 (I), 7-13 (P; 5.. = the I types)
        uint32_t want_type = 0;
        if (IO::kEncoding) {
            const uint32_t itype = m.kind == H264S_I16x16 ? 1u + m.i16_mode + 4u * m.cbp_chroma + (m.cbp_luma ? 12u : 0u) : 0u;
            switch (m.kind) {
            case H264S_P16x16: want_type = 0; break;
            case H264S_P16x8: want_type = 1; break;
            case H264S_P8x16: want_type = 2; break;
            case H264S_P8x8: want_type = 3; break;
            case H264S_I4x4: case H264S_I8x8: case H264S_I16x16: want_type = itype + (sp_.p_slice ? 5u : 0u); break;
            default: return -1;
            }
            if (!sp_.p_slice && m.kind != H264S_I4x4 && m.kind != H264S_I8x8 && m.kind != H264S_I16x16) return -1;
        }
        uint32_t t = io_.ue(want_type);
        bool ref0 = false;           // P_8x8ref0: every ref_idx_l0 is 0 and not sent
        if (sp_.p_slice && t < 5) {
            m.kind = (uint8_t)(t == 0 ? H264S_P16x16 : (t == 1 ? H264S_P16x8 : (t == 2 ? H264S_P8x16 : H264S_P8x8)));
            ref0 = t == 4;
        } else {
            if (sp_.p_slice) t -= 5;
            if (t == 25) return -2;
            if (t > 25) return -1;
            if (t == 0) { if (!IO::kEncoding) m.kind = H264S_I4x4; }
            else {
                m.kind = H264S_I16x16;
                m.i16_mode = (uint8_t)((t - 1) & 3);
                m.cbp_chroma = (uint8_t)(((t - 1) >> 2) % 3);
                m.cbp_luma = t >= 13 ? 15 : 0;
            }
        }
        intra_ = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        bool t8 = false;
        if (m.kind == H264S_I4x4 || m.kind == H264S_I8x8) {
            if (sp_.transform_8x8_mode) {
                t8 = io_.u(1, m.kind == H264S_I8x8) != 0;
                m.kind = t8 ? H264S_I8x8 : H264S_I4x4;
            } else if (m.kind == H264S_I8x8) return -1;
            const int nb = t8 ? 4 : 16;
            for (int i = 0; i < nb; i++) {
                m.prev_flag[i] = (uint8_t)io_.u(1, m.prev_flag[i] != 0);
                if (!m.prev_flag[i]) m.rem_mode[i] = (uint8_t)io_.u(3, m.rem_mode[i] & 7u);
            }
        }
        c.kind = m.kind;
        if (intra_) {
            const uint32_t cm = io_.ue(m.chroma_mode);
            if (cm > 3) return -1;
            m.chroma_mode = (uint8_t)cm;
            c.chroma_mode = m.chroma_mode;
        } else {
            const int rc = inter_prediction(m, ref0);
            if (rc) return rc;
        }
        if (m.kind != H264S_I16x16) {
            // coded_block_pattern, me(v) (Table 9-4)
            const uint8_t* map = intra_ ? kCbpIntra : kCbpInter;
            uint32_t want = 0;
            if (IO::kEncoding) {
                const int cbp = m.cbp_luma | (m.cbp_chroma << 4);
                while (want < 48 && map[want] != cbp) want++;
                if (want == 48) return -1;
            }
            const uint32_t code = io_.ue(want);
            if (code > 47) return -1;
            m.cbp_luma = map[code] & 15;
            m.cbp_chroma = map[code] >> 4;
            if (m.cbp_luma && sp_.transform_8x8_mode && !intra_ &&
                (m.kind != H264S_P8x8 || (m.sub_type[0] | m.sub_type[1] | m.sub_type[2] | m.sub_type[3]) == 0))
                t8 = io_.u(1, m.t8x8 != 0) != 0;
        }
        m.t8x8 = t8;
        c.t8x8 = t8;
        c.cbp_luma = m.cbp_luma;
        c.cbp_chroma = m.cbp_chroma;
        c.slice = slice_;
        if (m.cbp_luma || m.cbp_chroma || m.kind == H264S_I16x16) {
            const int32_t dq = io_.se(m.qp_delta);
            if (dq < -26 || dq > 25) return -1;
            m.qp_delta = (int8_t)dq;
            return residual(m);
        }
        if (IO::kEncoding && m.qp_delta != 0) return -1;
        m.qp_delta = 0;
        return 0;
    }

private:
    // te(v) with range cMax = num_ref_idx_active - 1 > 0
    int ref_idx_te(int want)
    {
        const int cmax = sp_.num_ref_idx_l0 - 1;
        if (cmax > 1) return (int)io_.ue((uint32_t)want);
        return io_.u(1, want ? 0u : 1u) ? 0 : 1;
    }
    void mvd_part(h264s_mb& m, int idx)
    {
        m.mvd[idx][0] = (int16_t)io_.se(m.mvd[idx][0]);
        m.mvd[idx][1] = (int16_t)io_.se(m.mvd[idx][1]);
    }
    int inter_prediction(h264s_mb& m, bool ref0)
    {
        const bool send_ref = sp_.num_ref_idx_l0 > 1 && !ref0;
        int n_parts = 1;
        if (m.kind == H264S_P8x8) {
            for (int q = 0; q < 4; q++) {
                const uint32_t st = io_.ue(m.sub_type[q]);       // Table 7-17: 0 8x8, 1 8x4, 2 4x8, 3 4x4
                if (st > 3) return -1;
                m.sub_type[q] = (uint8_t)st;
            }
            n_parts = 4;
        } else if (m.kind != H264S_P16x16) n_parts = 2;
        for (int p = 0; p < n_parts; p++) {
            int v = 0;
            if (send_ref) v = ref_idx_te(m.ref_idx[p]);
            if (v < 0 || v > 31) return -1;
            m.ref_idx[p] = (uint8_t)v;
        }
        if (m.kind == H264S_P8x8) {
            for (int q = 0; q < 4; q++) {
                const int n = m.sub_type[q] == 0 ? 1 : (m.sub_type[q] == 3 ? 4 : 2);
                for (int s = 0; s < n; s++) mvd_part(m, q * 4 + s);
            }
        } else {
            for (int p = 0; p < n_parts; p++) mvd_part(m, p * 4);
        }
        return 0;
    }

    // ---- residual -------------------------------------------------------------------------------------------------------------------
    // nC of a luma block at raster (bx, by) of the current macroblock / of chroma block (bx, by) of plane ch (9.2.1)
    int nc_from(int na, bool has_a, int nb, bool has_b) const
    {
        if (has_a && has_b) return (na + nb + 1) >> 1;
        return has_a ? na : (has_b ? nb : 0);
    }
    int luma_nc(int bx, int by) const
    {
        int na = 0, nb = 0;
        bool ha = true, hb = true;
        if (bx > 0) na = nnz_[cur_].luma[by * 4 + bx - 1];
        else if (a_ >= 0) na = nnz_[a_].luma[by * 4 + 3];
        else ha = false;
        if (by > 0) nb = nnz_[cur_].luma[(by - 1) * 4 + bx];
        else if (b_ >= 0) nb = nnz_[b_].luma[12 + bx];
        else hb = false;
        return nc_from(na, ha, nb, hb);
    }
    int chroma_nc(int ch, int bx, int by) const
    {
        int na = 0, nb = 0;
        bool ha = true, hb = true;
        if (bx > 0) na = nnz_[cur_].chroma[ch][by * 2];
        else if (a_ >= 0) na = nnz_[a_].chroma[ch][by * 2 + 1];
        else ha = false;
        if (by > 0) nb = nnz_[cur_].chroma[ch][bx];
        else if (b_ >= 0) nb = nnz_[b_].chroma[ch][2 + bx];
        else hb = false;
        return nc_from(na, ha, nb, hb);
    }

    // residual_block_cavlc (7.3.5.3.2): lv[0 .. n) in scan order, nC as derived by the caller (-1: chroma DC).  Returns total_coeff
    // or a negative error.
    int block(int16_t* lv, int n, int nc)
    {
        const uint8_t* tlen; const uint8_t* tbits; int tn;
        if (nc < 0) { tlen = kChromaDcTokenLen; tbits = kChromaDcTokenBits; tn = 4 * 5; }
        else {
            const int t = nc < 2 ? 0 : (nc < 4 ? 1 : (nc < 8 ? 2 : 3));
            tlen = kCoeffTokenLen[t]; tbits = kCoeffTokenBits[t]; tn = 4 * 17;
        }
        int level[16], run[16];
        int total = 0, t1 = 0, total_zeros = 0;
        if (IO::kEncoding) {
            // levels from the highest frequency down, the zeros in front of each (run_before), trailing +-1s (at most three)
            int z = 0;
            for (int i = n - 1; i >= 0; i--) {
                if (lv[i]) {
                    if (total) run[total - 1] = z;
                    level[total++] = lv[i];
                    z = 0;
                } else if (total) z++;
            }
            if (total) run[total - 1] = z;
            for (int i = 0; i < total; i++) total_zeros += run[i];
            while (t1 < total && t1 < 3 && (level[t1] == 1 || level[t1] == -1)) t1++;
        }
        const int tok = io_.vlc(tlen, tbits, tn, 4 * total + t1);
        if (tok < 0) return -1;
        total = tok >> 2; t1 = tok & 3;
        if (!IO::kEncoding) memset(lv, 0, sizeof(int16_t) * (size_t)n);
        if (total == 0) return 0;
        if (total > n) return -1;
        int suffix_len = total > 10 && t1 < 3 ? 1 : 0;
        for (int i = 0; i < total; i++) {
            if (i < t1) {
                const uint32_t s = io_.u(1, level[i] < 0);
                level[i] = s ? -1 : 1;
                continue;
            }
            int code = 0;
            if (IO::kEncoding) {
                code = level[i] > 0 ? 2 * level[i] - 2 : -2 * level[i] - 1;
                if (i == t1 && t1 < 3) code -= 2;
            }
            // level_prefix (unary) and level_suffix: prefix p with suffix size s covers codes base .. base + 2^s - 1
            int p = 0;
            if (IO::kEncoding) {
                for (;; p++) {
                    if (p > 28) return -1;
                    const int s = p == 14 && suffix_len == 0 ? 4 : (p >= 15 ? p - 3 : suffix_len);
                    const int base = ((p < 15 ? p : 15) << suffix_len) + (p >= 15 && suffix_len == 0 ? 15 : 0) + (p >= 16 ? (1 << (p - 3)) - 4096 : 0);
                    if (suffix_len == 0 && p < 14) { if (code == base) break; continue; }
                    if (code >= base && code < base + (1 << s)) break;
                }
                for (int k = 0; k < p; k++) io_.u(1, 0);
                io_.u(1, 1);
            } else {
                while (p < 32 && !io_.overrun() && io_.u(1, 0) == 0) p++;
                if (p > 28) return -1;
            }
            {
                const int s = p == 14 && suffix_len == 0 ? 4 : (p >= 15 ? p - 3 : suffix_len);
                const int base = ((p < 15 ? p : 15) << suffix_len) + (p >= 15 && suffix_len == 0 ? 15 : 0) + (p >= 16 ? (1 << (p - 3)) - 4096 : 0);
                const bool has_suffix = suffix_len > 0 || p >= 14;
                uint32_t sfx = 0;
                if (has_suffix && s > 0) sfx = io_.u(s, (uint32_t)(code - base));
                if (!IO::kEncoding) code = base + (int)sfx;
            }
            if (!IO::kEncoding) {
                int c2 = code;
                if (i == t1 && t1 < 3) c2 += 2;
                level[i] = (c2 & 1) ? (-c2 - 1) >> 1 : (c2 + 2) >> 1;
            }
            if (suffix_len == 0) suffix_len = 1;
            const int al = level[i] < 0 ? -level[i] : level[i];
            if (al > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
        }
        if (total < n) {
            int tz;
            if (n == 4) tz = io_.vlc(kChromaDcTotalZerosLen[total - 1], kChromaDcTotalZerosBits[total - 1], 4, total_zeros);
            else tz = io_.vlc(kTotalZerosLen[total - 1], kTotalZerosBits[total - 1], 16, total_zeros);
            if (tz < 0 || tz + total > n) return -1;
            total_zeros = tz;
        } else total_zeros = 0;
        int zeros_left = total_zeros;
        for (int i = 0; i < total - 1; i++) {
            if (zeros_left > 0) {
                const int zi = (zeros_left < 7 ? zeros_left : 7) - 1;
                const int r = io_.vlc(kRunLen[zi], kRunBits[zi], 16, run[i]);
                if (r < 0 || r > zeros_left) return -1;
                run[i] = r;
                zeros_left -= r;
            } else run[i] = 0;
        }
        run[total - 1] = zeros_left;
        if (!IO::kEncoding) {
            int pos = -1;
            for (int i = total - 1; i >= 0; i--) {
                pos += run[i] + 1;
                if (pos >= n) return -1;
                if (level[i] < -32768 || level[i] > 32767) return -1;
                lv[pos] = (int16_t)level[i];
            }
        }
        return total;
    }

    int residual(h264s_mb& m)
    {
        NnzInfo& z = nnz_[cur_];
        if (m.kind == H264S_I16x16) {
            if (block(m.luma_dc, 16, luma_nc(0, 0)) < 0) return -1;
        }
        for (int q = 0; q < 4; q++) {
            const int qx = (q & 1) * 2, qy = (q >> 1) * 2;
            if (!((m.cbp_luma >> q) & 1)) continue;          // levels and total_coeff stay 0
            int16_t split[4][16];
            if (m.t8x8 && IO::kEncoding)                      // 8x8 block as four interleaved 4x4 blocks: 8x8 scan index 4 i + k -> block k, index i
                for (int k = 0; k < 4; k++)
                    for (int i = 0; i < 16; i++) split[k][i] = m.luma[q * 4 + ((4 * i + k) >> 4)][(4 * i + k) & 15];
            for (int s = 0; s < 4; s++) {
                const int bx = qx + (s & 1), by = qy + (s >> 1);
                int16_t* lv = m.t8x8 ? split[s] : m.luma[q * 4 + s];
                const int tc = m.kind == H264S_I16x16 ? block(lv, 15, luma_nc(bx, by)) : block(lv, 16, luma_nc(bx, by));
                if (tc < 0) return -1;
                z.luma[by * 4 + bx] = (uint8_t)tc;
            }
            if (m.t8x8 && !IO::kEncoding)
                for (int k = 0; k < 4; k++)
                    for (int i = 0; i < 16; i++) m.luma[q * 4 + ((4 * i + k) >> 4)][(4 * i + k) & 15] = split[k][i];
        }
        if (m.cbp_chroma & 3)
            for (int ch = 0; ch < 2; ch++)
                if (block(m.chroma_dc[ch], 4, -1) < 0) return -1;
        if (m.cbp_chroma & 2)
            for (int ch = 0; ch < 2; ch++)
                for (int s = 0; s < 4; s++) {
                    const int tc = block(m.chroma_ac[ch][s], 15, chroma_nc(ch, s & 1, s >> 1));
                    if (tc < 0) return -1;
                    z.chroma[ch][s] = (uint8_t)tc;
                }
        return 0;
    }

    IO& io_;
    const SliceParams& sp_;
    MbInfo* info_;
    NnzInfo* nnz_;
    uint16_t slice_;
    int a_ = -1, b_ = -1, cur_ = 0;
    bool intra_ = false;
};

}  // namespace h264host
