// h264_derive.h -- what the host derives from a macroblock's syntax elements before it can cross the C ABI of include/h264r.h:
// QP'Y / QP'C, Intra4x4 / Intra8x8 prediction modes from prev_/rem_ flags, motion vectors from mvd + prediction, P_Skip
// inference, neighbour availability; and the inverse scans that put coefficient levels where the engine expects them.
// The same state serves the stream writer in the other direction (final vectors / modes -> mvd / prev / rem).
//
// Two rule sets:
//   spec        ITU-T H.264 8.3.1.1, 8.3.2.1, 8.4.1.1 - 8.4.1.3, 7.4.5 (QP), Table 8-15;
//   ref quirks  the reference decoder's own host-side derivations where they differ: D2 QPc = max(QPY' - 2, 0)
//               (h264/macroblock.cpp:487-488), D11 any inter neighbour forces the DC mode predictor
//               (h264/macroblock.cpp:1005-1009), D12 the P_Skip zero-vector test and the neighbour-partition lookup of
//               picture::get_PartNeighbour / get_MvNeighbour (h264/picture.cpp:185-394) with its sub-partition offsets,
//               its view of not-yet-decoded partitions of the current macroblock and its C -> D substitution.
//
// Replaces, on the reference's side: macroblock::ParseIntra4x4PredMode (h264/macroblock.cpp:992-1035), the QP lines of
// macroblock::Calc_residual (h264/macroblock.cpp:483-488), picture::get_MvNeighbour (h264/picture.cpp:280-394), macroblock::Init0
// for P_Skip (h264/macroblock.cpp:81-142) and Inverse4x4Scan (h264/residual.cpp:136-168).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../../include/h264_syntax.h"
#include "../../../include/h264r.h"
#include "mb_codec.h"

namespace h264host {

static const uint8_t kZigzag4x4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static const uint8_t kZigzag8x8[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// Table 8-15: QPC as a function of qPI
static const uint8_t kChromaQp[52] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                      26, 27, 28, 29, 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};

struct DeriveParams {
    int w_mbs = 0, h_mbs = 0;
    bool ref_quirks = false;
    bool constrained_intra_pred = false;
    int chroma_qp_offset[2] = {0, 0};
};

// per-picture state of the derivations
class PicDerive {
public:
    void begin_picture(const DeriveParams& dp)
    {
        dp_ = dp;
        w4_ = dp.w_mbs * 4; h4_ = dp.h_mbs * 4;
        mv_.assign((size_t)w4_ * h4_ * 2, 0);
        ref_.assign((size_t)w4_ * h4_, kNotYet);
        imode_.assign((size_t)w4_ * h4_, kNotINxN);
        const size_t n = (size_t)dp.w_mbs * dp.h_mbs;
        if (dp.ref_quirks) qmb_.assign(n, QuirkMb());
    }
    void begin_slice(int slice_qp) { qp_ = slice_qp; }

    // Syntax -> final values of macroblock `addr` (decode direction).  info[] is the codec's neighbour array (slice membership).
    // Writes the header record, the per-4x4 vectors (32 int16; untouched for intra) and the 384 levels in engine order.
    void derive_mb(int addr, const h264s_mb& m, const MbInfo* info, h264r_mb_hdr* hdr, int16_t* mv, int16_t* coeff)
    {
        setup(addr, info);
        memset(hdr, 0, sizeof(*hdr));
        hdr->flags = avail_flags(addr, info, m);
        // ---- QP (7.4.5: QPY = (QPY,PRED + mb_qp_delta + 52) % 52)
        qp_ = (qp_ + m.qp_delta + 52) % 52;
        hdr->qp_y = (uint8_t)qp_;
        if (dp_.ref_quirks) hdr->qp_cb = hdr->qp_cr = (uint8_t)(qp_ > 2 ? qp_ - 2 : 0);
        else {
            hdr->qp_cb = kChromaQp[clip(qp_ + dp_.chroma_qp_offset[0], 0, 51)];
            hdr->qp_cr = kChromaQp[clip(qp_ + dp_.chroma_qp_offset[1], 0, 51)];
        }
        hdr->cbp = (uint8_t)(m.cbp_luma | (m.cbp_chroma << 4));
        if (m.t8x8) hdr->flags |= H264R_T8x8;
        const bool intra = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        if (intra) {
            hdr->intra_modes = (uint8_t)((m.i16_mode & 3) | ((m.chroma_mode & 3) << 2));
            fill_ref(0, 0, 4, 4, kIntra);
            if (m.kind == H264S_I16x16) {
                hdr->mb_class = H264R_MB_I16x16;
                fill_imode(0, 0, 4, kNotINxN);
            } else if (m.kind == H264S_I4x4) {
                hdr->mb_class = H264R_MB_I4x4;
                for (int b = 0; b < 16; b++) {
                    const int r = kBlkRaster[b], bx = r & 3, by = r >> 2;
                    const int mode = intra_mode(bx, by, m.prev_flag[b], m.rem_mode[b]);
                    imode_at(bx, by) = (int8_t)mode;
                    hdr->u.pred4x4[b >> 1] |= (uint8_t)(mode << ((b & 1) * 4));
                }
            } else {
                hdr->mb_class = H264R_MB_I8x8;
                for (int q = 0; q < 4; q++) {
                    const int bx = (q & 1) * 2, by = (q >> 1) * 2;
                    const int mode = intra_mode(bx, by, m.prev_flag[q], m.rem_mode[q]);
                    fill_imode(bx, by, 2, (int8_t)mode);
                    hdr->u.pred4x4[q >> 1] |= (uint8_t)(mode << ((q & 1) * 4));
                }
            }
            if (dp_.ref_quirks) qmb_[addr].kind = m.kind;
        } else {
            fill_imode(0, 0, 4, kInter);
            inter_mb(addr, m, hdr, /*underive=*/false, nullptr);
            for (int r = 0; r < 16; r++) {
                mv[r * 2] = mv_at(r & 3, r >> 2)[0];
                mv[r * 2 + 1] = mv_at(r & 3, r >> 2)[1];
            }
        }
        scan_levels(m, coeff);
    }

    // The other direction, for the stream writer: m.mvd[][] holds FINAL vectors (indexed like mvd) and m.rem_mode[] FINAL intra
    // modes; rewrites them as mvd / prev_flag + rem_mode.  Must be called for every macroblock of the picture in order (it also
    // advances the neighbour state).  A P_Skip macroblock keeps its kind only if its inferred vector is wanted: the caller checks
    // with skip_vector().
    void underive_mb(int addr, h264s_mb& m, const MbInfo* info)
    {
        setup(addr, info);
        qp_ = (qp_ + m.qp_delta + 52) % 52;
        const bool intra = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        if (intra) {
            fill_ref(0, 0, 4, 4, kIntra);
            if (m.kind == H264S_I16x16) fill_imode(0, 0, 4, kNotINxN);
            else {
                const int nb = m.kind == H264S_I4x4 ? 16 : 4;
                for (int b = 0; b < nb; b++) {
                    int bx, by;
                    if (nb == 16) { bx = kBlkRaster[b] & 3; by = kBlkRaster[b] >> 2; } else { bx = (b & 1) * 2; by = (b >> 1) * 2; }
                    const int want = m.rem_mode[b], pred = predicted_intra_mode(bx, by);
                    if (want == pred) { m.prev_flag[b] = 1; m.rem_mode[b] = 0; }
                    else { m.prev_flag[b] = 0; m.rem_mode[b] = (uint8_t)(want < pred ? want : want - 1); }
                    if (nb == 16) imode_at(bx, by) = (int8_t)want; else fill_imode(bx, by, 2, (int8_t)want);
                }
            }
            if (dp_.ref_quirks) qmb_[addr].kind = m.kind;
        } else {
            fill_imode(0, 0, 4, kInter);
            h264r_mb_hdr hdr;
            memset(&hdr, 0, sizeof(hdr));
            inter_mb(addr, m, &hdr, /*underive=*/true, &m);
        }
    }
    // the vector a P_Skip macroblock at addr would get (the slice state must stand just before that macroblock)
    void skip_vector(int addr, const MbInfo* info, int16_t out[2])
    {
        setup(addr, info);
        predict_skip(addr, out);
    }
    int qp() const { return qp_; }

private:
    enum : int8_t { kNotYet = -2, kIntra = -1 };
    enum : int8_t { kNotINxN = -1, kInter = -2 };      // imode_: Intra16x16 ("mode 2, available") / inter macroblock
    static int clip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
    static int median(int a, int b, int c) { return a > b ? (b > c ? b : (a > c ? c : a)) : (a > c ? a : (b > c ? c : b)); }

    void setup(int addr, const MbInfo* info)
    {
        addr_ = addr; info_ = info;
        mbx_ = addr % dp_.w_mbs; mby_ = addr / dp_.w_mbs;
        slice_ = info[addr].slice;
        fill_ref(0, 0, 4, 4, kNotYet);
    }
    bool mb_avail(int mbx, int mby) const
    {
        if (mbx < 0 || mby < 0 || mbx >= dp_.w_mbs || mby >= dp_.h_mbs) return false;
        const int a = mby * dp_.w_mbs + mbx;
        return a <= addr_ && info_[a].slice == slice_;
    }
    bool is_inter(int mbx, int mby) const { return info_[mby * dp_.w_mbs + mbx].kind >= H264S_P16x16 && info_[mby * dp_.w_mbs + mbx].kind <= H264S_PSKIP; }
    uint8_t avail_flags(int addr, const MbInfo* info, const h264s_mb& m) const
    {
        (void)addr; (void)info;
        const bool intra = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
        const bool cip = dp_.constrained_intra_pred && intra && !dp_.ref_quirks;   // the reference ignores the flag (h264/macroblock.cpp:1455)
        auto ok = [&](int dx, int dy) { return mb_avail(mbx_ + dx, mby_ + dy) && !(cip && is_inter(mbx_ + dx, mby_ + dy)); };
        uint8_t f = 0;
        if (ok(-1, 0)) f |= H264R_AVAIL_A;
        if (ok(0, -1)) f |= H264R_AVAIL_B;
        if (ok(1, -1)) f |= H264R_AVAIL_C;
        if (ok(-1, -1)) f |= H264R_AVAIL_D;
        if (m.kind == H264S_PSKIP) f |= H264R_SKIP;
        return f;
    }
    // 4x4-block grids, coordinates relative to the current macroblock (may reach into the neighbours)
    size_t at(int bx, int by) const { return (size_t)(mby_ * 4 + by) * w4_ + (size_t)(mbx_ * 4 + bx); }
    int16_t* mv_at(int bx, int by) { return &mv_[at(bx, by) * 2]; }
    int8_t& ref_at(int bx, int by) { return ref_[at(bx, by)]; }
    int8_t& imode_at(int bx, int by) { return imode_[at(bx, by)]; }
    void fill_ref(int bx, int by, int bw, int bh, int8_t v)
    {
        for (int y = by; y < by + bh; y++)
            for (int x = bx; x < bx + bw; x++) ref_at(x, y) = v;
    }
    void fill_imode(int bx, int by, int n, int8_t v)
    {
        for (int y = by; y < by + n; y++)
            for (int x = bx; x < bx + n; x++) imode_at(x, y) = v;
    }
    void fill_mv(int bx, int by, int bw, int bh, int8_t ref, int x, int y)
    {
        for (int j = by; j < by + bh; j++)
            for (int i = bx; i < bx + bw; i++) {
                ref_at(i, j) = ref;
                mv_at(i, j)[0] = (int16_t)x;
                mv_at(i, j)[1] = (int16_t)y;
            }
    }
    // is the 4x4 block at (bx, by) relative to the current macroblock usable as a neighbour?
    bool blk_avail(int bx, int by)
    {
        const int X = mbx_ * 4 + bx, Y = mby_ * 4 + by;
        if (X < 0 || Y < 0 || X >= w4_ || Y >= h4_) return false;
        if (!mb_avail(X >> 2, Y >> 2)) return false;
        return ref_[(size_t)Y * w4_ + X] != kNotYet;
    }

    // ---- intra prediction modes (8.3.1.1 / 8.3.2.1) --------------------------------------------------------------------------------
    int predicted_intra_mode(int bx, int by)
    {
        int mode[2];
        const int nx[2] = {bx - 1, bx}, ny[2] = {by, by - 1};
        for (int k = 0; k < 2; k++) {
            const int X = mbx_ * 4 + nx[k], Y = mby_ * 4 + ny[k];
            if (X < 0 || Y < 0 || !mb_avail(X >> 2, Y >> 2)) return 2;                 // dcPredModePredictedFlag
            const int8_t v = imode_[(size_t)Y * w4_ + X];
            if (v == kInter) {
                if (dp_.constrained_intra_pred || dp_.ref_quirks) return 2;            // quirk D11: as if the flag were always set
                mode[k] = 2;
            } else mode[k] = v == kNotINxN ? 2 : v;
        }
        return mode[0] < mode[1] ? mode[0] : mode[1];
    }
    int intra_mode(int bx, int by, int prev_flag, int rem)
    {
        const int pred = predicted_intra_mode(bx, by);
        if (prev_flag) return pred;
        return rem < pred ? rem : rem + 1;
    }

    // ---- motion vectors --------------------------------------------------------------------------------------------------------------
    struct Nb { int ref; int mv[2]; bool avail; };
    Nb neighbour(int bx, int by)
    {
        Nb n;
        n.avail = blk_avail(bx, by);
        n.ref = -1; n.mv[0] = n.mv[1] = 0;
        if (n.avail) {
            const int8_t r = ref_at(bx, by);
            if (r >= 0) { n.ref = r; n.mv[0] = mv_at(bx, by)[0]; n.mv[1] = mv_at(bx, by)[1]; }
        }
        return n;
    }
    // 8.4.1.3: prediction for the partition whose top-left 4x4 block is (bx, by), bw x bh blocks; dir: 0 median, 1 prefer B, 2 prefer A,
    // 3 prefer C (the directional rules of 16x8 / 8x16)
    void predict_spec(int bx, int by, int bw, int ref, int dir, int out[2])
    {
        const Nb a = neighbour(bx - 1, by);
        Nb b = neighbour(bx, by - 1);
        Nb c = neighbour(bx + bw, by - 1);
        if (!c.avail) c = neighbour(bx - 1, by - 1);
        if (dir == 1 && b.ref == ref) { out[0] = b.mv[0]; out[1] = b.mv[1]; return; }
        if (dir == 2 && a.ref == ref) { out[0] = a.mv[0]; out[1] = a.mv[1]; return; }
        if (dir == 3 && c.ref == ref) { out[0] = c.mv[0]; out[1] = c.mv[1]; return; }
        if (!b.avail && !c.avail && a.avail) { b = a; c = a; }
        const int match = (a.ref == ref) + (b.ref == ref) + (c.ref == ref);
        if (match == 1) {
            const Nb& s = a.ref == ref ? a : (b.ref == ref ? b : c);
            out[0] = s.mv[0]; out[1] = s.mv[1];
            return;
        }
        out[0] = median(a.mv[0], b.mv[0], c.mv[0]);
        out[1] = median(a.mv[1], b.mv[1], c.mv[1]);
    }

    // ---- the reference's own neighbour lookup (ref quirks) ---------------------------------------------------------------------------
    struct QuirkMb {
        uint8_t kind = 255;        // H264S_*; 255 = not decoded
        uint8_t sub[4] = {0, 0, 0, 0};
        int8_t ref[4] = {0, 0, 0, 0};
        int16_t mv[4][4][2];       // [mbPartIdx][subMbPartIdx] as the reference stores them (zero until decoded)
        QuirkMb() { memset(mv, 0, sizeof(mv)); }
    };
    static int part_w(int kind) { return kind == H264S_P16x16 || kind == H264S_PSKIP || kind == H264S_P16x8 ? 16 : 8; }
    static int part_h(int kind) { return kind == H264S_P16x16 || kind == H264S_PSKIP || kind == H264S_P8x16 ? 16 : 8; }
    static int sub_w(int st) { return st == 0 || st == 1 ? 8 : 4; }
    static int sub_h(int st) { return st == 0 || st == 2 ? 8 : 4; }
    struct QNb { const QuirkMb* mb; bool avail; int part, sub; int ref; int mv[2]; };
    // picture::get_PartNeighbour (h264/picture.cpp:205-278) + the lambda f_1 of get_MvNeighbour (h264/picture.cpp:289-320)
    QNb quirk_neighbour(const QuirkMb& cur, int part, int sub, char dir)
    {
        const int pw = part_w(cur.kind), ph = part_h(cur.kind);
        const int per_row = 16 / pw;
        const int x = (part % per_row) * pw, y = (part / per_row) * ph;
        int xs = 0, ys = 0;
        if (cur.kind == H264S_P8x8) { xs = (sub % 8) * 8; ys = (sub / 8) * 8; }       // h264/picture.cpp:227
        int pred_w;
        if (cur.kind == H264S_PSKIP) pred_w = 16;
        else if (cur.kind == H264S_P8x8) pred_w = sub_w(cur.sub[part]);
        else pred_w = pw;
        int xd = 0, yd = 0;
        switch (dir) {
        case 'A': xd = -1; yd = 0; break;
        case 'B': xd = 0; yd = -1; break;
        case 'C': xd = pred_w; yd = -1; break;
        default: xd = -1; yd = -1; break;
        }
        const int xn = (int8_t)(x + xs + xd), yn = (int8_t)(y + ys + yd);
        int dx = 0, dy = 0;
        bool none = false;
        if (xn >= 0 && yn >= 0 && xn < 16) { dx = 0; dy = 0; }
        else if (xn < 0 && yn < 0) { dx = -1; dy = -1; }
        else if (xn < 0 && yn >= 0) { dx = -1; dy = 0; }
        else if (xn >= 0 && yn < 0 && xn < 16) { dx = 0; dy = -1; }
        else if (xn >= 16 && yn < 0) { dx = 1; dy = -1; }
        else none = true;
        QNb n;
        n.mb = nullptr; n.avail = false; n.part = -1; n.sub = -1; n.ref = -1; n.mv[0] = n.mv[1] = 0;
        if (none || !mb_avail(mbx_ + dx, mby_ + dy)) return n;
        const QuirkMb& N = qmb_[(size_t)(mby_ + dy) * dp_.w_mbs + (mbx_ + dx)];
        n.mb = &N; n.avail = true;
        const int xw = (xn + 16) % 16, yw = (yn + 16) % 16;
        const bool n_intra = N.kind == H264S_I4x4 || N.kind == H264S_I8x8 || N.kind == H264S_I16x16;
        if (n_intra) { n.part = 1; n.sub = 0; return n; }                             // Index_ToN (h264/picture.cpp:186-198)
        n.part = (16 / part_w(N.kind)) * (yw / part_h(N.kind)) + xw / part_w(N.kind);
        if (N.kind != H264S_P8x8) n.sub = 0;
        else {
            const int st = N.sub[n.part];
            n.sub = (8 / sub_w(st)) * ((yw % 8) / sub_h(st)) + (xw % 8) / sub_w(st);
        }
        n.ref = N.ref[n.part];
        n.mv[0] = N.mv[n.part][n.sub][0];
        n.mv[1] = N.mv[n.part][n.sub][1];
        return n;
    }
    // picture::get_MvNeighbour (h264/picture.cpp:280-394)
    void predict_quirk(int addr, int part, int sub, int ref, int out[2])
    {
        const QuirkMb& cur = qmb_[addr];
        QNb a = quirk_neighbour(cur, part, sub, 'A');
        QNb b = quirk_neighbour(cur, part, sub, 'B');
        QNb c = quirk_neighbour(cur, part, sub, 'C');
        const QNb d = quirk_neighbour(cur, part, sub, 'D');
        auto inter0 = [](const QNb& n) {       // (N->is_interpred() && (N->ref_idx_l0[0] && N->mv_l0[0][0] == (0, 0)))
            const bool intra = n.mb->kind == H264S_I4x4 || n.mb->kind == H264S_I8x8 || n.mb->kind == H264S_I16x16;
            return !intra && n.mb->ref[0] != 0 && n.mb->mv[0][0][0] == 0 && n.mb->mv[0][0][1] == 0;
        };
        if (cur.kind == H264S_PSKIP && (!a.avail || inter0(a) || !b.avail || inter0(b))) { out[0] = out[1] = 0; return; }
        if (!c.avail || c.part == -1 || c.sub == -1) {
            // C <- D: macroblock, indices and vector, but NOT the reference index (h264/picture.cpp:341-348)
            const int keep_ref = c.ref;
            c = d;
            c.ref = keep_ref;
        }
        const int pw = part_w(cur.kind), ph = part_h(cur.kind);
        if (pw == 16 && ph == 8 && part == 0 && b.ref == ref) { out[0] = b.mv[0]; out[1] = b.mv[1]; return; }
        if (((pw == 16 && ph == 8 && part == 1) || (pw == 8 && ph == 16 && part == 0)) && a.ref == ref) { out[0] = a.mv[0]; out[1] = a.mv[1]; return; }
        if (pw == 8 && ph == 16 && part == 1 && c.ref == ref) { out[0] = c.mv[0]; out[1] = c.mv[1]; return; }
        if (!b.avail && !c.avail) {
            b.mv[0] = c.mv[0] = a.mv[0];
            b.mv[1] = c.mv[1] = a.mv[1];
            b.ref = c.ref = a.ref;
        }
        if (a.ref == ref && b.ref != ref && c.ref != ref) { out[0] = a.mv[0]; out[1] = a.mv[1]; }
        else if (a.ref != ref && b.ref == ref && c.ref != ref) { out[0] = b.mv[0]; out[1] = b.mv[1]; }
        else if (a.ref != ref && b.ref != ref && c.ref == ref) { out[0] = c.mv[0]; out[1] = c.mv[1]; }
        else { out[0] = median(a.mv[0], b.mv[0], c.mv[0]); out[1] = median(a.mv[1], b.mv[1], c.mv[1]); }
    }

    void predict_skip(int addr, int16_t out[2])
    {
        int p[2];
        if (dp_.ref_quirks) {
            QuirkMb& q = qmb_[addr];
            q = QuirkMb();
            q.kind = H264S_PSKIP;
            predict_quirk(addr, 0, 0, 0, p);
        } else {
            const Nb a = neighbour(-1, 0), b = neighbour(0, -1);
            if (!a.avail || !b.avail || (a.ref == 0 && a.mv[0] == 0 && a.mv[1] == 0) || (b.ref == 0 && b.mv[0] == 0 && b.mv[1] == 0)) p[0] = p[1] = 0;
            else predict_spec(0, 0, 4, 0, 0, p);
        }
        out[0] = (int16_t)p[0]; out[1] = (int16_t)p[1];
    }

    // One partition: prediction + mvd (or final - prediction when underiving), state update.
    void partition(int addr, const h264s_mb& m, h264s_mb* wr, bool underive, int idx, int part, int sub, int bx, int by, int bw, int bh,
                   int ref, int dir)
    {
        int p[2];
        if (dp_.ref_quirks) predict_quirk(addr, part, sub, ref, p);
        else predict_spec(bx, by, bw, ref, dir, p);
        int fx, fy;
        if (underive) {
            fx = m.mvd[idx][0]; fy = m.mvd[idx][1];
            wr->mvd[idx][0] = (int16_t)(fx - p[0]);
            wr->mvd[idx][1] = (int16_t)(fy - p[1]);
        } else {
            fx = (int16_t)(p[0] + m.mvd[idx][0]); fy = (int16_t)(p[1] + m.mvd[idx][1]);
        }
        fill_mv(bx, by, bw, bh, (int8_t)ref, fx, fy);
        if (dp_.ref_quirks) { qmb_[addr].mv[part][sub][0] = (int16_t)fx; qmb_[addr].mv[part][sub][1] = (int16_t)fy; }
    }

    void inter_mb(int addr, const h264s_mb& m, h264r_mb_hdr* hdr, bool underive, h264s_mb* wr)
    {
        if (dp_.ref_quirks) {
            QuirkMb& q = qmb_[addr];
            q = QuirkMb();
            q.kind = m.kind;
            for (int i = 0; i < 4; i++) { q.sub[i] = m.sub_type[i]; q.ref[i] = (int8_t)m.ref_idx[i]; }
        }
        switch (m.kind) {
        case H264S_PSKIP: {
            int16_t v[2];
            predict_skip(addr, v);
            fill_mv(0, 0, 4, 4, 0, v[0], v[1]);
            if (dp_.ref_quirks) { qmb_[addr].mv[0][0][0] = v[0]; qmb_[addr].mv[0][0][1] = v[1]; }
            hdr->mb_class = H264R_MB_P16x16;
            break;
        }
        case H264S_P16x16:
            partition(addr, m, wr, underive, 0, 0, 0, 0, 0, 4, 4, m.ref_idx[0], 0);
            hdr->mb_class = H264R_MB_P16x16;
            for (int q = 0; q < 4; q++) hdr->u.ref_idx[q] = (int8_t)m.ref_idx[0];
            break;
        case H264S_P16x8:
            partition(addr, m, wr, underive, 0, 0, 0, 0, 0, 4, 2, m.ref_idx[0], 1);
            partition(addr, m, wr, underive, 4, 1, 0, 0, 2, 4, 2, m.ref_idx[1], 2);
            hdr->mb_class = H264R_MB_P16x8;
            for (int q = 0; q < 4; q++) hdr->u.ref_idx[q] = (int8_t)m.ref_idx[q >> 1];
            break;
        case H264S_P8x16:
            partition(addr, m, wr, underive, 0, 0, 0, 0, 0, 2, 4, m.ref_idx[0], 2);
            partition(addr, m, wr, underive, 4, 1, 0, 2, 0, 2, 4, m.ref_idx[1], 3);
            hdr->mb_class = H264R_MB_P8x16;
            for (int q = 0; q < 4; q++) hdr->u.ref_idx[q] = (int8_t)m.ref_idx[q & 1];
            break;
        default:
            hdr->mb_class = H264R_MB_P8x8;
            for (int q = 0; q < 4; q++) {
                const int bx = (q & 1) * 2, by = (q >> 1) * 2, st = m.sub_type[q] & 3, ref = m.ref_idx[q];
                hdr->sub_mb_type |= (uint8_t)(st << (2 * q));
                hdr->u.ref_idx[q] = (int8_t)ref;
                switch (st) {
                case 0: partition(addr, m, wr, underive, q * 4, q, 0, bx, by, 2, 2, ref, 0); break;
                case 1:
                    partition(addr, m, wr, underive, q * 4, q, 0, bx, by, 2, 1, ref, 0);
                    partition(addr, m, wr, underive, q * 4 + 1, q, 1, bx, by + 1, 2, 1, ref, 0);
                    break;
                case 2:
                    partition(addr, m, wr, underive, q * 4, q, 0, bx, by, 1, 2, ref, 0);
                    partition(addr, m, wr, underive, q * 4 + 1, q, 1, bx + 1, by, 1, 2, ref, 0);
                    break;
                default:
                    for (int s = 0; s < 4; s++) partition(addr, m, wr, underive, q * 4 + s, q, s, bx + (s & 1), by + (s >> 1), 1, 1, ref, 0);
                }
            }
        }
    }

    // ---- coefficient levels: scan order -> the engine's block layout ("Coefficients" in include/h264r.h) -----------------------------
    // (most blocks of a coded quadrant are empty: four 64-bit words say so before the sixteen levels are looked at)
    static bool block_is_zero(const int16_t* lv)
    {
        uint64_t q[4];
        memcpy(q, lv, 32);
        return !(q[0] | q[1] | q[2] | q[3]);
    }
    static void scan_levels(const h264s_mb& m, int16_t* coeff)
    {
        memset(coeff, 0, sizeof(int16_t) * H264R_COEFF_PER_MB);
        if (m.kind == H264S_PSKIP) return;
        if (m.kind == H264S_I16x16) {
            for (int k = 0; k < 16; k++)
                if (m.luma_dc[k]) coeff[kBlkRaster[kZigzag4x4[k]] * 16] = m.luma_dc[k];
            if (m.cbp_luma)
                for (int b = 0; b < 16; b++) {
                    if (block_is_zero(m.luma[b])) continue;
                    for (int k = 0; k < 15; k++)
                        if (m.luma[b][k]) coeff[b * 16 + kZigzag4x4[k + 1]] = m.luma[b][k];
                }
        } else if (m.t8x8) {
            for (int q = 0; q < 4; q++) {
                if (!((m.cbp_luma >> q) & 1)) continue;
                const int16_t* lv = m.luma[q * 4];
                for (int k = 0; k < 64; k++)
                    if (lv[k]) coeff[q * 64 + kZigzag8x8[k]] = lv[k];
            }
        } else {
            for (int b = 0; b < 16; b++) {
                if (!((m.cbp_luma >> (b >> 2)) & 1) || block_is_zero(m.luma[b])) continue;
                for (int k = 0; k < 16; k++)
                    if (m.luma[b][k]) coeff[b * 16 + kZigzag4x4[k]] = m.luma[b][k];
            }
        }
        if (m.cbp_chroma)
            for (int c = 0; c < 2; c++)
                for (int b = 0; b < 4; b++) {
                    int16_t* o = coeff + (16 + c * 4 + b) * 16;
                    o[0] = m.chroma_dc[c][b];
                    if (m.cbp_chroma == 2 && !block_is_zero(m.chroma_ac[c][b]))
                        for (int k = 0; k < 15; k++)
                            if (m.chroma_ac[c][b][k]) o[kZigzag4x4[k + 1]] = m.chroma_ac[c][b][k];
                }
    }

    DeriveParams dp_;
    int w4_ = 0, h4_ = 0;
    std::vector<int16_t> mv_;
    std::vector<int8_t> ref_, imode_;
    std::vector<QuirkMb> qmb_;
    const MbInfo* info_ = nullptr;
    int addr_ = 0, mbx_ = 0, mby_ = 0, qp_ = 26;
    uint16_t slice_ = 0;
};

}  // namespace h264host
