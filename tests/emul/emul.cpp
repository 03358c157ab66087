// TEST INFRASTRUCTURE -- host-side SIMT emulator for the kernel code of
// videodecoder_b200/csrc/cuda/h264r_core.cuh + h264r_seq.cuh.
//
// The CUDA kernels are written as per-warp phase sequences (see h264r_seq.cuh).  This file compiles
// the very same headers with g++ and runs each phase with a loop over the 32 lanes, macroblocks in
// raster order (which satisfies every dependency the device wavefront respects).  It lets the CPU
// test-suite compare the kernels' arithmetic with the oracle without a GPU.  It is NOT a CPU
// fallback: nothing in the product loads it.
#include "../../videodecoder_b200/csrc/cuda/h264r_seq.cuh"
#include "../../videodecoder_b200/csrc/cuda/h264r_mcgrp.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace h264r;

namespace {
struct HostExec {
    template <class F> void operator()(F f) { for (int lane = 0; lane < 32; lane++) f(lane); }
};

void fill_desc(PicDesc& pd, const h264r_pic_params* pp, const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
               uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* const* ref_y, const uint8_t* const* ref_u,
               const uint8_t* const* ref_v, uint8_t* exc_map)
{
    memset(&pd, 0, sizeof(pd));
    pd.hdr = hdr; pd.mv = mv; pd.coeff = coeff; pd.y = y; pd.u = u; pd.v = v;
    pd.blk_mask = nullptr; pd.blk_off = nullptr;        // set by use_dense() / use_packed()
    pd.level_fmt = H264R_LEVELS_INT16;
    pd.mv_off = nullptr;
    pd.residual = nullptr;
    pd.nzmask = nullptr;
    pd.tmaps = nullptr;
    pd.dbrec = nullptr;
    pd.mbox = nullptr;
    for (int i = 0; i < H264R_MAX_REFS; i++) {
        pd.ref_y[i] = ref_y ? ref_y[i] : nullptr;
        pd.ref_u[i] = ref_u ? ref_u[i] : nullptr;
        pd.ref_v[i] = ref_v ? ref_v[i] : nullptr;
        pd.ref_id[i] = pp->ref_list0[i];
        pd.luma_weight[i] = pp->luma_weight[i]; pd.luma_offset[i] = pp->luma_offset[i];
        for (int c = 0; c < 2; c++) { pd.chroma_weight[i][c] = pp->chroma_weight[i][c]; pd.chroma_offset[i][c] = pp->chroma_offset[i][c]; }
    }
    pd.slice_type = pp->slice_type;
    pd.weighted_pred = pp->weighted_pred;
    pd.luma_log2_wd = pp->luma_log2_weight_denom;
    pd.chroma_log2_wd = pp->chroma_log2_weight_denom;
    pd.deblock_on = pp->disable_deblocking_filter_idc != 1;
    pd.alpha_off = pp->slice_alpha_c0_offset_div2 * 2;
    pd.beta_off = pp->slice_beta_offset_div2 * 2;
    pd.excursion_map = exc_map;
    static uint32_t exc_counter;
    exc_counter = 0;
    pd.exc_counter = &exc_counter;
}
// coefficient storage of a picture as the device sees it
struct CoeffStore {
    std::vector<uint32_t> mask, off;
    std::vector<int16_t> blocks;
    void use_dense(PicDesc& pd, int n_mbs)
    {
        mask.assign(n_mbs, 0xFFFFFFu); off.resize(n_mbs);
        for (int a = 0; a < n_mbs; a++) off[a] = 24u * a;
        pd.blk_mask = mask.data(); pd.blk_off = off.data();
    }
    // what a host parser would submit through h264r_submit_mbs_packed: only the blocks with a non-zero level
    void use_packed(PicDesc& pd, int n_mbs, const int16_t* dense, bool compact = false)
    {
        mask.assign(n_mbs, 0); off.resize(n_mbs); blocks.clear();
        for (int a = 0; a < n_mbs; a++) {
            off[a] = (uint32_t)(blocks.size() / 16);
            for (int b = 0; b < 24; b++) {
                const int16_t* p = dense + ((size_t)a * 24 + b) * 16;
                bool nz = false;
                for (int k = 0; k < 16; k++) nz = nz || p[k] != 0;
                if (nz) { mask[a] |= 1u << b; blocks.insert(blocks.end(), p, p + 16); }
            }
        }
        blocks.resize(blocks.size() + 16);       // keeps data() non-null for all-zero pictures
        pd.coeff = blocks.data(); pd.blk_mask = mask.data(); pd.blk_off = off.data();
        // int8 transport when every level fits (h264r_picture_buf::level_bytes = 1)
        bool fits = true;
        for (int16_t v : blocks) fits = fits && v >= -128 && v <= 127;
        if (fits) {
            narrow.assign(blocks.begin(), blocks.end());
            pd.coeff = reinterpret_cast<const int16_t*>(narrow.data());
            pd.level_fmt = H264R_LEVELS_INT8;
        }
        if (fits && compact) {
            // H264R_LEVELS_COMPACT: 8-byte units; macroblocks whose blocks all have at most six levels store a significance
            // map + the levels per block (one unit), the others int8 blocks (two units)
            units.clear();
            size_t blk = 0;
            for (int a = 0; a < n_mbs; a++) {
                const int nb = __builtin_popcount(mask[a]);
                bool ok = nb > 0;
                for (int b = 0; b < nb; b++) {
                    int nz = 0;
                    for (int k = 0; k < 16; k++) nz += blocks[(blk + b) * 16 + k] != 0;
                    ok = ok && nz <= 6;
                }
                off[a] = (uint32_t)(units.size() / 8);
                for (int b = 0; b < nb; b++) {
                    const int16_t* p = &blocks[(blk + b) * 16];
                    if (ok) {
                        uint8_t u[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        int n = 0;
                        for (int k = 0; k < 16; k++) if (p[k]) { u[k >> 3] |= (uint8_t)(1u << (k & 7)); u[2 + n++] = (uint8_t)(int8_t)p[k]; }
                        units.insert(units.end(), u, u + 8);
                    } else
                        for (int k = 0; k < 16; k++) units.push_back((uint8_t)(int8_t)p[k]);
                }
                if (ok) mask[a] |= H264R_BLK_COMPACT;
                blk += nb;
            }
            units.resize(units.size() + 16);
            pd.coeff = reinterpret_cast<const int16_t*>(units.data());
            pd.level_fmt = H264R_LEVELS_COMPACT;
        }
    }
    std::vector<int8_t> narrow;
    std::vector<uint8_t> units;
};
// motion vectors in partition order, as h264r_submit_picture(..., n_mvs >= 0) delivers them
struct MvStore {
    std::vector<uint32_t> off;
    std::vector<int16_t> mvs;
    void use_packed(PicDesc& pd, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* dense)
    {
        off.resize(n_mbs); mvs.clear();
        for (int a = 0; a < n_mbs; a++) {
            off[a] = (uint32_t)(mvs.size() / 2);
            const int n = mv_count(hdr[a].mb_class, hdr[a].sub_mb_type);
            for (int k = 0; k < n; k++) {
                int blk = 0;        // first 4x4 block (raster) covered by partition k
                while (mv_part_index(hdr[a].mb_class, hdr[a].sub_mb_type, blk) != k) blk++;
                mvs.push_back(dense[(size_t)a * 32 + blk * 2]); mvs.push_back(dense[(size_t)a * 32 + blk * 2 + 1]);
            }
        }
        // the stream grows backward from pd.mv: vector i at pd.mv[-2(i+1)]
        std::vector<int16_t> rev(mvs.size() + 2, 0);
        const size_t nv = mvs.size() / 2;
        for (size_t i = 0; i < nv; i++) { rev[2 * (nv - 1 - i)] = mvs[2 * i]; rev[2 * (nv - 1 - i) + 1] = mvs[2 * i + 1]; }
        mvs.swap(rev);
        pd.mv = mvs.data() + 2 * nv; pd.mv_off = off.data();
    }
};
}  // namespace

extern "C" {

// Stream form (H264R_HDR_STREAM): what k_blk_scan + k_stream_expand leave in a slot, macroblocks in order -- the scans as running
// sums of stream_mb_counts, the expansion by stream_expand_mb (the functions the kernels call).  mv_end: end of a buffer of at
// least max_mvs words (vector i at mv_end[-1 - i]).  Returns the words walked (strict: -1 - a if macroblock a is not as long as mb_words[a] says; else minus the number of such
// macroblocks); totals through n_units / n_mvs / n_intra.
long h264emul_stream_expand(int n_mbs, const uint32_t* hdr8, const uint8_t* mb_words, const uint32_t* stream, h264r_mb_hdr* hdr16,
                            uint32_t* mask, uint32_t* mv_end, uint32_t* units, uint32_t* blk_off, uint32_t* mv_off, uint32_t* intra_list,
                            long* n_units, long* n_mvs, long* n_intra, int strict)
{
    uint32_t wo = 0, uo = 0, mo = 0, ni = 0;
    long bad = 0;
    for (int a = 0; a < n_mbs; a++) {
        const uint2 s8 = uint2{hdr8[2 * a], hdr8[2 * a + 1]};
        const uint32_t c = stream_mb_counts(s8, stream_mb_has_mask(s8) ? stream[wo] : 0u);
        blk_off[a] = uo; mv_off[a] = mo;
        if (c >> 16) intra_list[ni++] = (uint32_t)a;
        const uint32_t walked = stream_expand_mb(s8, stream + wo, uo, mo, reinterpret_cast<uint4*>(hdr16 + a), mask + a, mv_end, reinterpret_cast<uint2*>(units));
        if (walked != mb_words[a]) {                       // what k_stream_expand reports as H264R_STATUS_STREAM
            if (strict) return -1 - a;
            bad++;                                         // (the device goes on: whatever it reads, it stays inside the slot)
        }
        wo += mb_words[a]; uo += c & 255u; mo += (c >> 8) & 255u;
    }
    *n_units = uo; *n_mvs = mo; *n_intra = ni;
    return bad ? -bad : (long)wo;
}

// residual of every MB: out[n_mbs][384] int16
void h264emul_residual(int w_mbs, int h_mbs, unsigned flags, const h264r_mb_hdr* hdr, const int16_t* coeff, int16_t* out)
{
    h264r_pic_params pp;
    memset(&pp, 0, sizeof(pp));
    PicDesc pd;
    fill_desc(pd, &pp, hdr, nullptr, coeff, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    CoeffStore cs;
    if (flags & 0x40000000u) cs.use_packed(pd, w_mbs * h_mbs, coeff); else cs.use_dense(pd, w_mbs * h_mbs);
    HostExec ex;
    MbWork w;
    for (int a = 0; a < w_mbs * h_mbs; a++) {
        if (flags & H264R_REF_COMPAT) seq_residual<true>(ex, w, pd, a, out + (size_t)a * 384);
        else seq_residual<false>(ex, w, pd, a, out + (size_t)a * 384);
    }
}

// A frame of the emulated frame pool: planes with the replicated border of the device frame pool.
struct PaddedFrame {
    int W, H, py, pc;
    std::vector<uint8_t> buf_y, buf_u, buf_v;
    PaddedFrame(int W_, int H_) : W(W_), H(H_), py(W_ + 2 * PAD_X), pc(W_ / 2 + 2 * PAD_XC),
        buf_y((size_t)py * (H_ + 2 * PAD_Y)), buf_u((size_t)pc * (H_ / 2 + 2 * PAD_YC)), buf_v(buf_u.size()) {}
    uint8_t* y() { return buf_y.data() + (size_t)PAD_Y * py + PAD_X; }
    uint8_t* u() { return buf_u.data() + (size_t)PAD_YC * pc + PAD_XC; }
    uint8_t* v() { return buf_v.data() + (size_t)PAD_YC * pc + PAD_XC; }
    void load(const uint8_t* sy, const uint8_t* su, const uint8_t* sv)
    {
        for (int r = 0; r < H; r++) memcpy(y() + (size_t)r * py, sy + (size_t)r * W, W);
        for (int r = 0; r < H / 2; r++) { memcpy(u() + (size_t)r * pc, su + (size_t)r * (W / 2), W / 2); memcpy(v() + (size_t)r * pc, sv + (size_t)r * (W / 2), W / 2); }
    }
    void store(uint8_t* dy, uint8_t* du, uint8_t* dv)
    {
        for (int r = 0; r < H; r++) memcpy(dy + (size_t)r * W, y() + (size_t)r * py, W);
        for (int r = 0; r < H / 2; r++) { memcpy(du + (size_t)r * (W / 2), u() + (size_t)r * pc, W / 2); memcpy(dv + (size_t)r * (W / 2), v() + (size_t)r * pc, W / 2); }
    }
    void pad()      // what k_pad does on the device
    {
        for (int r = -PAD_Y; r < H + PAD_Y; r++) for (int lane = 0; lane < 32; lane++) pad_row_phase(y(), py, W, H, PAD_X, PAD_Y, r, lane);
        for (int r = -PAD_YC; r < H / 2 + PAD_YC; r++)
            for (int lane = 0; lane < 32; lane++) {
                pad_row_phase(u(), pc, W / 2, H / 2, PAD_XC, PAD_YC, r, lane);
                pad_row_phase(v(), pc, W / 2, H / 2, PAD_XC, PAD_YC, r, lane);
            }
    }
};

// whole picture; caller's planes have pitch w_mbs*16 / w_mbs*8 (copied into / out of padded frames).
// Returns the number of luma excursions.
unsigned long long h264emul_recon_picture(int w_mbs, int h_mbs, unsigned flags, const h264r_pic_params* pp,
                                          const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
                                          uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* const* ref_y,
                                          const uint8_t* const* ref_u, const uint8_t* const* ref_v, uint8_t* exc_map)
{
    const int W = w_mbs * 16, H = h_mbs * 16;
    PaddedFrame cur(W, H);
    std::vector<PaddedFrame> refs;
    const uint8_t* ry[H264R_MAX_REFS] = {nullptr}; const uint8_t* ru[H264R_MAX_REFS] = {nullptr}; const uint8_t* rv[H264R_MAX_REFS] = {nullptr};
    refs.reserve(H264R_MAX_REFS);
    for (int i = 0; i < H264R_MAX_REFS; i++) {
        if (!ref_y || !ref_y[i]) continue;
        refs.emplace_back(W, H);
        refs.back().load(ref_y[i], ref_u[i], ref_v[i]);
        refs.back().pad();
        ry[i] = refs.back().y(); ru[i] = refs.back().u(); rv[i] = refs.back().v();
    }
    for (int i = 0; i < H264R_MAX_REFS; i++) if (!ry[i] && ry[0]) { ry[i] = ry[0]; ru[i] = ru[0]; rv[i] = rv[0]; }
    Geometry g = make_geometry(w_mbs, h_mbs, cur.py, cur.pc);
    PicDesc pd;
    fill_desc(pd, pp, hdr, mv, coeff, cur.y(), cur.u(), cur.v(), ry, ru, rv, exc_map);
    CoeffStore cs;
    if (flags & 0x40000000u) cs.use_packed(pd, w_mbs * h_mbs, coeff, (flags & 0x20000000u) != 0); else cs.use_dense(pd, w_mbs * h_mbs);      // test switches: bits 30, 29
    MvStore ms;
    if ((flags & 0x40000000u) && mv) ms.use_packed(pd, w_mbs * h_mbs, hdr, mv);
    HostExec ex;
    WarpWork ww;
    MbWork& w = ww.mb;
    std::vector<uint32_t> nzmask((size_t)w_mbs * h_mbs, 0xdeadbeefu);
    pd.nzmask = nzmask.data();
    unsigned long long exc = 0;
    const bool compat = (flags & H264R_REF_COMPAT) != 0;
    const bool use_fast_intra = !(flags & 0x80000000u);     // test switch: bit 31 selects the generic intra phases
    int n = w_mbs * h_mbs;
    // same order as the device: all inter MBs of the picture first, then the intra MBs
    if ((flags & 0x10000000u) && mv && !(compat && pd.level_fmt == H264R_LEVELS_COMPACT && pd.mv_off)) {
        fprintf(stderr, "the group sequence needs REF_COMPAT, compact levels and partition-order vectors\n");
        abort();
    }
    if ((flags & 0x10000000u) && mv) {
        // test switch bit 28: the group sequence of k_inter_grp (h264r_mcgrp.cuh), GRP_MBS macroblocks per step
        static GroupWork gw;
        // bit 26: the tensor-copy variant's staging (aligned words, shifted by the consumer)
        for (int a0 = 0; a0 < n; a0 += GRP_MBS) {
            if (flags & 0x04000000u) seq_inter_group<true, true>(ex, gw, pd, g, a0, n, []() {});
            else seq_inter_group<true, false>(ex, gw, pd, g, a0, n, []() {});
        }
    } else
    for (int a = 0; a < n; a++) {
        h264r_mb_hdr h = load_hdr(pd, a);
        if (is_intra(h.mb_class)) continue;
        if (compat) seq_inter_any<true, false>(ex, ww, pd, g, h, a);
        else if (flags & 0x08000000u) {
            // test switch bit 27: the three spec-mode parts of k_inter (luma without / with the 8x8 transform, chroma), one after the other
            const MbPre pre = load_pre(pd, a);
            seq_inter_any<false, false, false, PART_LUMA4>(ex, ww, pd, g, h, a, pre);
            seq_inter_any<false, false, false, PART_LUMA8>(ex, ww, pd, g, h, a, pre);
            seq_inter_any<false, false, false, PART_CHROMA>(ex, ww, pd, g, h, a, pre);
        } else seq_inter_any<false, false>(ex, ww, pd, g, h, a);
    }
    // as on the device: the flat kernels leave the residual of the intra macroblocks in the picture's scratch
    std::vector<int16_t> scratch((size_t)n * 384);
    if (use_fast_intra) {
        pd.residual = scratch.data();
        for (int a = 0; a < n; a++) {
            h264r_mb_hdr h = load_hdr(pd, a);
            if (!is_intra(h.mb_class)) continue;
            if (compat) seq_residual_scratch<true>(ex, w, pd, h, a); else seq_residual_scratch<false>(ex, w, pd, h, a);
        }
    }
    static uint2 lut[INTRA4_LUT_SIZE];
    intra4_lut_fill(lut, 0, 1);
    for (int xf = 0; xf < 4; xf++)
        for (int yf = 0; yf < 4; yf++)
            if (mc_need(xf, yf) != mc_select(xf, yf).need) { fprintf(stderr, "mc_need table disagrees with mc_select at (%d,%d)\n", xf, yf); abort(); }
    int prev = -2;
    for (int a = 0; a < n; a++) {
        h264r_mb_hdr h = load_hdr(pd, a);
        if (!is_intra(h.mb_class)) continue;
        // as the row warps of k_intra_rows do: the left column comes from the warp's own tile when it has just done a - 1
        bool carry = use_fast_intra && prev == a - 1 && (a % w_mbs) != 0;
        if (!use_fast_intra) { if (compat) seq_intra_mb<true>(ex, w, pd, g, h, a); else seq_intra_mb<false>(ex, w, pd, g, h, a); }
        else {
            memcpy(w.res, pd.residual + (size_t)a * 384, 768);
            if (compat) seq_intra_any<true>(ex, w, lut, pd, g, h, a, carry ? w.colx : nullptr, true); else seq_intra_any<false>(ex, w, lut, pd, g, h, a, carry ? w.colx : nullptr, true);
        }
        prev = a;
    }
    if (!compat && pd.deblock_on) {
        if (use_fast_intra) {       // k_deblock_prep, then the row kernel's fast path (h264r_deblock.cuh), fed as k_deblock_rows feeds it
            std::vector<uint8_t> dbrec((size_t)n * DBREC_BYTES);
            pd.dbrec = dbrec.data();
            for (int a = 0; a < n; a++) {
                int mbx, mby;
                mb_xy(g, a, mbx, mby);
                h264r_mb_hdr h = load_hdr(pd, a), hl = mbx > 0 ? load_hdr(pd, a - 1) : h, ht = mby > 0 ? load_hdr(pd, a - w_mbs) : h;
                for (int lane = 0; lane < 32; lane++)
                    phase_deblock_prep(pd.dbrec + (size_t)a * DBREC_BYTES, pd, g, h, hl, ht, pd.nzmask[a], mbx > 0 ? pd.nzmask[a - 1] : 0u,
                                       mby > 0 ? pd.nzmask[a - w_mbs] : 0u, a, mbx, mby, lane);
            }
            static DeblockWork dw;
            static PerLane<DeblockIn> in;
            // hand-over records of the wavefront (rows run one after the other here, so a record is complete when it is read: the
            // poll checks the stamp once and aborts if the protocol left a word unwritten)
            std::vector<uint32_t> mbox((size_t)n * MBOX_WORDS, 0u);
            pd.mbox = mbox.data();
            const uint32_t stamp = 7;
            for (int a = 0; a < n; a++) {
                int mbx, mby;
                mb_xy(g, a, mbx, mby);
                for (int lane = 0; lane < 32; lane++) in(lane) = deblock_fetch(pd, g, pd.dbrec, a, mbx, mby, lane);
                auto poll = [&](int lane) -> uint32_t {
                    if (lane >= 24) return 0u;
                    const uint2 v = mbox_load(pd.mbox + (size_t)(a - w_mbs) * MBOX_WORDS + 2 * lane);
                    if (v.y != stamp) { fprintf(stderr, "deblock hand-over: macroblock %d word %d was never published\n", a - w_mbs, lane); abort(); }
                    return v.x;
                };
                seq_deblock_mb_fast(ex, dw, pd, g, in, mbx > 0, mbx, mby, stamp, poll);
            }
        } else
            for (int a = 0; a < n; a++) seq_deblock_mb(ex, w, pd, g, a);
    }
    if (compat && use_fast_intra) {
        // REF_COMPAT: the warps of the border macroblocks wrote the luma border themselves (phase_pad_mb_luma, device
        // sequences only: the generic intra phases are not run by any kernel in this mode); it must be what k_pad would
        // have written
        PaddedFrame chk = cur;
        chk.pad();
        if (chk.buf_y != cur.buf_y) { fprintf(stderr, "fused border replication disagrees with pad_row_phase\n"); abort(); }
    }
    cur.store(y, u, v);
    exc = *pd.exc_counter;
    return exc;
}

void h264emul_digest(const uint8_t* y, const uint8_t* u, const uint8_t* v, int w, int h, uint32_t out[4])
{
    uint32_t d[4] = {0, 0, 0, 0}, i = 0;
    auto word = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
    for (int k = 0; k < w * h / 4; k++) digest_accumulate(d, word(y + 4 * k), i++);
    for (int k = 0; k < w * h / 16; k++) digest_accumulate(d, word(u + 4 * k), i++);
    for (int k = 0; k < w * h / 16; k++) digest_accumulate(d, word(v + 4 * k), i++);
    memcpy(out, d, 16);
}
}
