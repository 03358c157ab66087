// h264_host.cpp -- the C interface of include/h264_host.h: the serial host side of the decoder (Annex-B -> per-picture buffers of
// include/h264r.h) and the CABAC stream writer.  The pieces: bits.h (NAL / RBSP / Exp-Golomb), h264_headers.h (SPS, PPS, slice
// header), cabac.h + cabac_tables.h (arithmetic coding engines, context initialisation), mb_codec.h (macroblock layer, one walker
// for both directions), h264_derive.h (vectors, intra modes, QP, scans), h264_dpb.h (frame slots, reference lists, output order).
//
// Reference call structure this stands in for: main.cpp:20-45 -> NAL::decode (NAL.cpp:14-30) -> decode_PIC (NAL.cpp:185-316) ->
// Slice::PraseSliceDataer (slice.cpp:278-389) -> macroblock::Parse (macroblock.cpp:143-475).  Where the reference calls
// macroblock::Calc per macroblock, this code appends the macroblock to the picture's buffers; reconstruction happens behind
// h264r_submit_* on the GPU.
#include <chrono>
#include <string>
#include <vector>

#include "../../../include/h264_host.h"
#include "../../../include/h264r_writer.h"
#include "bits.h"
#include "cabac.h"
#include "cavlc.h"
#include "h264_derive.h"
#include "h264_dpb.h"
#include "h264_headers.h"
#include "mb_codec.h"

using namespace h264host;

// =====================================================================================================================================
// stream writer
// =====================================================================================================================================
namespace {

void seq_to_sets(const h264s_seq& s, Sps* sps, Pps* pps)
{
    *sps = Sps();
    sps->profile_idc = s.profile_idc;
    sps->level_idc = 51;
    sps->log2_max_frame_num = s.log2_max_frame_num_minus4 + 4;
    sps->poc_type = s.pic_order_cnt_type;
    sps->log2_max_poc_lsb = s.log2_max_poc_lsb_minus4 + 4;
    sps->max_num_ref_frames = s.max_num_ref_frames;
    sps->w_mbs = s.w_mbs;
    sps->h_mbs = s.h_mbs;
    sps->valid = true;
    *pps = Pps();
    pps->entropy_coding_mode_flag = s.profile_idc == 66 ? 0 : 1;       // Baseline profile: CAVLC (it has no CABAC)
    if (s.profile_idc == 66) sps->constraint_flags = 0x80;              // constraint_set0_flag
    pps->num_ref_idx_l0_default = s.num_ref_idx_l0_default_active_minus1 + 1;
    pps->weighted_pred_flag = s.weighted_pred_flag;
    pps->pic_init_qp = s.pic_init_qp;
    pps->chroma_qp_index_offset = s.chroma_qp_index_offset;
    pps->second_chroma_qp_index_offset = Sps::is_high(s.profile_idc) ? s.second_chroma_qp_index_offset : s.chroma_qp_index_offset;
    pps->deblocking_filter_control_present_flag = s.deblocking_filter_control_present_flag;
    pps->constrained_intra_pred_flag = s.constrained_intra_pred_flag;
    pps->transform_8x8_mode_flag = Sps::is_high(s.profile_idc) ? s.transform_8x8_mode_flag : 0;
    pps->valid = true;
}

void pic_to_header(const h264s_seq& s, const h264s_pic& p, SliceHeader* h)
{
    *h = SliceHeader();
    const bool is_p = p.slice_type == 0;
    h->nal_unit_type = p.is_idr ? kNalIdr : kNalSlice;
    h->nal_ref_idc = p.nal_ref_idc;
    h->slice_type = is_p ? kSliceP : kSliceI;
    h->slice_type_raw = is_p ? 5 : 7;                    // the reference needs 5..9 (h264/slice.cpp:19)
    h->frame_num = p.frame_num;
    h->idr_pic_id = p.idr_pic_id;
    h->poc_lsb = p.poc_lsb;
    h->num_ref_idx_l0 = is_p ? p.num_ref_idx_l0_active_minus1 + 1 : 0;
    h->luma_log2_weight_denom = p.luma_log2_weight_denom;
    h->chroma_log2_weight_denom = p.chroma_log2_weight_denom;
    for (int i = 0; i < 16; i++) {
        h->luma_weight_flag[i] = p.luma_weight_flag[i];
        h->luma_weight[i] = p.luma_weight[i];
        h->luma_offset[i] = p.luma_offset[i];
        h->chroma_weight_flag[i] = p.chroma_weight_flag[i];
        for (int c = 0; c < 2; c++) { h->chroma_weight[i][c] = p.chroma_weight[i][c]; h->chroma_offset[i][c] = p.chroma_offset[i][c]; }
    }
    h->n_mods = p.n_mods < 16 ? p.n_mods : 16;
    for (int i = 0; i < h->n_mods; i++) { h->mods[i].idc = p.mods[i][0]; h->mods[i].value = p.mods[i][1]; }
    h->n_mmco = p.n_mmco < 16 ? p.n_mmco : 16;
    for (int i = 0; i < h->n_mmco; i++) { h->mmco[i].op = p.mmco[i][0]; h->mmco[i].a = p.mmco[i][1]; h->mmco[i].b = p.mmco[i][2]; }
    h->long_term_reference_flag = p.long_term_reference_flag;
    h->cabac_init_idc = p.cabac_init_idc;
    h->slice_qp_delta = p.slice_qp_delta;
    h->disable_deblocking_filter_idc = p.disable_deblocking_filter_idc;
    h->alpha_c0_offset_div2 = p.slice_alpha_c0_offset_div2;
    h->beta_offset_div2 = p.slice_beta_offset_div2;
    (void)s;
}

bool any_nonzero(const int16_t* p, int n)
{
    for (int i = 0; i < n; i++)
        if (p[i]) return true;
    return false;
}

// makes a record say only what the syntax can carry (see h264e_encode_picture in h264_host.h)
int normalise_mb(h264s_mb& m, bool p_slice, bool t8_mode, bool absolute)
{
    const bool intra = m.kind == H264S_I4x4 || m.kind == H264S_I8x8 || m.kind == H264S_I16x16;
    if (m.kind > H264S_I8x8) return kErrState;
    if (!p_slice && !intra) return kErrState;
    if (m.kind == H264S_I8x8 && !t8_mode) return kErrState;
    if (m.kind == H264S_PSKIP) {
        const uint8_t k = m.kind;
        memset(&m, 0, sizeof(m));
        m.kind = k;
        return kOk;
    }
    if (m.kind == H264S_I8x8) m.t8x8 = 1;
    if (m.kind == H264S_I4x4 || m.kind == H264S_I16x16 || !t8_mode) m.t8x8 = 0;
    if (m.kind == H264S_P8x8 && (m.sub_type[0] | m.sub_type[1] | m.sub_type[2] | m.sub_type[3])) m.t8x8 = 0;
    if (m.kind == H264S_I16x16) {
        bool ac = false;
        for (int b = 0; b < 16; b++) { m.luma[b][15] = 0; ac |= any_nonzero(m.luma[b], 15); }
        if (ac) m.cbp_luma = 15;
        else if (m.cbp_luma != 15) m.cbp_luma = 0;
        if (!m.cbp_luma) memset(m.luma, 0, sizeof(m.luma));
    } else {
        memset(m.luma_dc, 0, sizeof(m.luma_dc));
        for (int q = 0; q < 4; q++) {
            if (!((m.cbp_luma >> q) & 1)) memset(m.luma[q * 4], 0, sizeof(int16_t) * 64);
            else if (m.t8x8 && !any_nonzero(m.luma[q * 4], 64)) m.cbp_luma &= (uint8_t)~(1u << q);
        }
        m.cbp_luma &= 15;
        if (!m.cbp_luma && !intra) m.t8x8 = 0;
    }
    if (m.cbp_chroma > 2) m.cbp_chroma = 2;
    if (m.cbp_chroma == 0) memset(m.chroma_dc, 0, sizeof(m.chroma_dc));
    if (m.cbp_chroma < 2) memset(m.chroma_ac, 0, sizeof(m.chroma_ac));
    for (int c = 0; c < 2; c++)
        for (int b = 0; b < 4; b++) m.chroma_ac[c][b][15] = 0;
    if (!(m.cbp_luma || m.cbp_chroma || m.kind == H264S_I16x16)) m.qp_delta = 0;
    // prediction syntax the macroblock type does not carry
    if (!intra) {
        m.chroma_mode = 0; m.i16_mode = 0;
        memset(m.prev_flag, 0, sizeof(m.prev_flag));
        memset(m.rem_mode, 0, sizeof(m.rem_mode));
        bool used[16] = {false};
        if (m.kind == H264S_P16x16) used[0] = true;
        else if (m.kind == H264S_P16x8 || m.kind == H264S_P8x16) used[0] = used[4] = true;
        else
            for (int q = 0; q < 4; q++) {
                static const int cnt[4] = {1, 2, 2, 4};
                for (int k = 0; k < cnt[m.sub_type[q] & 3]; k++) used[q * 4 + k] = true;
            }
        for (int i = 0; i < 16; i++)
            if (!used[i]) m.mvd[i][0] = m.mvd[i][1] = 0;
        if (m.kind != H264S_P8x8) {
            memset(m.sub_type, 0, sizeof(m.sub_type));
            const int np = m.kind == H264S_P16x16 ? 1 : 2;
            for (int i = np; i < 4; i++) m.ref_idx[i] = 0;
        }
    } else {
        memset(m.sub_type, 0, sizeof(m.sub_type));
        memset(m.ref_idx, 0, sizeof(m.ref_idx));
        memset(m.mvd, 0, sizeof(m.mvd));
        if (m.kind != H264S_I16x16) m.i16_mode = 0;
        const int nb = m.kind == H264S_I4x4 ? 16 : (m.kind == H264S_I8x8 ? 4 : 0);
        for (int i = 0; i < 16; i++) {
            if (i >= nb) { m.prev_flag[i] = 0; m.rem_mode[i] = 0; }
            else if (absolute) { m.prev_flag[i] = 0; if (m.rem_mode[i] > 8) return kErrState; }     // rem_mode holds the final mode
            else if (m.prev_flag[i]) { m.prev_flag[i] = 1; m.rem_mode[i] = 0; }
            else m.rem_mode[i] &= 7;
        }
    }
    m.pad_ = 0;
    return kOk;
}

}  // namespace

extern "C" long h264e_write_sps_pps(const h264s_seq* s, uint8_t* out, long cap)
{
    Sps sps; Pps pps;
    seq_to_sets(*s, &sps, &pps);
    BitWriter a, b;
    write_sps(a, sps);
    write_pps(b, pps);
    std::vector<uint8_t> o;
    nal_emit(3, kNalSps, a.bytes, o);
    nal_emit(3, kNalPps, b.bytes, o);
    if ((long)o.size() > cap) return -1;
    memcpy(out, o.data(), o.size());
    return (long)o.size();
}

extern "C" long h264e_encode_picture(const h264s_seq* s, const h264s_pic* p, h264s_mb* mbs, int n_mbs, uint32_t flags, uint8_t* out, long cap)
{
    if (!s || !p || !mbs || n_mbs != s->w_mbs * s->h_mbs) return kErrState;
    Sps sps; Pps pps; SliceHeader h;
    seq_to_sets(*s, &sps, &pps);
    pic_to_header(*s, *p, &h);
    const bool p_slice = h.slice_type == kSliceP;
    BitWriter w;
    write_slice_header(w, sps, pps, h);
    SliceParams sp;
    sp.w_mbs = s->w_mbs; sp.h_mbs = s->h_mbs; sp.p_slice = p_slice;
    sp.num_ref_idx_l0 = h.num_ref_idx_l0; sp.transform_8x8_mode = pps.transform_8x8_mode_flag != 0;
    std::vector<MbInfo> info((size_t)n_mbs);
    memset(info.data(), 0, sizeof(MbInfo) * (size_t)n_mbs);
    const int slice_qp = pps.pic_init_qp + h.slice_qp_delta;
    if (!pps.entropy_coding_mode_flag) {
        // CAVLC slice data (7.3.4): mb_skip_run in front of every coded macroblock of a P slice (and at the end of the slice)
        std::vector<NnzInfo> nnz((size_t)n_mbs);
        VlcWriter io{w};
        CavlcSliceCodec<VlcWriter> codec(io, sp, info.data(), nnz.data(), 0);
        PicDerive drv;
        if (flags & H264E_ABSOLUTE) {
            DeriveParams dp;
            dp.w_mbs = s->w_mbs; dp.h_mbs = s->h_mbs;
            dp.ref_quirks = (flags & H264E_REF_QUIRKS) != 0;
            dp.constrained_intra_pred = pps.constrained_intra_pred_flag != 0;
            drv.begin_picture(dp);
            drv.begin_slice(slice_qp);
        }
        uint32_t skip_run = 0;
        for (int a = 0; a < n_mbs; a++) {
            h264s_mb& m = mbs[a];
            int rc = normalise_mb(m, p_slice, sp.transform_8x8_mode, (flags & H264E_ABSOLUTE) != 0);
            if (rc) return rc;
            if (flags & H264E_ABSOLUTE) {
                info[(size_t)a].slice = 1;
                info[(size_t)a].kind = m.kind;
                drv.underive_mb(a, m, info.data());
            }
            if (m.kind == H264S_PSKIP) {
                if (!p_slice) return kErrState;
                codec.skipped(a, m);
                skip_run++;
                continue;
            }
            if (p_slice) { w.ue(skip_run); skip_run = 0; }
            rc = codec.mb(a, m);
            if (rc) return rc;
        }
        if (p_slice && skip_run) w.ue(skip_run);
        w.trailing();
        std::vector<uint8_t> o;
        nal_emit(h.nal_ref_idc, h.nal_unit_type, w.bytes, o);
        if ((long)o.size() > cap) return kErrCapacity;
        memcpy(out, o.data(), o.size());
        return (long)o.size();
    }
    w.align_ones();                                        // cabac_alignment_one_bit
    CabacEncoder enc;
    enc.ctx.init(slice_qp, !p_slice, h.cabac_init_idc);
    enc.start();
    SliceCodec<CabacEncoder> codec(enc, sp, info.data(), 0);
    codec.begin_slice();
    PicDerive drv;
    if (flags & H264E_ABSOLUTE) {
        DeriveParams dp;
        dp.w_mbs = s->w_mbs; dp.h_mbs = s->h_mbs;
        dp.ref_quirks = (flags & H264E_REF_QUIRKS) != 0;
        dp.constrained_intra_pred = pps.constrained_intra_pred_flag != 0;
        drv.begin_picture(dp);
        drv.begin_slice(slice_qp);
    }
    for (int a = 0; a < n_mbs; a++) {
        h264s_mb& m = mbs[a];
        int rc = normalise_mb(m, p_slice, sp.transform_8x8_mode, (flags & H264E_ABSOLUTE) != 0);
        if (rc) return rc;
        if (flags & H264E_ABSOLUTE) {
            info[(size_t)a].slice = 1;
            info[(size_t)a].kind = m.kind;
            drv.underive_mb(a, m, info.data());
        }
        rc = codec.mb(a, m);
        if (rc) return rc;
        codec.end_of_slice(a == n_mbs - 1);
    }
    std::vector<uint8_t> rbsp = w.bytes;
    rbsp.insert(rbsp.end(), enc.out.begin(), enc.out.end());
    std::vector<uint8_t> o;
    nal_emit(h.nal_ref_idc, h.nal_unit_type, rbsp, o);
    if ((long)o.size() > cap) return kErrCapacity;
    memcpy(out, o.data(), o.size());
    return (long)o.size();
}

// =====================================================================================================================================
// decoder
// =====================================================================================================================================
struct h264d {
    const uint8_t* data = nullptr;
    size_t n = 0, pos = 0;
    uint32_t flags = 0;
    int pool = 0, reorder = 0;
    Sps sps[32];
    Pps pps[256];
    Dpb dpb;
    std::vector<MbInfo> info;
    std::vector<NnzInfo> nnz;               // CAVLC: total_coeff of the picture's blocks
    PicDerive drv;
    std::vector<h264r_mb_hdr> hdr;
    std::vector<int16_t> mv, coeff;
    std::vector<h264s_mb> syntax;
    std::vector<uint8_t> rbsp;
    std::string err;
    bool have_picture = false, last_was_p = false, flushed = false;
    int w_mbs = 0, h_mbs = 0, n_intra = 0, decoded = 0;
    int fail(int code, const std::string& why) { err = why; return code; }
};

namespace {

// slice_data() of one slice NAL into the picture's buffers.  Returns the address after the last macroblock or a negative error.
int decode_slice_data(h264d* d, BitReader& r, const Pps& pps, const SliceHeader& h, int slice_no, int64_t* bytes)
{
    const int n_mbs = d->w_mbs * d->h_mbs;
    if (h.first_mb >= n_mbs) return d->fail(kErrBitstream, "first_mb_in_slice beyond the picture");
    SliceParams sp;
    sp.w_mbs = d->w_mbs; sp.h_mbs = d->h_mbs; sp.p_slice = h.slice_type == kSliceP;
    sp.num_ref_idx_l0 = h.num_ref_idx_l0; sp.transform_8x8_mode = pps.transform_8x8_mode_flag != 0;
    const int slice_qp = pps.pic_init_qp + h.slice_qp_delta;
    if (slice_qp < 0 || slice_qp > 51) return d->fail(kErrBitstream, "slice QP out of range");
    if (!pps.entropy_coding_mode_flag) {
        // CAVLC slice data (7.3.4)
        *bytes += (int64_t)r.bytes_left();
        if (d->nnz.size() != (size_t)n_mbs) d->nnz.assign((size_t)n_mbs, NnzInfo());
        VlcReader io{r};
        CavlcSliceCodec<VlcReader> codec(io, sp, d->info.data(), d->nnz.data(), slice_no);
        d->drv.begin_slice(slice_qp);
        h264s_mb local;
        const bool keep = (d->flags & H264D_KEEP_SYNTAX) != 0;
        int addr = h.first_mb;
        auto finish_mb = [&](h264s_mb& m) -> int {
            const int lim = h.num_ref_idx_l0 + (h.num_ref_idx_l0 == 0);
            if (m.ref_idx[0] >= lim || m.ref_idx[1] >= lim || m.ref_idx[2] >= lim || m.ref_idx[3] >= lim)
                return d->fail(kErrBitstream, "ref_idx_l0 beyond num_ref_idx_l0_active");
            d->drv.derive_mb(addr, m, d->info.data(), &d->hdr[(size_t)addr], &d->mv[(size_t)addr * H264R_MV_PER_MB],
                             &d->coeff[(size_t)addr * H264R_COEFF_PER_MB]);
            if (d->hdr[(size_t)addr].mb_class <= H264R_MB_I16x16) d->n_intra++;
            addr++;
            return 0;
        };
        bool more = r.more_rbsp_data();
        while (more) {
            if (sp.p_slice) {
                const uint32_t run = r.ue();
                if (run > (uint32_t)(n_mbs - addr)) return d->fail(kErrBitstream, "mb_skip_run runs past the last macroblock");
                for (uint32_t k = 0; k < run; k++) {
                    if (d->info[(size_t)addr].slice != 0) return d->fail(kErrBitstream, "macroblock coded twice");
                    h264s_mb& m = keep ? d->syntax[(size_t)addr] : local;
                    codec.skipped(addr, m);
                    const int rc = finish_mb(m);
                    if (rc) return rc;
                }
                more = r.more_rbsp_data();
                if (!more) break;
            }
            if (addr >= n_mbs) return d->fail(kErrBitstream, "slice runs past the last macroblock");
            if (d->info[(size_t)addr].slice != 0) return d->fail(kErrBitstream, "macroblock coded twice");
            h264s_mb& m = keep ? d->syntax[(size_t)addr] : local;
            const int rc = codec.mb(addr, m);
            if (rc == -2) return d->fail(kErrUnsupported, "I_PCM macroblock");
            if (rc || r.overrun()) return d->fail(kErrBitstream, "macroblock layer");
            const int rc2 = finish_mb(m);
            if (rc2) return rc2;
            more = r.more_rbsp_data();
        }
        return addr;
    }
    r.align();                                             // cabac_alignment_one_bit
    CabacDecoder dec;
    dec.ctx.init(slice_qp, !sp.p_slice, h.cabac_init_idc);
    dec.start(r.byte_ptr(), r.bytes_left());
    *bytes += (int64_t)r.bytes_left();
    SliceCodec<CabacDecoder> codec(dec, sp, d->info.data(), slice_no);
    codec.begin_slice();
    d->drv.begin_slice(slice_qp);
    h264s_mb local;
    const bool keep = (d->flags & H264D_KEEP_SYNTAX) != 0;
    int addr = h.first_mb;
    for (;;) {
        if (d->info[(size_t)addr].slice != 0) return d->fail(kErrBitstream, "macroblock coded twice");
        h264s_mb& m = keep ? d->syntax[(size_t)addr] : local;
        const int rc = codec.mb(addr, m);
        if (rc == -2) return d->fail(kErrUnsupported, "I_PCM macroblock");
        if (rc) return d->fail(kErrBitstream, "macroblock layer");
        if (m.ref_idx[0] >= h.num_ref_idx_l0 + (h.num_ref_idx_l0 == 0) || m.ref_idx[1] >= h.num_ref_idx_l0 + (h.num_ref_idx_l0 == 0) ||
            m.ref_idx[2] >= h.num_ref_idx_l0 + (h.num_ref_idx_l0 == 0) || m.ref_idx[3] >= h.num_ref_idx_l0 + (h.num_ref_idx_l0 == 0))
            return d->fail(kErrBitstream, "ref_idx_l0 beyond num_ref_idx_l0_active");
        d->drv.derive_mb(addr, m, d->info.data(), &d->hdr[(size_t)addr], &d->mv[(size_t)addr * H264R_MV_PER_MB],
                         &d->coeff[(size_t)addr * H264R_COEFF_PER_MB]);
        if (d->hdr[(size_t)addr].mb_class <= H264R_MB_I16x16) d->n_intra++;
        addr++;
        const int end = codec.end_of_slice(0);
        if (dec.overrun()) return d->fail(kErrBitstream, "slice data ends inside a macroblock");
        if (end) break;
        if (addr >= n_mbs) return d->fail(kErrBitstream, "slice runs past the last macroblock");
    }
    return addr;
}

void fill_pic_params(const Pps& pps, const SliceHeader& h, const int* list0, int n_refs, h264r_pic_params* pp)
{
    memset(pp, 0, sizeof(*pp));
    pp->slice_type = h.slice_type == kSliceP ? H264R_SLICE_P : H264R_SLICE_I;
    pp->num_ref_idx_l0 = n_refs > H264R_MAX_REFS ? H264R_MAX_REFS : n_refs;
    for (int i = 0; i < H264R_MAX_REFS; i++) {
        pp->ref_list0[i] = i < n_refs ? list0[i] : -1;
        pp->luma_weight[i] = 1;
        pp->chroma_weight[i][0] = pp->chroma_weight[i][1] = 1;
    }
    if (h.slice_type == kSliceP && pps.weighted_pred_flag) {
        pp->weighted_pred = 1;
        pp->luma_log2_weight_denom = h.luma_log2_weight_denom;
        pp->chroma_log2_weight_denom = h.chroma_log2_weight_denom;
        for (int i = 0; i < pp->num_ref_idx_l0; i++) {
            pp->luma_weight[i] = (int16_t)h.luma_weight[i];
            pp->luma_offset[i] = (int16_t)h.luma_offset[i];
            for (int c = 0; c < 2; c++) {
                pp->chroma_weight[i][c] = (int16_t)h.chroma_weight[i][c];
                pp->chroma_offset[i][c] = (int16_t)h.chroma_offset[i][c];
            }
        }
    }
    pp->disable_deblocking_filter_idc = h.disable_deblocking_filter_idc;
    pp->slice_alpha_c0_offset_div2 = h.alpha_c0_offset_div2;
    pp->slice_beta_offset_div2 = h.beta_offset_div2;
}

void report(const std::vector<DpbFrame>& out, const std::vector<int>& rel, h264d_picture* pic)
{
    pic->n_output = 0;
    for (const DpbFrame& f : out)
        if (pic->n_output < 17) {
            pic->output[pic->n_output].frame_id = f.slot;
            pic->output[pic->n_output].poc = f.poc;
            pic->output[pic->n_output].decode_index = f.decode_index;
            pic->n_output++;
        }
    pic->n_release = 0;
    for (int s : rel)
        if (pic->n_release < 17) pic->release[pic->n_release++] = s;
}

}  // namespace

extern "C" h264d* h264d_open(const uint8_t* annexb, size_t n_bytes, uint32_t flags, int max_frames, int reorder_depth)
{
    h264d* d = new h264d();
    d->data = annexb; d->n = n_bytes; d->flags = flags;
    d->pool = max_frames < 0 ? 0 : max_frames;
    d->reorder = reorder_depth < 0 ? 0 : reorder_depth;
    d->dpb.configure(d->pool, d->reorder);
    d->dpb.reset();
    return d;
}
extern "C" void h264d_close(h264d* d) { delete d; }
extern "C" const char* h264d_error(const h264d* d) { return d ? d->err.c_str() : "no decoder"; }

extern "C" int h264d_next(h264d* d, h264d_picture* out)
{
    if (!d || !out) return kErrState;
    memset(out, 0, sizeof(*out));
    d->have_picture = false;
    bool in_picture = false;
    int next_mb = 0, slice_no = 0, n_mbs = 0;
    SliceHeader first;
    const Sps* sps = nullptr;
    const Pps* pps = nullptr;
    for (;;) {
        size_t start, len;
        if (!annexb_next(d->data, d->n, &d->pos, &start, &len)) {
            if (in_picture) return d->fail(kErrBitstream, "stream ends inside a picture");
            if (!d->flushed) {
                std::vector<DpbFrame> o; std::vector<int> rel;
                d->dpb.flush(&o, &rel);
                report(o, rel, out);
                d->flushed = true;
            }
            return 0;
        }
        const uint8_t* nal = d->data + start;
        if (nal[0] & 0x80) return d->fail(kErrBitstream, "forbidden_zero_bit set");
        const int ref_idc = (nal[0] >> 5) & 3, type = nal[0] & 31;
        if (type != kNalSps && type != kNalPps && type != kNalSlice && type != kNalIdr) {
            if (type >= 2 && type <= 4) return d->fail(kErrUnsupported, "data partitioning");
            continue;                                      // SEI, AUD, end of sequence, filler ...: nothing the reconstruction needs
        }
        nal_unescape(nal + 1, len - 1, d->rbsp);
        BitReader r(d->rbsp.data(), d->rbsp.size());
        const char* why = "";
        if (type == kNalSps) {
            Sps s;
            const int rc = parse_sps(r, &s, &why);
            if (rc) return d->fail(rc, std::string("SPS: ") + (rc == kErrUnsupported ? why : "damaged"));
            if (s.sps_id > 31) return d->fail(kErrBitstream, "seq_parameter_set_id");
            d->sps[s.sps_id] = s;
            continue;
        }
        if (type == kNalPps) {
            Pps p;
            const int rc = parse_pps(r, &p, &why);
            if (rc) return d->fail(rc, std::string("PPS: ") + (rc == kErrUnsupported ? why : "damaged"));
            d->pps[p.pps_id] = p;
            continue;
        }
        SliceHeader h;
        auto lookup = [d](int id, const Sps** s, const Pps** p) {
            if (id < 0 || id > 255 || !d->pps[id].valid || !d->sps[d->pps[id].sps_id].valid) return false;
            *p = &d->pps[id]; *s = &d->sps[d->pps[id].sps_id];
            return true;
        };
        const Sps* s2 = nullptr; const Pps* p2 = nullptr;
        int rc = parse_slice_header(r, type, ref_idc, lookup, &h, &s2, &p2, &why);
        if (rc) return d->fail(rc, std::string("slice header: ") + (rc == kErrUnsupported ? why : (*why ? why : "damaged")));
        if (h.first_mb == 0 || !in_picture) {
            if (in_picture) return d->fail(kErrBitstream, "new picture before the previous one was complete");
            if (h.first_mb != 0) return d->fail(kErrBitstream, "picture does not start at macroblock 0 (lost slice)");
            // ---- a new picture
            sps = s2; pps = p2; first = h;
            d->w_mbs = sps->w_mbs; d->h_mbs = sps->h_mbs;
            n_mbs = d->w_mbs * d->h_mbs;
            d->info.assign((size_t)n_mbs, MbInfo());
            memset(d->info.data(), 0, sizeof(MbInfo) * (size_t)n_mbs);
            d->hdr.resize((size_t)n_mbs);
            d->mv.assign((size_t)n_mbs * H264R_MV_PER_MB, 0);
            d->coeff.resize((size_t)n_mbs * H264R_COEFF_PER_MB);
            if (d->flags & H264D_KEEP_SYNTAX) d->syntax.resize((size_t)n_mbs);
            d->n_intra = 0;
            DeriveParams dp;
            dp.w_mbs = d->w_mbs; dp.h_mbs = d->h_mbs;
            dp.ref_quirks = (d->flags & H264D_REF_QUIRKS) != 0;
            dp.constrained_intra_pred = pps->constrained_intra_pred_flag != 0;
            dp.chroma_qp_offset[0] = pps->chroma_qp_index_offset;
            dp.chroma_qp_offset[1] = pps->second_chroma_qp_index_offset;
            d->drv.begin_picture(dp);
            int list0[32], n_refs = 0, slot = -1;
            rc = d->dpb.begin_picture(*sps, h, &slot, list0, &n_refs);
            if (rc) return d->fail(rc, d->dpb.error());
            out->w_mbs = d->w_mbs; out->h_mbs = d->h_mbs;
            out->frame_id = slot;
            out->is_idr = h.idr(); out->nal_ref_idc = h.nal_ref_idc; out->frame_num = h.frame_num;
            out->poc = d->dpb.current_poc();
            out->slice_qp = pps->pic_init_qp + h.slice_qp_delta;
            fill_pic_params(*pps, h, list0, n_refs, &out->pp);
            in_picture = true; next_mb = 0; slice_no = 0;
        } else {
            if (p2 != pps || h.frame_num != first.frame_num || h.slice_type != first.slice_type ||
                h.num_ref_idx_l0 != first.num_ref_idx_l0 || h.n_mods || first.n_mods ||
                h.disable_deblocking_filter_idc != first.disable_deblocking_filter_idc ||
                h.alpha_c0_offset_div2 != first.alpha_c0_offset_div2 || h.beta_offset_div2 != first.beta_offset_div2)
                return d->fail(kErrUnsupported, "slices of one picture with different types, reference lists or deblocking parameters");
            if (h.first_mb != next_mb) return d->fail(kErrBitstream, "slices out of order or missing");
        }
        rc = decode_slice_data(d, r, *pps, h, slice_no, &out->slice_data_bytes);
        if (rc < 0) return rc;
        next_mb = rc;
        slice_no++;
        if (next_mb < n_mbs) continue;
        // ---- picture complete
        std::vector<DpbFrame> o; std::vector<int> rel;
        rc = d->dpb.end_picture(*sps, first, &o, &rel);
        if (rc) return d->fail(rc, d->dpb.error());
        report(o, rel, out);
        out->decode_index = d->decoded++;
        out->hdr = d->hdr.data();
        out->mv = first.slice_type == kSliceP ? d->mv.data() : nullptr;
        out->coeff = d->coeff.data();
        out->syntax = (d->flags & H264D_KEEP_SYNTAX) ? d->syntax.data() : nullptr;
        out->n_intra = d->n_intra;
        d->have_picture = true;
        d->last_was_p = first.slice_type == kSliceP;
        return 1;
    }
}

extern "C" int h264w_pack_arrays(h264r_picture_buf* buf, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
                                 int level_fmt, size_t* n_blocks, long* n_mvs)
{
    if (!buf || !hdr || !coeff || !n_blocks || !n_mvs || n_mbs <= 0) return H264R_E_INVAL;
    h264r_writer w;
    h264r_writer_begin(&w, buf, n_mbs, level_fmt);
    for (int a = 0; a < n_mbs; a++) {
        const int16_t* v = (mv && hdr[a].mb_class >= H264R_MB_P16x16) ? mv + (size_t)a * H264R_MV_PER_MB : nullptr;
        const int rc = h264r_writer_mb(&w, &hdr[a], v, coeff + (size_t)a * H264R_COEFF_PER_MB);
        if (rc != H264R_OK) return rc;
    }
    return h264r_writer_end(&w, n_blocks, n_mvs);
}

extern "C" int h264d_fill_picture_buf(h264d* d, h264r_picture_buf* buf, int level_fmt, size_t* n_blocks, long* n_mvs)
{
    if (!d || !buf || !d->have_picture) return H264R_E_STATE;
    const int n = d->w_mbs * d->h_mbs;
    h264r_writer w;
    h264r_writer_begin(&w, buf, n, level_fmt);
    for (int a = 0; a < n; a++) {
        const h264r_mb_hdr* h = &d->hdr[(size_t)a];
        const int16_t* mv = (d->last_was_p && h->mb_class >= H264R_MB_P16x16) ? &d->mv[(size_t)a * H264R_MV_PER_MB] : nullptr;
        const int rc = h264r_writer_mb(&w, h, mv, &d->coeff[(size_t)a * H264R_COEFF_PER_MB]);
        if (rc != H264R_OK) return rc;
    }
    return h264r_writer_end(&w, n_blocks, n_mvs);
}

extern "C" long h264d_benchmark(const uint8_t* annexb, size_t n_bytes, uint32_t flags, int repeat, double* seconds)
{
    long pictures = 0;
    // H264D_BENCH_PACK: every picture also goes through the writer into a stream-form buffer (plain memory laid out by hand:
    // 8-byte headers, a word count per macroblock, the stream) -- what a host does before h264r_submit_picture.  (Packing as a
    // pass of its own over the finished picture measured faster than calling the writer per macroblock from inside the parser:
    // 2.8 ms against 3.9 ms per 1080p picture -- the parser's and the writer's branches do not share a predictor well.)
    const bool pack = (flags & H264D_BENCH_PACK) != 0;
    flags &= ~(uint32_t)H264D_BENCH_PACK;
    std::vector<h264r_mb_hdr8> hdr8;
    std::vector<uint8_t> words;
    std::vector<uint32_t> stream;
    std::vector<h264r_mb_hdr> hdr16;
    std::vector<uint32_t> mask;
    std::vector<int16_t> blocks, mvs;
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < repeat; k++) {
        h264d* d = h264d_open(annexb, n_bytes, flags, 0, 0);
        h264d_picture pic;
        int rc;
        while ((rc = h264d_next(d, &pic)) == 1) {
            pictures++;
            if (!pack) continue;
            const size_t n = (size_t)d->w_mbs * d->h_mbs;
            if (hdr8.size() != n) {
                hdr8.resize(n); words.resize(n); stream.resize(n * 116); hdr16.resize(n); mask.resize(n);
                blocks.resize(n * H264R_COEFF_PER_MB); mvs.resize(n * H264R_MV_PER_MB);
            }
            h264r_picture_buf b;
            memset(&b, 0, sizeof(b));
            b.hdr8 = hdr8.data(); b.mb_words = words.data(); b.stream = stream.data(); b.max_stream_words = stream.size();
            b.hdr_bytes = H264R_HDR_STREAM;
            size_t nb = 0; long nm = 0;
            int prc = h264d_fill_picture_buf(d, &b, H264R_LEVELS_COMPACT, &nb, &nm);
            if (prc == H264R_E_INVAL) {           // a level beyond int8: the picture travels with 16-byte headers and int16 levels
                b.hdr_bytes = 16; b.hdr = hdr16.data(); b.blk_mask = mask.data(); b.blocks = blocks.data(); b.max_blocks = n * 24;
                b.mv = mvs.data(); b.mv_end = mvs.data() + mvs.size();
                prc = h264d_fill_picture_buf(d, &b, H264R_LEVELS_INT16, &nb, &nm);
            }
            if (prc != H264R_OK) { h264d_close(d); return -100 + prc; }
        }
        h264d_close(d);
        if (rc < 0) return rc;
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    return pictures;
}
