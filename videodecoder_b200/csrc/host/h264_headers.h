// h264_headers.h -- sequence / picture parameter sets and slice headers of the host entropy path: one structure each, one
// reader and one writer each (ITU-T H.264 7.3.2.1, 7.3.2.2, 7.3.3).
//
// Replaces, on the reference's side, NAL::decode_SPS / decode_PPS (h264/NAL.cpp:31-184: fields into SPS_data / PPS_data,
// High-profile fields misread) and Slice::PraseSliceHeader (h264/slice.cpp:14-275).  Scope: progressive frames
// (frame_mbs_only_flag = 1), 4:2:0, 8 bit, one slice group, flat scaling lists, I and P slices -- what the reconstruction
// engine behind include/h264r.h covers; anything else is reported as unsupported, never guessed.
#pragma once
#include <stdint.h>
#include <string.h>

#include "bits.h"

namespace h264host {

enum { kNalSlice = 1, kNalIdr = 5, kNalSei = 6, kNalSps = 7, kNalPps = 8, kNalAud = 9 };
enum { kSliceP = 0, kSliceB = 1, kSliceI = 2 };

struct Sps {
    int profile_idc = 77, constraint_flags = 0, level_idc = 40, sps_id = 0;
    int chroma_format_idc = 1, bit_depth_luma = 8, bit_depth_chroma = 8;
    int log2_max_frame_num = 4, poc_type = 0, log2_max_poc_lsb = 4;
    int delta_pic_order_always_zero_flag = 0, offset_for_non_ref_pic = 0, offset_for_top_to_bottom_field = 0;
    int num_ref_frames_in_poc_cycle = 0, offset_for_ref_frame[256];
    int max_num_ref_frames = 1, gaps_in_frame_num_allowed = 0;
    int w_mbs = 0, h_mbs = 0, frame_mbs_only_flag = 1, direct_8x8_inference_flag = 1;
    int crop_flag = 0, crop[4] = {0, 0, 0, 0};
    int vui_present = 0;
    bool valid = false;
    static bool is_high(int p)
    {
        return p == 100 || p == 110 || p == 122 || p == 244 || p == 44 || p == 83 || p == 86 || p == 118 || p == 128 ||
               p == 138 || p == 139 || p == 134 || p == 135;
    }
};

struct Pps {
    int pps_id = 0, sps_id = 0, entropy_coding_mode_flag = 1, bottom_field_pic_order_in_frame_present_flag = 0;
    int num_slice_groups = 1, num_ref_idx_l0_default = 1, num_ref_idx_l1_default = 1;
    int weighted_pred_flag = 0, weighted_bipred_idc = 0, pic_init_qp = 26, pic_init_qs = 26;
    int chroma_qp_index_offset = 0, second_chroma_qp_index_offset = 0;
    int deblocking_filter_control_present_flag = 0, constrained_intra_pred_flag = 0, redundant_pic_cnt_present_flag = 0;
    int transform_8x8_mode_flag = 0;
    bool valid = false;
};

struct RefListMod { int idc, value; };
struct Mmco { int op, a, b; };          // a: difference_of_pic_nums_minus1 / long_term_pic_num / max_long_term_frame_idx_plus1, b: long_term_frame_idx

struct SliceHeader {
    int nal_unit_type = kNalSlice, nal_ref_idc = 0;
    int first_mb = 0, slice_type_raw = 7, slice_type = kSliceI, pps_id = 0;
    int frame_num = 0, idr_pic_id = 0, poc_lsb = 0, delta_poc_bottom = 0, delta_poc[2] = {0, 0};
    int redundant_pic_cnt = 0;
    int num_ref_idx_override = 0, num_ref_idx_l0 = 0;
    int n_mods = 0;
    RefListMod mods[66];
    // pred_weight_table
    int luma_log2_weight_denom = 0, chroma_log2_weight_denom = 0;
    int luma_weight_flag[32], luma_weight[32], luma_offset[32];
    int chroma_weight_flag[32], chroma_weight[32][2], chroma_offset[32][2];
    // dec_ref_pic_marking
    int no_output_of_prior_pics_flag = 0, long_term_reference_flag = 0, adaptive_marking = 0, n_mmco = 0;
    Mmco mmco[66];
    int cabac_init_idc = 0, slice_qp_delta = 0;
    int disable_deblocking_filter_idc = 0, alpha_c0_offset_div2 = 0, beta_offset_div2 = 0;
    bool idr() const { return nal_unit_type == kNalIdr; }
};

// ---- errors ------------------------------------------------------------------------------------------------------------------
enum {
    kOk = 0, kErrBitstream = -1, kErrUnsupported = -2, kErrState = -3, kErrCapacity = -4,
};

// ---- SPS -----------------------------------------------------------------------------------------------------------------------
inline int parse_sps(BitReader& r, Sps* s, const char** why)
{
    *s = Sps();
    s->profile_idc = (int)r.u(8);
    s->constraint_flags = (int)r.u(8);
    s->level_idc = (int)r.u(8);
    s->sps_id = (int)r.ue();
    if (Sps::is_high(s->profile_idc)) {
        s->chroma_format_idc = (int)r.ue();
        if (s->chroma_format_idc == 3) { *why = "4:4:4 (separate_colour_plane_flag)"; return kErrUnsupported; }
        s->bit_depth_luma = 8 + (int)r.ue();
        s->bit_depth_chroma = 8 + (int)r.ue();
        r.u1();                                            // qpprime_y_zero_transform_bypass_flag
        if (r.u1()) { *why = "seq_scaling_matrix_present_flag"; return kErrUnsupported; }
    }
    if (s->chroma_format_idc != 1 || s->bit_depth_luma != 8 || s->bit_depth_chroma != 8) { *why = "only 4:2:0 8 bit"; return kErrUnsupported; }
    s->log2_max_frame_num = 4 + (int)r.ue();
    s->poc_type = (int)r.ue();
    if (s->poc_type == 0) s->log2_max_poc_lsb = 4 + (int)r.ue();
    else if (s->poc_type == 1) {
        s->delta_pic_order_always_zero_flag = (int)r.u1();
        s->offset_for_non_ref_pic = r.se();
        s->offset_for_top_to_bottom_field = r.se();
        s->num_ref_frames_in_poc_cycle = (int)r.ue();
        if (s->num_ref_frames_in_poc_cycle > 255) return kErrBitstream;
        for (int i = 0; i < s->num_ref_frames_in_poc_cycle; i++) s->offset_for_ref_frame[i] = r.se();
    } else if (s->poc_type != 2) return kErrBitstream;
    s->max_num_ref_frames = (int)r.ue();
    s->gaps_in_frame_num_allowed = (int)r.u1();
    s->w_mbs = (int)r.ue() + 1;
    s->h_mbs = (int)r.ue() + 1;
    s->frame_mbs_only_flag = (int)r.u1();
    if (!s->frame_mbs_only_flag) { *why = "interlaced (frame_mbs_only_flag = 0)"; return kErrUnsupported; }
    s->direct_8x8_inference_flag = (int)r.u1();
    s->crop_flag = (int)r.u1();
    if (s->crop_flag) for (int i = 0; i < 4; i++) s->crop[i] = (int)r.ue();
    s->vui_present = (int)r.u1();                          // VUI carries nothing the reconstruction needs: not read
    if (r.overrun() || s->max_num_ref_frames > 16 || s->w_mbs > 1024 || s->h_mbs > 1024) return kErrBitstream;
    s->valid = true;
    return kOk;
}

inline void write_sps(BitWriter& w, const Sps& s)
{
    w.put((uint32_t)s.profile_idc, 8);
    w.put((uint32_t)s.constraint_flags, 8);
    w.put((uint32_t)s.level_idc, 8);
    w.ue((uint32_t)s.sps_id);
    if (Sps::is_high(s.profile_idc)) {
        w.ue(1); w.ue(0); w.ue(0);                         // 4:2:0, 8 bit, 8 bit
        w.put1(0); w.put1(0);                              // no transform bypass, flat scaling lists
    }
    w.ue((uint32_t)(s.log2_max_frame_num - 4));
    w.ue((uint32_t)s.poc_type);
    if (s.poc_type == 0) w.ue((uint32_t)(s.log2_max_poc_lsb - 4));
    w.ue((uint32_t)s.max_num_ref_frames);
    w.put1((uint32_t)s.gaps_in_frame_num_allowed);
    w.ue((uint32_t)(s.w_mbs - 1));
    w.ue((uint32_t)(s.h_mbs - 1));
    w.put1(1);                                             // frame_mbs_only_flag
    w.put1((uint32_t)s.direct_8x8_inference_flag);
    w.put1((uint32_t)s.crop_flag);
    if (s.crop_flag) for (int i = 0; i < 4; i++) w.ue((uint32_t)s.crop[i]);
    w.put1(0);                                             // vui_parameters_present_flag
    w.trailing();
}

// ---- PPS -----------------------------------------------------------------------------------------------------------------------
inline int parse_pps(BitReader& r, Pps* p, const char** why)
{
    *p = Pps();
    p->pps_id = (int)r.ue();
    p->sps_id = (int)r.ue();
    p->entropy_coding_mode_flag = (int)r.u1();
    p->bottom_field_pic_order_in_frame_present_flag = (int)r.u1();
    p->num_slice_groups = (int)r.ue() + 1;
    if (p->num_slice_groups != 1) { *why = "slice groups (FMO)"; return kErrUnsupported; }
    p->num_ref_idx_l0_default = (int)r.ue() + 1;
    p->num_ref_idx_l1_default = (int)r.ue() + 1;
    p->weighted_pred_flag = (int)r.u1();
    p->weighted_bipred_idc = (int)r.u(2);
    p->pic_init_qp = 26 + r.se();
    p->pic_init_qs = 26 + r.se();
    p->chroma_qp_index_offset = r.se();
    p->deblocking_filter_control_present_flag = (int)r.u1();
    p->constrained_intra_pred_flag = (int)r.u1();
    p->redundant_pic_cnt_present_flag = (int)r.u1();
    p->second_chroma_qp_index_offset = p->chroma_qp_index_offset;
    if (r.more_rbsp_data()) {
        p->transform_8x8_mode_flag = (int)r.u1();
        if (r.u1()) { *why = "pic_scaling_matrix_present_flag"; return kErrUnsupported; }
        p->second_chroma_qp_index_offset = r.se();
    }
    if (r.overrun() || p->pps_id > 255 || p->sps_id > 31 || p->num_ref_idx_l0_default > 32) return kErrBitstream;
    p->valid = true;
    return kOk;
}

inline void write_pps(BitWriter& w, const Pps& p)
{
    w.ue((uint32_t)p.pps_id);
    w.ue((uint32_t)p.sps_id);
    w.put1((uint32_t)p.entropy_coding_mode_flag);
    w.put1(0);                                             // bottom_field_pic_order_in_frame_present_flag
    w.ue(0);                                               // num_slice_groups_minus1
    w.ue((uint32_t)(p.num_ref_idx_l0_default - 1));
    w.ue((uint32_t)(p.num_ref_idx_l1_default - 1));
    w.put1((uint32_t)p.weighted_pred_flag);
    w.put((uint32_t)p.weighted_bipred_idc, 2);
    w.se(p.pic_init_qp - 26);
    w.se(p.pic_init_qs - 26);
    w.se(p.chroma_qp_index_offset);
    w.put1((uint32_t)p.deblocking_filter_control_present_flag);
    w.put1((uint32_t)p.constrained_intra_pred_flag);
    w.put1(0);                                             // redundant_pic_cnt_present_flag
    if (p.transform_8x8_mode_flag || p.second_chroma_qp_index_offset != p.chroma_qp_index_offset) {
        w.put1((uint32_t)p.transform_8x8_mode_flag);
        w.put1(0);                                         // pic_scaling_matrix_present_flag
        w.se(p.second_chroma_qp_index_offset);
    }
    w.trailing();
}

// ---- slice header (up to and including the deblocking fields; the caller aligns for CABAC) ---------------------------------------
// Two-stage read: the PPS id is in the header itself, so the caller hands in a lookup.
template <class Lookup>
inline int parse_slice_header(BitReader& r, int nal_unit_type, int nal_ref_idc, Lookup&& lookup, SliceHeader* h,
                              const Sps** sps_out, const Pps** pps_out, const char** why)
{
    *h = SliceHeader();
    h->nal_unit_type = nal_unit_type;
    h->nal_ref_idc = nal_ref_idc;
    h->first_mb = (int)r.ue();
    h->slice_type_raw = (int)r.ue();
    if (h->slice_type_raw > 9) return kErrBitstream;
    h->slice_type = h->slice_type_raw % 5;
    if (h->slice_type != kSliceP && h->slice_type != kSliceI) { *why = "B / SP / SI slices"; return kErrUnsupported; }
    h->pps_id = (int)r.ue();
    const Sps* sps = nullptr;
    const Pps* pps = nullptr;
    if (!lookup(h->pps_id, &sps, &pps)) { *why = "slice refers to a parameter set that was not sent"; return kErrBitstream; }
    *sps_out = sps; *pps_out = pps;
    h->frame_num = (int)r.u(sps->log2_max_frame_num);
    if (h->idr()) h->idr_pic_id = (int)r.ue();
    if (sps->poc_type == 0) {
        h->poc_lsb = (int)r.u(sps->log2_max_poc_lsb);
        if (pps->bottom_field_pic_order_in_frame_present_flag) h->delta_poc_bottom = r.se();
    } else if (sps->poc_type == 1 && !sps->delta_pic_order_always_zero_flag) {
        h->delta_poc[0] = r.se();
        if (pps->bottom_field_pic_order_in_frame_present_flag) h->delta_poc[1] = r.se();
    }
    if (pps->redundant_pic_cnt_present_flag) h->redundant_pic_cnt = (int)r.ue();
    h->num_ref_idx_l0 = 0;
    if (h->slice_type == kSliceP) {
        h->num_ref_idx_l0 = pps->num_ref_idx_l0_default;
        h->num_ref_idx_override = (int)r.u1();
        if (h->num_ref_idx_override) h->num_ref_idx_l0 = (int)r.ue() + 1;
        if (h->num_ref_idx_l0 > 32) return kErrBitstream;
        if (r.u1()) {                                      // ref_pic_list_modification_flag_l0
            for (;;) {
                const int idc = (int)r.ue();
                if (idc == 3) break;
                if (idc > 3 || h->n_mods >= 66 || r.overrun()) return kErrBitstream;
                h->mods[h->n_mods].idc = idc;
                h->mods[h->n_mods].value = (int)r.ue();
                h->n_mods++;
            }
        }
        if (pps->weighted_pred_flag) {
            h->luma_log2_weight_denom = (int)r.ue();
            h->chroma_log2_weight_denom = (int)r.ue();
            if (h->luma_log2_weight_denom > 7 || h->chroma_log2_weight_denom > 7) return kErrBitstream;
            for (int i = 0; i < h->num_ref_idx_l0; i++) {
                h->luma_weight[i] = 1 << h->luma_log2_weight_denom;
                h->luma_offset[i] = 0;
                h->luma_weight_flag[i] = (int)r.u1();
                if (h->luma_weight_flag[i]) { h->luma_weight[i] = r.se(); h->luma_offset[i] = r.se(); }
                h->chroma_weight_flag[i] = (int)r.u1();
                for (int c = 0; c < 2; c++) {
                    h->chroma_weight[i][c] = 1 << h->chroma_log2_weight_denom;
                    h->chroma_offset[i][c] = 0;
                    if (h->chroma_weight_flag[i]) { h->chroma_weight[i][c] = r.se(); h->chroma_offset[i][c] = r.se(); }
                }
            }
        }
    }
    if (nal_ref_idc != 0) {
        if (h->idr()) {
            h->no_output_of_prior_pics_flag = (int)r.u1();
            h->long_term_reference_flag = (int)r.u1();
        } else {
            h->adaptive_marking = (int)r.u1();
            if (h->adaptive_marking) {
                for (;;) {
                    const int op = (int)r.ue();
                    if (op == 0) break;
                    if (op > 6 || h->n_mmco >= 66 || r.overrun()) return kErrBitstream;
                    Mmco& m = h->mmco[h->n_mmco++];
                    m.op = op; m.a = 0; m.b = 0;
                    if (op == 1 || op == 3) m.a = (int)r.ue();          // difference_of_pic_nums_minus1
                    if (op == 2) m.a = (int)r.ue();                      // long_term_pic_num
                    if (op == 3 || op == 6) m.b = (int)r.ue();          // long_term_frame_idx
                    if (op == 4) m.a = (int)r.ue();                      // max_long_term_frame_idx_plus1
                }
            }
        }
    }
    if (pps->entropy_coding_mode_flag && h->slice_type != kSliceI) {
        h->cabac_init_idc = (int)r.ue();
        if (h->cabac_init_idc > 2) return kErrBitstream;
    }
    h->slice_qp_delta = r.se();
    if (pps->deblocking_filter_control_present_flag) {
        h->disable_deblocking_filter_idc = (int)r.ue();
        if (h->disable_deblocking_filter_idc > 2) return kErrBitstream;
        if (h->disable_deblocking_filter_idc != 1) {
            h->alpha_c0_offset_div2 = r.se();
            h->beta_offset_div2 = r.se();
        }
    }
    if (r.overrun()) return kErrBitstream;
    return kOk;
}

inline void write_slice_header(BitWriter& w, const Sps& sps, const Pps& pps, const SliceHeader& h)
{
    w.ue((uint32_t)h.first_mb);
    w.ue((uint32_t)h.slice_type_raw);
    w.ue((uint32_t)h.pps_id);
    w.put((uint32_t)h.frame_num, sps.log2_max_frame_num);
    if (h.idr()) w.ue((uint32_t)h.idr_pic_id);
    if (sps.poc_type == 0) w.put((uint32_t)h.poc_lsb, sps.log2_max_poc_lsb);
    if (h.slice_type == kSliceP) {
        const bool ovr = h.num_ref_idx_l0 != pps.num_ref_idx_l0_default;
        w.put1(ovr);
        if (ovr) w.ue((uint32_t)(h.num_ref_idx_l0 - 1));
        w.put1(h.n_mods > 0);
        if (h.n_mods > 0) {
            for (int i = 0; i < h.n_mods; i++) { w.ue((uint32_t)h.mods[i].idc); w.ue((uint32_t)h.mods[i].value); }
            w.ue(3);
        }
        if (pps.weighted_pred_flag) {
            w.ue((uint32_t)h.luma_log2_weight_denom);
            w.ue((uint32_t)h.chroma_log2_weight_denom);
            for (int i = 0; i < h.num_ref_idx_l0; i++) {
                w.put1(h.luma_weight_flag[i] != 0);
                if (h.luma_weight_flag[i]) { w.se(h.luma_weight[i]); w.se(h.luma_offset[i]); }
                w.put1(h.chroma_weight_flag[i] != 0);
                if (h.chroma_weight_flag[i])
                    for (int c = 0; c < 2; c++) { w.se(h.chroma_weight[i][c]); w.se(h.chroma_offset[i][c]); }
            }
        }
    }
    if (h.nal_ref_idc != 0) {
        if (h.idr()) { w.put1((uint32_t)h.no_output_of_prior_pics_flag); w.put1((uint32_t)h.long_term_reference_flag); }
        else {
            w.put1(h.n_mmco > 0);
            if (h.n_mmco > 0) {
                for (int i = 0; i < h.n_mmco; i++) {
                    const Mmco& m = h.mmco[i];
                    w.ue((uint32_t)m.op);
                    if (m.op == 1 || m.op == 3) w.ue((uint32_t)m.a);
                    if (m.op == 2) w.ue((uint32_t)m.a);
                    if (m.op == 3 || m.op == 6) w.ue((uint32_t)m.b);
                    if (m.op == 4) w.ue((uint32_t)m.a);
                }
                w.ue(0);
            }
        }
    }
    if (pps.entropy_coding_mode_flag && h.slice_type != kSliceI) w.ue((uint32_t)h.cabac_init_idc);
    w.se(h.slice_qp_delta);
    if (pps.deblocking_filter_control_present_flag) {
        w.ue((uint32_t)h.disable_deblocking_filter_idc);
        if (h.disable_deblocking_filter_idc != 1) { w.se(h.alpha_c0_offset_div2); w.se(h.beta_offset_div2); }
    }
}

}  // namespace h264host
