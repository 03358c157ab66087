// Stream synthesizer: abstract syntax (include/h264_syntax.h) -> Annex-B bytes + syntax-element queue.
//
// The image has no H.264 encoder, so every stream used by the tests and by bench.py's reference
// arm is written here.  Headers are laid out field by field as the reference reads them
// (NAL::decode_SPS h264/NAL.cpp:31-116, NAL::decode_PPS h264/NAL.cpp:117-184,
// Slice::PraseSliceHeader h264/slice.cpp:14-275); slice data is not arithmetic-coded: the
// macroblock layer is emitted as the sequence of values Parser::read_ae would return, in the
// order macroblock::Parse (h264/macroblock.cpp:143-475), macroblock::Calc_residual
// (h264/macroblock.cpp:476-514), residual::Parse (h264/residual.cpp:21-129) and
// residual_block_cabac (h264/residual.cpp:561-614) ask for them.
#include "../../../include/h264_syntax.h"
#include <vector>

namespace {

struct BitWriter {
    std::vector<uint8_t> bytes;
    uint32_t acc = 0;
    int nbits = 0;
    void put(uint32_t v, int n)
    {
        for (int i = n - 1; i >= 0; i--) {
            acc = (acc << 1) | ((v >> i) & 1u);
            if (++nbits == 8) { bytes.push_back((uint8_t)acc); acc = 0; nbits = 0; }
        }
    }
    void ue(uint32_t v)
    {
        uint32_t x = v + 1;
        int len = 0;
        while ((x >> len) > 1) len++;
        put(0, len);
        put(x, len + 1);
    }
    void se(int32_t v) { ue(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
    void trailing()            // rbsp_trailing_bits
    {
        put(1, 1);
        while (nbits) put(0, 1);
    }
    void align_ones()          // cabac_alignment_one_bit (h264/slice.cpp:290)
    {
        while (nbits) put(1, 1);
    }
};

// start code + NAL header + payload with emulation-prevention bytes (stripped again by
// Bitsbuf::bufin, h264/bitsbuf.cpp:124-170)
long emit_nal(int nal_ref_idc, int nal_unit_type, const std::vector<uint8_t>& rbsp, uint8_t* out, long cap)
{
    std::vector<uint8_t> o;
    o.push_back(0); o.push_back(0); o.push_back(0); o.push_back(1);
    o.push_back((uint8_t)((nal_ref_idc << 5) | nal_unit_type));
    int zeros = 0;
    for (uint8_t b : rbsp) {
        if (zeros >= 2 && b <= 3) { o.push_back(3); zeros = 0; }
        o.push_back(b);
        zeros = (b == 0) ? zeros + 1 : 0;
    }
    if ((long)o.size() > cap) return -1;
    for (size_t i = 0; i < o.size(); i++) out[i] = o[i];
    return (long)o.size();
}

struct Queue {
    int32_t* q; long cap; long n; bool overflow;
    void push(int code, int value)
    {
        if (n + 2 > cap) { overflow = true; return; }
        q[n++] = code; q[n++] = value;
    }
};

// residual_block_cabac (h264/residual.cpp:561-614): levels[0..len) in scan order, where
// len = endIdx + 1 is maxNumCoeff of the block.
void put_block(Queue& q, const int16_t* levels, int len)
{
    int last = -1;
    for (int i = 0; i < len; i++) if (levels[i]) last = i;
    q.push(H264S_SE_CBF, last >= 0);
    if (last < 0) return;
    int num = len;
    for (int i = 0; i < num - 1; i++) {
        int sig = levels[i] != 0;
        q.push(H264S_SE_SIG, sig);
        if (sig) {
            int is_last = (i == last);
            q.push(H264S_SE_LAST, is_last);
            if (is_last) num = i + 1;
        }
    }
    // num-1 is the last significant coefficient (inferred when it is the final position)
    for (int i = num - 1; i >= 0; i--) {
        if (i == num - 1 || levels[i]) {
            int v = levels[i];
            q.push(H264S_SE_ABS_M1, (v < 0 ? -v : v) - 1);
            q.push(H264S_SE_SIGN, v < 0);
        }
    }
}

const int kPartCount[7] = {0, 0, 1, 2, 2, 4, 1};
const int kSubCount[4] = {1, 2, 2, 4};

}  // namespace

extern "C" long h264s_write_sps_pps(const h264s_seq* s, uint8_t* out, long cap)
{
    BitWriter sps;
    sps.put(s->profile_idc, 8);
    sps.put(0, 8);                       // constraint_set0..5 + reserved_zero_2bits
    sps.put(40, 8);                      // level_idc
    sps.ue(0);                           // seq_parameter_set_id
    sps.ue(s->log2_max_frame_num_minus4);
    sps.ue(0);                           // pic_order_cnt_type
    sps.ue(s->log2_max_poc_lsb_minus4);
    sps.ue(s->max_num_ref_frames);
    sps.put(0, 1);                       // gaps_in_frame_num_value_allowed_flag
    sps.ue(s->w_mbs - 1);
    sps.ue(s->h_mbs - 1);
    sps.put(1, 1);                       // frame_mbs_only_flag
    sps.put(1, 1);                       // direct_8x8_inference_flag
    sps.put(0, 1);                       // frame_cropping_flag
    sps.put(0, 1);                       // vui_parameters_present_flag
    sps.trailing();
    BitWriter pps;
    pps.ue(0);                           // pic_parameter_set_id
    pps.ue(0);                           // seq_parameter_set_id
    pps.put(1, 1);                       // entropy_coding_mode_flag (the reference has no CAVLC)
    pps.put(0, 1);                       // bottom_field_pic_order_in_frame_present_flag
    pps.ue(0);                           // num_slice_groups_minus1
    pps.ue(s->num_ref_idx_l0_default_active_minus1);
    pps.ue(0);                           // num_ref_idx_l1_default_active_minus1
    pps.put(s->weighted_pred_flag, 1);
    pps.put(0, 2);                       // weighted_bipred_idc
    pps.se(s->pic_init_qp - 26);
    pps.se(0);                           // pic_init_qs_minus26
    pps.se(s->chroma_qp_index_offset);
    pps.put(s->deblocking_filter_control_present_flag, 1);
    pps.put(0, 1);                       // constrained_intra_pred_flag
    pps.put(0, 1);                       // redundant_pic_cnt_present_flag
    pps.trailing();
    long a = emit_nal(3, 7, sps.bytes, out, cap);
    if (a < 0) return -1;
    long b = emit_nal(3, 8, pps.bytes, out + a, cap - a);
    if (b < 0) return -1;
    return a + b;
}

extern "C" long h264s_write_slice_nal(const h264s_seq* s, const h264s_pic* p, uint8_t* out, long cap)
{
    BitWriter w;
    const bool is_p = p->slice_type == 0;
    w.ue(0);                                           // first_mb_in_slice
    w.ue(is_p ? 5 : 7);                                // slice_type: the reference needs 5..9 (h264/slice.cpp:19)
    w.ue(0);                                           // pic_parameter_set_id
    w.put(p->frame_num, s->log2_max_frame_num_minus4 + 4);
    if (p->is_idr) w.ue(p->idr_pic_id);
    w.put(p->poc_lsb, s->log2_max_poc_lsb_minus4 + 4);
    if (is_p) {
        bool ovr = p->num_ref_idx_l0_active_minus1 != s->num_ref_idx_l0_default_active_minus1;
        w.put(ovr, 1);                                 // num_ref_idx_active_override_flag
        if (ovr) w.ue(p->num_ref_idx_l0_active_minus1);
        w.put(0, 1);                                   // ref_pic_list_modification_flag_l0
        if (s->weighted_pred_flag) {                   // pred_weight_table (h264/slice.cpp:116-184)
            w.ue(p->luma_log2_weight_denom);
            w.ue(p->chroma_log2_weight_denom);
            for (int i = 0; i <= p->num_ref_idx_l0_active_minus1; i++) {
                w.put(p->luma_weight_flag[i] != 0, 1);
                if (p->luma_weight_flag[i]) { w.se(p->luma_weight[i]); w.se(p->luma_offset[i]); }
                w.put(0, 1);                           // chroma_weight_l0_flag
            }
        }
    }
    if (p->nal_ref_idc) {
        if (p->is_idr) { w.put(0, 1); w.put(0, 1); }   // no_output_of_prior_pics_flag, long_term_reference_flag
        else w.put(0, 1);                              // adaptive_ref_pic_marking_mode_flag: sliding window
    }
    if (is_p) w.ue(0);                                 // cabac_init_idc
    w.se(p->slice_qp_delta);
    if (s->deblocking_filter_control_present_flag) {
        w.ue(p->disable_deblocking_filter_idc);
        if (p->disable_deblocking_filter_idc != 1) {
            w.se(p->slice_alpha_c0_offset_div2);
            w.se(p->slice_beta_offset_div2);
        }
    }
    w.align_ones();
    // slice data travels in the queue; two filler bytes keep the NAL non-empty after the header
    w.put(0x80, 8);
    w.put(0x80, 8);
    return emit_nal(p->nal_ref_idc, p->is_idr ? 5 : 1, w.bytes, out, cap);
}

extern "C" long h264s_serialize_mbs(const h264s_seq* s, const h264s_pic* p, const h264s_mb* mbs, int n_mbs,
                                    int32_t* queue, long cap)
{
    Queue q{queue, cap, 0, false};
    const bool is_p = p->slice_type == 0;
    const bool send_ref = p->num_ref_idx_l0_active_minus1 > 0;
    for (int a = 0; a < n_mbs; a++) {
        const h264s_mb& m = mbs[a];
        if (is_p) q.push(H264S_SE_MB_SKIP, m.kind == H264S_PSKIP);
        if (m.kind != H264S_PSKIP) {
            int mb_type;                                   // MbTypeName value (h264/enums.h:13-73)
            if (m.kind == H264S_I4x4) mb_type = 0;
            else if (m.kind == H264S_I16x16) mb_type = 1 + m.i16_mode + 4 * m.cbp_chroma + (m.cbp_luma ? 12 : 0);
            else mb_type = 40 + (m.kind - H264S_P16x16);
            q.push(H264S_SE_MB_TYPE, mb_type);
            if (m.kind == H264S_P8x8) {
                for (int i = 0; i < 4; i++) q.push(H264S_SE_SUB_MB_TYPE, m.sub_type[i]);
                if (send_ref) for (int i = 0; i < 4; i++) q.push(H264S_SE_REF_IDX, m.ref_idx[i]);
                for (int i = 0; i < 4; i++)
                    for (int j = 0; j < kSubCount[m.sub_type[i] & 3]; j++) {
                        q.push(H264S_SE_MVD, m.mvd[i * 4 + j][0]);
                        q.push(H264S_SE_MVD, m.mvd[i * 4 + j][1]);
                    }
            } else if (m.kind == H264S_I4x4) {
                for (int i = 0; i < 16; i++) {
                    q.push(H264S_SE_PREV_I4_FLAG, m.prev_flag[i] != 0);
                    if (!m.prev_flag[i]) q.push(H264S_SE_REM_I4_MODE, m.rem_mode[i]);
                }
                q.push(H264S_SE_CHROMA_MODE, m.chroma_mode);
            } else if (m.kind == H264S_I16x16) {
                q.push(H264S_SE_CHROMA_MODE, m.chroma_mode);
            } else {
                int np = kPartCount[m.kind];
                if (send_ref) for (int i = 0; i < np; i++) q.push(H264S_SE_REF_IDX, m.ref_idx[i]);
                for (int i = 0; i < np; i++) {
                    q.push(H264S_SE_MVD, m.mvd[i * 4][0]);
                    q.push(H264S_SE_MVD, m.mvd[i * 4][1]);
                }
            }
            if (m.kind != H264S_I16x16) q.push(H264S_SE_CBP, m.cbp_luma + 16 * m.cbp_chroma);
            if (m.cbp_luma || m.cbp_chroma || m.kind == H264S_I16x16) {
                q.push(H264S_SE_QP_DELTA, m.qp_delta);
                if (m.kind == H264S_I16x16) {
                    put_block(q, m.luma_dc, 16);
                    for (int b = 0; b < 16; b++)
                        if (m.cbp_luma & (1 << (b >> 2))) put_block(q, m.luma[b], 15);
                } else {
                    for (int b = 0; b < 16; b++)
                        if (m.cbp_luma & (1 << (b >> 2))) put_block(q, m.luma[b], 16);
                }
                if (m.cbp_chroma & 3) for (int c = 0; c < 2; c++) put_block(q, m.chroma_dc[c], 4);
                if (m.cbp_chroma & 2)
                    for (int c = 0; c < 2; c++)
                        for (int b = 0; b < 4; b++) put_block(q, m.chroma_ac[c][b], 15);
            }
        }
        q.push(H264S_SE_END_OF_SLICE, a == n_mbs - 1);
    }
    return q.overflow ? -1 : q.n;
}
