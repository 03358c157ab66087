/*
 * h264_syntax.h -- host-side "abstract syntax" of one picture: the values of the syntax
 * elements in the order and granularity at which the reference's macroblock layer consumes
 * them (macroblock::Parse, h264/macroblock.cpp:143-475; residual::Parse +
 * residual_block_cabac, h264/residual.cpp:21-129, 561-614).
 *
 * Two things use it:
 *   - the stream synthesizer (videodecoder_b200/csrc/host/h264_synth.cpp) turns it into an
 *     Annex-B stream (SPS/PPS/slice headers exactly as NAL::decode_SPS / decode_PPS /
 *     Slice::PraseSliceHeader read them, h264/NAL.cpp:31-184, h264/slice.cpp:14-275) plus a
 *     syntax-element queue: (code, value) int32 pairs in the order the reference calls
 *     Parser::read_ae (h264/Parser.cpp:18).  There is no H.264 encoder in this image, so this
 *     is how every test stream is made;
 *   - the host parser (h264_host.cpp) consumes the same queue through the SyntaxSource
 *     interface and emits the h264r_* buffers of include/h264r.h.
 *
 * Plain C, fixed layout (mirrored as a numpy dtype in videodecoder_b200/synth.py).
 */
#ifndef H264_SYNTAX_H_
#define H264_SYNTAX_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* syntax-element codes of the queue = the reference's request codes (h264/cabac.cpp:55-78);
 * codes >= 0x1000 are reduced to their class byte (code >> 12) & 0xFF. */
#define H264S_SE_MB_SKIP        11
#define H264S_SE_MB_TYPE        21
#define H264S_SE_T8X8           22
#define H264S_SE_CBP            23
#define H264S_SE_QP_DELTA       25
#define H264S_SE_PREV_I4_FLAG   31
#define H264S_SE_REM_I4_MODE    32
#define H264S_SE_CHROMA_MODE    35
#define H264S_SE_SUB_MB_TYPE    51
#define H264S_SE_REF_IDX        0x41
#define H264S_SE_MVD            0x43
#define H264S_SE_CBF            0x61
#define H264S_SE_SIG            0x62
#define H264S_SE_LAST           0x63
#define H264S_SE_ABS_M1         0x64
#define H264S_SE_SIGN           65
#define H264S_SE_END_OF_SLICE   276

/* h264s_mb.kind */
#define H264S_I4x4    0
#define H264S_I16x16  1
#define H264S_P16x16  2
#define H264S_P16x8   3
#define H264S_P8x16   4
#define H264S_P8x8    5
#define H264S_PSKIP   6
#define H264S_I8x8    7      /* I_NxN with transform_size_8x8_flag (no reference counterpart: Prediction_Intra8x8 is an empty stub,
                                h264/macroblock.cpp:1369-1372) */

typedef struct h264s_mb {
    uint8_t kind;
    uint8_t i16_mode;        /* Intra16x16PredMode 0..3                                     */
    uint8_t cbp_luma;        /* CodedBlockPatternLuma (Intra16x16: 0 or 15)                 */
    uint8_t cbp_chroma;      /* CodedBlockPatternChroma 0..2                                */
    uint8_t chroma_mode;     /* intra_chroma_pred_mode 0..3                                 */
    int8_t  qp_delta;        /* mb_qp_delta (sent only when the MB has residual syntax)     */
    uint8_t t8x8;            /* transform_size_8x8_flag: luma[4*i .. 4*i+3] then hold the 64 levels of 8x8 block i in 8x8 zig-zag order */
    uint8_t pad_;
    uint8_t prev_flag[16];   /* prev_intra4x4_pred_mode_flag[luma4x4BlkIdx] (Intra8x8: [luma8x8BlkIdx], 4 used) */
    uint8_t rem_mode[16];    /* rem_intra4x4_pred_mode[luma4x4BlkIdx]       (Intra8x8: rem_intra8x8_pred_mode)  */
    uint8_t sub_type[4];     /* sub_mb_type per 8x8 (0 8x8, 1 8x4, 2 4x8, 3 4x4)            */
    uint8_t ref_idx[4];      /* ref_idx_l0 per partition                                    */
    int16_t mvd[16][2];      /* mvd_l0[mbPartIdx*4 + subMbPartIdx][x,y]                     */
    int16_t luma_dc[16];     /* Intra16x16DCLevel, scan (zig-zag) order                     */
    int16_t luma[16][16];    /* [luma4x4BlkIdx][scan idx]; Intra16x16 AC uses idx 0..14     */
    int16_t chroma_dc[2][4]; /* [iCbCr][chroma4x4BlkIdx]                                    */
    int16_t chroma_ac[2][4][16]; /* [iCbCr][blk][scan idx 0..14 of the AC run]              */
} h264s_mb;                  /* 928 bytes */

typedef struct h264s_seq {
    int32_t w_mbs, h_mbs;
    int32_t max_num_ref_frames;
    int32_t pic_init_qp;               /* 26 + pic_init_qp_minus26                          */
    int32_t chroma_qp_index_offset;
    int32_t num_ref_idx_l0_default_active_minus1;
    int32_t weighted_pred_flag;
    int32_t deblocking_filter_control_present_flag;
    int32_t log2_max_frame_num_minus4;
    int32_t log2_max_poc_lsb_minus4;
    int32_t profile_idc;               /* 77: the reference misparses High SPS fields (h264/NAL.cpp:45-57) */
    /* fields below are read by the h264e_* writer only (streams for spec-conformant decoders; the reference has no 8x8 transform) */
    int32_t transform_8x8_mode_flag;
    int32_t constrained_intra_pred_flag;
    int32_t second_chroma_qp_index_offset;
    int32_t pic_order_cnt_type;        /* 0 or 2 */
} h264s_seq;

typedef struct h264s_pic {
    int32_t is_idr;
    int32_t slice_type;                /* 0 = P, 2 = I (reference Slicetype, h264/enums.h:4-10) */
    int32_t nal_ref_idc;
    int32_t frame_num;
    int32_t idr_pic_id;
    int32_t poc_lsb;
    int32_t slice_qp_delta;
    int32_t num_ref_idx_l0_active_minus1;
    int32_t luma_log2_weight_denom;
    int32_t chroma_log2_weight_denom;
    int32_t luma_weight_flag[16];
    int32_t luma_weight[16];
    int32_t luma_offset[16];
    int32_t disable_deblocking_filter_idc;
    int32_t slice_alpha_c0_offset_div2;
    int32_t slice_beta_offset_div2;
    /* read by the h264e_* writer only */
    int32_t cabac_init_idc;
    int32_t long_term_reference_flag;
    int32_t chroma_weight_flag[16];
    int32_t chroma_weight[16][2];
    int32_t chroma_offset[16][2];
    int32_t n_mods;                    /* ref_pic_list_modification: (modification_of_pic_nums_idc, value) pairs */
    int32_t mods[16][2];
    int32_t n_mmco;                    /* dec_ref_pic_marking: (operation, first operand, second operand) */
    int32_t mmco[16][3];
} h264s_pic;

/* All writers return the number of bytes / int32 values produced, or -1 if cap is too small. */
long h264s_write_sps_pps(const h264s_seq* seq, uint8_t* out, long cap);
long h264s_write_slice_nal(const h264s_seq* seq, const h264s_pic* pic, uint8_t* out, long cap);
long h264s_serialize_mbs(const h264s_seq* seq, const h264s_pic* pic, const h264s_mb* mbs, int n_mbs,
                         int32_t* queue, long cap);

#ifdef __cplusplus
}
#endif
#endif
