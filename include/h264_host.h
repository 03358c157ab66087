/*
 * h264_host.h -- C interface of the host side that sits ABOVE the reconstruction ABI (include/h264r.h): the serial part of an
 * H.264 decoder -- Annex-B / NAL parsing, parameter sets, slice headers, CABAC and CAVLC entropy decoding, motion-vector /
 * intra-mode / QP derivation, decoded-picture-buffer bookkeeping -- producing, per picture, exactly the buffers h264r_submit_*
 * takes; and the inverse, a stream writer for both entropy codes (test and benchmark tooling: the image has no H.264 encoder;
 * profile_idc 66, Baseline, selects CAVLC).  libh264host.so, plain C ABI.
 *
 * What it stands in for in the reference (Terice/VideoDecoder, h264/):
 *   h264d_open / h264d_next        main.cpp:20-45 NAL loop, Parser::find_nextNAL (Parser.cpp:21-53), NAL::decode (NAL.cpp:14-30),
 *                                  decode_SPS / decode_PPS (NAL.cpp:31-184), decode_PIC (NAL.cpp:185-316),
 *                                  Slice::PraseSliceHeader / PraseSliceDataer (slice.cpp:14-389), cabac (cabac.cpp),
 *                                  macroblock::Parse / Init0 (macroblock.cpp:81-475), residual::Parse (residual.cpp:21-129, 561-614),
 *                                  picture::get_MvNeighbour (picture.cpp:280-394), Decoder (Decoder.cpp:133-497)
 *   h264d_picture                  what macroblock::Calc + residual::Decode would consume next: the hand-over to h264r_*
 *   h264e_*                        no counterpart (the reference only decodes)
 *
 * The reconstruction itself is NOT here: there is no CPU path for it (include/h264r.h).
 */
#ifndef H264_HOST_H_
#define H264_HOST_H_
#include <stddef.h>
#include <stdint.h>

#include "h264_syntax.h"
#include "h264r.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- decoder ------------------------------------------------------------------------------------------------------------------- */
#define H264D_OK             0
#define H264D_E_BITSTREAM   -1   /* damaged or truncated stream                                                   */
#define H264D_E_UNSUPPORTED -2   /* valid H.264 outside the scope of the engine (see h264d_error for which tool)  */
#define H264D_E_STATE       -3
#define H264D_E_CAPACITY    -4   /* frame pool too small for the stream's reference structure                     */

/* flags of h264d_open */
#define H264D_REF_QUIRKS   0x1u  /* derive vectors / intra modes / chroma QP the way the reference's own host side does where it
                                    deviates from the standard (D2, D11, D12, picture.cpp:205-394): the pictures then match
                                    the reference decoder's, and an engine created with H264R_REF_COMPAT reconstructs them            */
#define H264D_KEEP_SYNTAX  0x2u  /* keep the syntax-element records of the last picture (h264d_picture.syntax)                     */

typedef struct h264d h264d;

typedef struct h264d_frame_ref { int32_t frame_id; int32_t poc; int32_t decode_index; } h264d_frame_ref;

typedef struct h264d_picture {
    int32_t w_mbs, h_mbs;
    int32_t frame_id;                 /* frame-pool slot to reconstruct into (h264r_frame_begin)                    */
    int32_t decode_index;             /* 0, 1, 2 ... in decoding order                                              */
    int32_t is_idr, nal_ref_idc, frame_num, poc, slice_qp;
    h264r_pic_params pp;              /* ref_list0 holds frame-pool slots                                           */
    const h264r_mb_hdr* hdr;          /* [n_mbs]                  valid until the next h264d_next / h264d_close      */
    const int16_t* mv;                /* [n_mbs * 32] per-4x4 layout (P pictures; NULL for I)                        */
    const int16_t* coeff;             /* [n_mbs * 384] dense layout                                                  */
    const h264s_mb* syntax;           /* [n_mbs] with H264D_KEEP_SYNTAX, else NULL                                   */
    int32_t n_intra;                  /* intra macroblocks in hdr[] (h264r_picture_buf.n_intra)                      */
    int32_t n_output;                 /* pictures that became ready for output with this one, display order          */
    h264d_frame_ref output[17];
    int32_t n_release;                /* slots that are neither referenced nor waiting for output any more           */
    int32_t release[17];
    int64_t slice_data_bytes;         /* CABAC bytes consumed for this picture                                       */
} h264d_picture;

/* max_frames: size of the frame pool the DPB may use; 0 = never reuse a slot (frame_id = decode_index; the caller sizes the
 * engine's pool for the whole stream -- what the tests do).  reorder_depth: pictures held back for output reordering (0 for the
 * I / P streams in scope, whose decoding order is their display order). */
h264d* h264d_open(const uint8_t* annexb, size_t n_bytes, uint32_t flags, int max_frames, int reorder_depth);
void   h264d_close(h264d* d);
/* Parses up to and including the next picture.  Returns 1 with *out filled, 0 at the end of the stream (out->n_output / n_release
 * then carry what was still held), or a negative H264D_E_* code. */
int    h264d_next(h264d* d, h264d_picture* out);
const char* h264d_error(const h264d* d);
/* Fills a picture buffer of the engine from the picture returned last (include/h264r_writer.h does the packing): level_fmt =
 * H264R_LEVELS_COMPACT or H264R_LEVELS_INT16.  Returns H264R_OK or an H264R_E_* code (H264R_E_INVAL: a level outside int8 under
 * COMPACT -- call again with INT16). */
int    h264d_fill_picture_buf(h264d* d, h264r_picture_buf* buf, int level_fmt, size_t* n_blocks, long* n_mvs);
/* The same packing for a caller that has a picture's syntax in the engine's dense per-macroblock arrays (hdr [n_mbs], mv
 * [n_mbs * 32] or NULL, coeff [n_mbs * 384], include/h264r.h): runs include/h264r_writer.h over them, in whatever form the buffer
 * was formatted for (h264r_picture_buf_format).  Returns H264R_OK or the writer's code. */
int    h264w_pack_arrays(h264r_picture_buf* buf, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
                         int level_fmt, size_t* n_blocks, long* n_mvs);
/* Host-side throughput: decodes the whole stream `repeat` times without keeping anything; returns pictures parsed, *seconds = CPU
 * wall time of the parsing alone -- or, with H264D_BENCH_PACK in flags, of parsing + packing every picture into the stream form
 * of a picture buffer (plain memory; h264d_fill_picture_buf), i.e. everything a host thread does per picture before
 * h264r_submit_picture. */
#define H264D_BENCH_PACK 0x100u
long   h264d_benchmark(const uint8_t* annexb, size_t n_bytes, uint32_t flags, int repeat, double* seconds);

/* ---- stream writer (tooling) ----------------------------------------------------------------------------------------------------- */
#define H264E_ABSOLUTE     0x1u  /* mbs[].mvd holds FINAL vectors and mbs[].rem_mode FINAL intra prediction modes (prev_flag ignored); the
                                    writer derives mvd / prev_intra*_pred_mode_flag / rem_* against its own prediction and stores them
                                    back into mbs[]                                                                                  */
#define H264E_REF_QUIRKS   0x2u  /* ...using the reference's derivations instead of the standard's (see H264D_REF_QUIRKS)            */

/* SPS + PPS NAL units.  profile_idc 100 writes the High-profile SPS fields (4:2:0, 8 bit, flat scaling lists) and, when
 * transform_8x8_mode_flag or second_chroma_qp_index_offset ask for it, the PPS extension. */
long h264e_write_sps_pps(const h264s_seq* seq, uint8_t* out, long cap);
/* One picture = one slice NAL unit, CABAC (CAVLC when seq->profile_idc is 66: Baseline profile has no CABAC).  mbs[] is normalised in place where the syntax cannot express it (coded_block_pattern
 * bits of empty 8x8 blocks, mb_qp_delta of macroblocks without residual, 8x8 transform flags that the macroblock type forbids).
 * Returns bytes written or a negative H264D_E_* code. */
long h264e_encode_picture(const h264s_seq* seq, const h264s_pic* pic, h264s_mb* mbs, int n_mbs, uint32_t flags, uint8_t* out, long cap);

#ifdef __cplusplus
}
#endif
#endif
