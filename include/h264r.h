/*
 * h264r.h -- C ABI of the B200 macroblock-reconstruction engine (libh264r.so).
 *
 * This is the drop-in boundary for the hot path of Terice/VideoDecoder.  The reference has no
 * plugin/FFI interface: it interleaves entropy decoding and reconstruction per macroblock inside
 * Slice::PraseSliceDataer (h264/slice.cpp:278-389).  The seam is cut where nothing downstream feeds
 * back into entropy decoding (CABAC contexts read neighbour *syntax* only): the host keeps
 * NAL/SPS/PPS/slice-header parsing, CABAC and MV / intra-mode / QP derivation, and hands this
 * library flat per-picture syntax buffers.  Each entry point below names the reference call it
 * replaces or is driven by.
 *
 * Conventions
 *  - every function returns 0 (H264R_OK) or a negative H264R_E_* code; the library never calls
 *    exit() (the reference does: h264/macroblock.cpp:656-661, h264/block.cpp:50-51);
 *  - the caller owns every pointer it passes and may reuse it as soon as the call returns,
 *    except buffers obtained from h264r_host_alloc() / h264r_picture_buf_alloc(), which are read by an
 *    asynchronous copy and must stay untouched until h264r_wait_uploads(), h264r_sync() or a read-back of the
 *    picture returns (h264r_flush() alone does not wait for the copy);
 *  - a context is bound to one CUDA device and must be used from one host thread at a time
 *    (the reference is single-threaded); different contexts are independent;
 *  - there is NO CPU fallback: if no CUDA device is usable h264r_create() fails.
 *
 * Plain C, no CUDA or torch types.
 */
#ifndef H264R_H_
#define H264R_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H264R_ABI_VERSION 7

/* ---- status codes ---------------------------------------------------------------------- */
#define H264R_OK            0
#define H264R_E_INVAL      -1   /* bad argument                                              */
#define H264R_E_CUDA       -2   /* CUDA runtime / driver error (see h264r_last_error)         */
#define H264R_E_NOMEM      -3   /* device or pinned-host allocation failed                    */
#define H264R_E_STATE      -4   /* call sequence violated (e.g. submit without frame_begin)   */
#define H264R_E_NODEVICE   -5   /* no usable CUDA device: the engine has no CPU path          */
#define H264R_E_CAPACITY   -6   /* more pending pictures / frames than the context was sized for */

/* ---- context flags (h264r_create) ---------------------------------------------------------
 * H264R_REF_COMPAT reproduces, bit for bit, the reference decoder's deviations from the
 * standard on the path it implements:
 *   D1 chroma DC "transform" = c[1][1]*[[1,-1],[-1,1]]          (h264/matrix.cpp:150-171 via h264/residual.cpp:214)
 *   D3 no Clip1 after pred+residual; planes are the low 8 bits  (h264/macroblock.cpp:1467-1481, h264/picture.cpp:396-409)
 *   D4 no chroma prediction of any kind: U/V = residual         (h264/macroblock.cpp:1477-1480)
 *   D5 Intra16x16 DC, top-only: sums column 15 of the MB above  (h264/macroblock.cpp:1410-1411)
 *   D6/D7 quarter-pel tap coordinates                           (h264/staticcharts.h:1505, h264/macroblock.cpp:843-846)
 *   D8 source offset of 4x4 sub-partitions 2,3                  (h264/macroblock.cpp:786-789)
 *   no deblocking filter, 4x4 transform only.
 * (D2, QPc = max(QPY'-2,0), is host-side: the caller passes qp_cb/qp_cr per macroblock.)
 * Without the flag the engine follows ITU-T H.264 (03/2014): Clip1, 2x2 Hadamard, chroma
 * intra prediction and MC, standard interpolation, 8x8 transform, Intra8x8, deblocking.
 */
#define H264R_REF_COMPAT   0x1u
/* H264R_DIGESTS: compute the digest of every picture as part of its reconstruction (kernel K5 queued with each wave),
 * so that h264r_frame_digests() only reads results back.  For pipelines that check or ship digests of every frame
 * (GOP-sharded decoding across GPUs). */
#define H264R_DIGESTS      0x2u

/* ---- macroblock classes (h264r_mb_hdr.mb_class) --------------------------------------------
 * Collapsed from the reference's MbTypeName (h264/enums.h:13-73): the Intra16x16 mode/cbp
 * variants I_16x16_<mode>_<cbpC>_<cbpL> travel in i16_mode / cbp; P_Skip is P_16x16 with zero
 * coefficients (macroblock::Init0, h264/macroblock.cpp:81-142). */
#define H264R_MB_I4x4      0
#define H264R_MB_I8x8      1
#define H264R_MB_I16x16    2
#define H264R_MB_P16x16    4
#define H264R_MB_P16x8     5
#define H264R_MB_P8x16     6
#define H264R_MB_P8x8      7

/* h264r_mb_hdr.flags */
#define H264R_AVAIL_A      0x01u  /* left MB usable for intra prediction  (macroblock::is_avaiable, h264/macroblock.cpp:1446) */
#define H264R_AVAIL_B      0x02u  /* MB above                                                    */
#define H264R_AVAIL_C      0x04u  /* MB above-right                                              */
#define H264R_AVAIL_D      0x08u  /* MB above-left                                               */
#define H264R_T8x8         0x10u  /* transform_size_8x8_flag                                     */
#define H264R_SKIP         0x20u  /* informational: P_Skip                                       */

/* sub_mb_type codes (2 bits per 8x8 quadrant, quadrant q in bits 2q..2q+1), P_8x8 only;
 * same numbering as the reference's SubMbTypeName P_L0_* (h264/enums.h:89-93). */
#define H264R_SUB_8x8      0
#define H264R_SUB_8x4      1
#define H264R_SUB_4x8      2
#define H264R_SUB_4x4      3

/* One record per macroblock, raster order within the picture.  16 bytes. */
typedef struct h264r_mb_hdr {
    uint8_t mb_class;      /* H264R_MB_*                                                       */
    uint8_t flags;         /* H264R_AVAIL_* | H264R_T8x8 | H264R_SKIP                          */
    uint8_t qp_y;          /* QP'Y used for luma dequantisation and deblocking (macroblock::QPY_, h264/macroblock.cpp:485) */
    uint8_t qp_cb;         /* QP'C for Cb (reference: macroblock::QPC_, h264/macroblock.cpp:487-488) */
    uint8_t qp_cr;         /* QP'C for Cr                                                      */
    uint8_t cbp;           /* bits 0-3 CodedBlockPatternLuma, bits 4-5 CodedBlockPatternChroma */
    uint8_t intra_modes;   /* bits 0-1 Intra16x16PredMode, bits 2-3 intra_chroma_pred_mode     */
    uint8_t sub_mb_type;   /* 4 x 2 bits, H264R_SUB_*                                          */
    union {
        uint8_t pred4x4[8];   /* I4x4: Intra4x4PredMode[luma4x4BlkIdx], 4 bits each, even idx in the low nibble;
                                 I8x8: Intra8x8PredMode[0..3] in pred4x4[0..1] the same way     */
        int8_t  ref_idx[8];   /* P: ref_idx_l0 per 8x8 quadrant in [0..3] (index into ref_list0) */
    } u;
} h264r_mb_hdr;

/* The same record in 8 bytes, for picture buffers (h264r_picture_buf.hdr_bytes = 8): half the header bytes on the wire; the
 * engine expands it on the device (k_blk_scan) into the 16-byte record its kernels read.
 *   w0: bits 0-2 mb_class, 3-8 flags, 9-12 intra_modes, 13-18 qp_y, 19-24 qp_cb, 25-30 qp_cr
 *   w1: bits 0-5 cbp, 6-13 sub_mb_type, 14-29 ref_idx[0..3], 4 bits each
 * The prediction modes of an intra macroblock (u.pred4x4, 8 bytes) travel as the FIRST 8-byte unit of the macroblock's
 * coefficient blocks, in front of the units its blk_mask announces. */
typedef struct h264r_mb_hdr8 { uint32_t w0, w1; } h264r_mb_hdr8;
/* STREAM form of a picture buffer (hdr_bytes = H264R_HDR_STREAM; compact levels only): the fewest bytes on the wire.  8-byte
 * headers, one byte per macroblock with the number of 32-bit words the macroblock has in `stream`, and the stream itself, per
 * macroblock and in this order:
 *   - its blk_mask word (bits 0-23 present blocks, H264R_BLK_COMPACT)            only if cbp != 0 or the macroblock is Intra16x16
 *                                                                                 (else the mask is 0)
 *   - u.pred4x4, two words                                                        intra macroblocks
 *   - its vectors in partition order, one word each (x | y << 16)                 inter macroblocks; none for a P_L0_16x16 /
 *     P_Skip macroblock whose vector fits 13 bits per component and rides in the header: w0 bit 31 (H264R_HDR8_MV_INLINE) set,
 *     x = w1 bits 6-13 | w0 bits 9-12 << 8 | w1 bit 30 << 12, y = w1 bits 18-29 | w1 bit 31 << 12, both two's complement
 *     With w1 bit 30 (H264R_HDR8_MV_SHORT; a macroblock without an inline vector) the vectors are int8 pairs (x, y), two bytes
 *     each, possible when every component of the macroblock's vectors lies in [-128, 127]; they then open the byte run below.
 *   - its present blocks in ascending block index: H264R_BLK_COMPACT macroblocks as a BYTE run -- per block the 16-bit
 *     significance map (low byte first) and one byte per level in position order, i.e. the 8-byte unit of
 *     H264R_LEVELS_COMPACT without its padding; other macroblocks four words per block (16 int8 levels, raster), word-aligned.
 *     The bytes of a macroblock (short vectors, byte run) are zero-padded to the next 32-bit word where its int8 blocks begin
 *     and at its end.
 * The engine expands the stream on the device (k_blk_scan) into the forms its kernels read. */
#define H264R_HDR_STREAM       9
#define H264R_HDR8_MV_INLINE   (1u << 31)    /* in w0 */
#define H264R_HDR8_MV_SHORT    (1u << 30)    /* in w1 */

/* Motion vectors: int16 (x, y) in quarter-sample units, one pair per 4x4 luma block in RASTER
 * order inside the macroblock (block b = 4*(y/4) + x/4), 64 bytes per MB; the final vectors
 * mv_l0 = mvp + mvd as computed by picture::get_MvNeighbour + macroblock::Parse
 * (h264/picture.cpp:280-394, h264/macroblock.cpp:391-398).  P pictures only.
 *
 * Coefficients: 384 int16 levels per MB (768 bytes) = 24 blocks x 16:
 *   blocks 0..15  luma, luma4x4BlkIdx order as parsed (h264/residual.cpp:39-47),
 *   blocks 16..19 Cb, 20..23 Cr, chroma4x4BlkIdx (raster 2x2) order.
 * Inside a block the 16 levels are in raster (row-major) position, i.e. the host has already
 * applied the inverse zig-zag scan (Inverse4x4Scan, h264/residual.cpp:136-168).  Intra16x16 luma
 * DC levels and chroma DC levels are stored, before any Hadamard transform, in slot 0 of the
 * block they belong to (luma: DC matrix element (r,c) goes to the block at raster position (r,c);
 * chroma: DC level i goes to chroma block i).  With H264R_T8x8 the four blocks of an 8x8
 * quadrant hold the 64 levels of that 8x8 transform block in raster 8x8 order.
 */
#define H264R_COEFF_PER_MB 384
#define H264R_MV_PER_MB    32    /* int16 values: 16 blocks x (x,y) */

#define H264R_SLICE_P 0          /* numbering of the reference's Slicetype (h264/enums.h:4-10) */
#define H264R_SLICE_I 2

#define H264R_MAX_REFS 16

/* Per-picture parameters: what NAL::decode_PIC (h264/NAL.cpp:185-316) knows after the slice
 * header and reference-list construction. */
typedef struct h264r_pic_params {
    int32_t  slice_type;                      /* H264R_SLICE_I / H264R_SLICE_P                       */
    int32_t  num_ref_idx_l0;                  /* entries used in ref_list0                          */
    int32_t  ref_list0[H264R_MAX_REFS];       /* frame ids (Decoder::list_Ref0, h264/Decoder.h:39)   */
    /* explicit weighted prediction (Weight_CoefficWeight, h264/macroblock.cpp:524-577)              */
    int32_t  weighted_pred;                   /* pps weighted_pred_flag                             */
    int32_t  luma_log2_weight_denom;
    int32_t  chroma_log2_weight_denom;
    int16_t  luma_weight[H264R_MAX_REFS], luma_offset[H264R_MAX_REFS];
    int16_t  chroma_weight[H264R_MAX_REFS][2], chroma_offset[H264R_MAX_REFS][2];
    /* in-loop deblocking (slice header fields parsed at h264/slice.cpp:247-256; the reference has
     * no filter, so these are ignored under H264R_REF_COMPAT)                                      */
    int32_t  disable_deblocking_filter_idc;   /* 1 = off                                            */
    int32_t  slice_alpha_c0_offset_div2;
    int32_t  slice_beta_offset_div2;
} h264r_pic_params;

typedef struct h264r_ctx h264r_ctx;

/* Creates a context on CUDA device `device` for pictures of w_mbs x h_mbs macroblocks.
 * max_frames  = size of the device frame pool (decoded + reference pictures alive at once);
 * max_pending = pictures that may be begun/submitted between two flushes (their syntax buffers
 *               stay resident in HBM until the flush executes).
 * Replaces: the per-picture `new picture(...)` + per-MB `new macroblock` heap graph
 * (h264/NAL.cpp:194, h264/slice.cpp:295) with a planar device frame pool. */
int h264r_create(h264r_ctx** out, int device, int w_mbs, int h_mbs,
                 int max_frames, int max_pending, uint32_t flags);
void h264r_destroy(h264r_ctx* ctx);

/* Pinned host memory for syntax buffers (so submit can DMA straight from the parser's output). */
void* h264r_host_alloc(h264r_ctx* ctx, size_t bytes);
void  h264r_host_free(h264r_ctx* ctx, void* p);

/* Picture bracket.  frame_id names a slot of the frame pool (0 <= frame_id < max_frames) and is
 * what ref_list0 entries of later pictures refer to.
 * Driven by: NAL::decode_PIC after Slice::PraseSliceHeader and Decoder::init_RefPicList /
 * opra_RefModfication (h264/NAL.cpp:204-224). */
int h264r_frame_begin(h264r_ctx* ctx, int frame_id, const h264r_pic_params* pp);

/* Syntax of macroblocks [first_mb, first_mb + n_mbs) of the picture begun last, raster order.
 * mv may be NULL for I pictures.  May be called several times per picture (chunked / per slice).
 * Driven by: the macroblock loop of Slice::PraseSliceDataer (h264/slice.cpp:292-386); carries what
 * macroblock::Parse / Init0 and residual::Parse produced (h264/macroblock.cpp:143-475, 81-142,
 * h264/residual.cpp:21-129). */
int h264r_submit_mbs(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs,
                     const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff);

/* The same, with the coefficients in packed form: blk_mask[i] (bits 0..23) says which of the 24 blocks of macroblock
 * first_mb + i are present; `blocks` holds the present blocks only, 16 levels each (same order inside a block as
 * above), macroblock after macroblock in ascending block index; n_blocks = total number of set bits.  A block that
 * is not present is all zero -- this is what an entropy decoder has in hand after coded_block_flag
 * (residual::residual_block_cabac, h264/residual.cpp:561-614) without first expanding into the reference's dense
 * per-macroblock arrays, and it is what crosses PCIe.  With H264R_T8x8 the four bits of an 8x8 quadrant go together.
 * Chunks of a picture must arrive in macroblock order; dense and packed submits cannot be mixed within a picture. */
int h264r_submit_mbs_packed(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs,
                            const h264r_mb_hdr* hdr, const int16_t* mv,
                            const uint32_t* blk_mask, const int16_t* blocks, size_t n_blocks);

/* Whole pictures with ONE transfer.  A picture buffer is pinned host memory laid out like the engine's device-side
 * syntax slot; the host parser writes headers, vectors, block masks and the present blocks straight into it (same
 * formats as h264r_submit_mbs_packed, whole picture, macroblock order) and h264r_submit_picture() ships it with a
 * single DMA and queues the picture (no h264r_frame_end).  The buffer must stay untouched until the next
 * h264r_sync() (or a read-back of the picture) returns, and may then be refilled.  max_blocks = 0 sizes for the worst case (24 per MB). */
#define H264R_LEVELS_INT16   2
#define H264R_LEVELS_INT8    1
#define H264R_LEVELS_COMPACT 3
#define H264R_BLK_COMPACT    (1u << 24)     /* in blk_mask[a], H264R_LEVELS_COMPACT only */
typedef struct h264r_picture_buf {
    h264r_mb_hdr* hdr;        /* [n_mbs]                    */
    int16_t*      mv;         /* per-4x4 layout: [n_mbs * H264R_MV_PER_MB], forward from here                        */
    int16_t*      mv_end;     /* partition layout: vector i is written at mv_end[-2(i+1)] (x), mv_end[-2(i+1)+1] (y):
                                 the stream grows backward, so that it ends where the headers begin and the picture
                                 is one contiguous range                                                            */
    uint32_t*     blk_mask;   /* [n_mbs]                    */
    int16_t*      blocks;     /* [max_blocks * 16]          */
    size_t        max_blocks;
    int32_t       n_intra;    /* number of intra macroblocks in hdr[], as counted by the parser: exact, or an upper bound below
                                 the picture's macroblock count (it sizes the work list of the intra kernel: macroblocks
                                 beyond the count would not be reconstructed; the device checks the count and the classes
                                 and reports through h264r_device_status());
                                 -1 = unknown: the library then walks hdr[] itself (and validates mb_class while there) */
    int32_t       level_bytes;/* storage of blocks[], one of H264R_LEVELS_*:
                                 _INT16 (2, default): int16 levels, 32 bytes per block;
                                 _INT8 (1): the parser wrote int8 levels, 16 bytes per block, through (int8_t*)blocks -- allowed
                                   when every level of the picture is in [-128, 127] (the common case at the QPs video is
                                   coded at), and half the bytes on the wire;
                                 _COMPACT (3): int8 levels in 8-byte units.  A macroblock whose blocks all have at most six
                                   levels sets H264R_BLK_COMPACT in its blk_mask word and stores each present block as ONE
                                   unit: bytes 0-1 the significance map (bit k = position k, raster, holds a level -- what
                                   significant_coeff_flag says after the inverse scan), bytes 2-7 the levels of the set
                                   positions in ascending position order.  Any other macroblock stores its blocks as in
                                   _INT8, two units each.  n_blocks of h264r_submit_picture counts UNITS.  A quarter of the
                                   _INT16 bytes at two levels per block                                                   */
    void*         base;       /* owned by the library       */
    int32_t       hdr_bytes;  /* 16 (default): hdr[] holds h264r_mb_hdr records.  8: hdr8[] holds h264r_mb_hdr8 records, blk_mask and
                                 blocks follow them directly (set by h264r_picture_buf_format(); H264R_LEVELS_COMPACT only) and every
                                 intra macroblock's first unit holds its u.pred4x4 bytes (counted in n_blocks)                  */
    int32_t       ring_index; /* 0: a buffer of its own; i + 1: member i of a ring (h264r_picture_ring_alloc)                       */
    h264r_mb_hdr8* hdr8;      /* [n_mbs], hdr_bytes = 8 or H264R_HDR_STREAM */
    uint8_t*      mb_words;   /* [n_mbs], H264R_HDR_STREAM: 32-bit words of each macroblock in stream[]                          */
    uint32_t*     stream;     /* H264R_HDR_STREAM: see above; n_blocks of h264r_submit_picture counts its WORDS                  */
    size_t        max_stream_words;
} h264r_picture_buf;
/* (include/h264r_writer.h: header-only helper that fills a picture buffer macroblock by macroblock in these formats.) */
int  h264r_picture_buf_alloc(h264r_ctx* ctx, size_t max_blocks, h264r_picture_buf* out);
/* n picture buffers in ONE pinned allocation, at a constant stride: what a host that parses several streams round-robin fills
 * per round.  h264r_submit_pictures() then ships consecutive members with one strided transfer (a round of half-megabyte pictures
 * moves faster as one DMA than as sixteen).  h264r_picture_buf_free(&out[0]) frees the ring; freeing other members is a no-op. */
int  h264r_picture_ring_alloc(h264r_ctx* ctx, int n, size_t max_blocks, h264r_picture_buf* out /* [n] */);
/* Chooses the header format of a picture buffer before it is filled: hdr_bytes = 16, 8 or H264R_HDR_STREAM.  Re-points hdr / hdr8 / blk_mask /
 * blocks inside the buffer (with 8-byte headers the masks and the blocks move up so that the picture stays one contiguous
 * range, 8 bytes per macroblock shorter). */
int  h264r_picture_buf_format(h264r_ctx* ctx, h264r_picture_buf* buf, int hdr_bytes);
void h264r_picture_buf_free(h264r_ctx* ctx, h264r_picture_buf* buf);
/* n_mvs < 0: buf->mv holds the per-4x4 layout described above (16 vectors per macroblock).
 * n_mvs >= 0: n_mvs vectors in PARTITION order, stored backward from buf->mv_end -- per macroblock, in macroblock order, one vector per
 * partition as the bitstream carries them (P_L0_16x16 and P_Skip: 1; 16x8, 8x16: 2; P_8x8: per quadrant 0..3 its
 * sub-macroblock partitions, 8x8: 1, 8x4 / 4x8: 2, 4x4: 4, in raster order; intra: none; mvd loops of
 * macroblock::Parse, h264/macroblock.cpp:360-398).  The engine derives each macroblock's first vector from
 * mb_class / sub_mb_type on the device. */
int  h264r_submit_picture(h264r_ctx* ctx, int frame_id, const h264r_picture_buf* buf, size_t n_blocks, long n_mvs);

/* h264r_submit_picture for n pictures begun with h264r_frame_begin.  Pictures whose buffers are consecutive members of a ring and
 * which were begun back to back (their syntax slots are then neighbours) travel together in one strided transfer; any other
 * picture travels on its own, so the call is always equivalent to n h264r_submit_picture calls.  On an error at picture k the
 * pictures before it are submitted and the code of picture k is returned. */
int  h264r_submit_pictures(h264r_ctx* ctx, int n, const int* frame_ids, const h264r_picture_buf* bufs, const size_t* n_blocks, const long* n_mvs);

/* Gives up a picture that was begun but not completed (a failed submit, a damaged slice): its syntax slot and its frame
 * slot are free again.  The reference has no counterpart (it exit()s). */
int h264r_frame_abort(h264r_ctx* ctx, int frame_id);

/* Marks the picture complete and queues its reconstruction.  Asynchronous; work is issued to the
 * GPU at the next h264r_flush().  Replaces: residual::Decode + macroblock::Calc for every MB of the
 * picture (h264/residual.cpp:184-201, h264/macroblock.cpp:18-47). */
int h264r_frame_end(h264r_ctx* ctx, int frame_id);

/* Issues all queued pictures.  Pictures whose references are complete (or queued earlier) are
 * grouped into dependency waves; every wave is reconstructed by batched kernel launches, so
 * independent GOPs / streams submitted together fill the GPU. */
int h264r_flush(h264r_ctx* ctx);
/* flush + wait for the device. */
int h264r_sync(h264r_ctx* ctx);

/* Number of independent kernel pipelines ("lanes", one CUDA stream each) the scheduler spreads prediction chains over;
 * 1 = every kernel of a flush in one stream, in order (what per-kernel timing wants).  Clamped to the number the
 * context was created with (1 unless H264R_LANES says otherwise).  Synchronises the context. */
int h264r_set_lanes(h264r_ctx* ctx, int lanes);

/* Waits until every syntax buffer submitted so far is resident in device memory (uploads run on their own
 * stream and overlap the kernels of earlier waves and flushes). */
int h264r_wait_uploads(h264r_ctx* ctx);

/* Reads back the reconstructed planes of frame_id (flushes and waits as needed).  Strides in
 * bytes.  Replaces: picture::get_SampleXY (h264/picture.cpp:396-409), the reference's only sample
 * accessor (used by its ASCII printer, h264/picture.cpp:53-76, and by inter prediction). */
int h264r_read_planes(h264r_ctx* ctx, int frame_id, uint8_t* y, int y_stride,
                      uint8_t* u, uint8_t* v, int c_stride);

/* Asynchronous read-back of whole frames, for pipelines that ship every reconstructed picture to the host: the frame slot -- the
 * three planes with their replicated borders, h264r_frame_layout() says where the samples lie -- is copied to `dst` (frame_bytes
 * bytes; pinned memory from h264r_host_alloc() for a truly asynchronous copy) on a read-back stream, behind the kernels that
 * produce the frame and beside the reconstruction of later pictures.  A later picture that overwrites the slot is ordered behind
 * the copy.  h264r_wait_readbacks() returns when every copy queued so far has landed.
 * Replaces: picture::get_SampleXY over a whole picture (the reference's output path, h264/picture.cpp:53-76, 396-409). */
typedef struct h264r_frame_layout_t {
    size_t frame_bytes;                 /* bytes of one frame slot                                    */
    size_t y_offset, u_offset, v_offset;/* byte offset of sample (0,0) of each plane inside the slot  */
    int32_t y_pitch, c_pitch;           /* bytes between rows                                         */
    int32_t width, height;              /* luma samples (coded size: multiples of 16)                 */
} h264r_frame_layout_t;
int h264r_frame_layout(h264r_ctx* ctx, h264r_frame_layout_t* out);
int h264r_read_frame_async(h264r_ctx* ctx, int frame_id, void* dst);
int h264r_wait_readbacks(h264r_ctx* ctx);

/* 16-byte position-dependent checksum of the three planes, computed on the device (kernel K5);
 * see DESIGN.md for the definition.  Used to compare GOP-sharded runs across GPUs without
 * moving pictures. */
int h264r_frame_digest(h264r_ctx* ctx, int frame_id, uint8_t out[16]);
/* Digests of many frames with one device->host copy. */
int h264r_frame_digests(h264r_ctx* ctx, const int* frame_ids, int n, uint8_t* out /* n*16 */);

/* Counter of luma samples that left [0,255] before the store since the last call (reset on read).
 * Under H264R_REF_COMPAT such samples mean the content relies on the reference's unclipped int
 * storage (D3), which 8-bit planes cannot represent: parity tests assert it is zero. */
int h264r_range_excursions(h264r_ctx* ctx, uint64_t* count);

/* Findings of the device-side checks of caller-filled picture buffers since the last call (flushes, waits, resets):
 * what h264r_submit_picture() accepts without a pass over the headers when the parser supplies n_intra. */
#define H264R_STATUS_BAD_CLASS   0x1u   /* a header with an unknown mb_class                                          */
#define H264R_STATUS_INTER_IN_I  0x2u   /* an inter macroblock in an I picture                                        */
#define H264R_STATUS_N_INTRA     0x4u   /* more intra macroblocks than n_intra said: some were not reconstructed       */
#define H264R_STATUS_STREAM      0x8u   /* stream form: mb_words[] does not add up to the submitted word count, or a
                                           macroblock's words are not what its header and mask announce (the picture is
                                           reconstructed from what was read; every access stays inside the picture's slot)   */
int h264r_device_status(h264r_ctx* ctx, uint32_t* flags);

/* Per-macroblock map of the same event for one frame: out[mb_addr] = 1 where a luma sample of that
 * macroblock left [0,255].  Content generators use it to keep synthetic pictures in range. */
int h264r_excursion_map(h264r_ctx* ctx, int frame_id, uint8_t* out /* w_mbs*h_mbs */);

/* Kernel K1 on its own: dequantisation + inverse transforms of n_mbs macroblocks (any n_mbs, no
 * picture structure needed).  out receives 384 int16 per macroblock: residual_Y 16x16 row-major,
 * then residual_U 8x8, then residual_V 8x8, saturated to int16.  Host pointers; synchronous.
 * Replaces: residual::Decode (h264/residual.cpp:184-201). */
int h264r_residual_mbs(h264r_ctx* ctx, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* coeff, int16_t* out);

/* Frame slot may be reused.  Replaces: Decoder::ctrl_Memory / clear_DecodedPic
 * (h264/Decoder.cpp:38-82).  Safe to call while pictures that read the slot are queued or in flight: a picture that
 * later overwrites the slot is ordered behind every queued reader (a later wave of the flush; across kernel pipelines an
 * event).  Only a reader that is still OPEN (begun, not yet ended) pins the slot: h264r_frame_begin() on it fails with
 * H264R_E_STATE until that picture is ended or aborted. */
int h264r_release(h264r_ctx* ctx, int frame_id);

/* Device-time of the kernels issued by the last h264r_flush(), in milliseconds (CUDA events on the
 * context's stream), and the number of kernel launches it made. */
int h264r_last_flush_stats(h264r_ctx* ctx, float* kernel_ms, int* launches);

/* Measurement support.  h264r_set_profiling(1) makes h264r_flush() bracket every kernel launch with
 * CUDA events on the context's stream; h264r_last_flush_kernel_ms() then returns, per kernel family
 * (0 = inter K1+K3, 1 = intra K1+K2, 2 = deblock K4, 3 = border replication, 4 = digest K5), the summed device time
 * and the number of launches of the last flush.  h264r_mark()/h264r_elapsed_ms() record and compare user events
 * (slot 0..3) on the same stream, e.g. around a submit..flush..digest sequence. */
int h264r_set_profiling(h264r_ctx* ctx, int on);
#define H264R_KERNEL_FAMILIES 5
int h264r_last_flush_kernel_ms(h264r_ctx* ctx, float ms[H264R_KERNEL_FAMILIES], int launches[H264R_KERNEL_FAMILIES]);
/* Per-launch device times of the last flush in issue order (needs h264r_set_profiling(1)): fills ms[] / family[]
 * with up to cap entries and returns the number written (>= 0) or a negative error code. */
int h264r_last_flush_launch_ms(h264r_ctx* ctx, float* ms, int* family, int cap);
int h264r_mark(h264r_ctx* ctx, int slot);
int h264r_elapsed_ms(h264r_ctx* ctx, int slot_from, int slot_to, float* ms);

const char* h264r_last_error(h264r_ctx* ctx);
int h264r_abi_version(void);

#ifdef H264R_TRACE
/* Development builds only (-DH264R_TRACE): clock stamps of the row kernels, see tools/trace_rows.py. */
int h264r_debug_trace(long long* out, int cap);
#endif

#ifdef __cplusplus
}
#endif
#endif /* H264R_H_ */
