/*
 * TEST INFRASTRUCTURE -- CPU restatement of the reconstruction path ("the oracle").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libh264r.so) never links or calls it.
 *
 * Pinning status:
 *   - with H264R_REF_COMPAT every function here is pinned against the compiled reference
 *     (oracle/_ref/h264ref_harness) by tests/test_oracle_vs_reference.py and by the fixtures under
 *     tests/golden/ that were produced by running that binary;
 *   - the spec-mode pieces the reference does not implement (Clip1, 2x2 chroma Hadamard, chroma
 *     intra prediction and MC, standard qpel taps, 8x8 transform, Intra8x8, deblocking) restate
 *     ITU-T H.264 (03/2014) clauses 8.3.2, 8.3.4, 8.4.2.2, 8.5.11-13, 8.7 and are
 *     PARITY UNPINNED by the reference (it has no such code); see DESIGN.md.
 */
#ifndef H264_RESTATE_H_
#define H264_RESTATE_H_
#include <stdint.h>
#include "../../include/h264r.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h264o_pic {
    int32_t  w_mbs, h_mbs;
    uint32_t flags;                 /* H264R_REF_COMPAT or 0 */
    uint8_t* y; uint8_t* u; uint8_t* v;   /* current planes, pitch w_mbs*16 / w_mbs*8 */
    int32_t* raw_y;                 /* optional: unclipped int luma, as the reference stores it (REF_COMPAT; may be NULL) */
    const h264r_mb_hdr* hdr;        /* [n_mbs] */
    const int16_t* mv;              /* [n_mbs][32] or NULL */
    const int16_t* coeff;           /* [n_mbs][384] */
    const h264r_pic_params* pp;
    const uint8_t* ref_y[H264R_MAX_REFS];  /* planes of ref_list0[i] */
    const uint8_t* ref_u[H264R_MAX_REFS];
    const uint8_t* ref_v[H264R_MAX_REFS];
    int32_t  ref_id[H264R_MAX_REFS];       /* identity of ref_list0[i] (deblocking compares pictures, not indices) */
    uint64_t excursions;            /* luma samples outside [0,255] before the store */
    uint8_t* excursion_map;         /* optional [n_mbs]: 1 where the MB had an excursion */
} h264o_pic;

/* residual::Decode for one MB (h264/residual.cpp:184-201): out_y[256], out_u[64], out_v[64] (row-major). */
void h264o_residual_mb(const h264o_pic* p, int mb_addr, int32_t* out_y, int32_t* out_u, int32_t* out_v);
/* macroblock::Calc for one MB (h264/macroblock.cpp:18-47): prediction + reconstruction into p->y/u/v. */
void h264o_recon_mb(h264o_pic* p, int mb_addr);
/* spec 8.7 over the whole picture (no reference counterpart). */
void h264o_deblock_picture(h264o_pic* p);
/* all MBs in raster order, then deblocking when enabled (spec mode, idc != 1). */
void h264o_recon_picture(h264o_pic* p);
/* K5 checksum, same definition as the device kernel. */
void h264o_digest(const uint8_t* y, const uint8_t* u, const uint8_t* v, int w, int h, uint8_t out[16]);

#ifdef __cplusplus
}
#endif
#endif
