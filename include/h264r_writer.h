/* h264r_writer.h -- host-side helper for C / C++ parsers: fills an h264r_picture_buf macroblock by macroblock in the
 * transport formats of h264r.h (block masks, partition-order vectors growing backward from mv_end, present blocks as
 * H264R_LEVELS_COMPACT units or int16 blocks), so that a parser that has a macroblock's syntax the way the reference
 * has it -- one final vector per 4x4 block, 24 x 16 levels, raster inside the block (see "Coefficients" in h264r.h) --
 * gets the one-DMA transport without knowing its byte layout.  Header-only, plain C99, no dependency on the library.
 *
 *   h264r_writer w;
 *   h264r_writer_begin(&w, &buf, n_mbs, H264R_LEVELS_COMPACT);
 *   for every macroblock in address order:
 *       rc = h264r_writer_mb(&w, &hdr, mv, coeff);      (mv: 32 int16 or NULL for intra; coeff: 384 int16)
 *       rc == H264R_E_INVAL on a level outside int8 in a COMPACT picture: begin again with H264R_LEVELS_INT16
 *   h264r_writer_end(&w, &n_blocks, &n_mvs);
 *   h264r_submit_picture(ctx, frame_id, &buf, n_blocks, n_mvs);
 *
 * Replaces, on the reference's side, nothing: it is what Slice::PraseSliceDataer (h264/slice.cpp:357-367) would call per
 * macroblock instead of macroblock::Calc once macroblock::Parse has filled the macroblock; the Python
 * twin is videodecoder_b200/pack.py (pack_blocks, pack_compact, pack_mvs_partition) and tests/test_adapter.py compares
 * the two byte for byte.
 */
#ifndef H264R_WRITER_H
#define H264R_WRITER_H

#include <string.h>
#include "h264r.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct h264r_writer {
    h264r_picture_buf* buf;
    int    n_mbs, next_mb;
    int    level_fmt;        /* H264R_LEVELS_COMPACT or H264R_LEVELS_INT16 */
    size_t n_units;          /* 8-byte units (COMPACT) or 32-byte blocks (INT16) written so far */
    long   n_mvs;
    int    n_intra;
} h264r_writer;

static inline void h264r_writer_begin(h264r_writer* w, h264r_picture_buf* buf, int n_mbs, int level_fmt)
{
    w->buf = buf; w->n_mbs = n_mbs; w->next_mb = 0; w->level_fmt = level_fmt;
    w->n_units = 0; w->n_mvs = 0; w->n_intra = 0;
}

/* the 8-byte form of a header (h264r_mb_hdr8 in h264r.h) */
static inline h264r_mb_hdr8 h264r_hdr8_pack(const h264r_mb_hdr* h)
{
    h264r_mb_hdr8 s;
    s.w0 = (uint32_t)(h->mb_class & 7u) | ((uint32_t)(h->flags & 63u) << 3) | ((uint32_t)(h->intra_modes & 15u) << 9) |
           ((uint32_t)(h->qp_y & 63u) << 13) | ((uint32_t)(h->qp_cb & 63u) << 19) | ((uint32_t)(h->qp_cr & 63u) << 25);
    s.w1 = (uint32_t)(h->cbp & 63u) | ((uint32_t)h->sub_mb_type << 6);
    if (h->mb_class > H264R_MB_I16x16)
        s.w1 |= ((uint32_t)(h->u.ref_idx[0] & 15) << 14) | ((uint32_t)(h->u.ref_idx[1] & 15) << 18) |
                ((uint32_t)(h->u.ref_idx[2] & 15) << 22) | ((uint32_t)(h->u.ref_idx[3] & 15) << 26);
    return s;
}

/* signed 13-bit range of the inline vector of the stream form */
static inline int h264r_mv_fits_inline(int x, int y) { return x >= -4096 && x <= 4095 && y >= -4096 && y <= 4095; }

static inline int h264r_writer_popcount(unsigned v)
{
#if defined(__GNUC__)
    return __builtin_popcount(v);
#else
    int n = 0;
    for (; v; v &= v - 1) n++;
    return n;
#endif
}
/* one 8-byte unit of H264R_LEVELS_COMPACT: bytes 0-1 the significance map, bytes 2-7 the (at most six) levels of the set
 * positions in ascending order; returns their number */
static inline int h264r_writer_compact_unit(const int16_t* c, unsigned map, uint8_t u[8])
{
    int n = 0;
    memset(u, 0, 8);
    u[0] = (uint8_t)map; u[1] = (uint8_t)(map >> 8);
    for (; map; map &= map - 1) {
#if defined(__GNUC__)
        const unsigned k = (unsigned)__builtin_ctz(map);
#else
        unsigned k = 0, t = map & (0u - map);
        while (t >>= 1) k++;
#endif
        u[2 + n++] = (uint8_t)(int8_t)c[k];
    }
    return n;
}
/* number of vectors the bitstream carries for a macroblock (mvd loops of macroblock::Parse, h264/macroblock.cpp:360-398) and
 * the first 4x4 block (raster) of each of them */
static inline int h264r_writer_partitions(const h264r_mb_hdr* h, int first_block[16])
{
    static const int sub_count[4] = {1, 2, 2, 4};                                   /* H264R_SUB_8x8, 8x4, 4x8, 4x4 */
    static const int sub_first[4][4] = {{0, 0, 0, 0}, {0, 4, 0, 0}, {0, 1, 0, 0}, {0, 1, 4, 5}};
    int n = 0, q, k;
    switch (h->mb_class) {
    case H264R_MB_P16x16: first_block[n++] = 0; break;
    case H264R_MB_P16x8:  first_block[n++] = 0; first_block[n++] = 8; break;
    case H264R_MB_P8x16:  first_block[n++] = 0; first_block[n++] = 2; break;
    case H264R_MB_P8x8:
        for (q = 0; q < 4; q++) {
            const int s = (h->sub_mb_type >> (2 * q)) & 3, b0 = (q >> 1) * 8 + (q & 1) * 2;
            for (k = 0; k < sub_count[s]; k++) first_block[n++] = b0 + sub_first[s][k];
        }
        break;
    default: break;                                                                 /* intra: none */
    }
    return n;
}

/* One macroblock: hdr as it goes on the wire; mv = the final vector of each 4x4 block (32 int16, raster) or NULL for an intra
 * macroblock; coeff = 24 blocks x 16 levels as described under "Coefficients" in h264r.h.  Returns H264R_OK, H264R_E_STATE
 * (more macroblocks than begun with), H264R_E_CAPACITY (the buffer's block area is full) or H264R_E_INVAL (a level outside
 * [-128, 127] in a COMPACT picture). */
static inline int h264r_writer_mb(h264r_writer* w, const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff)
{
    h264r_picture_buf* b = w->buf;
    uint32_t mask = 0;
    int blk, k, n_present = 0, compact = 1, part[16], n_part;
    uint16_t map[24];
    if (w->next_mb >= w->n_mbs) return H264R_E_STATE;
    for (blk = 0; blk < 24; blk++) {
        /* most blocks are empty: four 64-bit words say so; of the others, the significance map (bit k = position k holds a level) */
        uint64_t q[4];
        unsigned m16 = 0, wide = 0;
        memcpy(q, coeff + blk * 16, 32);
        map[blk] = 0;
        if (!(q[0] | q[1] | q[2] | q[3])) continue;
        for (k = 0; k < 16; k++) {
            const int v = coeff[blk * 16 + k];
            m16 |= (unsigned)(v != 0) << k;
            wide |= (unsigned)(v + 128) > 255u;
        }
        if (wide && w->level_fmt == H264R_LEVELS_COMPACT) return H264R_E_INVAL;
        map[blk] = (uint16_t)m16;
        mask |= 1u << blk; n_present++;
        if (h264r_writer_popcount(m16) > 6) compact = 0;
    }
    if ((b->hdr_bytes == 8 || b->hdr_bytes == H264R_HDR_STREAM) && w->level_fmt != H264R_LEVELS_COMPACT) return H264R_E_INVAL;
    if (b->hdr_bytes == H264R_HDR_STREAM) {
        /* stream form: everything of the macroblock behind its 8-byte header and its word count, see h264r.h */
        uint32_t* s = b->stream + w->n_units;              /* n_units counts stream WORDS here */
        uint32_t words[2 + 1 + 16 + 24 * 4 + 1];
        int nw = 0, inline_mv = 0;
        const int intra = hdr->mb_class <= H264R_MB_I16x16;
        h264r_mb_hdr8 h8 = h264r_hdr8_pack(hdr);
        if (compact && n_present) mask |= H264R_BLK_COMPACT;
        if (hdr->cbp || hdr->mb_class == H264R_MB_I16x16) words[nw++] = mask;      /* (Intra16x16 DC levels come without a pattern bit) */
        else if (mask & 0xFFFFFFu) return H264R_E_INVAL;   /* levels without coded_block_pattern */
        if (intra) { memcpy(&words[nw], hdr->u.pred4x4, 8); nw += 2; }
        n_part = (mv && !intra) ? h264r_writer_partitions(hdr, part) : 0;
        if (hdr->mb_class == H264R_MB_P16x16 && n_part == 1 && h264r_mv_fits_inline(mv[0], mv[1])) {
            const uint32_t x = (uint32_t)mv[0] & 0x1FFFu, y = (uint32_t)mv[1] & 0x1FFFu;
            h8.w0 |= H264R_HDR8_MV_INLINE | (((x >> 8) & 15u) << 9);
            h8.w1 = (h8.w1 & 0x3C03Fu) | ((x & 255u) << 6) | (((x >> 12) & 1u) << 30) | ((y & 0xFFFu) << 18) | (((y >> 12) & 1u) << 31);
            inline_mv = 1;
        }
        {
            /* behind the words: a byte run -- the vectors as int8 pairs when all of them fit (H264R_HDR8_MV_SHORT; else as words),
             * then per block of a compact macroblock the 16-bit significance map and one byte per level -- zero-padded to the next
             * word (also in front of the int8 blocks of any other macroblock) */
            uint8_t* q = (uint8_t*)&words[nw];
            int nb = 0;
            if (!inline_mv && n_part) {
                int fits = 1;
                for (k = 0; k < n_part; k++) fits &= mv[part[k] * 2] >= -128 && mv[part[k] * 2] <= 127 && mv[part[k] * 2 + 1] >= -128 && mv[part[k] * 2 + 1] <= 127;
                if (fits) {
                    h8.w1 |= H264R_HDR8_MV_SHORT;
                    for (k = 0; k < n_part; k++) { q[nb++] = (uint8_t)(int8_t)mv[part[k] * 2]; q[nb++] = (uint8_t)(int8_t)mv[part[k] * 2 + 1]; }
                } else
                    for (k = 0; k < n_part; k++) {
                        const uint32_t v = (uint32_t)(uint16_t)mv[part[k] * 2] | ((uint32_t)(uint16_t)mv[part[k] * 2 + 1] << 16);
                        memcpy(q + nb, &v, 4);
                        nb += 4;
                    }
            }
            if (compact && n_present) {
                for (blk = 0; blk < 24; blk++) {
                    uint8_t u[8];
                    int n;
                    if (!((mask >> blk) & 1u)) continue;
                    n = h264r_writer_compact_unit(coeff + blk * 16, map[blk], u);
                    memcpy(q + nb, u, (size_t)(2 + n));
                    nb += 2 + n;
                }
            } else {
                while (nb & 3) q[nb++] = 0;
                for (blk = 0; blk < 24; blk++) {
                    const int16_t* c = coeff + blk * 16;
                    if (!((mask >> blk) & 1u)) continue;
                    for (k = 0; k < 16; k++) q[nb++] = (uint8_t)(int8_t)c[k];
                }
            }
            while (nb & 3) q[nb++] = 0;
            nw += nb >> 2;
        }
        if (nw > 255) return H264R_E_INVAL;
        if (w->n_units + (size_t)nw > b->max_stream_words) return H264R_E_CAPACITY;
        memcpy(s, words, (size_t)nw * 4);
        w->n_units += (size_t)nw;
        b->hdr8[w->next_mb] = h8;
        b->mb_words[w->next_mb] = (uint8_t)nw;
        w->n_mvs += n_part;
        (void)inline_mv;
        if (intra) w->n_intra++;
        w->next_mb++;
        return H264R_OK;
    }
    if (w->level_fmt == H264R_LEVELS_COMPACT) {
        uint8_t* units = (uint8_t*)b->blocks;
        const int ext = b->hdr_bytes == 8 && hdr->mb_class <= H264R_MB_I16x16;      /* prediction modes in front of the blocks */
        const size_t need = (size_t)n_present * (compact ? 1 : 2) + (size_t)ext;
        if (w->n_units + need > b->max_blocks * 4) return H264R_E_CAPACITY;
        if (ext) { memcpy(units + w->n_units * 8, hdr->u.pred4x4, 8); w->n_units += 1; }
        for (blk = 0; blk < 24; blk++) {
            uint8_t* u = units + w->n_units * 8;
            const int16_t* c = coeff + blk * 16;
            if (!((mask >> blk) & 1u)) continue;
            if (compact) {                     /* significance map, then the levels in position order */
                h264r_writer_compact_unit(c, map[blk], u);
                w->n_units += 1;
            } else {                           /* int8 levels, raster: two units */
                for (k = 0; k < 16; k++) u[k] = (uint8_t)(int8_t)c[k];
                w->n_units += 2;
            }
        }
        if (compact && n_present) mask |= H264R_BLK_COMPACT;
    } else {
        if (w->n_units + (size_t)n_present > b->max_blocks) return H264R_E_CAPACITY;
        for (blk = 0; blk < 24; blk++)
            if ((mask >> blk) & 1u) { memcpy(b->blocks + w->n_units * 16, coeff + blk * 16, 32); w->n_units++; }
    }
    if (b->hdr_bytes == 8) b->hdr8[w->next_mb] = h264r_hdr8_pack(hdr); else b->hdr[w->next_mb] = *hdr;
    b->blk_mask[w->next_mb] = mask;
    n_part = mv ? h264r_writer_partitions(hdr, part) : 0;
    for (k = 0; k < n_part; k++) {             /* vector i of the picture at mv_end[-2(i+1)], mv_end[-2(i+1)+1] */
        b->mv_end[-2 * (w->n_mvs + 1)] = mv[part[k] * 2];
        b->mv_end[-2 * (w->n_mvs + 1) + 1] = mv[part[k] * 2 + 1];
        w->n_mvs++;
    }
    if (hdr->mb_class <= H264R_MB_I16x16) w->n_intra++;
    w->next_mb++;
    return H264R_OK;
}

/* Finishes the picture: sets buf->n_intra and buf->level_bytes and returns what h264r_submit_picture wants. */
static inline int h264r_writer_end(h264r_writer* w, size_t* n_blocks, long* n_mvs)
{
    if (w->next_mb != w->n_mbs) return H264R_E_STATE;
    w->buf->n_intra = w->n_intra;
    w->buf->level_bytes = w->level_fmt;
    *n_blocks = w->n_units;
    *n_mvs = w->n_mvs;
    return H264R_OK;
}

#ifdef __cplusplus
}
#endif
#endif
