// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
//
// The binding of INTEGRATION.md, compiled against the reference's own classes: the reference decoder parses a stream
// (NAL -> slice -> macroblock::Parse, untouched sources, queue-fed cabac shim), and after every picture this program
// packs what the reference's HOST side produced -- macroblock types, prediction modes, final motion vectors, QPs and
// the coefficient levels still sitting in the residual block trees -- into the buffers of include/h264r.h, hands them
// to the engine through the C ABI and compares the planes that come back with the reference's own CPU reconstruction
// of the same picture (picture::get_SampleXY, h264/picture.cpp:396-409).
//
//   h264ref_adapter <stream.264> <queue.bin>                      engine vs reference, one JSON line, exit 0 iff bit-exact
//   h264ref_adapter <stream.264> <queue.bin> --pack-only out.bin  no GPU: writes per picture hdr[n][16], mv[n][32] int16,
//                                                                 coeff[n][384] int16 (CPU test compares them with pack.py)
//   ... --transport buf                                           the one-DMA transport: include/h264r_writer.h fills a
//                                                                 picture buffer (compact levels, partition-order vectors),
//                                                                 h264r_submit_picture ships it; with --pack-only the buffer's
//                                                                 used parts are written instead (hdr, masks, units, vectors)
//
// pack_macroblock() below is the stub a maintainer would add to Slice::PraseSliceDataer (h264/slice.cpp:357-367) in place
// of macroblock::Calc; here it runs after the reference has also reconstructed the picture, so that both results exist.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "Parser.h"
#include "Decoder.h"
#include "Debug.h"
#include "NAL.h"
#include "picture.h"
#include "macroblock.h"
#include "residual.h"
#include "block.h"
#include "matrix.h"
#include "slice.h"

#include "../../include/h264r.h"
#include "../../include/h264r_writer.h"

extern std::vector<int> g_h264o_queue;
extern size_t g_h264o_queue_pos;

// scan index -> raster position inside a 4x4 block (what Inverse4x4Scan does, h264/residual.cpp:136-141)
static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
// luma4x4BlkIdx <-> raster position of the block inside the macroblock (h264/staticcharts.h:1378-1383; an involution)
static const uint8_t kBlkRaster[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};

static void pack_macroblock(macroblock* mb, int mbx, int mby, int w_mbs, h264r_mb_hdr& h, int16_t* mv, int16_t* c)
{
    memset(&h, 0, sizeof h);
    memset(mv, 0, H264R_MV_PER_MB * sizeof(int16_t));
    memset(c, 0, H264R_COEFF_PER_MB * sizeof(int16_t));
    // one slice per picture (the reference starts a new picture per slice NAL): availability is geometry
    // (macroblock::is_avaiable, h264/macroblock.cpp:1446-1457)
    if (mbx > 0) h.flags |= H264R_AVAIL_A;
    if (mby > 0) h.flags |= H264R_AVAIL_B;
    if (mby > 0 && mbx < w_mbs - 1) h.flags |= H264R_AVAIL_C;
    if (mbx > 0 && mby > 0) h.flags |= H264R_AVAIL_D;
    h.qp_y = (uint8_t)mb->QPY_;                               // h264/macroblock.cpp:485
    h.qp_cb = h.qp_cr = (uint8_t)mb->QPC_;                    // h264/macroblock.cpp:487-488 (D2 stays on the host side)
    const bool skip = mb->type == P_Skip;
    const int cbp_l = skip ? 0 : mb->CodedBlockPatternLuma, cbp_c = skip ? 0 : mb->CodedBlockPatternChroma;
    h.cbp = (uint8_t)(cbp_l | (cbp_c << 4));
    residual* re = mb->re;
    if (mb->type == I_NxN) {
        h.mb_class = H264R_MB_I4x4;
        h.intra_modes = (uint8_t)((mb->intra_chroma_pred_mode & 3) << 2);
        for (int b = 0; b < 16; b++) h.u.pred4x4[b >> 1] |= (uint8_t)((mb->Intra4x4PredMode[b] & 15) << ((b & 1) * 4));   // h264/macroblock.cpp:992-1035
    } else if (mb->type >= I_16x16_0_0_0 && mb->type <= I_16x16_3_2_1) {
        h.mb_class = H264R_MB_I16x16;
        h.intra_modes = (uint8_t)((((int)mb->type - 1) & 3) | ((mb->intra_chroma_pred_mode & 3) << 2));                // staticcharts.h:1194-1226, column 2
    } else {
        switch (mb->type) {
        case P_L0_16x16: case P_Skip: h.mb_class = H264R_MB_P16x16; break;
        case P_L0_L0_16x8: h.mb_class = H264R_MB_P16x8; break;
        case P_L0_L0_8x16: h.mb_class = H264R_MB_P8x16; break;
        default: h.mb_class = H264R_MB_P8x8; break;            // P_8x8, P_8x8ref0
        }
        if (skip) h.flags |= H264R_SKIP;
        int sub[4] = {0, 0, 0, 0};
        if (h.mb_class == H264R_MB_P8x8)
            for (int q = 0; q < 4; q++) { sub[q] = (int)mb->sub_mb_type[q] & 3; h.sub_mb_type |= (uint8_t)(sub[q] << (2 * q)); }   // P_L0_8x8.. = H264R_SUB_8x8..
        // one final vector per 4x4 block (raster) from mv_l0[partition][sub-partition] (h264/macroblock.cpp:391-398),
        // one reference index per 8x8 quadrant
        for (int ras = 0; ras < 16; ras++) {
            const int bx = ras & 3, by = ras >> 2, q = (by >> 1) * 2 + (bx >> 1);
            int part = 0, s = 0;
            if (h.mb_class == H264R_MB_P16x8) part = by >> 1;
            else if (h.mb_class == H264R_MB_P8x16) part = bx >> 1;
            else if (h.mb_class == H264R_MB_P8x8) {
                part = q;
                s = sub[q] == H264R_SUB_8x8 ? 0 : (sub[q] == H264R_SUB_8x4 ? (by & 1) : (sub[q] == H264R_SUB_4x8 ? (bx & 1) : (by & 1) * 2 + (bx & 1)));
            }
            mv[ras * 2 + 0] = (int16_t)mb->mv_l0[part][s][0];
            mv[ras * 2 + 1] = (int16_t)mb->mv_l0[part][s][1];
            h.u.ref_idx[q] = (int8_t)(mb->ref_idx_l0 ? mb->ref_idx_l0[part] : 0);
        }
    }
    // coefficient levels: scan order in the reference's block trees (h264/residual.cpp:36-127) -> raster inside the block,
    // blocks in luma4x4BlkIdx order; Intra16x16 / chroma DC levels into slot 0 of the block they belong to
    if (h.mb_class == H264R_MB_I16x16) {
        for (int k = 0; k < 16; k++) c[kBlkRaster[kZigzag[k]] * 16] = (int16_t)re->luma[1]->value[k];
        for (int b = 0; b < 16; b++)
            if (cbp_l & (1 << (b >> 2)))
                for (int k = 0; k < 15; k++) c[b * 16 + kZigzag[k + 1]] = (int16_t)re->luma[0]->get_childBlock(b >> 2)->get_childBlock(b & 3)->value[k];
    } else {
        for (int b = 0; b < 16; b++)
            if (cbp_l & (1 << (b >> 2)))
                for (int k = 0; k < 16; k++) c[b * 16 + kZigzag[k]] = (int16_t)re->luma[0]->get_childBlock(b >> 2)->get_childBlock(b & 3)->value[k];
    }
    for (int p = 0; p < 2; p++)
        for (int b = 0; b < 4; b++) {
            if (cbp_c & 3) c[(16 + p * 4 + b) * 16] = (int16_t)re->chroma[0]->get_childBlock(p)->value[b];
            if (cbp_c & 2)
                for (int k = 0; k < 15; k++) c[(16 + p * 4 + b) * 16 + kZigzag[k + 1]] = (int16_t)re->chroma[1]->get_childBlock(p)->get_childBlock(b)->value[k];
        }
}

int main(int argc, char* argv[])
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s <stream.264> <queue.bin> [--pack-only out.bin]\n", argv[0]);
        return 2;
    }
    const bool pack_only = argc >= 5 && !strcmp(argv[3], "--pack-only");
    bool transport_buf = false;
    for (int i = 3; i + 1 < argc; i++) if (!strcmp(argv[i], "--transport") && !strcmp(argv[i + 1], "buf")) transport_buf = true;
    FILE* fp = fopen(argv[1], "r");
    FILE* fq = fopen(argv[2], "rb");
    if (!fp || !fq) { fprintf(stderr, "cannot open input\n"); return 2; }
    fseek(fq, 0, SEEK_END);
    long qbytes = ftell(fq);
    fseek(fq, 0, SEEK_SET);
    g_h264o_queue.resize(qbytes / 4);
    if (qbytes && fread(g_h264o_queue.data(), 4, qbytes / 4, fq) != (size_t)(qbytes / 4)) return 2;
    fclose(fq);
    FILE* out = pack_only ? fopen(argv[4], "wb") : NULL;
    if (pack_only && !out) return 2;

    Debug debug("");
    Parser parser(fp, &debug);
    Decoder decoder;
    h264r_ctx* ctx = NULL;
    const int max_frames = 64;
    std::map<int, int> frame_of_decnum;        // picture::DecNum (h264/picture.cpp:412-415) -> frame id handed to the engine
    int n_pic = 0;
    long bad_y = 0, bad_c = 0, excursions_total = 0;
    std::vector<h264r_mb_hdr> hdr;
    std::vector<int16_t> mv, coeff;
    std::vector<uint8_t> y, u, v;
    h264r_picture_buf pbuf;
    memset(&pbuf, 0, sizeof pbuf);
    while (parser.find_nextNAL()) {
        NAL nal(&parser, &decoder);
        nal.pic = NULL;
        nal.decode();
        if (!(nal.type == IDR || nal.type == Non_IDR)) continue;
        picture* pic = nal.pic;
        const int wm = parser.pV->PicWidthInMbs, hm = parser.pV->PicHeightInMbs, n = wm * hm, W = wm * 16, H = hm * 16;
        hdr.resize(n); mv.resize((size_t)n * H264R_MV_PER_MB); coeff.resize((size_t)n * H264R_COEFF_PER_MB);
        bool any_inter = false;
        for (int r = 0; r < hm; r++)
            for (int c = 0; c < wm; c++) {
                const int a = r * wm + c;
                pack_macroblock(pic->get_MBXY(r, c), c, r, wm, hdr[a], &mv[(size_t)a * H264R_MV_PER_MB], &coeff[(size_t)a * H264R_COEFF_PER_MB]);
                any_inter = any_inter || hdr[a].mb_class >= H264R_MB_P16x16;
            }
        if (pack_only && transport_buf) {
            // the writer on a plain host buffer laid out by hand (h264r_picture_buf_alloc needs a device)
            std::vector<h264r_mb_hdr> bh(n);
            std::vector<uint32_t> bm(n);
            std::vector<int16_t> bmv((size_t)n * H264R_MV_PER_MB), bb((size_t)n * H264R_COEFF_PER_MB);
            h264r_picture_buf pb;
            memset(&pb, 0, sizeof pb);
            pb.hdr = bh.data(); pb.blk_mask = bm.data(); pb.mv = bmv.data(); pb.mv_end = bmv.data() + bmv.size();
            pb.blocks = bb.data(); pb.max_blocks = (size_t)24 * n;
            h264r_writer w;
            size_t n_units = 0;
            long n_mvs = 0;
            for (int fmt = H264R_LEVELS_COMPACT;; fmt = H264R_LEVELS_INT16) {
                h264r_writer_begin(&w, &pb, n, fmt);
                int rc = H264R_OK;
                for (int a = 0; a < n && rc == H264R_OK; a++)
                    rc = h264r_writer_mb(&w, &hdr[a], hdr[a].mb_class >= H264R_MB_P16x16 ? &mv[(size_t)a * H264R_MV_PER_MB] : NULL, &coeff[(size_t)a * H264R_COEFF_PER_MB]);
                if (rc == H264R_E_INVAL && fmt == H264R_LEVELS_COMPACT) continue;
                if (rc != H264R_OK || h264r_writer_end(&w, &n_units, &n_mvs) != H264R_OK) return 4;
                break;
            }
            int32_t head[4] = {pb.level_bytes, (int32_t)n_units, (int32_t)n_mvs, pb.n_intra};
            fwrite(head, 4, 4, out);
            fwrite(bh.data(), sizeof(h264r_mb_hdr), n, out);
            fwrite(bm.data(), 4, n, out);
            fwrite(pb.blocks, pb.level_bytes == H264R_LEVELS_INT16 ? 32 : 8, n_units, out);
            fwrite(pb.mv_end - 2 * n_mvs, 4, n_mvs, out);
            n_pic++;
            continue;
        }
        if (pack_only) {
            fwrite(hdr.data(), sizeof(h264r_mb_hdr), n, out);
            fwrite(mv.data(), 2, mv.size(), out);
            fwrite(coeff.data(), 2, coeff.size(), out);
            n_pic++;
            continue;
        }
        if (!ctx && h264r_create(&ctx, 0, wm, hm, max_frames, max_frames, H264R_REF_COMPAT) != H264R_OK) {
            fprintf(stderr, "h264r_create: %s\n", h264r_last_error(NULL));
            return 3;
        }
        // NAL::decode_PIC after the slice header and the reference list (h264/NAL.cpp:204-224)
        h264r_pic_params pp;
        memset(&pp, 0, sizeof pp);
        const bool is_p = !pic->is_IDR() && any_inter;
        pp.slice_type = is_p ? H264R_SLICE_P : H264R_SLICE_I;
        pp.disable_deblocking_filter_idc = 1;
        if (is_p) {
            pp.num_ref_idx_l0 = (int)decoder.list_Ref0.size() < H264R_MAX_REFS ? (int)decoder.list_Ref0.size() : H264R_MAX_REFS;
            for (int i = 0; i < pp.num_ref_idx_l0; i++) pp.ref_list0[i] = decoder.list_Ref0[i] ? frame_of_decnum[decoder.list_Ref0[i]->DecNum] : 0;
        }
        const int fid = n_pic % max_frames;
        frame_of_decnum[pic->DecNum] = fid;
        y.resize((size_t)W * H); u.resize((size_t)W * H / 4); v.resize((size_t)W * H / 4);
        int rc = h264r_frame_begin(ctx, fid, &pp);
        if (transport_buf) {
            if (!pbuf.base && rc == H264R_OK) rc = h264r_picture_buf_alloc(ctx, 0, &pbuf);
            h264r_writer w;
            size_t n_units = 0;
            long n_mvs = 0;
            for (int fmt = H264R_LEVELS_COMPACT; rc == H264R_OK; fmt = H264R_LEVELS_INT16) {
                h264r_writer_begin(&w, &pbuf, n, fmt);
                for (int a = 0; a < n && rc == H264R_OK; a++)
                    rc = h264r_writer_mb(&w, &hdr[a], hdr[a].mb_class >= H264R_MB_P16x16 ? &mv[(size_t)a * H264R_MV_PER_MB] : NULL, &coeff[(size_t)a * H264R_COEFF_PER_MB]);
                if (rc == H264R_E_INVAL && fmt == H264R_LEVELS_COMPACT) { rc = H264R_OK; continue; }
                if (rc == H264R_OK) rc = h264r_writer_end(&w, &n_units, &n_mvs);
                break;
            }
            if (rc == H264R_OK) rc = h264r_submit_picture(ctx, fid, &pbuf, n_units, n_mvs);       // (read_planes below syncs: the buffer may be refilled)
        } else {
            if (rc == H264R_OK) rc = h264r_submit_mbs(ctx, fid, 0, n, hdr.data(), is_p ? mv.data() : NULL, coeff.data());
            if (rc == H264R_OK) rc = h264r_frame_end(ctx, fid);
        }
        if (rc == H264R_OK) rc = h264r_read_planes(ctx, fid, y.data(), W, u.data(), v.data(), W / 2);
        if (rc != H264R_OK) { fprintf(stderr, "engine: %s\n", h264r_last_error(ctx)); return 3; }
        // the reference's own reconstruction of the same picture
        for (int r = 0; r < H; r++)
            for (int c = 0; c < W; c++) bad_y += y[(size_t)r * W + c] != pic->get_SampleXY(c, r, 0);
        for (int r = 0; r < H / 2; r++)
            for (int c = 0; c < W / 2; c++) {
                macroblock* mb = pic->get_MBXY(r / 8, c / 8);
                bad_c += u[(size_t)r * (W / 2) + c] != (unsigned char)(*(mb->sample_U))[r % 8][c % 8];
                bad_c += v[(size_t)r * (W / 2) + c] != (unsigned char)(*(mb->sample_V))[r % 8][c % 8];
            }
        n_pic++;
    }
    if (out) fclose(out);
    fclose(fp);
    if (ctx) {
        uint64_t exc = 0;
        h264r_range_excursions(ctx, &exc);
        excursions_total = (long)exc;
        if (pbuf.base) h264r_picture_buf_free(ctx, &pbuf);
        h264r_destroy(ctx);
    }
    printf("{\"pictures\": %d, \"luma_mismatches\": %ld, \"chroma_mismatches\": %ld, \"luma_range_excursions\": %ld, \"pack_only\": %s}\n",
           n_pic, bad_y, bad_c, excursions_total, pack_only ? "true" : "false");
    return (bad_y || bad_c) ? 1 : 0;
}
