// TEST INFRASTRUCTURE (oracle build only) -- not part of the product.
//
// Driver for the compiled reference decoder.  The reference's own main (h264/main.cpp:20-45)
// only prints ASCII art; this harness keeps its NAL loop shape (Parser::find_nextNAL ->
// NAL::decode, h264/Parser.cpp:21, h264/NAL.cpp:14) and, after every picture NAL, walks the
// public members of the decoded picture to dump what the parity tests need:
//   * the Y/U/V planes as the reference defines them: low 8 bits of each MB's int samples
//     (picture::get_SampleXY, h264/picture.cpp:396-409; chroma read straight from
//     macroblock::sample_U/V because get_SampleXY's /16 addressing is luma-only),
//   * the raw int luma samples (to detect values outside [0,255]),
//   * per-MB residual_Y/U/V (residual::Decode output, h264/residual.cpp:184-201),
//   * per-MB syntax as derived by the reference host side (type, QP, intra modes, MVs).
// The dump happens immediately because the picture may be freed at the next decode_PIC
// (Decoder::ctrl_Memory, h264/Decoder.cpp:67-82).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <time.h>

#include "Parser.h"
#include "Decoder.h"
#include "Debug.h"
#include "NAL.h"
#include "picture.h"
#include "macroblock.h"
#include "residual.h"
#include "matrix.h"
#include "slice.h"

#ifdef H264O_REAL_CABAC
// h264ref_cabac: the reference's own arithmetic decoder (h264/cabac.cpp) is linked instead of the queue-fed stand-in, so the
// whole reference -- bit reader, CABAC, macroblock layer, reconstruction -- runs unmodified on a genuine bitstream.
std::vector<int> g_h264o_queue;
size_t g_h264o_queue_pos = 0;
#else
extern std::vector<int> g_h264o_queue;
extern size_t g_h264o_queue_pos;
#endif

static double now_s()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void put_i32(FILE* f, int v) { fwrite(&v, 4, 1, f); }

static void dump_picture(FILE* out, NAL& nal, Parser& parser, Decoder& decoder, int pic_index, int level)
{
    picture* pic = nal.pic;
    int wm = parser.pV->PicWidthInMbs, hm = parser.pV->PicHeightInMbs;
    int W = wm * 16, H = hm * 16;
    put_i32(out, 0x44363248);  // "H26D"
    put_i32(out, pic_index);
    put_i32(out, wm);
    put_i32(out, hm);
    put_i32(out, (int)nal.type);
    put_i32(out, pic->POC);
    put_i32(out, level);
    // reference picture list 0 as left by Decoder::init_RefPicList / opra_RefModfication
    // (h264/Decoder.cpp:133-338): decode indices (picture::DecNum, h264/picture.cpp:412-415)
    put_i32(out, pic->DecNum);
    int nref = nal.pic && !pic->is_IDR() ? (int)decoder.list_Ref0.size() : 0;
    put_i32(out, nref);
    for (int i = 0; i < 16; i++) put_i32(out, i < nref && decoder.list_Ref0[i] ? decoder.list_Ref0[i]->DecNum : -1);
    std::vector<unsigned char> y(W * H), u(W * H / 4), v(W * H / 4);
    std::vector<int> raw(W * H);
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++) {
            y[r * W + c] = pic->get_SampleXY(c, r, 0);
            raw[r * W + c] = (*(pic->get_MBXY(r / 16, c / 16)->sample_Y))[r % 16][c % 16];
        }
    for (int r = 0; r < H / 2; r++)
        for (int c = 0; c < W / 2; c++) {
            macroblock* mb = pic->get_MBXY(r / 8, c / 8);
            u[r * (W / 2) + c] = (unsigned char)(*(mb->sample_U))[r % 8][c % 8];
            v[r * (W / 2) + c] = (unsigned char)(*(mb->sample_V))[r % 8][c % 8];
        }
    fwrite(y.data(), 1, y.size(), out);
    fwrite(u.data(), 1, u.size(), out);
    fwrite(v.data(), 1, v.size(), out);
    if (level >= 1) fwrite(raw.data(), 4, raw.size(), out);
    if (level >= 2) {
        for (int r = 0; r < hm; r++)
            for (int c = 0; c < wm; c++) {
                macroblock* mb = pic->get_MBXY(r, c);
                int rec[64];
                memset(rec, 0, sizeof(rec));
                rec[0] = (int)mb->type;
                rec[1] = (int)mb->premode;
                rec[2] = mb->QPY_;
                rec[3] = mb->QPC_;
                rec[4] = mb->CodedBlockPatternLuma;
                rec[5] = mb->CodedBlockPatternChroma;
                rec[6] = mb->intra_chroma_pred_mode;
                rec[7] = mb->transform_size_8x8_flag;
                rec[8] = mb->num_MBpart;
                rec[9] = mb->mb_skip_flag;
                for (int i = 0; i < 16; i++) rec[10 + i] = mb->Intra4x4PredMode ? mb->Intra4x4PredMode[i] : -1;
                for (int i = 0; i < 4; i++) rec[26 + i] = mb->sub_mb_type ? (int)mb->sub_mb_type[i] : -1;
                if (mb->is_interpred() && mb->mv_l0) {
                    int np = mb->num_MBpart;
                    for (int p = 0; p < np && p < 4; p++) {
                        rec[30 + p] = mb->ref_idx_l0 ? mb->ref_idx_l0[p] : 0;
                    }
                }
                fwrite(rec, 4, 34, out);
                int mv[32];
                for (int i = 0; i < 32; i++) mv[i] = 0x7fffffff;
                if (mb->is_interpred() && mb->mv_l0) {
                    int np = mb->num_MBpart;
                    for (int p = 0; p < np && p < 4; p++) {
                        int ns = mb->mv_l0[p].x_length;
                        for (int s = 0; s < ns && s < 4; s++) {
                            mv[(p * 4 + s) * 2 + 0] = mb->mv_l0[p][s][0];
                            mv[(p * 4 + s) * 2 + 1] = mb->mv_l0[p][s][1];
                        }
                    }
                }
                fwrite(mv, 4, 32, out);
                fwrite(mb->re->residual_Y->data, 4, 256, out);
                fwrite(mb->re->residual_U->data, 4, 64, out);
                fwrite(mb->re->residual_V->data, 4, 64, out);
                fwrite(mb->pred_Y->data, 4, 256, out);
            }
    }
}

int main(int argc, char* argv[])
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s <stream.264> <queue.bin> <dump.bin|-> [dump_level=1]\n", argv[0]);
        return 2;
    }
    int level = argc > 4 ? atoi(argv[4]) : 1;
    FILE* fp = fopen(argv[1], "r");
    if (!fp) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    if (strcmp(argv[2], "-")) {
        FILE* fq = fopen(argv[2], "rb");
        if (!fq) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
        fseek(fq, 0, SEEK_END);
        long qbytes = ftell(fq);
        fseek(fq, 0, SEEK_SET);
        g_h264o_queue.resize(qbytes / 4);
        if (qbytes && fread(g_h264o_queue.data(), 4, qbytes / 4, fq) != (size_t)(qbytes / 4)) return 2;
        fclose(fq);
    }
    FILE* out = strcmp(argv[3], "-") ? fopen(argv[3], "wb") : NULL;

    Debug debug("");
    Parser parser(fp, &debug);
    Decoder decoder;

    int n_pic = 0;
    double t_decode = 0.0;
    while (parser.find_nextNAL()) {
        NAL nal(&parser, &decoder);
        nal.pic = NULL;
        double t0 = now_s();
        nal.decode();
        double t1 = now_s();
        if (nal.type == IDR || nal.type == Non_IDR) {
            t_decode += t1 - t0;
            if (out) dump_picture(out, nal, parser, decoder, n_pic, level);
            n_pic++;
        }
    }
    if (out) fclose(out);
    fclose(fp);
    fprintf(stderr, "{\"pictures\": %d, \"decode_seconds\": %.6f, \"queue_used\": %zu, \"queue_total\": %zu}\n",
            n_pic, t_decode, g_h264o_queue_pos / 2, g_h264o_queue.size() / 2);
    return g_h264o_queue_pos == g_h264o_queue.size() ? 0 : 5;
}
