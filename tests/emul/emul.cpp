// TEST INFRASTRUCTURE -- host-side SIMT emulator for the kernel code of
// videodecoder_b200/csrc/cuda/h264r_core.cuh + h264r_seq.cuh.
//
// The CUDA kernels are written as per-warp phase sequences (see h264r_seq.cuh).  This file compiles
// the very same headers with g++ and runs each phase with a loop over the 32 lanes, macroblocks in
// raster order (which satisfies every dependency the device wavefront respects).  It lets the CPU
// test-suite compare the kernels' arithmetic with the oracle without a GPU.  It is NOT a CPU
// fallback: nothing in the product loads it.
#include "../../videodecoder_b200/csrc/cuda/h264r_seq.cuh"
#include <cstring>
#include <vector>

using namespace h264r;

namespace {
struct HostExec {
    template <class F> void operator()(F f) { for (int lane = 0; lane < 32; lane++) f(lane); }
};

void fill_desc(PicDesc& pd, const h264r_pic_params* pp, const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
               uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* const* ref_y, const uint8_t* const* ref_u,
               const uint8_t* const* ref_v, uint8_t* exc_map)
{
    memset(&pd, 0, sizeof(pd));
    pd.hdr = hdr; pd.mv = mv; pd.coeff = coeff; pd.y = y; pd.u = u; pd.v = v;
    for (int i = 0; i < H264R_MAX_REFS; i++) {
        pd.ref_y[i] = ref_y ? ref_y[i] : nullptr;
        pd.ref_u[i] = ref_u ? ref_u[i] : nullptr;
        pd.ref_v[i] = ref_v ? ref_v[i] : nullptr;
        pd.ref_id[i] = pp->ref_list0[i];
        pd.luma_weight[i] = pp->luma_weight[i]; pd.luma_offset[i] = pp->luma_offset[i];
        for (int c = 0; c < 2; c++) { pd.chroma_weight[i][c] = pp->chroma_weight[i][c]; pd.chroma_offset[i][c] = pp->chroma_offset[i][c]; }
    }
    pd.slice_type = pp->slice_type;
    pd.weighted_pred = pp->weighted_pred;
    pd.luma_log2_wd = pp->luma_log2_weight_denom;
    pd.chroma_log2_wd = pp->chroma_log2_weight_denom;
    pd.deblock_on = pp->disable_deblocking_filter_idc != 1;
    pd.alpha_off = pp->slice_alpha_c0_offset_div2 * 2;
    pd.beta_off = pp->slice_beta_offset_div2 * 2;
    pd.excursion_map = exc_map;
}
}  // namespace

extern "C" {

// residual of every MB: out[n_mbs][384] int16
void h264emul_residual(int w_mbs, int h_mbs, unsigned flags, const h264r_mb_hdr* hdr, const int16_t* coeff, int16_t* out)
{
    h264r_pic_params pp;
    memset(&pp, 0, sizeof(pp));
    PicDesc pd;
    fill_desc(pd, &pp, hdr, nullptr, coeff, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    HostExec ex;
    MbWork w;
    for (int a = 0; a < w_mbs * h_mbs; a++) {
        if (flags & H264R_REF_COMPAT) seq_residual<true>(ex, w, pd, a, out + (size_t)a * 384);
        else seq_residual<false>(ex, w, pd, a, out + (size_t)a * 384);
    }
}

// whole picture; planes have pitch w_mbs*16 / w_mbs*8.  Returns the number of luma excursions.
unsigned long long h264emul_recon_picture(int w_mbs, int h_mbs, unsigned flags, const h264r_pic_params* pp,
                                          const h264r_mb_hdr* hdr, const int16_t* mv, const int16_t* coeff,
                                          uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* const* ref_y,
                                          const uint8_t* const* ref_u, const uint8_t* const* ref_v, uint8_t* exc_map)
{
    Geometry g{w_mbs, h_mbs, w_mbs * 16, w_mbs * 8};
    PicDesc pd;
    fill_desc(pd, pp, hdr, mv, coeff, y, u, v, ref_y, ref_u, ref_v, exc_map);
    HostExec ex;
    WarpWork ww;
    MbWork& w = ww.mb;
    unsigned long long exc = 0;
    const bool compat = (flags & H264R_REF_COMPAT) != 0;
    int n = w_mbs * h_mbs;
    // same order as the device: all inter MBs of the picture first, then the intra MBs
    for (int a = 0; a < n; a++) {
        h264r_mb_hdr h = load_hdr(pd, a);
        if (is_intra(h.mb_class)) continue;
        exc += compat ? seq_inter_any<true>(ex, ww, pd, g, h, a) : seq_inter_any<false>(ex, ww, pd, g, h, a);
    }
    for (int a = 0; a < n; a++) {
        h264r_mb_hdr h = load_hdr(pd, a);
        if (!is_intra(h.mb_class)) continue;
        if (compat) seq_intra_mb<true>(ex, w, pd, g, h, a); else seq_intra_mb<false>(ex, w, pd, g, h, a);
        exc += w.excursions;
    }
    if (!compat && pd.deblock_on)
        for (int a = 0; a < n; a++) seq_deblock_mb(ex, w, pd, g, a);
    return exc;
}

void h264emul_digest(const uint8_t* y, const uint8_t* u, const uint8_t* v, int w, int h, uint32_t out[4])
{
    uint32_t d[4] = {0, 0, 0, 0}, i = 0;
    for (int k = 0; k < w * h; k++) digest_accumulate(d, y[k], i++);
    for (int k = 0; k < w * h / 4; k++) digest_accumulate(d, u[k], i++);
    for (int k = 0; k < w * h / 4; k++) digest_accumulate(d, v[k], i++);
    memcpy(out, d, 16);
}
}
