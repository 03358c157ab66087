// h264r_seq.cuh -- the order in which the phases of h264r_core.cuh are run for one macroblock.
//
// `Exec` abstracts "run this phase on all 32 lanes, then synchronise the warp":
//   device: every lane calls f(lane) and the warp meets at __syncwarp();
//   host emulator: a loop over lanes 0..31.
// Keeping the sequences here means the CPU tests run exactly the phase order the kernels use.
#pragma once
#include "h264r_core.cuh"
#include "h264r_mc.cuh"
#include "h264r_intra.cuh"
#include "h264r_deblock.cuh"

namespace h264r {

HD h264r_mb_hdr load_hdr(const PicDesc& pd, int a)
{
    const int4* p = reinterpret_cast<const int4*>(pd.hdr + a);
    int4 raw = H264R_LDG(p);
    return *reinterpret_cast<h264r_mb_hdr*>(&raw);
}

// K1 alone: residual of one MB into out[384] (int16: Y 16x16, U 8x8, V 8x8)
template <bool COMPAT, class Exec>
HD void seq_residual(Exec& ex, MbWork& w, const PicDesc& pd, int a, int16_t* out)
{
    h264r_mb_hdr h = load_hdr(pd, a);
    ex([&](int lane) { phase_load_coeff(w, pd, a, lane); });
    ex([&](int lane) { phase_residual<COMPAT>(w, h, lane); });
    ex([&](int lane) { for (int k = lane; k < 384; k += 32) out[k] = w.res[k]; });
}

// K1 + K3: one inter macroblock
template <bool COMPAT, class Exec>
HD void seq_inter_mb(Exec& ex, MbWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a)
{
    ex([&](int lane) { phase_load_coeff(w, pd, a, lane); });
    ex([&](int lane) {
        phase_residual<COMPAT>(w, h, lane);
        if (!COMPAT && pd.nzmask && lane < 16) w.bs[lane] = blk_coded(w, h, lane) ? 1 : 0;       // bs[] is free until K4
    });
    ex([&](int lane) {
        if (!COMPAT && pd.nzmask && lane == 0) {
            uint32_t m = 0;
            for (int r = 0; r < 16; r++) m |= (uint32_t)w.bs[r] << r;
            pd.nzmask[a] = m;
        }
        phase_inter_luma<COMPAT>(w, pd, g, h, a, lane);
        phase_inter_chroma<COMPAT>(w, pd, g, h, a, lane);
    });
    ex([&](int lane) { phase_store(w, pd, g, a, lane); });
}

// K1 ahead of the intra wavefront: residual of one intra macroblock into the picture's scratch (pd.residual).
// 4x4 transform: register version of h264r_intra.cuh; 8x8 transform / Intra8x8: generic phases.
template <bool COMPAT, class Exec>
HD void seq_residual_scratch(Exec& ex, MbWork& w, const PicDesc& pd, const h264r_mb_hdr& h, int a)
{
    if (!COMPAT && (h.mb_class == H264R_MB_I8x8 || (h.flags & H264R_T8x8))) {      // REF_COMPAT has no 8x8 transform
        ex([&](int lane) { phase_load_coeff(w, pd, a, lane); });
        ex([&](int lane) { phase_residual<COMPAT>(w, h, lane); });
    } else {
        ex([&](int lane) { phase_residual_fast<COMPAT>(w.res, pd, h, a, lane); });
    }
    ex([&](int lane) {
        if (lane < 24) {
            const uint4* src = reinterpret_cast<const uint4*>(w.res) + lane * 2;
            uint4* dst = reinterpret_cast<uint4*>(pd.residual + (size_t)a * 384) + lane * 2;
            dst[0] = src[0]; dst[1] = src[1];
        }
    });
}

// working set of a warp: the generic phases and the fast inter path never run at the same time
struct WarpWork {
    union {
        MbWork mb;
        InterWork in;
    };
    TmaState tma;
    HD WarpWork() {}
};

// K1 + K3 for any inter macroblock: fast path for the 4x4 transform, generic phases for the 8x8 transform.
template <bool COMPAT, bool TMA, bool LEAN, int PART = PART_ALL, class Exec>
HD void seq_inter_any(Exec& ex, WarpWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, const MbPre& pre)
{
    seq_inter_mb_fast<COMPAT, TMA, LEAN, PART>(ex, w.in, w.tma, pd, g, h, a, pre);
}
template <bool COMPAT, bool TMA, class Exec>
HD void seq_inter_any(Exec& ex, WarpWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a)
{
    seq_inter_mb_fast<COMPAT, TMA, false>(ex, w.in, w.tma, pd, g, h, a, load_pre(pd, a));
}

// K1 + K2: one intra macroblock (its neighbours A, B, C, D are complete)
template <bool COMPAT, class Exec, class RowSync>
HD void seq_intra_mb(Exec& ex, MbWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, RowSync& rs)
{
    rs.plan_done();
    ex([&](int lane) {
        phase_load_coeff(w, pd, a, lane);
        phase_intra_edges(w, pd, g, h, a, lane, !COMPAT);
    });
    ex([&](int lane) { phase_residual<COMPAT>(w, h, lane); });
    // a left neighbour reconstructed by another warp of the row: its column replaces what the edge phase read from the frame
    const uint8_t* left_col = rs.top_done(nullptr);
    if (left_col && (h.flags & H264R_AVAIL_A)) ex([&](int lane) { patch_left_col(w, left_col, !COMPAT, lane); });
    rs.edges_done();
    if (h.mb_class == H264R_MB_I16x16) {
        ex([&](int lane) {
            phase_intra16<COMPAT>(w, pd, g, h, a, lane);
            phase_intra_chroma<COMPAT>(w, h, lane);
        });
    } else if (h.mb_class == H264R_MB_I4x4) {
        ex([&](int lane) {
            phase_intra4<COMPAT>(w, pd, h, a, 0, lane);
            phase_intra_chroma<COMPAT>(w, h, lane);
        });
        for (int ras = 1; ras < 16; ras++) ex([&](int lane) { phase_intra4<COMPAT>(w, pd, h, a, ras, lane); });
    } else {
        for (int q = 0; q < 4; q++) {
            ex([&](int lane) {
                phase_intra8_filter(w, h, q, lane);
                if (q == 0) phase_intra_chroma<COMPAT>(w, h, lane);
            });
            ex([&](int lane) { phase_intra8(w, pd, h, a, q, lane); });
        }
    }
    ex([&](int lane) { phase_store(w, pd, g, a, lane); });
}

// Hooks of a row pair (k_intra_rows: two warps take the intra macroblocks of a row in turn): called by every lane between phases.
struct NoRowSync {
    HD void plan_done() {}                                             // next: the samples of the row above are read
    HD const uint8_t* top_done(const uint8_t* left_col) { return left_col; }  // next: the left column is read, from what this returns
    HD void edges_done() {}                                            // the left neighbour's column has been read
    HD void tile_final() {}                                            // MbWork::colx holds this macroblock's right column
};

template <bool COMPAT, class Exec>
HD void seq_intra_mb(Exec& ex, MbWork& w, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a)
{
    NoRowSync rs;
    seq_intra_mb<COMPAT>(ex, w, pd, g, h, a, rs);
}

// K1 + K2, latency version (h264r_intra.cuh) for Intra4x4 / Intra16x16; Intra8x8 takes the generic phases.
// left_col: MbWork::colx of the warp that reconstructed the macroblock to the left immediately before, or null (from the frame).
// `have_res`: w.res already holds the macroblock's residual (the caller moved it in from pd.residual).
template <bool COMPAT, class Exec, class RowSync>
HD void seq_intra_any(Exec& ex, MbWork& w, const uint2* lut, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, const uint8_t* left_col,
                      bool have_res, RowSync& rs)
{
    if (!COMPAT && (h.mb_class == H264R_MB_I8x8 || (h.flags & H264R_T8x8))) {      // REF_COMPAT has no 8x8 transform
        seq_intra_mb<COMPAT>(ex, w, pd, g, h, a, rs);
        ex([&](int lane) { save_right_col(w, lane); });
        rs.tile_final();
        return;
    }
    if (!have_res) ex([&](int lane) { phase_residual_fast<COMPAT>(w.res, pd, h, a, lane); });
    if (h.mb_class == H264R_MB_I16x16) {
        rs.plan_done();
        ex([&](int lane) { phase_intra_edges_fast(w, pd, g, h, a, lane, !COMPAT, nullptr, 1); });
        left_col = rs.top_done(left_col);
        ex([&](int lane) { phase_intra_edges_fast(w, pd, g, h, a, lane, !COMPAT, left_col, 2); });
        rs.edges_done();
        ex([&](int lane) {
            phase_intra16_fast<COMPAT>(w, pd, g, h, a, lane);
            if (COMPAT) store_chroma<true>(w, pd, g, a, lane);
            else phase_intra_chroma<false>(w, h, lane);
        });
        rs.tile_final();
        if (!COMPAT) ex([&](int lane) { store_chroma<false>(w, pd, g, a, lane); });
    } else {
        // everything that does not depend on the neighbours first: the plan; then the row above; then the left column, the
        // last thing to become final (the macroblock to the left is the previous link of the row's chain)
        PerLane<I4Plan> plan;
        ex([&](int lane) { intra4_plan(w, h, lut, lane, plan(lane)); });
        rs.plan_done();
        ex([&](int lane) { phase_intra_edges_fast(w, pd, g, h, a, lane, !COMPAT, nullptr, 1); });
        left_col = rs.top_done(left_col);
        ex([&](int lane) { phase_intra_edges_fast(w, pd, g, h, a, lane, !COMPAT, left_col, 2); });
        rs.edges_done();
        ex([&](int lane) {
            phase_intra4_step<COMPAT, 0>(w, pd, a, plan(lane), lane);
            if (!COMPAT) phase_intra_chroma<false>(w, h, lane);
        });
        ex([&](int lane) { phase_intra4_step<COMPAT, 1>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 2>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 3>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 4>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 5>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 6>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 7>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 8>(w, pd, a, plan(lane), lane); });
        ex([&](int lane) { phase_intra4_step<COMPAT, 9>(w, pd, a, plan(lane), lane); });
        rs.tile_final();
        ex([&](int lane) {
            store_luma_tile(w, pd, g, a, lane);
            store_chroma<COMPAT>(w, pd, g, a, lane);
        });
    }
    if (COMPAT) {        // as in seq_inter_mb_fast: the frame border next to a border macroblock
        int mbx, mby;
        mb_xy(g, a, mbx, mby);
        if (is_border_mb(g, mbx, mby)) ex([&](int lane) { phase_pad_mb_luma(pd.y, g.pitch_y, g.w_mbs, g.h_mbs, mbx, mby, lane); });
    }
}
template <bool COMPAT, class Exec>
HD void seq_intra_any(Exec& ex, MbWork& w, const uint2* lut, const PicDesc& pd, const Geometry& g, const h264r_mb_hdr& h, int a, const uint8_t* left_col,
                      bool have_res = false)
{
    NoRowSync rs;
    seq_intra_any<COMPAT>(ex, w, lut, pd, g, h, a, left_col, have_res, rs);
}

// K4: deblock one macroblock in place (its neighbours A, B, C are completely filtered)
template <class Exec>
HD void seq_deblock_mb(Exec& ex, MbWork& w, const PicDesc& pd, const Geometry& g, int a)
{
    h264r_mb_hdr h = load_hdr(pd, a);
    int mbx, mby;
    mb_xy(g, a, mbx, mby);
    h264r_mb_hdr hl = mbx > 0 ? load_hdr(pd, a - 1) : h;
    h264r_mb_hdr ht = mby > 0 ? load_hdr(pd, a - g.w_mbs) : h;
    ex([&](int lane) {
        phase_load_coeff(w, pd, a, lane);
        phase_deblock_load(w, pd, g, a, lane);
    });
    ex([&](int lane) { phase_deblock_bs(w, pd, g, h, a, lane); });
    ex([&](int lane) { phase_deblock_dir(w, pd, h, hl, 0, lane); });
    ex([&](int lane) { phase_deblock_dir(w, pd, h, ht, 1, lane); });
    ex([&](int lane) { phase_deblock_store(w, pd, g, a, lane); });
}

}  // namespace h264r
