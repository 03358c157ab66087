// h264r_kernels.cuh -- the sm_100a kernels: orchestration (work distribution, wavefront
// dependencies, synchronisation) around the per-macroblock phase sequences of h264r_seq.cuh.
#pragma once
#include "h264r_seq.cuh"
#include "h264r_mcgrp.cuh"

namespace h264r {

// Programmatic dependent launch (see launch_pdl in h264r_api.cu).  launch_dependents: the next kernel of the stream may be
// scheduled from now on.  wait: everything the previous kernel of the stream wrote is complete and visible (returns at once
// when the launch was not programmatic).
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct WarpExec {
    int lane;
    template <class F> __device__ __forceinline__ void operator()(F f)
    {
        f(lane);
        __syncwarp();
    }
};

#ifdef H264R_TRACE
// Development build (-DH264R_TRACE): the row kernels time their phases.  The warp of row 0 of the first picture of a
// launch never waits for anybody, so its clock stamps show what one macroblock costs one warp, phase by phase.
__device__ long long g_trace[4096];
__device__ int g_trace_n;
struct WarpExecTrace {
    int lane; bool on;
    long long t[20]; int k;          // stamps of the current macroblock stay in registers until mark(2)
    template <class F> __device__ __forceinline__ void operator()(F f)
    {
        f(lane);
        __syncwarp();
        if (on && k < 20) t[k++] = clock64();
    }
    __device__ __forceinline__ void mark(long long tag)
    {
        if (!on) return;
        if (tag == 1) { k = 0; t[k++] = clock64(); return; }
        long long now = clock64();
        if (lane == 0 && g_trace_n + k + 4 < 4096) {
            int n = g_trace_n;
            g_trace[n++] = -1; g_trace[n++] = t[0];
            for (int i = 1; i < k; i++) g_trace[n++] = t[i];
            g_trace[n++] = -2; g_trace[n++] = now;
            g_trace_n = n;
        }
    }
};
#ifndef H264R_TRACE_ROW
#define H264R_TRACE_ROW 0      // which row of the first picture is timed (0: never waits; others show the hand-over)
#endif
#define ROW_EXEC(flat) WarpExecTrace ex; ex.lane = (int)(threadIdx.x & 31); ex.on = (flat) == H264R_TRACE_ROW; ex.k = 0
#define ROW_MARK(tag) ex.mark(tag)
#else
#define ROW_EXEC(flat) WarpExec ex{(int)(threadIdx.x & 31)}
#define ROW_MARK(tag)
#endif

// The pictures a launch works on: indices into the device descriptor table (one descriptor per syntax-arena slot),
// passed BY VALUE in the kernel parameters.  Descriptors are uploaded with their picture's syntax (copy stream), so
// a flush itself copies nothing: a host-to-device copy issued at flush time would queue behind every syntax upload
// still in flight and serialise the kernels after them.
constexpr int PIC_LIST_MAX = 940;
struct PicList { int n; int slot[PIC_LIST_MAX]; };
// descriptor of arena slot s: the first bytes of the slot itself (it is uploaded with, and in front of, the syntax)
struct DescTable {
    const uint8_t* base; size_t stride;
    __device__ __forceinline__ const PicDesc& operator[](int slot) const { return *reinterpret_cast<const PicDesc*>(base + (size_t)slot * stride); }
};

#ifndef H264R_INTER_MIN_CTAS
#define H264R_INTER_MIN_CTAS 8       // resident CTAs per SM the inter kernel is compiled for (register cap 65536 / (128 * this))
#endif
constexpr int INTER_WARPS = 4;       // warps per CTA of the residual kernels
#ifndef H264R_KINTER_WARPS
#define H264R_KINTER_WARPS 1         // warps per CTA of k_inter.  1: the warp's working set is THE shared block of its CTA, a constant
#endif                               // address -- with several warps per CTA the compiler recomputes `work + warp * size` from
constexpr int KINTER_WARPS = H264R_KINTER_WARPS;      // threadIdx.x dozens of times per macroblock (it will not spend a register on it)

// ------------------------------------------------------------------------------------------------
// K1 alone: dequant + inverse transform of n_mbs macroblocks -> int16 residual (384 per MB).
// One warp per macroblock, grid-stride.
// ------------------------------------------------------------------------------------------------
template <bool COMPAT>
__global__ void __launch_bounds__(INTER_WARPS * 32) k_residual(PicDesc pd, int n_mbs, int16_t* out)
{
    __shared__ MbWork work[INTER_WARPS];
    int warp = threadIdx.x >> 5;
    WarpExec ex{(int)(threadIdx.x & 31)};
    for (int a = blockIdx.x * INTER_WARPS + warp; a < n_mbs; a += gridDim.x * INTER_WARPS)
        seq_residual<COMPAT>(ex, work[warp], pd, a, out + (size_t)a * 384);
}

// ------------------------------------------------------------------------------------------------
// K1+K3: all inter macroblocks of the P pictures of one wave.  No dependencies between MBs (their
// references are complete pictures), so this is a flat grid: one warp per MB, blockIdx.y = picture,
// blockIdx.x strides over its macroblocks.
// ------------------------------------------------------------------------------------------------
// TMA: reference windows of 16x16 / 8x8-quadrant blocks arrive by cp.async.bulk.tensor (h264r_mc.cuh) instead of loads.
// LEAN: every picture of the launch arrived as a picture buffer with partition-order vectors and compact levels and
// without weighted prediction (what a host parser produces for ordinary P pictures; the scheduler checks).  The other
// storages and the weighting then compile away: 300 instructions less in a kernel whose hot code does not fit the 32 KB
// L1.5 instruction cache (ncu: "no instruction" is its second largest stall).
// PART (h264r_mc.cuh): spec-mode lean launches run as three kernels -- luma of the macroblocks without the 8x8 transform, luma of
// those with it, chroma -- each of which fits the instruction cache; every other launch is one kernel (PART_ALL).
template <bool COMPAT, bool TMA, bool LEAN, int PART = PART_ALL>
__global__ void __launch_bounds__(KINTER_WARPS * 32, H264R_INTER_MIN_CTAS * 4 / KINTER_WARPS) k_inter(const DescTable descs, const __grid_constant__ PicList list, Geometry g)
{
    __shared__ WarpWork work[KINTER_WARPS];
    const int warp = KINTER_WARPS == 1 ? 0 : (int)(threadIdx.x >> 5);
    WarpExec ex{KINTER_WARPS == 1 ? (int)threadIdx.x : (int)(threadIdx.x & 31)};
    WarpWork& w = work[warp];
    if (TMA) {
        if (ex.lane == 0) { w.tma.parity = 1u; mbar_init(&w.tma.mbar, 1); }       // first use waits for parity 0
        __syncwarp();
    }
    const int n_mbs = g.w_mbs * g.h_mbs;
    const PicDesc& pd = descs[list.slot[blockIdx.y]];
    // the header of the warp's NEXT macroblock is fetched while the current one is reconstructed: it heads the
    // dependent chain header -> vectors -> window addresses -> reference loads
    const int stride = gridDim.x * KINTER_WARPS;
    int a = blockIdx.x * KINTER_WARPS + warp;
    int4 hraw = a < n_mbs ? __ldg(reinterpret_cast<const int4*>(pd.hdr + a)) : int4{0, 0, 0, 0};
    MbPre pre_next = a < n_mbs ? load_pre(pd, a) : MbPre{0u, 0u, 0u};
    // syntax does not depend on the previous kernel of the stream, reference pictures do
    grid_launch_dependents();
    grid_dependency_wait();
    for (; a < n_mbs; a += stride) {
        h264r_mb_hdr h = *reinterpret_cast<h264r_mb_hdr*>(&hraw);
        const MbPre pre = pre_next;
        if (a + stride < n_mbs) {
            hraw = __ldg(reinterpret_cast<const int4*>(pd.hdr + a + stride));
            pre_next = load_pre(pd, a + stride);
        }
        if (is_intra(h.mb_class)) continue;       // warp-uniform; the intra kernels reconstruct it afterwards
        seq_inter_any<COMPAT, TMA, LEAN, PART>(ex, w, pd, g, h, a, pre);
    }
}

// ------------------------------------------------------------------------------------------------
// K1+K3, group version (h264r_mcgrp.cuh): REF_COMPAT launches whose pictures are all lean (see k_inter).  One warp per CTA takes
// GRP_MBS consecutive macroblocks at a time; blockIdx.y = picture, blockIdx.x strides over its groups.  Header decoding runs with
// lane = macroblock, the residual with lane = coded block (dense passes), both before the dependency on the previous kernel of
// the stream is awaited: only staging reads reference samples.
// ------------------------------------------------------------------------------------------------
#ifndef H264R_GRP_MIN_CTAS
#define H264R_GRP_MIN_CTAS 32        // resident CTAs (= warps) per SM k_inter_grp is compiled for (32: 64 registers)
#endif
// TMA: the windows of 16x16 macroblocks arrive as tensor copies (cp.async.bulk.tensor.2d of the 48 x 22 box of 16-byte columns
// that contains the window, completion on an mbarrier), two deep; every other window as aligned word loads; the consumer shifts.
// Opt-in (H264R_GROUP_TMA=1): measured 0.147 ms per wave against 0.134 ms for load staging with pre-shifted words.
template <bool COMPAT, bool TMA>
__global__ void __launch_bounds__(32, H264R_GRP_MIN_CTAS) k_inter_grp(const DescTable descs, const __grid_constant__ PicList list, Geometry g)
{
    __shared__ GroupWork w;
    __shared__ GroupTma tws[1];                          // (the default kernel never touches it and the compiler drops it)
    GroupTma* tw = TMA ? &tws[0] : nullptr;
    WarpExec ex{(int)threadIdx.x};
    uint32_t par[2] = {0u, 0u};
    if (TMA) {
        if (ex.lane == 0) { mbar_init(&tw->mbar[0], 1); mbar_init(&tw->mbar[1], 1); }
        __syncwarp();
    }
    const int n_mbs = g.w_mbs * g.h_mbs;
    const PicDesc& pd = descs[list.slot[blockIdx.y]];
    grid_launch_dependents();
    bool waited = false;
    for (int a0 = (int)blockIdx.x * GRP_MBS; a0 < n_mbs; a0 += (int)gridDim.x * GRP_MBS)
        seq_inter_group<COMPAT, TMA>(ex, w, pd, g, a0, n_mbs, [&]() { if (!waited) { grid_dependency_wait(); waited = true; } }, par, tw);
    if (!waited) grid_dependency_wait();      // the grid must not retire before its predecessor (see k_intra_list)
}

// ------------------------------------------------------------------------------------------------
// K1 ahead of the intra wavefront for pictures without inter macroblocks (I pictures): residual of every macroblock
// into the picture's scratch, flat grid like k_inter.  Keeps dequantisation and the transforms off the wavefront's
// critical path, which is (W + 2H) macroblocks long.
// ------------------------------------------------------------------------------------------------
template <bool COMPAT>
__global__ void __launch_bounds__(INTER_WARPS * 32) k_residual_pics(const DescTable descs, const __grid_constant__ PicList list, Geometry g)
{
    __shared__ MbWork work[INTER_WARPS];
    int warp = threadIdx.x >> 5;
    WarpExec ex{(int)(threadIdx.x & 31)};
    const int n_mbs = g.w_mbs * g.h_mbs;
    const PicDesc& pd = descs[list.slot[blockIdx.y]];
    // a warp looks at 32 macroblocks at a time (one mb_class byte per lane) and transforms the intra ones
    for (int a0 = (blockIdx.x * INTER_WARPS + warp) * 32; a0 < n_mbs; a0 += gridDim.x * INTER_WARPS * 32) {
        const int al = a0 + ex.lane;
        uint32_t m = __ballot_sync(0xffffffffu, al < n_mbs && is_intra(H264R_LDG(&pd.hdr[al].mb_class)));
        while (m) {
            const int a = a0 + __ffs(m) - 1;
            m &= m - 1;
            h264r_mb_hdr h = load_hdr(pd, a);
            seq_residual_scratch<COMPAT>(ex, work[warp], pd, h, a);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Wavefront scheduling shared by K2 (intra) and K4 (deblock): one WARP per macroblock row.
//
// Macroblock (x, y) needs A (x-1, y), B (x, y-1), C (x+1, y-1), D (x-1, y-1).  A is the previous
// macroblock of the same warp; B, C, D belong to the warp of row y-1, which publishes in
// progress[y-1] how many leading macroblocks of its row are final.  The warp of row y may start
// (x, y) once progress[y-1] >= min(x + 2, W): rows run two macroblocks behind each other, the
// classic anti-diagonal wavefront, with up to min(H, W/2) rows of every picture of the wave in
// flight at once and all pictures of the wave side by side (n_pics * H warps spread over all SMs).
//
// Row warps are numbered (picture, row) -> picture * H + row and handed to CTAs through an atomic
// ticket, so the warp a row waits for always belongs to the same CTA or to a CTA that took an
// earlier ticket, i.e. one that is already running: the spin-wait cannot deadlock whatever order
// the hardware starts CTAs in.  The hand-over is st.release.gpu / ld.acquire.gpu on the progress
// word; neighbouring samples are then read with ld.global.cg (L2) by the phases.
//
// In P pictures only the intra macroblocks take part (k_inter reconstructed the rest before this
// launch): the warp finds them with a ballot over the row's mb_class bytes and publishes, after each
// one, the position of the NEXT intra macroblock of its row (everything before it is final).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Polls with relaxed loads (an acquire load invalidates the SM's L1 every time it is issued, for every warp of the
// SM) and fences once when the value has arrived: relaxed load + fence.acq_rel is an acquire pattern.
__device__ __forceinline__ void wait_progress(const int* p, int need)
{
    if (ld_relaxed_gpu(p) < need) {
        int spins = 0;
        do {
#ifndef H264R_POLL_NOSLEEP
            __nanosleep(20);
#else
            if (spins > H264R_POLL_NOSLEEP) __nanosleep(20);      // tuning: the first polls back to back
#endif
            if (++spins > (1 << 24)) __trap();       // seconds: a hand-over that never comes must end the launch, not hang the GPU
        } while (ld_relaxed_gpu(p) < need);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// does intra macroblock h read samples of its neighbour C (above right)?
__device__ __forceinline__ bool intra_reads_c(const h264r_mb_hdr& h)
{
    if (!(h.flags & H264R_AVAIL_C) || h.mb_class == H264R_MB_I16x16) return false;
    if (h.mb_class == H264R_MB_I4x4) {
        const int m = hdr_pred_mode(h, 5);       // luma4x4BlkIdx 5 = the block at raster (0, 3)
        return m == 3 || m == 7;
    }
    return true;
}

constexpr int ROW_WARPS = 4;          // row warps per CTA of the wavefront kernels
constexpr int ROW_MASK_WORDS = 16;    // intra bitmap of a row: up to 512 macroblocks (8192 samples) wide

// Takes the CTA's ticket and returns the flat row index of the calling warp.
__device__ __forceinline__ int row_ticket(int* ticket)
{
    __shared__ int s_ticket;
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1);
    __syncthreads();
    return s_ticket * ROW_WARPS + (int)(threadIdx.x >> 5);
}
// the same for CTAs whose rows are not one per warp: returns the flat index of the CTA's first row
__device__ __forceinline__ int row_ticket(int* ticket, int rows_per_cta)
{
    __shared__ int s_ticket2;
    if (threadIdx.x == 0) s_ticket2 = atomicAdd(ticket, 1);
    __syncthreads();
    return s_ticket2 * rows_per_cta;
}

// Two warps share a macroblock row and take its intra macroblocks in turn (warp `half` the even / odd ones of the row's list):
// while one runs the dependent part of macroblock n -- neighbour samples, the ten prediction steps -- the other stores
// macroblock n - 1 and prepares n + 1 (header, residual, Intra4x4 plan), so a row advances one macroblock per "steps" time
// instead of per whole macroblock.  The right column of a finished macroblock goes to the partner through MbWork::colx in shared
// memory; two counters per warp order the pair (all in shared memory, no fences beyond the CTA):
//   done   macroblocks whose right column is in colx (set right after the last step, before the stores)
//   edges  macroblocks whose neighbour samples have been read (the partner may overwrite its colx again)
// Each warp publishes its own progress word to the row below: "every macroblock of MINE left of this x is final".
struct RowPairFlags { int done[2], edges[2]; };
__device__ __forceinline__ void pair_wait(const volatile int* p, int need)
{
    for (int spins = 0; *p < need; spins++)
        if (spins > (1 << 26)) __trap();       // a partner that never arrives must end the launch, not hang the GPU
    __threadfence_block();
}
struct RowPairSync {
    volatile RowPairFlags* f;
    int half, k, need_other, lane;
    const int* up0; const int* up1; int up_need;      // progress words of the row above and how far it must be (0: nothing to wait for)
    const uint8_t* partner_col; bool left_is_partner;
    __device__ __forceinline__ void plan_done()
    {
        if (up_need > 0) { wait_progress(up0, up_need); wait_progress(up1, up_need); }
    }
    // the macroblock to the left, if it is the partner's previous one: its column comes through shared memory
    __device__ __forceinline__ const uint8_t* top_done(const uint8_t*)
    {
        if (!left_is_partner) return nullptr;
        pair_wait(&f->done[half ^ 1], need_other);
        return partner_col;
    }
    __device__ __forceinline__ void set(volatile int* p, int v)
    {
        __threadfence_block();
        __syncwarp();
        if (lane == 0) *p = v;
    }
    __device__ __forceinline__ void edges_done()
    {
        set(&f->edges[half], k + 1);
        pair_wait(&f->edges[half ^ 1], need_other);      // the partner has read this warp's last column: colx may be rewritten
    }
    __device__ __forceinline__ void tile_final() { set(&f->done[half], k + 1); }
};

template <bool COMPAT>
__global__ void __launch_bounds__(ROW_WARPS * 64, 2) k_intra_rows(const DescTable descs, const __grid_constant__ PicList list,
                                                              Geometry g, int* progress, int* ticket)
{
    __shared__ MbWork work[ROW_WARPS * 2];
    __shared__ uint32_t s_mask[ROW_WARPS * 2][2][ROW_MASK_WORDS + 1];     // intra bitmaps of this row and of the row above (per warp)
    __shared__ uint2 s_lut[INTRA4_LUT_SIZE];
    __shared__ RowPairFlags s_pair[ROW_WARPS];
    intra4_lut_fill(s_lut, threadIdx.x, blockDim.x);
    if (threadIdx.x < ROW_WARPS) s_pair[threadIdx.x] = RowPairFlags{{0, 0}, {0, 0}};
    const int W = g.w_mbs, H = g.h_mbs;
    const int warp = threadIdx.x >> 5, slot = warp >> 1, half = warp & 1;
    const int flat = row_ticket(ticket, ROW_WARPS) + slot;          // __syncthreads() inside: table and flags are complete
    if (flat >= list.n * H) return;
    ROW_EXEC(flat * 2 + half);
    const int p = flat / H, row = flat - p * H;
    const PicDesc& pd = descs[list.slot[p]];
    int* prog = progress + ((size_t)p * H) * 2;              // two words per row: one per warp of the pair
    MbWork& w = work[warp];
    const MbWork& partner = work[warp ^ 1];
    uint32_t* mask = s_mask[warp][0];
    uint32_t* above = s_mask[warp][1];
    const int n_words = (W + 31) >> 5;
    for (int k = 0; k < n_words; k++) {
        int x = k * 32 + ex.lane;
        bool intra = x < W && is_intra(H264R_LDG(&pd.hdr[row * W + x].mb_class));
        bool intra_up = row > 0 && x < W && is_intra(H264R_LDG(&pd.hdr[(row - 1) * W + x].mb_class));
        uint32_t m = __ballot_sync(0xffffffffu, intra), mu = __ballot_sync(0xffffffffu, intra_up);
        if (ex.lane == 0) { mask[k] = m; above[k] = mu; }
    }
    if (ex.lane == 0) above[n_words] = 0;
    __syncwarp();
    auto next_intra = [&](int from) -> int {      // first intra macroblock at x >= from, or W
        for (int k = from >> 5; k < n_words; k++) {
            uint32_t m = mask[k];
            if (k == (from >> 5)) m &= 0xffffffffu << (from & 31);
            if (m) return k * 32 + __ffs(m) - 1;
        }
        return W;
    };
    // this warp's macroblocks: entries half, half + 2, ... of the row's list; prev_x = the entry before (the partner's)
    int prev_x = -2, x = next_intra(0);
    if (half) { prev_x = x; x = x < W ? next_intra(x + 1) : W; }
    if (ex.lane == 0 && x > 0) st_release_gpu(&prog[row * 2 + half], x);
    // software pipeline: header and residual of this warp's NEXT macroblock travel in registers while the current one is
    // predicted (their latency is off the critical path)
    const bool have_res = pd.residual != nullptr;
    int4 hraw = int4{0, 0, 0, 0};
    uint4 r0 = uint4{0, 0, 0, 0}, r1 = uint4{0, 0, 0, 0};
    auto fetch = [&](int xn) {
        if (xn >= W) return;
        const int an = row * W + xn;
        hraw = __ldg(reinterpret_cast<const int4*>(pd.hdr + an));
        if (have_res && ex.lane < 24) {
            const uint4* src = reinterpret_cast<const uint4*>(pd.residual + (size_t)an * 384) + ex.lane * 2;
            r0 = __ldg(src); r1 = __ldg(src + 1);
        }
    };
    fetch(x);
    RowPairSync rs;
    rs.f = &s_pair[slot]; rs.half = half; rs.lane = ex.lane;
    for (int k = 0; x < W; k++) {
        const int a = row * W + x;
        h264r_mb_hdr h = *reinterpret_cast<h264r_mb_hdr*>(&hraw);
        if (have_res && ex.lane < 24) {           // the previous macroblock ended with __syncwarp(): w.res is free
            uint4* dst = reinterpret_cast<uint4*>(w.res) + ex.lane * 2;
            dst[0] = r0; dst[1] = r1;
        }
        __syncwarp();                             // the first phase already reads w.res (Intra4x4 plan)
        const int x_mid = next_intra(x + 1);      // the partner's next macroblock
        const int x_next = x_mid < W ? next_intra(x_mid + 1) : W;
        fetch(x_next);
        rs.up_need = 0;
        if (row > 0) {
            // B, C, D are final unless they are intra macroblocks the row above has not reached yet: the sequence waits (after
            // its plan) for the last intra one among (x-1, x, x+1) of the row above -- on both warps of that row
            // (C's samples are read by the block in the macroblock's top right corner only, and only by the two Intra4x4 modes that
            // look up-right -- diagonal down-left and vertical-left -- or by Intra8x8, whose reference-sample filter takes them: any
            // other macroblock follows the row above at a distance of ONE macroblock instead of two, which shortens the wavefront's
            // critical path from W + 2H towards W + H steps)
            int last = -1;
            for (int xx = intra_reads_c(h) ? x + 1 : x; xx >= x - 1 && xx >= 0; xx--)
                if (xx < W && ((above[xx >> 5] >> (xx & 31)) & 1)) { last = xx; break; }
            rs.up_need = last + 1;
            rs.up0 = &prog[(row - 1) * 2]; rs.up1 = &prog[(row - 1) * 2 + 1];
        }
        rs.k = k;
        rs.need_other = half ? k + 1 : k;         // how many macroblocks the partner has finished when its entry before mine is done
        rs.left_is_partner = prev_x == x - 1 && rs.need_other > 0;
        rs.partner_col = partner.colx;
        ROW_MARK(1);
        seq_intra_any<COMPAT>(ex, w, s_lut, pd, g, h, a, nullptr, have_res, rs);
        prev_x = x_mid;
        x = x_next;
        // ex() ended with __syncwarp(): every lane's stores precede lane 0's release; the partner learns that the samples are in the frame
        if (ex.lane == 0) st_release_gpu(&prog[row * 2 + half], x);
        ROW_MARK(2);
    }
}

// ------------------------------------------------------------------------------------------------
// K1+K2 for the intra macroblocks of P pictures, where they are few and scattered: one warp TICKET per intra
// macroblock instead of one warp per row.  k_blk_scan (prep stream, long before) left every picture's intra
// macroblocks as a list in address order; tickets run over the lists of the launch's pictures, so a warp
// reconstructs one macroblock -- residual included: k_residual_pics and the scratch round trip disappear for these
// pictures -- and takes the next ticket.  The inter macroblocks are final (k_inter ran before this launch); an intra
// neighbour A, D, B, C has a LOWER address, hence a lower ticket: it is finished or held by a running warp, which
// stamps intra_done[address] with the launch's stamp when it is through (only if one of the four macroblocks that read
// its samples is an intra macroblock -- otherwise nobody will ask).  With 3 % intra macroblocks nearly all of them
// start at once and the launch lasts about two macroblocks.
// ------------------------------------------------------------------------------------------------
constexpr int SPARSE_PICS = 64;
struct SparseList { int n; int slot[SPARSE_PICS]; int cum[SPARSE_PICS + 1]; };      // cum[i] = tickets before picture i
#ifndef H264R_LIST_MIN_CTAS
#define H264R_LIST_MIN_CTAS 4        // resident CTAs per SM k_intra_list is compiled for in spec mode (4: 128 registers)
#endif
#ifndef H264R_LIST_MIN_CTAS_COMPAT
#define H264R_LIST_MIN_CTAS_COMPAT 8 // ... under REF_COMPAT (8: 64 registers, 16 bytes of spills): the benchmark's 3 900 intra macroblocks per
#endif                               // wave then have a resident warp each instead of two in a row per warp (0.14 ms less per step)
constexpr int list_ctas_per_sm(bool compat) { return compat ? H264R_LIST_MIN_CTAS_COMPAT : H264R_LIST_MIN_CTAS; }
template <bool COMPAT>
__global__ void __launch_bounds__(ROW_WARPS * 32, list_ctas_per_sm(COMPAT)) k_intra_list(const DescTable descs, const __grid_constant__ SparseList list,
                                                              Geometry g, int* ticket, int stamp)
{
    __shared__ MbWork work[ROW_WARPS];
    __shared__ uint2 s_lut[INTRA4_LUT_SIZE];
    grid_launch_dependents();
    intra4_lut_fill(s_lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int W = g.w_mbs, H = g.h_mbs, total = list.cum[list.n];
    WarpExec ex{(int)(threadIdx.x & 31)};
    MbWork& w = work[threadIdx.x >> 5];
    int p = 0;
    for (;;) {
        int t = 0;
        if (ex.lane == 0) t = atomicAdd(ticket, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) break;
        while (t >= list.cum[p + 1]) p++;          // tickets only grow
        const PicDesc& pd = descs[list.slot[p]];
        const int idx = t - list.cum[p];
        if (idx >= (int)H264R_LDG(pd.intra_list)) continue;      // the host's count was an upper bound
        const int a = (int)H264R_LDG(pd.intra_list + 1 + idx);
        const h264r_mb_hdr h = load_hdr(pd, a);
        int mbx, mby;
        mb_xy(g, a, mbx, mby);
        // the residual needs no sample: computed before this launch's dependency on k_inter is awaited (the generic
        // phases of the 8x8 transform compute their own)
        const bool res_early = COMPAT || !(h.mb_class == H264R_MB_I8x8 || (h.flags & H264R_T8x8));
        if (res_early) ex([&](int lane) { phase_residual_fast<COMPAT>(w.res, pd, h, a, lane); });
        grid_dependency_wait();       // k_inter's samples (the first ticket pays, later calls return at once)
        // lanes 0..3 look at A, D, B, C (an intra one must have been stamped), lanes 4..7 at the four macroblocks that
        // read this one's samples (if none is intra, the stamp -- a fence -- is skipped)
        bool publish;
        {
            const int l = ex.lane & 3;
            const bool up = ex.lane < 4;
            const int d = l == 0 ? 1 : (l == 1 ? W + 1 : (l == 2 ? W : W - 1));
            const int nb = up ? a - d : a + d;
            bool exists;
            if (up) exists = l == 0 ? mbx > 0 : (l == 1 ? mbx > 0 && mby > 0 : (l == 2 ? mby > 0 : mby > 0 && mbx < W - 1));
            else exists = l == 0 ? mbx < W - 1 : (l == 1 ? mbx < W - 1 && mby < H - 1 : (l == 2 ? mby < H - 1 : mby < H - 1 && mbx > 0));
            const bool hit = ex.lane < 8 && exists && is_intra(H264R_LDG(&pd.hdr[nb].mb_class));
            const uint32_t hm = __ballot_sync(0xffffffffu, hit);
            publish = (hm & 0xf0u) != 0;
            uint32_t wm = hm & 0xfu;
            while (wm) {          // every lane waits (one address per neighbour): its own acquire orders its own loads
                const int j = __ffs(wm) - 1;
                wm &= wm - 1;
                const int nbj = __shfl_sync(0xffffffffu, nb, j);
                wait_progress(pd.intra_done + nbj, stamp);
            }
        }
        seq_intra_any<COMPAT>(ex, w, s_lut, pd, g, h, a, nullptr, res_early);
        // ex() ended with __syncwarp(): every lane's stores precede lane 0's release
        if (publish && ex.lane == 0) st_release_gpu(pd.intra_done + a, stamp);
    }
    // A CTA that never took a real macroblock (the host's counts are upper bounds) has not met its predecessor yet: the
    // grid must not retire -- and release the next k_inter, which waits on THIS grid only -- while the k_inter before it
    // still runs.  Free when the dependency is already satisfied.
    grid_dependency_wait();
}

// K4, row wavefront: one warp per macroblock row, all pictures of the launch side by side.  Rows hand over through stamped
// per-macroblock records (h264r_deblock.cuh, MBOX): the row below polls the record of the macroblock above, which its row
// publishes as soon as the macroblock to ITS right has filtered its left edge -- no progress counters, no fences.  Tickets are
// taken in (picture, row) order, so the row a warp waits for is always running.
__global__ void __launch_bounds__(ROW_WARPS * 32, 4) k_deblock_rows(const DescTable descs, const __grid_constant__ PicList list,
                                                                   Geometry g, uint32_t stamp, int* ticket)
{
    __shared__ DeblockWork work[ROW_WARPS];
    const int W = g.w_mbs, H = g.h_mbs;
    const int flat = row_ticket(ticket);
    if (flat >= list.n * H) return;
    const int warp = threadIdx.x >> 5;
    ROW_EXEC(flat);
    const int p = flat / H, row = flat - p * H;
    const PicDesc& pd = descs[list.slot[p]];
    DeblockWork& w = work[warp];
    PerLane<DeblockIn> in;
    // the macroblock's record and its own samples do not depend on the neighbours: fetched one macroblock ahead
    DeblockIn next = deblock_fetch(pd, g, pd.dbrec, row * W, 0, row, ex.lane);
    for (int x = 0; x < W; x++) {
        const int a = row * W + x;
        in(ex.lane) = next;
        if (x + 1 < W) next = deblock_fetch(pd, g, pd.dbrec, a + 1, x + 1, row, ex.lane);
        // first look at the record of the macroblock above, issued before the vertical edges so that its round trip hides
        // behind them; the poll inside the sequence takes over if it has not been published yet
        const uint32_t* top = pd.mbox + (size_t)(a - W) * MBOX_WORDS + 2 * (ex.lane < 24 ? ex.lane : 0);
        uint2 first = uint2{0u, 0u};
        if (row > 0 && ex.lane < 24) first = mbox_load(top);
        auto poll = [&](int lane) -> uint32_t {
            if (lane >= 24) return 0u;
            uint2 v = first;
            for (int spins = 0; v.y != stamp; spins++) {
                if (spins > (1 << 22)) __trap();       // a hand-over that never comes must end the launch, not hang the GPU
                v = mbox_load(top);
            }
            return v.x;
        };
        ROW_MARK(1);
        seq_deblock_mb_fast(ex, w, pd, g, in, x > 0, x, row, stamp, poll);
        ROW_MARK(2);
    }
}

// K4 preparation: boundary strengths and filter parameters of every macroblock of the listed pictures (flat grid like
// k_inter; nothing here depends on sample values, so it stays off the wavefront's critical path).
__global__ void __launch_bounds__(128) k_deblock_prep(const DescTable descs, const __grid_constant__ PicList list, Geometry g)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mbs = g.w_mbs * g.h_mbs;
    const PicDesc& pd = descs[list.slot[blockIdx.y]];
    for (int a = blockIdx.x * 4 + warp; a < n_mbs; a += gridDim.x * 4) {
        int mbx, mby;
        mb_xy(g, a, mbx, mby);
        const h264r_mb_hdr h = load_hdr(pd, a), hl = mbx > 0 ? load_hdr(pd, a - 1) : h, ht = mby > 0 ? load_hdr(pd, a - g.w_mbs) : h;
        phase_deblock_prep(pd.dbrec + (size_t)a * DBREC_BYTES, pd, g, h, hl, ht, pd.nzmask[a], mbx > 0 ? pd.nzmask[a - 1] : 0u,
                           mby > 0 ? pd.nzmask[a - g.w_mbs] : 0u, a, mbx, mby, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Per-picture index structures, one CTA per picture, on the prep stream (behind the picture's upload, ahead of the
// kernels that read them):
//   packed storage: blk_off[a] = number of stored coefficient blocks before macroblock a = exclusive scan of
//     popcount(blk_mask); mv_off[a] (partition-order vectors only) = exclusive scan of the vector counts that follow
//     from mb_class / sub_mb_type;
//   intra_list (P pictures that go to k_intra_list): the addresses of the intra macroblocks in ascending order, their
//     number in front.
// The three scans share the passes (block count in the low, vector count in the high half of a 64-bit word; intra
// count in a second word).
// ------------------------------------------------------------------------------------------------
// exclusive prefix of v over the CTA's 1024 threads (sm: 33 words; total = the sum over the CTA); all threads call it
__device__ __forceinline__ unsigned long long cta_excl_prefix(unsigned long long v, unsigned long long* sm, unsigned long long& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long x = sm[lane];
        unsigned long long ix = x;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= o) ix += t;
        }
        sm[lane] = ix - x;
        if (lane == 31) sm[32] = ix;
    }
    __syncthreads();
    const unsigned long long r = sm[warp] + incl - v;
    total = sm[32];
    __syncthreads();
    return r;
}

// k_blk_scan for a picture in STREAM form (H264R_HDR_STREAM, include/h264r.h): the wire carries 8-byte headers, a word count per
// macroblock and one stream with each macroblock's mask, prediction modes, vectors and blocks; this expands it into what the
// kernels read -- 16-byte headers, dense masks, the vector stream in partition order, 8-byte units -- next to the usual index
// structures.  Two prefix sums per tile: the stream offsets (from the word counts), then unit / vector / intra counts (the
// masks are IN the stream, so they can only be read once the offsets are known).
__device__ void blk_scan_stream(const PicDesc& pd, int n_mbs)
{
    __shared__ unsigned long long sm[33];
    uint32_t* off = const_cast<uint32_t*>(pd.blk_off);
    uint32_t* mvoff = const_cast<uint32_t*>(pd.mv_off);
    uint32_t* ilist = pd.intra_list;
    unsigned long long wbase = 0, base = 0, ibase = 0;
    for (int t0 = 0; t0 < n_mbs; t0 += 8192) {
        const int first = t0 + (int)threadIdx.x * 8;
        // the eight word counts of this thread's macroblocks (the array is 16-byte aligned, `first` a multiple of 8)
        uint2 wc = uint2{0u, 0u};
        if (first + 8 <= n_mbs) wc = __ldg(reinterpret_cast<const uint2*>(pd.mb_words + first));
        else for (int k = 0; k < 8; k++) if (first + k < n_mbs) { const uint32_t b = __ldg(pd.mb_words + first + k); if (k < 4) wc.x |= b << (8 * k); else wc.y |= b << (8 * (k - 4)); }
        const uint32_t s4 = (wc.x & 0x00ff00ffu) + ((wc.x >> 8) & 0x00ff00ffu) + (wc.y & 0x00ff00ffu) + ((wc.y >> 8) & 0x00ff00ffu);
        unsigned long long tot;
        const uint32_t woff0 = (uint32_t)(wbase + cta_excl_prefix((s4 & 0xffffu) + (s4 >> 16), sm, tot));
        wbase += tot;
        // counts: units in the low, vectors in the high half; intra separately
        uint32_t c[8];
        unsigned long long sum = 0;
        uint32_t n_intra = 0;
        {
            uint32_t wo = woff0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                c[k] = 0;
                if (first + k < n_mbs) {
                    const uint2 s8 = __ldg(pd.hdr8 + first + k);
                    const int cls = (int)(s8.x & 7u);
                    c[k] = stream_mb_counts(s8, stream_mb_has_mask(s8) ? __ldg(pd.stream + wo) : 0u);
                    sum += (unsigned long long)(c[k] & 255u) | ((unsigned long long)((c[k] >> 8) & 255u) << 32);
                    n_intra += c[k] >> 16;
                    if (cls == 3 || cls > H264R_MB_P8x8) atomicOr(pd.exc_counter + 1, H264R_STATUS_BAD_CLASS);
                    else if (!is_intra(cls) && pd.slice_type != H264R_SLICE_P) atomicOr(pd.exc_counter + 1, H264R_STATUS_INTER_IN_I);
                }
                wo += (k < 4 ? wc.x >> (8 * k) : wc.y >> (8 * (k - 4))) & 255u;
            }
        }
        unsigned long long run = base + cta_excl_prefix(sum, sm, tot);
        base += tot;
        uint32_t irun = (uint32_t)(ibase + cta_excl_prefix(n_intra, sm, tot));
        ibase += tot;
        // where every macroblock's words, units and vectors begin; the expansion itself is k_stream_expand's (one thread per macroblock)
        uint32_t wo = woff0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int a = first + k;
            if (a < n_mbs) {
                pd.stream_off[a] = wo;
                off[a] = (uint32_t)run;
                mvoff[a] = (uint32_t)(run >> 32);
                if ((c[k] >> 16) & 1u) ilist[1 + irun++] = (uint32_t)a;
                run += (unsigned long long)(c[k] & 255u) | ((unsigned long long)((c[k] >> 8) & 255u) << 32);
            }
            wo += (k < 4 ? wc.x >> (8 * k) : wc.y >> (8 * (k - 4))) & 255u;
        }
    }
    if (threadIdx.x == 0) {
        ilist[0] = (uint32_t)ibase;
        if (pd.n_intra_host >= 0 && ibase > (unsigned long long)pd.n_intra_host) atomicOr(pd.exc_counter + 1, H264R_STATUS_N_INTRA);
        if (wbase != (unsigned long long)pd.stream_words) atomicOr(pd.exc_counter + 1, H264R_STATUS_STREAM);
    }
}

__global__ void __launch_bounds__(1024) k_blk_scan(const DescTable descs, const __grid_constant__ PicList list, int n_mbs)
{
    const PicDesc& pd = descs[list.slot[blockIdx.x]];
    if (pd.stream) { blk_scan_stream(pd, n_mbs); return; }
    const uint32_t* mask = pd.blk_mask;
    uint32_t* off = pd.packed ? const_cast<uint32_t*>(pd.blk_off) : nullptr;
    uint32_t* mvoff = pd.packed ? const_cast<uint32_t*>(pd.mv_off) : nullptr;
    uint32_t* ilist = pd.intra_list;
    const uint2* hdr8 = pd.hdr8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ unsigned long long wsum[32];
    __shared__ uint32_t isum[32];
    __shared__ unsigned long long s_total;
    __shared__ uint32_t s_itotal;
    unsigned long long base = 0;          // counts of the tiles already done
    uint32_t ibase = 0;
    // tiles of 8192 macroblocks: every thread owns 8 consecutive ones, all their loads in flight at once
    for (int t0 = 0; t0 < n_mbs; t0 += 8192) {
        const int first = t0 + threadIdx.x * 8;
        unsigned long long c[8];
        uint32_t intra_bits = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            c[k] = 0;
            if (first + k < n_mbs) {
                if (off) {
                    const uint32_t mw = __ldg(mask + first + k);       // units: one per block, or two for the plain blocks of a compact picture
                    c[k] = (unsigned long long)(__popc(mw & 0xFFFFFFu) << ((pd.level_fmt == H264R_LEVELS_COMPACT && !(mw & H264R_BLK_COMPACT)) ? 1 : 0));
                }
                if (mvoff || ilist || hdr8) {
                    uint2 hw;
                    if (hdr8) {
                        // 8-byte header (h264r_mb_hdr8) -> the 16-byte record of the kernels; the prediction modes of an intra
                        // macroblock follow below, once the scan knows where its units begin
                        const uint2 s8 = __ldg(hdr8 + first + k);
                        hw.x = (s8.x & 7u) | (((s8.x >> 3) & 63u) << 8) | (((s8.x >> 13) & 63u) << 16) | (((s8.x >> 19) & 63u) << 24);
                        hw.y = ((s8.x >> 25) & 63u) | ((s8.y & 63u) << 8) | (((s8.x >> 9) & 15u) << 16) | (((s8.y >> 6) & 255u) << 24);
                        const uint32_t r = s8.y >> 14;
                        const uint32_t refs = (r & 15u) | (((r >> 4) & 15u) << 8) | (((r >> 8) & 15u) << 16) | (((r >> 12) & 15u) << 24);
                        *reinterpret_cast<uint4*>(const_cast<h264r_mb_hdr*>(pd.hdr) + first + k) = uint4{hw.x, hw.y, is_intra((int)(hw.x & 255u)) ? 0u : refs, 0u};
                        if (is_intra((int)(hw.x & 255u))) c[k] += 1ull;       // the unit with its prediction modes
                    } else
                        hw = __ldg(reinterpret_cast<const uint2*>(pd.hdr + first + k));     // mb_class = byte 0, sub_mb_type = byte 7
                    const int cls = (int)(hw.x & 255u);
                    if (mvoff) c[k] += (unsigned long long)mv_count(cls, (int)(hw.y >> 24)) << 32;
                    if (is_intra(cls)) intra_bits |= 1u << k;
                    // headers of picture buffers are the caller's: what the host could not check without a pass over them
                    if (cls == 3 || cls > H264R_MB_P8x8) atomicOr(pd.exc_counter + 1, H264R_STATUS_BAD_CLASS);
                    else if (!is_intra(cls) && pd.slice_type != H264R_SLICE_P) atomicOr(pd.exc_counter + 1, H264R_STATUS_INTER_IN_I);
                }
            }
        }
        unsigned long long sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += c[k];
        const uint32_t isumv = (uint32_t)__popc(intra_bits);
        unsigned long long incl = sum;
        uint32_t iincl = isumv;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            uint32_t ti = __shfl_up_sync(0xffffffffu, iincl, o);
            if (lane >= o) { incl += t; iincl += ti; }
        }
        if (lane == 31) { wsum[warp] = incl; isum[warp] = iincl; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long v = wsum[lane], iv = v;
            uint32_t vi = isum[lane], ivi = vi;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xffffffffu, iv, o);
                uint32_t ti = __shfl_up_sync(0xffffffffu, ivi, o);
                if (lane >= o) { iv += t; ivi += ti; }
            }
            wsum[lane] = iv - v;
            isum[lane] = ivi - vi;
            if (lane == 31) { s_total = iv; s_itotal = ivi; }
        }
        __syncthreads();
        unsigned long long run = base + wsum[warp] + incl - sum;
        uint32_t irun = ibase + isum[warp] + iincl - isumv;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (first + k < n_mbs) {
                const bool ext = hdr8 && ((intra_bits >> k) & 1u);
                if (off) off[first + k] = (uint32_t)run + (ext ? 1u : 0u);
                if (ext)       // u.pred4x4 of the expanded header <- the macroblock's first unit
                    *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(const_cast<h264r_mb_hdr*>(pd.hdr) + first + k) + 8) =
                        __ldg(reinterpret_cast<const uint2*>(pd.coeff) + (uint32_t)run);
                if (mvoff) mvoff[first + k] = (uint32_t)(run >> 32);
                run += c[k];
                if (ilist && ((intra_bits >> k) & 1u)) ilist[1 + irun++] = (uint32_t)(first + k);
            }
        base += s_total;
        ibase += s_itotal;
        __syncthreads();
    }
    if (ilist && threadIdx.x == 0) {
        ilist[0] = ibase;
        if (pd.n_intra_host >= 0 && ibase > (uint32_t)pd.n_intra_host) atomicOr(pd.exc_counter + 1, H264R_STATUS_N_INTRA);      // macroblocks left unreconstructed
    }
}

// The expansion of stream-form pictures (see blk_scan_stream): one thread per macroblock, behind k_blk_scan on the prep stream.
// A thread reads its macroblock's 8-byte header and its words of the stream and writes the 16-byte header, the dense mask, the
// vectors into the partition-order stream and the blocks as 8-byte units, at the offsets the scan left.
__global__ void __launch_bounds__(256) k_stream_expand(const DescTable descs, const __grid_constant__ PicList list, int n_mbs)
{
    const PicDesc& pd = descs[list.slot[blockIdx.y]];
    if (!pd.stream) return;
    const int a = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (a >= n_mbs) return;
    const uint32_t walked = stream_expand_mb(__ldg(pd.hdr8 + a), pd.stream + pd.stream_off[a], __ldg(pd.blk_off + a), __ldg(pd.mv_off + a),
                     reinterpret_cast<uint4*>(const_cast<h264r_mb_hdr*>(pd.hdr) + a), const_cast<uint32_t*>(pd.blk_mask) + a,
                     reinterpret_cast<uint32_t*>(const_cast<int16_t*>(pd.mv)), reinterpret_cast<uint2*>(const_cast<int16_t*>(pd.coeff)));
    if (walked != (uint32_t)__ldg(pd.mb_words + a)) atomicOr(pd.exc_counter + 1, H264R_STATUS_STREAM);
}

// ------------------------------------------------------------------------------------------------
// Border replication of finished pictures (so that later P pictures can stage reference windows without
// coordinate clamps).  One warp per padded row; luma always, chroma when `chroma` (spec mode: the
// reference has no chroma motion compensation, D4).
// ------------------------------------------------------------------------------------------------
__global__ void k_zero_digests(const __grid_constant__ PicList frames, uint32_t* table)
{
    for (int i = threadIdx.x; i < frames.n * 4; i += blockDim.x) table[4 * (size_t)frames.slot[i >> 2] + (i & 3)] = 0;
}

__global__ void __launch_bounds__(128) k_pad(const DescTable descs, const __grid_constant__ PicList list, Geometry g, int chroma,
                                             uint32_t* digest_table)
{
    const int Wl = g.w_mbs * 16, Hl = g.h_mbs * 16, Wc = Wl >> 1, Hc = Hl >> 1;
    const int rows_l = Hl + 2 * PAD_Y, rows_c = chroma ? Hc + 2 * PAD_YC : 0, rows = rows_l + 2 * rows_c;
    const int lane = threadIdx.x & 31;
    const long total = (long)list.n * rows;
    for (long i = (long)blockIdx.x * 4 + (threadIdx.x >> 5); i < total; i += (long)gridDim.x * 4) {
        int p = (int)(i / rows), r = (int)(i - (long)p * rows);
        const PicDesc& pd = descs[list.slot[p]];
        if (digest_table && r == 0 && lane < 4) digest_table[4 * (size_t)pd.frame_id + lane] = 0;     // K5 accumulates into it next
        if (r < rows_l) pad_row_phase(pd.y, g.pitch_y, Wl, Hl, PAD_X, PAD_Y, r - PAD_Y, lane);
        else if (r < rows_l + rows_c) pad_row_phase(pd.u, g.pitch_c, Wc, Hc, PAD_XC, PAD_YC, r - rows_l - PAD_YC, lane);
        else pad_row_phase(pd.v, g.pitch_c, Wc, Hc, PAD_XC, PAD_YC, r - rows_l - rows_c - PAD_YC, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// K5: frame digest (definition: oracle/restate/h264_restate.c:h264o_digest).  Four order-independent sums over
// the words of the three planes: grid = (DIGEST_CTAS, pictures), every CTA folds a band of rows of each plane,
// reduces, and adds its four partial sums to the picture's entry (zeroed by the host side before the launch).
// ------------------------------------------------------------------------------------------------
constexpr int DIGEST_CTAS = 48;
struct FramePool { uint8_t* base; size_t frame_bytes, off_y, off_u, off_v; };
// One plane's share of this CTA: a band of rows.  A thread owns a column of 8-byte pairs and walks down the band eight
// rows at a time, so that every thread keeps eight independent loads in flight (the kernel is pure streaming: what
// bounds it is bytes in flight per SM, not arithmetic).  The planes were written by earlier launches: plain loads.
__device__ __forceinline__ void digest_rows(uint32_t d[4], const uint8_t* plane, int pitch, int words_per_row, int rows, uint32_t first_index)
{
    const int band = (rows + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * band, r1 = min(rows, r0 + band);
    const int pairs = words_per_row >> 1;            // plane widths are multiples of 8 samples
    const int cols = min(pairs, (int)blockDim.x), rpp = blockDim.x / cols;      // thread grid: rpp rows x cols pairs per pass
    const int tr = threadIdx.x / cols, tc = threadIdx.x - tr * cols;
    if (tr >= rpp) return;
    for (int c = tc; c < pairs; c += cols) {
        const uint8_t* col = plane + (size_t)c * 8;
        int r = r0 + tr;
        for (; r + 7 * rpp < r1; r += 8 * rpp) {
            uint2 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = __ldg(reinterpret_cast<const uint2*>(col + (size_t)(r + k * rpp) * pitch));
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t idx = first_index + (uint32_t)(r + k * rpp) * (uint32_t)words_per_row + 2u * (uint32_t)c;
                digest_accumulate(d, v[k].x, idx);
                digest_accumulate(d, v[k].y, idx + 1u);
            }
        }
        for (; r < r1; r += rpp) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(col + (size_t)r * pitch));
            const uint32_t idx = first_index + (uint32_t)r * (uint32_t)words_per_row + 2u * (uint32_t)c;
            digest_accumulate(d, v.x, idx);
            digest_accumulate(d, v.y, idx + 1u);
        }
    }
}
__global__ void __launch_bounds__(256) k_digest(FramePool pool, const __grid_constant__ PicList frames, Geometry g, uint32_t* out)
{
    const uint8_t* f = pool.base + (size_t)frames.slot[blockIdx.y] * pool.frame_bytes;
    const int Wl = g.w_mbs * 16, Hl = g.h_mbs * 16, Wc = Wl / 2, Hc = Hl / 2;
    uint32_t d[4] = {0, 0, 0, 0};
    digest_rows(d, f + pool.off_y, g.pitch_y, Wl / 4, Hl, 0u);
    digest_rows(d, f + pool.off_u, g.pitch_c, Wc / 4, Hc, (uint32_t)(Wl / 4 * Hl));
    digest_rows(d, f + pool.off_v, g.pitch_c, Wc / 4, Hc, (uint32_t)(Wl / 4 * Hl + Wc / 4 * Hc));
    __shared__ uint32_t part[8][4];
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 4; k++) part[threadIdx.x >> 5][k] = d[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        uint32_t s = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) s += part[wv][threadIdx.x];
        atomicAdd(out + 4 * (size_t)frames.slot[blockIdx.y] + threadIdx.x, s);       // table indexed by frame id
    }
}

}  // namespace h264r
