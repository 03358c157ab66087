// h264r_kernels.cuh -- the sm_100a kernels: orchestration (work distribution, wavefront
// dependencies, synchronisation) around the per-macroblock phase sequences of h264r_seq.cuh.
#pragma once
#include "h264r_seq.cuh"

namespace h264r {

struct WarpExec {
    int lane;
    template <class F> __device__ __forceinline__ void operator()(F f)
    {
        f(lane);
        __syncwarp();
    }
};

constexpr int INTER_WARPS = 4;       // warps per CTA of the inter / residual kernels
constexpr int WAVE_WARPS = 32;       // warps per CTA of the wavefront kernels (one CTA per picture)

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
// references are complete pictures), so this is a flat grid: one warp per MB, grid-stride over
// (picture, MB) pairs.
// ------------------------------------------------------------------------------------------------
template <bool COMPAT>
__global__ void __launch_bounds__(INTER_WARPS * 32) k_inter(const PicDesc* __restrict__ descs, const int* __restrict__ list,
                                                            int n_pics, Geometry g, uint32_t* excursions)
{
    __shared__ WarpWork work[INTER_WARPS];
    int warp = threadIdx.x >> 5;
    WarpExec ex{(int)(threadIdx.x & 31)};
    WarpWork& w = work[warp];
    const int n_mbs = g.w_mbs * g.h_mbs;
    const long total = (long)n_pics * n_mbs;
    for (long i = (long)blockIdx.x * INTER_WARPS + warp; i < total; i += (long)gridDim.x * INTER_WARPS) {
        int p = (int)(i / n_mbs), a = (int)(i - (long)p * n_mbs);
        const PicDesc& pd = descs[list[p]];
        h264r_mb_hdr h = load_hdr(pd, a);
        if (is_intra(h.mb_class)) continue;       // warp-uniform
        uint32_t exc = seq_inter_any<COMPAT>(ex, w, pd, g, h, a);
        if (ex.lane == 0 && exc) atomicAdd(excursions, exc);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Wavefront scheduling shared by K2 (intra) and K4 (deblock): one CTA per picture.  Macroblocks
// are handed to warps in anti-diagonal order (d = x + 2y, then y), so a macroblock's neighbours
// A (d-1), B (d-2), C (d-1), D (d-3) were handed out earlier; a warp waits for them on per-MB
// flags in shared memory.  All warps of a CTA are resident, and every dependency points to an
// earlier ticket, so the spin-wait cannot deadlock.
// ------------------------------------------------------------------------------------------------
// ticket -> macroblock address in anti-diagonal order
__device__ __forceinline__ int wave_mb(int ticket, int W, int H, const int* __restrict__ diag_start, int n_diag)
{
    // binary search for the diagonal that contains the ticket
    int lo = 0, hi = n_diag - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (diag_start[mid] <= ticket) lo = mid; else hi = mid - 1;
    }
    int d = lo, k = ticket - diag_start[d];
    int ymin = d - (W - 1) <= 0 ? 0 : (d - (W - 1) + 1) >> 1;
    int y = ymin + k, x = d - 2 * y;
    return y * W + x;
}

__device__ __forceinline__ void wait_flag(volatile uint8_t* f)
{
    while (!*f) __nanosleep(32);
}

template <bool COMPAT>
__global__ void __launch_bounds__(WAVE_WARPS * 32) k_intra(const PicDesc* __restrict__ descs, const int* __restrict__ list,
                                                          Geometry g, const int* __restrict__ diag_start, int n_diag,
                                                          uint32_t* excursions)
{
    extern __shared__ __align__(16) unsigned char smem[];
    MbWork* work = reinterpret_cast<MbWork*>(smem);
    volatile uint8_t* done = smem + sizeof(MbWork) * WAVE_WARPS;
    __shared__ int ticket_counter;
    const PicDesc& pd = descs[list[blockIdx.x]];
    const int W = g.w_mbs, H = g.h_mbs, n_mbs = W * H;
    // inter macroblocks were reconstructed by k_inter before this launch
    for (int a = threadIdx.x; a < n_mbs; a += blockDim.x) done[a] = !is_intra(H264R_LDG(&pd.hdr[a].mb_class));
    if (threadIdx.x == 0) ticket_counter = 0;
    __syncthreads();
    int warp = threadIdx.x >> 5;
    WarpExec ex{(int)(threadIdx.x & 31)};
    MbWork& w = work[warp];
    for (;;) {
        int t = 0;
        if (ex.lane == 0) t = atomicAdd(&ticket_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_mbs) break;
        int a = wave_mb(t, W, H, diag_start, n_diag);
        if (done[a]) continue;                    // not an intra MB
        h264r_mb_hdr h = load_hdr(pd, a);
        if (h.flags & H264R_AVAIL_A) wait_flag(&done[a - 1]);
        if (h.flags & H264R_AVAIL_B) wait_flag(&done[a - W]);
        if (h.flags & H264R_AVAIL_C) wait_flag(&done[a - W + 1]);
        if (h.flags & H264R_AVAIL_D) wait_flag(&done[a - W - 1]);
        __threadfence();
        seq_intra_mb<COMPAT>(ex, w, pd, g, h, a);
        if (ex.lane == 0 && w.excursions) atomicAdd(excursions, w.excursions);
        __threadfence();                          // samples visible before the flag
        __syncwarp();
        if (ex.lane == 0) done[a] = 1;
    }
}

__global__ void __launch_bounds__(WAVE_WARPS * 32) k_deblock(const PicDesc* __restrict__ descs, const int* __restrict__ list,
                                                            Geometry g, const int* __restrict__ diag_start, int n_diag)
{
    extern __shared__ __align__(16) unsigned char smem[];
    MbWork* work = reinterpret_cast<MbWork*>(smem);
    volatile uint8_t* done = smem + sizeof(MbWork) * WAVE_WARPS;
    __shared__ int ticket_counter;
    const PicDesc& pd = descs[list[blockIdx.x]];
    const int W = g.w_mbs, H = g.h_mbs, n_mbs = W * H;
    for (int a = threadIdx.x; a < n_mbs; a += blockDim.x) done[a] = 0;
    if (threadIdx.x == 0) ticket_counter = 0;
    __syncthreads();
    int warp = threadIdx.x >> 5;
    WarpExec ex{(int)(threadIdx.x & 31)};
    MbWork& w = work[warp];
    for (;;) {
        int t = 0;
        if (ex.lane == 0) t = atomicAdd(&ticket_counter, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_mbs) break;
        int a = wave_mb(t, W, H, diag_start, n_diag);
        int mbx = a % W, mby = a / W;
        if (mbx > 0) wait_flag(&done[a - 1]);
        if (mby > 0) wait_flag(&done[a - W]);
        if (mby > 0 && mbx < W - 1) wait_flag(&done[a - W + 1]);
        __threadfence();
        seq_deblock_mb(ex, w, pd, g, a);
        __threadfence();
        __syncwarp();
        if (ex.lane == 0) done[a] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: frame digest.  One CTA per picture; every thread folds 4-sample words, block reduction, the
// CTA writes the four sums (no zeroing pass, no atomics on global memory).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_digest(const PicDesc* __restrict__ descs, const int* __restrict__ list, Geometry g)
{
    const PicDesc& pd = descs[list[blockIdx.x]];
    const int Wl = g.w_mbs * 16, Hl = g.h_mbs * 16, Wc = Wl / 2, Hc = Hl / 2;
    uint32_t d[4] = {0, 0, 0, 0};
    const int lw = Wl / 4, cw = Wc / 4;
    const int n_l = lw * Hl, n_c = cw * Hc;
    for (int i = threadIdx.x; i < n_l + 2 * n_c; i += blockDim.x) {
        const uint8_t* base; int pitch, wpr, k; uint32_t first;
        if (i < n_l) { base = pd.y; pitch = g.pitch_y; wpr = lw; k = i; first = 0; }
        else if (i < n_l + n_c) { base = pd.u; pitch = g.pitch_c; wpr = cw; k = i - n_l; first = (uint32_t)(Wl * Hl); }
        else { base = pd.v; pitch = g.pitch_c; wpr = cw; k = i - n_l - n_c; first = (uint32_t)(Wl * Hl + Wc * Hc); }
        int r = k / wpr, c = k - r * wpr;
        uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(base + (size_t)r * pitch) + c);
        uint32_t idx = first + (uint32_t)(r * wpr + c) * 4u;
        digest_accumulate(d, v & 255u, idx);
        digest_accumulate(d, (v >> 8) & 255u, idx + 1);
        digest_accumulate(d, (v >> 16) & 255u, idx + 2);
        digest_accumulate(d, v >> 24, idx + 3);
    }
    __shared__ uint32_t part[32][4];
    for (int k = 0; k < 4; k++)
        for (int o = 16; o > 0; o >>= 1) d[k] += __shfl_xor_sync(0xffffffffu, d[k], o);
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 4; k++) part[threadIdx.x >> 5][k] = d[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        uint32_t s = 0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) s += part[wv][threadIdx.x];
        pd.digest[threadIdx.x] = s;
    }
}

}  // namespace h264r
