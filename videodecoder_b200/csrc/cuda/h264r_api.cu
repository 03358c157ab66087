// h264r_api.cu -- the C ABI of include/h264r.h: context, device frame pool, syntax arena,
// dependency-wave scheduler and kernel launches.  There is no CPU path: every entry point that
// computes anything launches the sm_100a kernels of h264r_kernels.cuh.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "h264r_kernels.cuh"

using namespace h264r;

namespace {

enum FrameState : uint8_t { FRAME_FREE = 0, FRAME_OPEN = 1, FRAME_QUEUED = 2, FRAME_DONE = 3 };

struct Pending {
    int frame_id;
    int slot;               // syntax arena slot
    h264r_pic_params pp;
    bool ended;
    int n_intra;            // intra macroblocks seen by submit
    int n_submitted;
    int wave;
};

}  // namespace

struct h264r_ctx {
    int device = 0;
    Geometry g{};
    int n_mbs = 0, max_frames = 0, max_pending = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_ctrl = nullptr;
    bool ctrl_in_flight = false;
    bool profiling = false;
    std::vector<cudaEvent_t> kev;           // profiling events: pairs per launch
    std::vector<int> kev_family;
    int kev_used = 0;
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
    // frame pool: per slot Y, U, V
    uint8_t* d_frames = nullptr;
    size_t y_bytes = 0, c_bytes = 0, frame_bytes = 0;
    uint32_t* d_digests = nullptr;          // staging for K5 results: max_pending * 4
    uint8_t* d_exc_maps = nullptr;          // max_frames * n_mbs
    uint32_t* d_exc_counter = nullptr;
    // syntax arena: per slot hdr | mv | coeff
    uint8_t* d_syntax = nullptr;
    size_t slot_bytes = 0, off_mv = 0, off_coeff = 0;
    // per-flush control block (descriptors + picture lists), pinned host mirror + device copy
    uint8_t* h_ctrl = nullptr;
    uint8_t* d_ctrl = nullptr;
    size_t ctrl_bytes = 0;
    // wavefront hand-over words, zeroed at the start of every flush:
    // progress words indexed by (position in the flush's picture lists, macroblock row): 4*max_pending*H, then 2*max_pending tickets
    int* d_sync = nullptr;
    size_t sync_ints = 0;
    std::vector<uint8_t> frame_state;
    std::vector<int> frame_wave;            // wave of a queued frame inside the current flush
    std::vector<Pending> pending;
    std::vector<int> free_slots;
    bool timing_valid = false;
    int last_launches = 0;
    std::string err;

    // sample (0,0) of each plane; every plane is surrounded by a replicated border (PAD_* of h264r_mc.cuh)
    uint8_t* frame_y(int f) const { return d_frames + (size_t)f * frame_bytes + (size_t)PAD_Y * g.pitch_y + PAD_X; }
    uint8_t* frame_u(int f) const { return d_frames + (size_t)f * frame_bytes + y_bytes + (size_t)PAD_YC * g.pitch_c + PAD_XC; }
    uint8_t* frame_v(int f) const { return frame_u(f) + c_bytes; }
    uint8_t* slot_ptr(int s) const { return d_syntax + (size_t)s * slot_bytes; }
};

namespace {

thread_local std::string g_create_error;

int fail(h264r_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? H264R_E_NOMEM : H264R_E_CUDA,                  \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                                  \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Pending* find_pending(h264r_ctx* c, int frame_id)
{
    for (auto& p : c->pending) if (p.frame_id == frame_id) return &p;
    return nullptr;
}

void fill_desc(h264r_ctx* c, const Pending& p, PicDesc& pd)
{
    memset(&pd, 0, sizeof(pd));
    uint8_t* s = c->slot_ptr(p.slot);
    pd.hdr = reinterpret_cast<const h264r_mb_hdr*>(s);
    pd.mv = reinterpret_cast<const int16_t*>(s + c->off_mv);
    pd.coeff = reinterpret_cast<const int16_t*>(s + c->off_coeff);
    pd.y = c->frame_y(p.frame_id); pd.u = c->frame_u(p.frame_id); pd.v = c->frame_v(p.frame_id);
    for (int i = 0; i < H264R_MAX_REFS; i++) {
        int r = i < p.pp.num_ref_idx_l0 ? p.pp.ref_list0[i] : -1;
        if (r >= 0 && r < c->max_frames) {
            pd.ref_y[i] = c->frame_y(r); pd.ref_u[i] = c->frame_u(r); pd.ref_v[i] = c->frame_v(r);
        } else {
            // unused list entries point at the picture's first reference so that a stray index stays in bounds
            int r0 = p.pp.num_ref_idx_l0 > 0 ? p.pp.ref_list0[0] : p.frame_id;
            pd.ref_y[i] = c->frame_y(r0); pd.ref_u[i] = c->frame_u(r0); pd.ref_v[i] = c->frame_v(r0);
        }
        pd.ref_id[i] = r;
        pd.luma_weight[i] = p.pp.luma_weight[i]; pd.luma_offset[i] = p.pp.luma_offset[i];
        for (int k = 0; k < 2; k++) { pd.chroma_weight[i][k] = p.pp.chroma_weight[i][k]; pd.chroma_offset[i][k] = p.pp.chroma_offset[i][k]; }
    }
    pd.slice_type = p.pp.slice_type;
    pd.weighted_pred = p.pp.weighted_pred;
    pd.luma_log2_wd = p.pp.luma_log2_weight_denom;
    pd.chroma_log2_wd = p.pp.chroma_log2_weight_denom;
    pd.deblock_on = !(c->flags & H264R_REF_COMPAT) && p.pp.disable_deblocking_filter_idc != 1;
    pd.alpha_off = p.pp.slice_alpha_c0_offset_div2 * 2;
    pd.beta_off = p.pp.slice_beta_offset_div2 * 2;
    pd.excursion_map = c->d_exc_maps + (size_t)p.frame_id * c->n_mbs;
    pd.digest = nullptr;
    pd.frame_id = p.frame_id;
}

}  // namespace

extern "C" {

int h264r_abi_version(void) { return H264R_ABI_VERSION; }

const char* h264r_last_error(h264r_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int h264r_create(h264r_ctx** out, int device, int w_mbs, int h_mbs, int max_frames, int max_pending, uint32_t flags)
{
    h264r_ctx* ctx = nullptr;      // errors before the context exists go to the thread-local string
    if (!out || w_mbs <= 0 || h_mbs <= 0 || max_frames <= 0 || max_pending <= 0 || (flags & ~H264R_REF_COMPAT))
        return fail(nullptr, H264R_E_INVAL, "h264r_create: bad argument");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev)
        return fail(nullptr, H264R_E_NODEVICE, "h264r_create: no usable CUDA device (this engine has no CPU path)");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, H264R_E_NODEVICE, "h264r_create: device is not sm_100 class; kernels are built for sm_100a only");
    h264r_ctx* c = new h264r_ctx();
    ctx = c;
    c->device = device;
    c->flags = flags;
    c->g.w_mbs = w_mbs; c->g.h_mbs = h_mbs;
    c->g.pitch_y = (int)align_up((size_t)w_mbs * 16 + 2 * PAD_X, 128);
    c->g.pitch_c = (int)align_up((size_t)w_mbs * 8 + 2 * PAD_XC, 128);
    c->n_mbs = w_mbs * h_mbs;
    c->max_frames = max_frames; c->max_pending = max_pending;
    c->y_bytes = (size_t)c->g.pitch_y * (h_mbs * 16 + 2 * PAD_Y);
    c->c_bytes = (size_t)c->g.pitch_c * (h_mbs * 8 + 2 * PAD_YC);
    c->frame_bytes = align_up(c->y_bytes + 2 * c->c_bytes, 256);
    c->off_mv = align_up((size_t)c->n_mbs * sizeof(h264r_mb_hdr), 256);
    c->off_coeff = c->off_mv + align_up((size_t)c->n_mbs * H264R_MV_PER_MB * 2, 256);
    c->slot_bytes = c->off_coeff + align_up((size_t)c->n_mbs * H264R_COEFF_PER_MB * 2, 256);
    c->ctrl_bytes = align_up((size_t)max_pending * sizeof(PicDesc), 256) + (size_t)max_pending * 4 * sizeof(int) + 256;
    auto cleanup = [&](int code) { h264r_destroy(c); return code; };
#define CUC(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            fail(nullptr, 0, std::string(#call) + ": " + cudaGetErrorString(e_));                             \
            return cleanup(e_ == cudaErrorMemoryAllocation ? H264R_E_NOMEM : H264R_E_CUDA);                   \
        }                                                                                                     \
    } while (0)
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&c->ev0));
    CUC(cudaEventCreate(&c->ev1));
    CUC(cudaEventCreateWithFlags(&c->ev_ctrl, cudaEventDisableTiming));
    CUC(cudaMalloc(&c->d_frames, c->frame_bytes * max_frames));
    CUC(cudaMalloc(&c->d_digests, (size_t)max_pending * 16));
    CUC(cudaMalloc(&c->d_exc_maps, (size_t)max_frames * c->n_mbs));
    CUC(cudaMalloc(&c->d_exc_counter, 256));
    CUC(cudaMalloc(&c->d_syntax, c->slot_bytes * max_pending));
    CUC(cudaMalloc(&c->d_ctrl, c->ctrl_bytes));
    CUC(cudaMallocHost(&c->h_ctrl, c->ctrl_bytes));
    CUC(cudaMemsetAsync(c->d_exc_counter, 0, 256, c->stream));
    CUC(cudaMemsetAsync(c->d_exc_maps, 0, (size_t)max_frames * c->n_mbs, c->stream));
    c->sync_ints = (size_t)4 * max_pending * h_mbs + (size_t)2 * max_pending;
    CUC(cudaMalloc(&c->d_sync, c->sync_ints * sizeof(int)));
    CUC(cudaStreamSynchronize(c->stream));
    c->frame_state.assign(max_frames, FRAME_FREE);
    c->frame_wave.assign(max_frames, -1);
    for (int s = max_pending - 1; s >= 0; s--) c->free_slots.push_back(s);
    *out = c;
    return H264R_OK;
#undef CUC
}

void h264r_destroy(h264r_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_frames); cudaFree(c->d_digests); cudaFree(c->d_exc_maps); cudaFree(c->d_exc_counter);
    cudaFree(c->d_syntax); cudaFree(c->d_ctrl); cudaFree(c->d_sync);
    if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_ctrl) cudaEventDestroy(c->ev_ctrl);
    for (auto e : c->kev) cudaEventDestroy(e);
    for (auto e : c->marks) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* h264r_host_alloc(h264r_ctx* ctx, size_t bytes)
{
    if (!ctx) return nullptr;
    void* p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { ctx->err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void h264r_host_free(h264r_ctx* ctx, void* p)
{
    if (ctx && p) { cudaSetDevice(ctx->device); cudaFreeHost(p); }
}

int h264r_frame_begin(h264r_ctx* ctx, int frame_id, const h264r_pic_params* pp)
{
    if (!ctx || !pp || frame_id < 0 || frame_id >= ctx->max_frames) return fail(ctx, H264R_E_INVAL, "frame_begin: bad argument");
    if (pp->slice_type != H264R_SLICE_I && pp->slice_type != H264R_SLICE_P) return fail(ctx, H264R_E_INVAL, "frame_begin: slice_type must be I or P");
    if (pp->num_ref_idx_l0 < 0 || pp->num_ref_idx_l0 > H264R_MAX_REFS) return fail(ctx, H264R_E_INVAL, "frame_begin: num_ref_idx_l0 out of range");
    if (pp->slice_type == H264R_SLICE_P && pp->num_ref_idx_l0 < 1) return fail(ctx, H264R_E_INVAL, "frame_begin: P picture without references");
    uint8_t st = ctx->frame_state[frame_id];
    if (st == FRAME_OPEN || st == FRAME_QUEUED) return fail(ctx, H264R_E_STATE, "frame_begin: frame slot is still being produced");
    for (int i = 0; i < pp->num_ref_idx_l0; i++) {
        int r = pp->ref_list0[i];
        if (r < 0 || r >= ctx->max_frames || r == frame_id) return fail(ctx, H264R_E_INVAL, "frame_begin: bad reference frame id");
        uint8_t rs = ctx->frame_state[r];
        if (rs != FRAME_DONE && rs != FRAME_QUEUED) return fail(ctx, H264R_E_STATE, "frame_begin: reference frame is not reconstructed or queued");
    }
    if (ctx->free_slots.empty()) return fail(ctx, H264R_E_CAPACITY, "frame_begin: max_pending pictures already queued; call h264r_flush");
    Pending p{};
    p.frame_id = frame_id;
    p.slot = ctx->free_slots.back();
    ctx->free_slots.pop_back();
    p.pp = *pp;
    ctx->pending.push_back(p);
    ctx->frame_state[frame_id] = FRAME_OPEN;
    return H264R_OK;
}

int h264r_submit_mbs(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv,
                     const int16_t* coeff)
{
    if (!ctx || !hdr || !coeff || n_mbs <= 0 || first_mb < 0) return fail(ctx, H264R_E_INVAL, "submit_mbs: bad argument");
    Pending* p = find_pending(ctx, frame_id);
    if (!p || p->ended) return fail(ctx, H264R_E_STATE, "submit_mbs: no open picture with this frame id");
    if (first_mb + n_mbs > ctx->n_mbs) return fail(ctx, H264R_E_INVAL, "submit_mbs: macroblock range exceeds the picture");
    if (p->pp.slice_type == H264R_SLICE_P && !mv) return fail(ctx, H264R_E_INVAL, "submit_mbs: P picture needs motion vectors");
    for (int i = 0; i < n_mbs; i++) {
        uint8_t cls = hdr[i].mb_class;
        if (cls == 3 || cls > H264R_MB_P8x8) return fail(ctx, H264R_E_INVAL, "submit_mbs: unknown mb_class");
        if (cls <= H264R_MB_I16x16) p->n_intra++;
        else if (p->pp.slice_type != H264R_SLICE_P) return fail(ctx, H264R_E_INVAL, "submit_mbs: inter macroblock in an I picture");
        if ((ctx->flags & H264R_REF_COMPAT) && (cls == H264R_MB_I8x8 || (hdr[i].flags & H264R_T8x8)))
            return fail(ctx, H264R_E_INVAL, "submit_mbs: 8x8 transform is not part of REF_COMPAT (the reference has none)");
    }
    CU(cudaSetDevice(ctx->device));
    uint8_t* s = ctx->slot_ptr(p->slot);
    CU(cudaMemcpyAsync(s + (size_t)first_mb * sizeof(h264r_mb_hdr), hdr, (size_t)n_mbs * sizeof(h264r_mb_hdr), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s + ctx->off_coeff + (size_t)first_mb * H264R_COEFF_PER_MB * 2, coeff, (size_t)n_mbs * H264R_COEFF_PER_MB * 2,
                       cudaMemcpyHostToDevice, ctx->stream));
    if (mv && p->pp.slice_type == H264R_SLICE_P)
        CU(cudaMemcpyAsync(s + ctx->off_mv + (size_t)first_mb * H264R_MV_PER_MB * 2, mv, (size_t)n_mbs * H264R_MV_PER_MB * 2,
                           cudaMemcpyHostToDevice, ctx->stream));
    p->n_submitted += n_mbs;
    return H264R_OK;
}

int h264r_frame_end(h264r_ctx* ctx, int frame_id)
{
    if (!ctx) return H264R_E_INVAL;
    Pending* p = find_pending(ctx, frame_id);
    if (!p || p->ended) return fail(ctx, H264R_E_STATE, "frame_end: no open picture with this frame id");
    if (p->n_submitted != ctx->n_mbs) return fail(ctx, H264R_E_STATE, "frame_end: picture is missing macroblocks");
    p->ended = true;
    ctx->frame_state[frame_id] = FRAME_QUEUED;
    return H264R_OK;
}

int h264r_flush(h264r_ctx* ctx)
{
    if (!ctx) return H264R_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    // pictures whose frame_end has been called, in submission order
    std::vector<Pending*> run;
    for (auto& p : ctx->pending) if (p.ended) run.push_back(&p);
    if (run.empty()) return H264R_OK;       // keeps the statistics of the last real flush
    ctx->last_launches = 0;
    ctx->timing_valid = false;
    const bool compat = (ctx->flags & H264R_REF_COMPAT) != 0;
    // dependency waves
    int n_waves = 0;
    for (Pending* p : run) {
        int wv = 0;
        for (int i = 0; i < p->pp.num_ref_idx_l0; i++) {
            int r = p->pp.ref_list0[i];
            if (ctx->frame_state[r] == FRAME_QUEUED && ctx->frame_wave[r] >= 0 && ctx->frame_wave[r] + 1 > wv) wv = ctx->frame_wave[r] + 1;
        }
        p->wave = wv;
        ctx->frame_wave[p->frame_id] = wv;
        if (wv + 1 > n_waves) n_waves = wv + 1;
    }
    // control block: descriptors, then per wave three lists (inter / intra / deblock)
    const int n = (int)run.size();
    PicDesc* h_desc = reinterpret_cast<PicDesc*>(ctx->h_ctrl);
    size_t list_off = align_up((size_t)ctx->max_pending * sizeof(PicDesc), 256);
    int* h_lists = reinterpret_cast<int*>(ctx->h_ctrl + list_off);
    const PicDesc* d_desc = reinterpret_cast<const PicDesc*>(ctx->d_ctrl);
    const int* d_lists = reinterpret_cast<const int*>(ctx->d_ctrl + list_off);
    if (ctx->ctrl_in_flight) {               // the previous flush's upload of the pinned control block
        CU(cudaEventSynchronize(ctx->ev_ctrl));
        ctx->ctrl_in_flight = false;
    }
    for (int i = 0; i < n; i++) fill_desc(ctx, *run[i], h_desc[i]);
    struct WaveLists { int inter_off, n_inter, intra_off, n_intra, deb_off, n_deb, all_off, n_all; };
    std::vector<WaveLists> wl(n_waves);
    int pos = 0;
    for (int wv = 0; wv < n_waves; wv++) {
        wl[wv].inter_off = pos;
        for (int i = 0; i < n; i++) if (run[i]->wave == wv && run[i]->pp.slice_type == H264R_SLICE_P && run[i]->n_intra < ctx->n_mbs) h_lists[pos++] = i;
        wl[wv].n_inter = pos - wl[wv].inter_off;
        wl[wv].intra_off = pos;
        for (int i = 0; i < n; i++) if (run[i]->wave == wv && run[i]->n_intra > 0) h_lists[pos++] = i;
        wl[wv].n_intra = pos - wl[wv].intra_off;
        wl[wv].deb_off = pos;
        for (int i = 0; i < n; i++) if (run[i]->wave == wv && h_desc[i].deblock_on) h_lists[pos++] = i;
        wl[wv].n_deb = pos - wl[wv].deb_off;
        wl[wv].all_off = pos;
        for (int i = 0; i < n; i++) if (run[i]->wave == wv) h_lists[pos++] = i;
        wl[wv].n_all = pos - wl[wv].all_off;
    }
    CU(cudaMemcpyAsync(ctx->d_ctrl, ctx->h_ctrl, list_off + (size_t)pos * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(ctx->ev_ctrl, ctx->stream));
    ctx->ctrl_in_flight = true;
    CU(cudaMemsetAsync(ctx->d_sync, 0, ctx->sync_ints * sizeof(int), ctx->stream));
    int* tickets = ctx->d_sync + (size_t)4 * ctx->max_pending * ctx->g.h_mbs;
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    ctx->kev_used = 0;
    ctx->kev_family.clear();
    auto prof_begin = [&](int family) {
        if (!ctx->profiling) return;
        if ((int)ctx->kev.size() < ctx->kev_used + 2) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            ctx->kev.push_back(a); ctx->kev.push_back(b);
        }
        ctx->kev_family.push_back(family);
        cudaEventRecord(ctx->kev[ctx->kev_used], ctx->stream);
    };
    auto prof_end = [&]() {
        if (!ctx->profiling) return;
        cudaEventRecord(ctx->kev[ctx->kev_used + 1], ctx->stream);
        ctx->kev_used += 2;
    };
    for (int wv = 0; wv < n_waves; wv++) {
        const WaveLists& L = wl[wv];
        if (L.n_inter) {
            long total = (long)L.n_inter * ctx->n_mbs;
            long want = (total + INTER_WARPS - 1) / INTER_WARPS;
            long cap = (long)sm_count * 16 * 4;                   // 16 resident CTAs per SM, four passes
            int blocks = (int)(want < cap ? want : cap);
            prof_begin(0);
            if (compat) k_inter<true><<<blocks, INTER_WARPS * 32, 0, ctx->stream>>>(d_desc, d_lists + L.inter_off, L.n_inter, ctx->g, ctx->d_exc_counter);
            else k_inter<false><<<blocks, INTER_WARPS * 32, 0, ctx->stream>>>(d_desc, d_lists + L.inter_off, L.n_inter, ctx->g, ctx->d_exc_counter);
            prof_end();
            ctx->last_launches++;
        }
        if (L.n_intra) {
            int rows = L.n_intra * ctx->g.h_mbs, blocks = (rows + ROW_WARPS - 1) / ROW_WARPS;
            int* prog = ctx->d_sync + (size_t)L.intra_off * ctx->g.h_mbs;
            prof_begin(1);
            if (compat) k_intra_rows<true><<<blocks, ROW_WARPS * 32, 0, ctx->stream>>>(d_desc, d_lists + L.intra_off, L.n_intra, ctx->g, prog, tickets + 2 * wv, ctx->d_exc_counter);
            else k_intra_rows<false><<<blocks, ROW_WARPS * 32, 0, ctx->stream>>>(d_desc, d_lists + L.intra_off, L.n_intra, ctx->g, prog, tickets + 2 * wv, ctx->d_exc_counter);
            prof_end();
            ctx->last_launches++;
        }
        if (L.n_deb) {
            int rows = L.n_deb * ctx->g.h_mbs, blocks = (rows + ROW_WARPS - 1) / ROW_WARPS;
            int* prog = ctx->d_sync + (size_t)L.deb_off * ctx->g.h_mbs;
            prof_begin(2);
            k_deblock_rows<<<blocks, ROW_WARPS * 32, 0, ctx->stream>>>(d_desc, d_lists + L.deb_off, L.n_deb, ctx->g, prog, tickets + 2 * wv + 1);
            prof_end();
            ctx->last_launches++;
        }
        {
            // border replication of the finished pictures of this wave
            int rows = ctx->g.h_mbs * 16 + 2 * PAD_Y + (compat ? 0 : 2 * (ctx->g.h_mbs * 8 + 2 * PAD_YC));
            long warps = (long)L.n_all * rows;
            long cap = (long)sm_count * 16;
            int blocks = (int)((warps + 3) / 4 < cap ? (warps + 3) / 4 : cap);
            prof_begin(3);
            k_pad<<<blocks, 128, 0, ctx->stream>>>(d_desc, d_lists + L.all_off, L.n_all, ctx->g, compat ? 0 : 1);
            prof_end();
            ctx->last_launches++;
        }
    }
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaGetLastError());
    ctx->timing_valid = true;
    std::vector<Pending> keep;
    for (auto& p : ctx->pending) {
        if (p.ended) {
            ctx->frame_state[p.frame_id] = FRAME_DONE;
            ctx->frame_wave[p.frame_id] = -1;
            ctx->free_slots.push_back(p.slot);
        } else keep.push_back(p);
    }
    ctx->pending.swap(keep);
    return H264R_OK;
}

int h264r_sync(h264r_ctx* ctx)
{
    if (!ctx) return H264R_E_INVAL;
    int r = h264r_flush(ctx);
    if (r != H264R_OK) return r;
    CU(cudaStreamSynchronize(ctx->stream));
    return H264R_OK;
}

static int require_done(h264r_ctx* ctx, int frame_id, const char* who)
{
    if (!ctx || frame_id < 0 || frame_id >= ctx->max_frames) return fail(ctx, H264R_E_INVAL, std::string(who) + ": bad frame id");
    if (ctx->frame_state[frame_id] == FRAME_QUEUED) {
        int r = h264r_flush(ctx);
        if (r != H264R_OK) return r;
    }
    if (ctx->frame_state[frame_id] != FRAME_DONE) return fail(ctx, H264R_E_STATE, std::string(who) + ": frame is not reconstructed");
    return H264R_OK;
}

int h264r_read_planes(h264r_ctx* ctx, int frame_id, uint8_t* y, int y_stride, uint8_t* u, uint8_t* v, int c_stride)
{
    int r = require_done(ctx, frame_id, "read_planes");
    if (r != H264R_OK) return r;
    CU(cudaSetDevice(ctx->device));
    int W = ctx->g.w_mbs * 16, H = ctx->g.h_mbs * 16;
    if (y) CU(cudaMemcpy2DAsync(y, y_stride, ctx->frame_y(frame_id), ctx->g.pitch_y, W, H, cudaMemcpyDeviceToHost, ctx->stream));
    if (u) CU(cudaMemcpy2DAsync(u, c_stride, ctx->frame_u(frame_id), ctx->g.pitch_c, W / 2, H / 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (v) CU(cudaMemcpy2DAsync(v, c_stride, ctx->frame_v(frame_id), ctx->g.pitch_c, W / 2, H / 2, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return H264R_OK;
}

int h264r_frame_digests(h264r_ctx* ctx, const int* frame_ids, int n, uint8_t* out)
{
    if (!ctx || !frame_ids || !out || n <= 0) return fail(ctx, H264R_E_INVAL, "frame_digests: bad argument");
    for (int i = 0; i < n; i++) {
        int r = require_done(ctx, frame_ids[i], "frame_digests");
        if (r != H264R_OK) return r;
    }
    CU(cudaSetDevice(ctx->device));
    // K5 runs on demand: minimal descriptors (planes + digest slot) in chunks of max_pending
    PicDesc* h_desc = reinterpret_cast<PicDesc*>(ctx->h_ctrl);
    size_t list_off = align_up((size_t)ctx->max_pending * sizeof(PicDesc), 256);
    int* h_lists = reinterpret_cast<int*>(ctx->h_ctrl + list_off);
    for (int base = 0; base < n; base += ctx->max_pending) {
        int m = n - base < ctx->max_pending ? n - base : ctx->max_pending;
        CU(cudaStreamSynchronize(ctx->stream));      // the pinned control block may still be in flight
        ctx->ctrl_in_flight = false;
        for (int i = 0; i < m; i++) {
            int f = frame_ids[base + i];
            memset(&h_desc[i], 0, sizeof(PicDesc));
            h_desc[i].y = ctx->frame_y(f); h_desc[i].u = ctx->frame_u(f); h_desc[i].v = ctx->frame_v(f);
            h_desc[i].digest = ctx->d_digests + (size_t)i * 4;
            h_lists[i] = i;
        }
        CU(cudaMemcpyAsync(ctx->d_ctrl, ctx->h_ctrl, list_off + (size_t)m * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        k_digest<<<m, 1024, 0, ctx->stream>>>(reinterpret_cast<const PicDesc*>(ctx->d_ctrl),
                                              reinterpret_cast<const int*>(ctx->d_ctrl + list_off), ctx->g);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + (size_t)base * 16, ctx->d_digests, (size_t)m * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return H264R_OK;
}

int h264r_frame_digest(h264r_ctx* ctx, int frame_id, uint8_t out[16]) { return h264r_frame_digests(ctx, &frame_id, 1, out); }

int h264r_range_excursions(h264r_ctx* ctx, uint64_t* count)
{
    if (!ctx || !count) return H264R_E_INVAL;
    int r = h264r_flush(ctx);
    if (r != H264R_OK) return r;
    CU(cudaSetDevice(ctx->device));
    uint32_t v = 0;
    CU(cudaMemcpyAsync(&v, ctx->d_exc_counter, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_exc_counter, 0, 4, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *count = v;
    return H264R_OK;
}

int h264r_excursion_map(h264r_ctx* ctx, int frame_id, uint8_t* out)
{
    int r = require_done(ctx, frame_id, "excursion_map");
    if (r != H264R_OK) return r;
    if (!out) return fail(ctx, H264R_E_INVAL, "excursion_map: null output");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(out, ctx->d_exc_maps + (size_t)frame_id * ctx->n_mbs, ctx->n_mbs, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_exc_maps + (size_t)frame_id * ctx->n_mbs, 0, ctx->n_mbs, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return H264R_OK;
}

int h264r_residual_mbs(h264r_ctx* ctx, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* coeff, int16_t* out)
{
    if (!ctx || !hdr || !coeff || !out || n_mbs <= 0) return fail(ctx, H264R_E_INVAL, "residual_mbs: bad argument");
    CU(cudaSetDevice(ctx->device));
    uint8_t* d = nullptr;
    size_t hb = align_up((size_t)n_mbs * sizeof(h264r_mb_hdr), 256), cb = align_up((size_t)n_mbs * 768, 256);
    CU(cudaMalloc(&d, hb + 2 * cb));
    PicDesc pd;
    memset(&pd, 0, sizeof(pd));
    pd.hdr = reinterpret_cast<const h264r_mb_hdr*>(d);
    pd.coeff = reinterpret_cast<const int16_t*>(d + hb);
    int16_t* d_out = reinterpret_cast<int16_t*>(d + hb + cb);
    cudaError_t e = cudaMemcpyAsync(d, hdr, (size_t)n_mbs * sizeof(h264r_mb_hdr), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + hb, coeff, (size_t)n_mbs * 768, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        int blocks = (n_mbs + INTER_WARPS - 1) / INTER_WARPS;
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (ctx->flags & H264R_REF_COMPAT) k_residual<true><<<blocks, INTER_WARPS * 32, 0, ctx->stream>>>(pd, n_mbs, d_out);
        else k_residual<false><<<blocks, INTER_WARPS * 32, 0, ctx->stream>>>(pd, n_mbs, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n_mbs * 768, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, H264R_E_CUDA, std::string("residual_mbs: ") + cudaGetErrorString(e));
    return H264R_OK;
}

int h264r_release(h264r_ctx* ctx, int frame_id)
{
    if (!ctx || frame_id < 0 || frame_id >= ctx->max_frames) return fail(ctx, H264R_E_INVAL, "release: bad frame id");
    uint8_t st = ctx->frame_state[frame_id];
    if (st == FRAME_OPEN) return fail(ctx, H264R_E_STATE, "release: picture is still open");
    if (st == FRAME_QUEUED) {
        int r = h264r_flush(ctx);
        if (r != H264R_OK) return r;
    }
    ctx->frame_state[frame_id] = FRAME_FREE;     // stream order protects work that still reads the slot
    return H264R_OK;
}

int h264r_last_flush_stats(h264r_ctx* ctx, float* kernel_ms, int* launches)
{
    if (!ctx) return H264R_E_INVAL;
    if (launches) *launches = ctx->last_launches;
    if (kernel_ms) {
        *kernel_ms = 0.f;
        if (ctx->timing_valid) {
            CU(cudaSetDevice(ctx->device));
            CU(cudaEventSynchronize(ctx->ev1));
            CU(cudaEventElapsedTime(kernel_ms, ctx->ev0, ctx->ev1));
        }
    }
    return H264R_OK;
}

int h264r_set_profiling(h264r_ctx* ctx, int on)
{
    if (!ctx) return H264R_E_INVAL;
    ctx->profiling = on != 0;
    return H264R_OK;
}

int h264r_last_flush_kernel_ms(h264r_ctx* ctx, float ms[H264R_KERNEL_FAMILIES], int launches[H264R_KERNEL_FAMILIES])
{
    if (!ctx || !ms || !launches) return H264R_E_INVAL;
    for (int k = 0; k < H264R_KERNEL_FAMILIES; k++) { ms[k] = 0.f; launches[k] = 0; }
    CU(cudaSetDevice(ctx->device));
    for (int i = 0; i < ctx->kev_used; i += 2) {
        float t = 0.f;
        CU(cudaEventSynchronize(ctx->kev[i + 1]));
        CU(cudaEventElapsedTime(&t, ctx->kev[i], ctx->kev[i + 1]));
        int f = ctx->kev_family[i / 2];
        ms[f] += t; launches[f]++;
    }
    return H264R_OK;
}

int h264r_last_flush_launch_ms(h264r_ctx* ctx, float* ms, int* family, int cap)
{
    if (!ctx || !ms || !family || cap < 0) return H264R_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    int n = 0;
    for (int i = 0; i < ctx->kev_used && n < cap; i += 2, n++) {
        CU(cudaEventSynchronize(ctx->kev[i + 1]));
        CU(cudaEventElapsedTime(&ms[n], ctx->kev[i], ctx->kev[i + 1]));
        family[n] = ctx->kev_family[i / 2];
    }
    return n;
}

int h264r_mark(h264r_ctx* ctx, int slot)
{
    if (!ctx || slot < 0 || slot > 3) return H264R_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->marks[slot]) CU(cudaEventCreate(&ctx->marks[slot]));
    CU(cudaEventRecord(ctx->marks[slot], ctx->stream));
    return H264R_OK;
}

int h264r_elapsed_ms(h264r_ctx* ctx, int slot_from, int slot_to, float* ms)
{
    if (!ctx || !ms || slot_from < 0 || slot_from > 3 || slot_to < 0 || slot_to > 3 || !ctx->marks[slot_from] || !ctx->marks[slot_to])
        return fail(ctx, H264R_E_INVAL, "elapsed_ms: bad slot");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventSynchronize(ctx->marks[slot_to]));
    CU(cudaEventElapsedTime(ms, ctx->marks[slot_from], ctx->marks[slot_to]));
    return H264R_OK;
}

}  // extern "C"
