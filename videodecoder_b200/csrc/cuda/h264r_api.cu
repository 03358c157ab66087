// h264r_api.cu -- the C ABI of include/h264r.h: context, device frame pool, syntax arena,
// dependency-wave scheduler and kernel launches.  There is no CPU path: every entry point that
// computes anything launches the sm_100a kernels of h264r_kernels.cuh.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
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
    int n_intra_claimed;    // h264r_submit_picture with buf->n_intra >= 0: the caller's count (k_blk_scan checks it on the device), else -1
    int n_submitted;
    int wave;
    int lane;
    int layout;             // 0 = nothing submitted yet, 1 = dense (h264r_submit_mbs), 2 = packed (h264r_submit_mbs_packed)
    size_t n_blocks;        // packed: coefficient blocks stored so far
    bool mv_partition;      // vectors in partition order (h264r_submit_picture with n_mvs >= 0)
    bool short_hdr;         // picture buffer with 8-byte headers (h264r_picture_buf.hdr_bytes = 8): k_blk_scan expands them
    bool stream;            // ... in stream form (H264R_HDR_STREAM): k_blk_scan expands headers, masks, vectors and blocks
    int level_bytes;        // H264R_LEVELS_*: int16 levels (32-byte blocks); picture buffers only: int8 levels (16-byte blocks), compact (8-byte units)
    uint64_t seq;           // order of frame_end calls (= order of the uploads on the copy stream)
    bool early;             // flush: the upload had landed when the flush began (index structures made up front, nothing to wait for)
    bool deblock_on;        // K4 runs for this picture (spec mode, disable_deblocking_filter_idc != 1)
};

}  // namespace

struct h264r_ctx {
    int device = 0;
    Geometry g{};
    int n_mbs = 0, max_frames = 0, max_pending = 0;
    uint32_t flags = 0;
    cudaStream_t stream = nullptr;          // kernels, read-back
    cudaStream_t copy_stream = nullptr;     // syntax uploads: overlap the kernels of earlier waves / flushes
    int n_lanes = 1;                        // independent kernel pipelines (see h264r_flush) in use
    int n_lanes_created = 1;
    int inter_ctas_per_sm = 8;
    bool sparse_intra = true;               // k_intra_sparse for P pictures with few intra macroblocks (H264R_SPARSE_INTRA=0: row wavefront)
    cudaStream_t lane_stream[8] = {};
    cudaEvent_t ev_lane[8] = {};
    // side streams: work that nothing in a lane's chain picture -> next picture waits for stays out of the lanes.
    // prep: block / vector offset scans of uploaded pictures (run ahead of the lanes, behind the uploads);
    // dig: K5 of finished pictures (runs behind the lanes)
    cudaStream_t prep_stream = nullptr, dig_stream = nullptr;
    cudaEvent_t ev_scan[8] = {}, ev_wave[8] = {};
    int digest_every = 1;                   // K5 launches cover this many waves of a flush (H264R_DIGEST_EVERY; measured: 1 is best)
    bool lean_inter = true;                 // k_inter<LEAN> for pictures that qualify (H264R_LEAN=0: the generic instantiation)
    bool pdl = true;                        // programmatic dependent launch of the P-wave kernels (H264R_PDL=0: plain launches)
    bool inter_parts = true;                // spec-mode lean launches as three kernels (H264R_INTER_PARTS=0: one)
    bool tma_old = false, group_tma = false; // tensor-copy window staging in k_inter (H264R_TMA=1) / in k_inter_grp (H264R_GROUP_TMA=1): opt-in, see DESIGN.md
    bool any_stream = false;                // a stream-form picture has been submitted: scans are followed by k_stream_expand
    bool inter_group = true;                // k_inter_grp for REF_COMPAT launches of lean pictures (H264R_INTER_GROUP=0: k_inter<LEAN>)
    int inter_group_gens = 4;               // its grid = SMs x 32 x this (H264R_GROUP_GENS)
    int inter_generations = 8;              // k_inter grid = SMs x inter_ctas_per_sm x this: a few CTA generations balance the tail
    cudaEvent_t ev_dig[64] = {};            // ring: digest launch number n is done
    uint64_t dig_seq = 0;
    uint64_t lane_dig_waited[8] = {};
    std::vector<uint64_t> frame_dig_seq;    // digest launch that last READ a frame slot (a new picture in the slot waits for it)
    // read-back stream: h264r_read_frame_async copies finished frames to the host behind the flush that produced them, while
    // later flushes reconstruct; a picture that overwrites a slot waits for the slot's last read-back (ev_out ring)
    cudaStream_t out_stream = nullptr;
    cudaEvent_t ev_out[64] = {};
    uint64_t out_seq = 0, out_waited_flush = 0;
    uint64_t lane_out_waited[8] = {};
    std::vector<uint64_t> frame_out_seq;    // read-back that last READ a frame slot
    cudaEvent_t ev_main = nullptr;
    bool main_dirty = false;                // the main stream holds work the lanes must wait for
    int next_lane = 0;
    std::vector<int> frame_lane;            // lane that produced / last used a frame slot, -1 = never used
    std::vector<uint32_t> frame_reader_lanes;   // lanes (bit mask) that carry pictures READING the slot's current picture as a reference:
                                            // a picture that overwrites the slot waits for them (write after read)
    size_t sync_pos = 0;                    // ring position in d_sync
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_copied;     // per arena slot: its picture's syntax is resident
    cudaEvent_t ev_flush[64] = {};           // ring: all kernels of flush number n are done
    uint64_t flush_seq = 0, copy_waited_seq = 0, end_seq = 0;
    std::vector<uint64_t> slot_flush_seq;   // flush that last read the slot
    void* d_tmaps = nullptr;                // TMA tensor maps of the luma planes: [frame][2] (box 32x22, box 16x14), 128 B each
    uint32_t* d_dense_mask = nullptr;       // blk_mask / blk_off of the dense layout (shared by all dense pictures)
    uint32_t* d_dense_off = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> kev;           // profiling events: pairs per launch
    std::vector<int> kev_family;
    int kev_used = 0;
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
    // frame pool: per slot Y, U, V
    uint8_t* d_frames = nullptr;
    size_t y_bytes = 0, c_bytes = 0, frame_bytes = 0;
    uint32_t* d_digests = nullptr;          // K5 results, 4 words per frame id
    std::vector<uint8_t> digest_valid;      // per frame id: the table entry belongs to the frame's current picture
    std::vector<uint32_t> h_digests;        // read-back staging
    uint8_t* d_exc_maps = nullptr;          // max_frames * n_mbs
    uint32_t* d_exc_counter = nullptr;
    // syntax arena: per slot hdr | mv | coeff
    uint8_t* d_syntax = nullptr;
    // arena slot = what the uploads carry, in this order: descriptor | hdr | blk_mask | mv | coefficient blocks;
    // then blk_off and mv_off (written by k_blk_scan, never uploaded)
    // then the residual scratch of the intra macroblocks (device only)
    size_t slot_bytes = 0, off_desc = 0, off_hdr = 0, off_mv = 0, off_mask = 0, off_boff = 0, off_mvoff = 0, off_coeff = 0, off_res = 0, off_nz = 0, off_dbrec = 0, off_ilist = 0, off_idone = 0, off_mbox = 0, off_hdr16 = 0;
    // 8-byte headers: masks and blocks of the upload move up behind the shorter header array
    size_t off_mask_s = 0, off_coeff_s = 0;
    // stream form: word counts and stream behind the 8-byte headers (upload); expanded masks / units in device-only areas
    size_t off_words = 0, off_stream = 0, off_xmask = 0, off_xunits = 0, off_xwoff = 0;
    uint32_t mbox_stamp = 0;                // launch number of k_deblock_rows (stamp of its hand-over records)
    int sparse_stamp = 0;                   // launch stamp of k_intra_list (intra_done words of older launches are smaller)
    // per-flush control block (descriptors + picture lists), pinned host mirror + device copy
    uint8_t* h_ctrl = nullptr;
    size_t ctrl_bytes = 0;
    // wavefront hand-over words, zeroed at the start of every flush:
    // progress words indexed by (position in the flush's picture lists, macroblock row): 4*max_pending*H, then 2*max_pending tickets
    int* d_sync = nullptr;
    size_t sync_ints = 0;
    std::vector<uint8_t> frame_state;
    std::vector<int> frame_wave;            // wave of a queued frame inside the current flush
    std::vector<Pending> pending;
    std::vector<int> free_slots;            // FIFO: the slot released longest ago is reused first
    size_t free_head = 0;
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

// Launch with programmatic stream serialisation: the kernel may be scheduled while the kernel before it in the stream
// is still draining; it runs what does not depend on that kernel (headers, residual, work lists) and meets it at
// griddepcontrol.wait (grid_dependency_wait() in h264r_kernels.cuh).  The kernels of a P wave -- k_inter, k_intra_list,
// then k_inter of the next wave -- follow each other this way: their launch latencies and tails overlap.
template <class... KArgs, class... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

FramePool digest_pool(const h264r_ctx* c)
{
    return FramePool{c->d_frames, c->frame_bytes, (size_t)PAD_Y * c->g.pitch_y + PAD_X, c->y_bytes + (size_t)PAD_YC * c->g.pitch_c + PAD_XC,
                     c->y_bytes + c->c_bytes + (size_t)PAD_YC * c->g.pitch_c + PAD_XC};
}

Pending* find_pending(h264r_ctx* c, int frame_id)
{
    for (auto& p : c->pending) if (p.frame_id == frame_id) return &p;
    return nullptr;
}

void fill_desc(h264r_ctx* c, const Pending& p, PicDesc& pd)
{
    memset(&pd, 0, sizeof(pd));
    uint8_t* s = c->slot_ptr(p.slot);
    pd.hdr = reinterpret_cast<const h264r_mb_hdr*>(s + (p.short_hdr ? c->off_hdr16 : c->off_hdr));
    pd.hdr8 = p.short_hdr ? reinterpret_cast<const uint2*>(s + c->off_hdr) : nullptr;
    pd.mb_words = p.stream ? s + c->off_words : nullptr;
    pd.stream = p.stream ? reinterpret_cast<const uint32_t*>(s + c->off_stream) : nullptr;
    pd.stream_off = reinterpret_cast<uint32_t*>(s + c->off_xwoff);
    pd.stream_words = p.stream ? (uint32_t)p.n_blocks : 0u;
    pd.mv = reinterpret_cast<const int16_t*>(s + (p.mv_partition ? c->off_desc : c->off_mv));
    pd.coeff = reinterpret_cast<const int16_t*>(s + (p.stream ? c->off_xunits : (p.short_hdr ? c->off_coeff_s : c->off_coeff)));
    pd.level_fmt = p.level_bytes;
    if (p.layout == 2) {
        pd.blk_mask = reinterpret_cast<const uint32_t*>(s + (p.stream ? c->off_xmask : (p.short_hdr ? c->off_mask_s : c->off_mask)));
        pd.blk_off = reinterpret_cast<const uint32_t*>(s + c->off_boff);
    } else { pd.blk_mask = c->d_dense_mask; pd.blk_off = c->d_dense_off; }
    pd.nzmask = reinterpret_cast<uint32_t*>(s + c->off_nz);
    pd.dbrec = s + c->off_dbrec;
    pd.tmaps = c->d_tmaps;
    pd.mv_off = p.mv_partition ? reinterpret_cast<const uint32_t*>(s + c->off_mvoff) : nullptr;
    pd.packed = p.layout == 2;
    pd.intra_list = reinterpret_cast<uint32_t*>(s + c->off_ilist);
    pd.intra_done = reinterpret_cast<int32_t*>(s + c->off_idone);
    pd.mbox = (c->flags & H264R_REF_COMPAT) ? nullptr : reinterpret_cast<uint32_t*>(s + c->off_mbox);
    // k_residual_pics leaves the residual of the intra macroblocks here, ahead of the wavefront
    pd.residual = p.n_intra > 0 ? reinterpret_cast<int16_t*>(s + c->off_res) : nullptr;
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
    pd.exc_counter = c->d_exc_counter;
    pd.n_intra_host = p.n_intra_claimed;
    pd.digest = nullptr;
    pd.frame_id = p.frame_id;
}

}  // namespace

extern "C" {

int h264r_abi_version(void) { return H264R_ABI_VERSION; }

#ifdef H264R_TRACE
// development build only: clock stamps of the row kernels (see WarpExecTrace); resets the buffer
int h264r_debug_trace(long long* out, int cap)
{
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int));
    if (n > cap) n = cap;
    cudaMemcpyFromSymbol(out, g_trace, (size_t)n * sizeof(long long));
    int zero = 0;
    cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int));
    return n;
}
#endif

const char* h264r_last_error(h264r_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int h264r_create(h264r_ctx** out, int device, int w_mbs, int h_mbs, int max_frames, int max_pending, uint32_t flags)
{
    h264r_ctx* ctx = nullptr;      // errors before the context exists go to the thread-local string
    if (!out || w_mbs <= 0 || h_mbs <= 0 || max_frames <= 0 || max_pending <= 0 || (flags & ~(H264R_REF_COMPAT | H264R_DIGESTS)))
        return fail(nullptr, H264R_E_INVAL, "h264r_create: bad argument");
    if (w_mbs > 32 * ROW_MASK_WORDS)        // the row kernels keep one bit per macroblock of a row in shared memory
        return fail(nullptr, H264R_E_INVAL, "h264r_create: pictures wider than 512 macroblocks (8192 samples) are not supported");
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
    c->g = make_geometry(w_mbs, h_mbs, (int)align_up((size_t)w_mbs * 16 + 2 * PAD_X, 128), (int)align_up((size_t)w_mbs * 8 + 2 * PAD_XC, 128));
    c->n_mbs = w_mbs * h_mbs;
    c->max_frames = max_frames; c->max_pending = max_pending;
    c->y_bytes = (size_t)c->g.pitch_y * (h_mbs * 16 + 2 * PAD_Y);
    c->c_bytes = (size_t)c->g.pitch_c * (h_mbs * 8 + 2 * PAD_YC);
    c->frame_bytes = align_up(c->y_bytes + 2 * c->c_bytes, 256);
    c->off_mv = 0;
    c->off_desc = align_up((size_t)c->n_mbs * H264R_MV_PER_MB * 2, 256);
    c->off_hdr = c->off_desc + align_up(sizeof(PicDesc), 256);
    c->off_mask = c->off_hdr + align_up((size_t)c->n_mbs * sizeof(h264r_mb_hdr), 256);
    c->off_coeff = c->off_mask + align_up((size_t)c->n_mbs * 4, 256);
    c->off_boff = c->off_coeff + align_up((size_t)c->n_mbs * H264R_COEFF_PER_MB * 2, 256);
    c->off_mvoff = c->off_boff + align_up((size_t)c->n_mbs * 4, 256);
    c->off_res = c->off_mvoff + align_up((size_t)c->n_mbs * 4, 256);
    c->off_nz = c->off_res + align_up((size_t)c->n_mbs * 768, 256);
    c->off_dbrec = c->off_nz + align_up((size_t)c->n_mbs * 4, 256);
    c->off_ilist = c->off_dbrec + ((flags & H264R_REF_COMPAT) ? 0 : align_up((size_t)c->n_mbs * DBREC_BYTES, 256));
    c->off_idone = c->off_ilist + align_up((size_t)(c->n_mbs + 1) * 4, 256);
    c->off_mbox = c->off_idone + align_up((size_t)c->n_mbs * 4, 256);
    c->off_hdr16 = c->off_mbox + ((flags & H264R_REF_COMPAT) ? 0 : align_up((size_t)c->n_mbs * MBOX_WORDS * 4, 256));      // device only: headers expanded from the 8-byte form
    c->off_xmask = c->off_hdr16 + align_up((size_t)c->n_mbs * sizeof(h264r_mb_hdr), 256);
    c->off_xunits = c->off_xmask + align_up((size_t)c->n_mbs * 4, 256);
    c->off_xwoff = c->off_xunits + align_up((size_t)c->n_mbs * 48 * 8, 256);           // 24 blocks x two 8-byte units
    c->slot_bytes = c->off_xwoff + align_up((size_t)c->n_mbs * 4, 256);
    c->off_words = c->off_hdr + (size_t)c->n_mbs * sizeof(h264r_mb_hdr8);
    c->off_stream = align_up(c->off_words + (size_t)c->n_mbs, 16);
    c->off_mask_s = c->off_hdr + (size_t)c->n_mbs * sizeof(h264r_mb_hdr8);
    c->off_coeff_s = align_up(c->off_mask_s + (size_t)c->n_mbs * 4, 16);
    c->ctrl_bytes = align_up((size_t)max_pending * sizeof(PicDesc), 256);      // descriptor table, one entry per arena slot
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
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    {
        const char* e = getenv("H264R_LANES");      // tuning knob; 1 = a single pipeline
        int nl = e ? atoi(e) : 1;
        c->n_lanes = nl < 1 ? 1 : (nl > 8 ? 8 : nl);
        c->n_lanes_created = c->n_lanes;
        const char* e5 = getenv("H264R_INTER_GENS");
        if (e5 && atoi(e5) > 0) c->inter_generations = atoi(e5);
        const char* e7 = getenv("H264R_DIGEST_EVERY");
        if (e7 && atoi(e7) > 0) c->digest_every = atoi(e7);
        const char* e12 = getenv("H264R_INTER_PARTS");
        if (e12) c->inter_parts = atoi(e12) != 0;
        const char* e10 = getenv("H264R_INTER_GROUP");
        if (e10) c->inter_group = atoi(e10) != 0;
        const char* e11 = getenv("H264R_GROUP_GENS");
        if (e11 && atoi(e11) > 0) c->inter_group_gens = atoi(e11);
        const char* e9 = getenv("H264R_LEAN");
        if (e9) c->lean_inter = atoi(e9) != 0;
        const char* e8 = getenv("H264R_PDL");
        if (e8) c->pdl = atoi(e8) != 0;
        const char* e3 = getenv("H264R_SPARSE_INTRA");
        if (e3) c->sparse_intra = atoi(e3) != 0;
        const char* e2 = getenv("H264R_INTER_CTAS");
        if (e2 && atoi(e2) > 0) c->inter_ctas_per_sm = atoi(e2);
    }
    for (int l = 0; l < c->n_lanes; l++) {
        CUC(cudaStreamCreateWithFlags(&c->lane_stream[l], cudaStreamNonBlocking));
        CUC(cudaEventCreateWithFlags(&c->ev_lane[l], cudaEventDisableTiming));
    }
    CUC(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    {
        // H264R_PREP_PRIO=1: the scans / expansions of later waves at the lowest stream priority, so that their CTAs take what the
        // wave kernels leave free instead of an equal share (tuning knob; default 0 -- measured: resident 82.2 k against 82.1 k
        // frames/s, end to end 75.8 k against 78.0 k, where a round's scan is on the path to its kernels)
        int lo = 0, hi = 0;
        const char* ep = getenv("H264R_PREP_PRIO");
        if (ep && atoi(ep) != 0 && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess && lo != hi)
            CUC(cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, lo));
        else
            CUC(cudaStreamCreateWithFlags(&c->prep_stream, cudaStreamNonBlocking));
    }
    CUC(cudaStreamCreateWithFlags(&c->dig_stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 64; i++) CUC(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    c->frame_out_seq.assign(max_frames, 0);
    for (int l = 0; l < 8; l++) {
        CUC(cudaEventCreateWithFlags(&c->ev_scan[l], cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->ev_wave[l], cudaEventDisableTiming));
    }
    for (int i = 0; i < 64; i++) CUC(cudaEventCreateWithFlags(&c->ev_dig[i], cudaEventDisableTiming));

    c->frame_dig_seq.assign(max_frames, 0);
    c->frame_lane.assign(max_frames, -1);
    c->frame_reader_lanes.assign(max_frames, 0);
    c->ev_copied.assign(max_pending, nullptr);
    for (int i = 0; i < max_pending; i++) CUC(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
    for (int i = 0; i < 64; i++) CUC(cudaEventCreateWithFlags(&c->ev_flush[i], cudaEventDisableTiming));
    c->slot_flush_seq.assign(max_pending, 0);
    {
        std::vector<uint32_t> m(c->n_mbs, 0xFFFFFFu), o(c->n_mbs);
        for (int a = 0; a < c->n_mbs; a++) o[a] = 24u * (uint32_t)a;
        CUC(cudaMalloc(&c->d_dense_mask, (size_t)c->n_mbs * 4));
        CUC(cudaMalloc(&c->d_dense_off, (size_t)c->n_mbs * 4));
        CUC(cudaMemcpy(c->d_dense_mask, m.data(), (size_t)c->n_mbs * 4, cudaMemcpyHostToDevice));
        CUC(cudaMemcpy(c->d_dense_off, o.data(), (size_t)c->n_mbs * 4, cudaMemcpyHostToDevice));
    }
    CUC(cudaEventCreate(&c->ev0));
    CUC(cudaEventCreate(&c->ev1));
    CUC(cudaMalloc(&c->d_frames, c->frame_bytes * max_frames));
    CUC(cudaMalloc(&c->d_digests, (size_t)max_frames * 16));
    // Tensor maps of the padded luma planes: the tensor-copy variants of the inter kernels stage the windows of 16x16 macroblocks with them
    // (H264R_GROUP_TMA=1: k_inter_grp<.., TMA>; H264R_TMA=1: the per-macroblock kernels k_inter<.., TMA, ..>)
    c->tma_old = getenv("H264R_TMA") && atoi(getenv("H264R_TMA")) != 0;
    if (getenv("H264R_GROUP_TMA")) c->group_tma = atoi(getenv("H264R_GROUP_TMA")) != 0;
    {
        // one tensor map per frame slot and window shape over the PADDED luma plane (u8, pitch x rows); k_inter issues
        // cp.async.bulk.tensor copies of 48 x 22 (16x16 blocks) and 32 x 14 (8x8 quadrants) boxes from them
        // the encoder is a driver entry point: resolved through the runtime so that the library itself does not link
        // libcuda (it must load, for symbol checks, on machines without a driver)
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CUC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) {
            fail(nullptr, 0, "h264r_create: the driver does not provide cuTensorMapEncodeTiled");
            return cleanup(H264R_E_CUDA);
        }
        const EncodeTiled encode = reinterpret_cast<EncodeTiled>(fn);
        std::vector<CUtensorMap> maps((size_t)max_frames * 2);
        const cuuint64_t dims[2] = {(cuuint64_t)c->g.pitch_y, (cuuint64_t)(h_mbs * 16 + 2 * PAD_Y)};
        const cuuint64_t strides[1] = {(cuuint64_t)c->g.pitch_y};
        const cuuint32_t estr[2] = {1, 1};
        for (int f = 0; f < max_frames; f++)
            for (int k = 0; k < 2; k++) {
                const cuuint32_t box[2] = {k == 0 ? (cuuint32_t)TMA0_W : (cuuint32_t)TMA1_W, k == 0 ? (cuuint32_t)TMA0_H : (cuuint32_t)TMA1_H};
                CUresult r = encode(&maps[(size_t)f * 2 + k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->d_frames + (size_t)f * c->frame_bytes,
                                                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    fail(nullptr, 0, "h264r_create: cuTensorMapEncodeTiled failed");
                    return cleanup(H264R_E_CUDA);
                }
            }
        static_assert(sizeof(CUtensorMap) == 128, "tensor maps are indexed as 128-byte records on the device");
        CUC(cudaMalloc(&c->d_tmaps, maps.size() * sizeof(CUtensorMap)));
        CUC(cudaMemcpy(c->d_tmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    }
    c->digest_valid.assign(max_frames, 0);
    c->h_digests.assign((size_t)max_frames * 4, 0);
    CUC(cudaMalloc(&c->d_exc_maps, (size_t)max_frames * c->n_mbs));
    CUC(cudaMalloc(&c->d_exc_counter, 256));
    CUC(cudaMalloc(&c->d_syntax, c->slot_bytes * max_pending));
    CUC(cudaMallocHost(&c->h_ctrl, c->ctrl_bytes));
    CUC(cudaMemsetAsync(c->d_exc_counter, 0, 256, c->stream));
    for (int s = 0; s < max_pending; s++) CUC(cudaMemsetAsync(c->slot_ptr(s) + c->off_idone, 0, (size_t)c->n_mbs * 4, c->stream));
    if (!(flags & H264R_REF_COMPAT))        // stamps of the deblocking hand-over records: launch numbers start at 1
        for (int s = 0; s < max_pending; s++) CUC(cudaMemsetAsync(c->slot_ptr(s) + c->off_mbox, 0, (size_t)c->n_mbs * MBOX_WORDS * 4, c->stream));
    CUC(cudaMemsetAsync(c->d_exc_maps, 0, (size_t)max_frames * c->n_mbs, c->stream));
    // ring of hand-over words: in flight are at most max_pending pictures x 2 wavefront kernels x (H + 1) words
    c->sync_ints = (size_t)8 * max_pending * (h_mbs + 1) + 1024;
    CUC(cudaMalloc(&c->d_sync, c->sync_ints * sizeof(int)));
    CUC(cudaMemsetAsync(c->d_sync, 0, c->sync_ints * sizeof(int), c->stream));
    CUC(cudaStreamSynchronize(c->stream));
    c->frame_state.assign(max_frames, FRAME_FREE);
    c->frame_wave.assign(max_frames, -1);
    for (int s = 0; s < max_pending; s++) c->free_slots.push_back(s);
    *out = c;
    return H264R_OK;
#undef CUC
}

void h264r_destroy(h264r_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (int l = 0; l < 8; l++) if (c->lane_stream[l]) cudaStreamSynchronize(c->lane_stream[l]);
    if (c->prep_stream) cudaStreamSynchronize(c->prep_stream);
    if (c->dig_stream) cudaStreamSynchronize(c->dig_stream);
    if (c->out_stream) cudaStreamSynchronize(c->out_stream);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int l = 0; l < 8; l++) {
        if (c->lane_stream[l]) cudaStreamDestroy(c->lane_stream[l]);
        if (c->ev_lane[l]) cudaEventDestroy(c->ev_lane[l]);
        if (c->ev_scan[l]) cudaEventDestroy(c->ev_scan[l]);
        if (c->ev_wave[l]) cudaEventDestroy(c->ev_wave[l]);
    }
    for (auto e : c->ev_dig) if (e) cudaEventDestroy(e);

    if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
    if (c->dig_stream) cudaStreamDestroy(c->dig_stream);
    if (c->out_stream) cudaStreamDestroy(c->out_stream);
    for (auto e : c->ev_out) if (e) cudaEventDestroy(e);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    cudaFree(c->d_dense_mask); cudaFree(c->d_dense_off); cudaFree(c->d_tmaps);
    for (auto e : c->ev_copied) if (e) cudaEventDestroy(e);
    for (auto e : c->ev_flush) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    cudaFree(c->d_frames); cudaFree(c->d_digests); cudaFree(c->d_exc_maps); cudaFree(c->d_exc_counter);
    cudaFree(c->d_syntax); cudaFree(c->d_sync);
    if (c->h_ctrl) cudaFreeHost(c->h_ctrl);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
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
    // A picture still OPEN may be issued by a later flush than this one: overwriting one of its references in between
    // cannot be ordered.  (Queued pictures are fine: the flush puts the overwriting picture in a later wave.)
    for (const auto& q : ctx->pending)
        if (!q.ended)
            for (int i = 0; i < q.pp.num_ref_idx_l0; i++)
                if (q.pp.ref_list0[i] == frame_id) return fail(ctx, H264R_E_STATE, "frame_begin: frame slot is a reference of a picture that is still open");
    if (ctx->free_head >= ctx->free_slots.size()) return fail(ctx, H264R_E_CAPACITY, "frame_begin: max_pending pictures already queued; call h264r_flush");
    Pending p{};
    p.level_bytes = 2;
    p.n_intra_claimed = -1;
    p.deblock_on = !(ctx->flags & H264R_REF_COMPAT) && pp->disable_deblocking_filter_idc != 1;
    p.frame_id = frame_id;
    p.slot = ctx->free_slots[ctx->free_head++];
    if (ctx->free_head > 4096 && ctx->free_head * 2 > ctx->free_slots.size()) {
        ctx->free_slots.erase(ctx->free_slots.begin(), ctx->free_slots.begin() + ctx->free_head);
        ctx->free_head = 0;
    }
    p.pp = *pp;
    ctx->pending.push_back(p);
    ctx->frame_state[frame_id] = FRAME_OPEN;
    ctx->digest_valid[frame_id] = 0;
    return H264R_OK;
}

// common part of the two submit calls: validation, header and motion-vector upload
static int submit_common(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv, int layout,
                         Pending** out)
{
    if (!ctx || !hdr || n_mbs <= 0 || first_mb < 0) return fail(ctx, H264R_E_INVAL, "submit_mbs: bad argument");
    Pending* p = find_pending(ctx, frame_id);
    if (!p || p->ended) return fail(ctx, H264R_E_STATE, "submit_mbs: no open picture with this frame id");
    if (first_mb + n_mbs > ctx->n_mbs) return fail(ctx, H264R_E_INVAL, "submit_mbs: macroblock range exceeds the picture");
    if (p->pp.slice_type == H264R_SLICE_P && !mv) return fail(ctx, H264R_E_INVAL, "submit_mbs: P picture needs motion vectors");
    if (p->layout && p->layout != layout) return fail(ctx, H264R_E_STATE, "submit_mbs: dense and packed submits cannot be mixed within a picture");
    // chunks in macroblock order, each macroblock once: n_submitted == n_mbs at frame_end then means "all of them"
    if (first_mb != p->n_submitted) return fail(ctx, H264R_E_STATE, "submit_mbs: chunks must arrive in macroblock order, without gaps or repeats");
    int n_intra = 0;          // committed only if the whole chunk is accepted (a failed chunk may be retried)
    for (int i = 0; i < n_mbs; i++) {
        uint8_t cls = hdr[i].mb_class;
        if (cls == 3 || cls > H264R_MB_P8x8) return fail(ctx, H264R_E_INVAL, "submit_mbs: unknown mb_class");
        if (cls <= H264R_MB_I16x16) n_intra++;
        else if (p->pp.slice_type != H264R_SLICE_P) return fail(ctx, H264R_E_INVAL, "submit_mbs: inter macroblock in an I picture");
        if ((ctx->flags & H264R_REF_COMPAT) && (cls == H264R_MB_I8x8 || (hdr[i].flags & H264R_T8x8)))
            return fail(ctx, H264R_E_INVAL, "submit_mbs: 8x8 transform is not part of REF_COMPAT (the reference has none)");
    }
    CU(cudaSetDevice(ctx->device));
    // the slot may still be read by the kernels of the flush that used it last
    if (ctx->slot_flush_seq[p->slot] > ctx->copy_waited_seq) {
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_flush[ctx->slot_flush_seq[p->slot] & 63], 0));
        ctx->copy_waited_seq = ctx->slot_flush_seq[p->slot];
    }
    p->layout = layout;
    p->n_intra += n_intra;
    uint8_t* s = ctx->slot_ptr(p->slot);
    CU(cudaMemcpyAsync(s + ctx->off_hdr + (size_t)first_mb * sizeof(h264r_mb_hdr), hdr, (size_t)n_mbs * sizeof(h264r_mb_hdr), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (mv && p->pp.slice_type == H264R_SLICE_P)
        CU(cudaMemcpyAsync(s + ctx->off_mv + (size_t)first_mb * H264R_MV_PER_MB * 2, mv, (size_t)n_mbs * H264R_MV_PER_MB * 2,
                           cudaMemcpyHostToDevice, ctx->copy_stream));
    *out = p;
    return H264R_OK;
}

int h264r_submit_mbs(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv,
                     const int16_t* coeff)
{
    if (!coeff) return fail(ctx, H264R_E_INVAL, "submit_mbs: bad argument");
    Pending* p = nullptr;
    int r = submit_common(ctx, frame_id, first_mb, n_mbs, hdr, mv, 1, &p);
    if (r != H264R_OK) return r;
    uint8_t* s = ctx->slot_ptr(p->slot);
    CU(cudaMemcpyAsync(s + ctx->off_coeff + (size_t)first_mb * H264R_COEFF_PER_MB * 2, coeff, (size_t)n_mbs * H264R_COEFF_PER_MB * 2,
                       cudaMemcpyHostToDevice, ctx->copy_stream));
    p->n_submitted += n_mbs;
    return H264R_OK;
}

int h264r_submit_mbs_packed(h264r_ctx* ctx, int frame_id, int first_mb, int n_mbs, const h264r_mb_hdr* hdr, const int16_t* mv,
                            const uint32_t* blk_mask, const int16_t* blocks, size_t n_blocks)
{
    if (!blk_mask || (n_blocks && !blocks)) return fail(ctx, H264R_E_INVAL, "submit_mbs_packed: bad argument");
    if (n_blocks > (size_t)24 * (n_mbs > 0 ? n_mbs : 0)) return fail(ctx, H264R_E_INVAL, "submit_mbs_packed: more blocks than 24 per macroblock");
    Pending* p = nullptr;
    int r = submit_common(ctx, frame_id, first_mb, n_mbs, hdr, mv, 2, &p);
    if (r != H264R_OK) return r;
    uint8_t* s = ctx->slot_ptr(p->slot);
    CU(cudaMemcpyAsync(s + ctx->off_mask + (size_t)first_mb * 4, blk_mask, (size_t)n_mbs * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (n_blocks)
        CU(cudaMemcpyAsync(s + ctx->off_coeff + p->n_blocks * 32, blocks, n_blocks * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
    p->n_blocks += n_blocks;
    p->n_submitted += n_mbs;
    return H264R_OK;
}

int h264r_picture_buf_alloc(h264r_ctx* ctx, size_t max_blocks, h264r_picture_buf* out)
{
    if (!ctx || !out) return fail(ctx, H264R_E_INVAL, "picture_buf_alloc: bad argument");
    if (max_blocks == 0 || max_blocks > (size_t)24 * ctx->n_mbs) max_blocks = (size_t)24 * ctx->n_mbs;
    CU(cudaSetDevice(ctx->device));
    uint8_t* b = nullptr;
    if (cudaMallocHost(&b, ctx->off_coeff + max_blocks * 32) != cudaSuccess) return fail(ctx, H264R_E_NOMEM, "picture_buf_alloc: pinned allocation failed");
    memset(b, 0, ctx->off_coeff);
    out->hdr = reinterpret_cast<h264r_mb_hdr*>(b + ctx->off_hdr);
    out->mv = reinterpret_cast<int16_t*>(b + ctx->off_mv);
    out->mv_end = reinterpret_cast<int16_t*>(b + ctx->off_desc);
    out->blk_mask = reinterpret_cast<uint32_t*>(b + ctx->off_mask);
    out->blocks = reinterpret_cast<int16_t*>(b + ctx->off_coeff);
    out->max_blocks = max_blocks;
    out->n_intra = -1;
    out->level_bytes = 2;
    out->base = b;
    out->hdr_bytes = 16;
    out->ring_index = 0;
    out->hdr8 = nullptr;
    out->mb_words = nullptr; out->stream = nullptr; out->max_stream_words = 0;
    return H264R_OK;
}

int h264r_picture_ring_alloc(h264r_ctx* ctx, int n, size_t max_blocks, h264r_picture_buf* out)
{
    if (!ctx || !out || n <= 0) return fail(ctx, H264R_E_INVAL, "picture_ring_alloc: bad argument");
    if (max_blocks == 0 || max_blocks > (size_t)24 * ctx->n_mbs) max_blocks = (size_t)24 * ctx->n_mbs;
    CU(cudaSetDevice(ctx->device));
    const size_t stride = align_up(ctx->off_coeff + max_blocks * 32, 4096);
    uint8_t* b = nullptr;
    if (cudaMallocHost(&b, stride * (size_t)n) != cudaSuccess) return fail(ctx, H264R_E_NOMEM, "picture_ring_alloc: pinned allocation failed");
    for (int i = 0; i < n; i++, b += stride) {
        memset(b, 0, ctx->off_coeff);
        out[i].hdr = reinterpret_cast<h264r_mb_hdr*>(b + ctx->off_hdr);
        out[i].mv = reinterpret_cast<int16_t*>(b + ctx->off_mv);
        out[i].mv_end = reinterpret_cast<int16_t*>(b + ctx->off_desc);
        out[i].blk_mask = reinterpret_cast<uint32_t*>(b + ctx->off_mask);
        out[i].blocks = reinterpret_cast<int16_t*>(b + ctx->off_coeff);
        out[i].max_blocks = max_blocks;
        out[i].n_intra = -1;
        out[i].level_bytes = 2;
        out[i].base = b;
        out[i].hdr_bytes = 16;
        out[i].ring_index = i + 1;
        out[i].hdr8 = nullptr;
        out[i].mb_words = nullptr; out[i].stream = nullptr; out[i].max_stream_words = 0;
    }
    return H264R_OK;
}

int h264r_picture_buf_format(h264r_ctx* ctx, h264r_picture_buf* buf, int hdr_bytes)
{
    if (!ctx || !buf || !buf->base || (hdr_bytes != 16 && hdr_bytes != 8 && hdr_bytes != H264R_HDR_STREAM))
        return fail(ctx, H264R_E_INVAL, "picture_buf_format: bad argument");
    uint8_t* b = static_cast<uint8_t*>(buf->base);
    const bool s8 = hdr_bytes == 8, st = hdr_bytes == H264R_HDR_STREAM;
    buf->hdr_bytes = hdr_bytes;
    buf->hdr = reinterpret_cast<h264r_mb_hdr*>(b + ctx->off_hdr);
    buf->hdr8 = (s8 || st) ? reinterpret_cast<h264r_mb_hdr8*>(b + ctx->off_hdr) : nullptr;
    buf->blk_mask = reinterpret_cast<uint32_t*>(b + (s8 ? ctx->off_mask_s : ctx->off_mask));
    buf->blocks = reinterpret_cast<int16_t*>(b + (s8 ? ctx->off_coeff_s : ctx->off_coeff));
    buf->mb_words = st ? b + ctx->off_words : nullptr;
    buf->stream = st ? reinterpret_cast<uint32_t*>(b + ctx->off_stream) : nullptr;
    buf->max_stream_words = st ? (ctx->off_coeff + buf->max_blocks * 32 - ctx->off_stream) / 4 : 0;
    return H264R_OK;
}

void h264r_picture_buf_free(h264r_ctx* ctx, h264r_picture_buf* buf)
{
    if (!ctx || !buf || !buf->base) return;
    cudaSetDevice(ctx->device);
    if (buf->ring_index <= 1) cudaFreeHost(buf->base);       // a ring goes with its first member; the others only forget their pointers
    memset(buf, 0, sizeof(*buf));
}

// Everything h264r_submit_picture does short of the transfer: checks, picture state, the descriptor written into the buffer.
// *first / *end: the byte range of the buffer (= of the arena slot) that has to travel.
static int prepare_picture(h264r_ctx* ctx, int frame_id, const h264r_picture_buf* buf, size_t n_blocks, long n_mvs, Pending** out, size_t* first_out, size_t* end_out)
{
    const bool stream = buf && buf->hdr_bytes == H264R_HDR_STREAM;
    if (!ctx || !buf || !buf->base || (!stream && n_blocks > buf->max_blocks * (buf->level_bytes == H264R_LEVELS_COMPACT ? 4 : 1)) ||
        (stream && (n_blocks > buf->max_stream_words || n_mvs < 0)) || n_mvs > (long)16 * (ctx ? ctx->n_mbs : 0) ||
        (buf->level_bytes != H264R_LEVELS_INT8 && buf->level_bytes != H264R_LEVELS_INT16 && buf->level_bytes != H264R_LEVELS_COMPACT))
        return fail(ctx, H264R_E_INVAL, "submit_picture: bad argument");
    Pending* p = find_pending(ctx, frame_id);
    if (!p || p->ended) return fail(ctx, H264R_E_STATE, "submit_picture: no open picture with this frame id");
    if (p->layout || p->n_submitted) return fail(ctx, H264R_E_STATE, "submit_picture: the picture already has macroblocks");
    const bool short_hdr = buf->hdr_bytes == 8 || stream;
    if (stream && buf->stream != reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(buf->base) + ctx->off_stream))
        return fail(ctx, H264R_E_INVAL, "submit_picture: the stream form needs a buffer formatted by h264r_picture_buf_format");
    if (short_hdr && (buf->level_bytes != H264R_LEVELS_COMPACT || buf->hdr8 != reinterpret_cast<h264r_mb_hdr8*>(static_cast<uint8_t*>(buf->base) + ctx->off_hdr)))
        return fail(ctx, H264R_E_INVAL, "submit_picture: 8-byte headers need H264R_LEVELS_COMPACT and a buffer formatted by h264r_picture_buf_format");
    if (buf->n_intra >= 0) {
        // the parser counted while parsing: no second pass over the headers
        if (buf->n_intra > ctx->n_mbs || (p->pp.slice_type != H264R_SLICE_P && buf->n_intra != ctx->n_mbs))
            return fail(ctx, H264R_E_INVAL, "submit_picture: n_intra inconsistent with the picture");
        p->n_intra = buf->n_intra;
        p->n_intra_claimed = buf->n_intra;
    } else {
        for (int i = 0; i < ctx->n_mbs; i++) {
            const uint8_t cls = short_hdr ? (uint8_t)(buf->hdr8[i].w0 & 7u) : buf->hdr[i].mb_class;
            const uint8_t mbf = short_hdr ? (uint8_t)((buf->hdr8[i].w0 >> 3) & 63u) : buf->hdr[i].flags;
            if (cls == 3 || cls > H264R_MB_P8x8) return fail(ctx, H264R_E_INVAL, "submit_picture: unknown mb_class");
            if (cls <= H264R_MB_I16x16) p->n_intra++;
            else if (p->pp.slice_type != H264R_SLICE_P) return fail(ctx, H264R_E_INVAL, "submit_picture: inter macroblock in an I picture");
            if ((ctx->flags & H264R_REF_COMPAT) && (cls == H264R_MB_I8x8 || (mbf & H264R_T8x8)))
                return fail(ctx, H264R_E_INVAL, "submit_picture: 8x8 transform is not part of REF_COMPAT (the reference has none)");
        }
    }
    CU(cudaSetDevice(ctx->device));
    if (ctx->slot_flush_seq[p->slot] > ctx->copy_waited_seq) {
        CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_flush[ctx->slot_flush_seq[p->slot] & 63], 0));
        ctx->copy_waited_seq = ctx->slot_flush_seq[p->slot];
    }
    p->layout = 2;
    p->n_blocks = n_blocks;
    p->n_submitted = ctx->n_mbs;
    p->mv_partition = n_mvs >= 0;
    p->short_hdr = short_hdr;
    p->stream = stream;
    if (stream) ctx->any_stream = true;
    p->level_bytes = buf->level_bytes;
    // ONE transfer: the vectors present (backward from the descriptor), descriptor, headers, masks, the blocks present
    uint8_t* hb = static_cast<uint8_t*>(buf->base);
    fill_desc(ctx, *p, *reinterpret_cast<PicDesc*>(hb + ctx->off_desc));
    const bool is_p = p->pp.slice_type == H264R_SLICE_P;
    const size_t mv_bytes = (!is_p || stream) ? 0 : (n_mvs >= 0 ? (size_t)n_mvs * 4 : ctx->off_desc);      // per-4x4 layout: the whole area; stream form: inside the stream
    const size_t first = ctx->off_desc - mv_bytes;
    const size_t unit_bytes = buf->level_bytes == H264R_LEVELS_INT16 ? 32 : (buf->level_bytes == H264R_LEVELS_INT8 ? 16 : 8);
    *first_out = first;
    *end_out = stream ? ctx->off_stream + n_blocks * 4 : (buf->hdr_bytes == 8 ? ctx->off_coeff_s : ctx->off_coeff) + n_blocks * unit_bytes;
    *out = p;
    return H264R_OK;
}
static void picture_sent(h264r_ctx* ctx, Pending* p)
{
    cudaEventRecord(ctx->ev_copied[p->slot], ctx->copy_stream);
    p->seq = ++ctx->end_seq;
    p->ended = true;
    ctx->frame_state[p->frame_id] = FRAME_QUEUED;
}

int h264r_submit_picture(h264r_ctx* ctx, int frame_id, const h264r_picture_buf* buf, size_t n_blocks, long n_mvs)
{
    Pending* p = nullptr;
    size_t first = 0, end = 0;
    const int rc = prepare_picture(ctx, frame_id, buf, n_blocks, n_mvs, &p, &first, &end);
    if (rc != H264R_OK) return rc;
    const uint8_t* hb = static_cast<const uint8_t*>(buf->base);
    CU(cudaMemcpyAsync(ctx->slot_ptr(p->slot) + first, hb + first, end - first, cudaMemcpyHostToDevice, ctx->copy_stream));
    picture_sent(ctx, p);
    return H264R_OK;
}

int h264r_submit_pictures(h264r_ctx* ctx, int n, const int* frame_ids, const h264r_picture_buf* bufs, const size_t* n_blocks, const long* n_mvs)
{
    if (!ctx || n <= 0 || !frame_ids || !bufs || !n_blocks || !n_mvs) return fail(ctx, H264R_E_INVAL, "submit_pictures: bad argument");
    struct Item { Pending* p; size_t first, end; const uint8_t* hb; size_t cap; };
    std::vector<Item> it((size_t)n);
    int rc = H264R_OK, m = 0;
    for (; m < n; m++) {
        rc = prepare_picture(ctx, frame_ids[m], &bufs[m], n_blocks[m], n_mvs[m], &it[m].p, &it[m].first, &it[m].end);
        if (rc != H264R_OK) break;          // the pictures before this one still travel
        it[m].hb = static_cast<const uint8_t*>(bufs[m].base);
        it[m].cap = ctx->off_coeff + bufs[m].max_blocks * 32;
    }
    // runs of pictures whose buffers lie at a constant stride (members of a ring, in order) and whose arena slots are consecutive:
    // ONE strided transfer per run, over the union of the byte ranges (what it carries beyond a picture's own range is unused space
    // of that picture's slot on both sides)
    for (int i = 0; i < m;) {
        int j = i + 1;
        ptrdiff_t stride = 0;
        size_t lo = it[i].first, hi = it[i].end, cap = it[i].cap;
        while (j < m && j - i < 64) {
            const ptrdiff_t d = it[j].hb - it[j - 1].hb;
            if (it[j].p->slot != it[j - 1].p->slot + 1 || d <= 0 || (stride && d != stride)) break;
            const size_t lo2 = std::min(lo, it[j].first), hi2 = std::max(hi, it[j].end), cap2 = std::min(cap, it[j].cap);
            if (hi2 > cap2 || hi2 - lo2 > (size_t)d) break;
            stride = d; lo = lo2; hi = hi2; cap = cap2;
            j++;
        }
        cudaError_t e;
        if (j - i == 1) e = cudaMemcpyAsync(ctx->slot_ptr(it[i].p->slot) + it[i].first, it[i].hb + it[i].first, it[i].end - it[i].first, cudaMemcpyHostToDevice, ctx->copy_stream);
        else e = cudaMemcpy2DAsync(ctx->slot_ptr(it[i].p->slot) + lo, ctx->slot_bytes, it[i].hb + lo, (size_t)stride, hi - lo, (size_t)(j - i), cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e != cudaSuccess) return fail(ctx, H264R_E_CUDA, std::string("submit_pictures: ") + cudaGetErrorString(e));
        for (int k = i; k < j; k++) picture_sent(ctx, it[k].p);
        i = j;
    }
    return rc;
}


int h264r_frame_end(h264r_ctx* ctx, int frame_id)
{
    if (!ctx) return H264R_E_INVAL;
    Pending* p = find_pending(ctx, frame_id);
    if (!p || p->ended) return fail(ctx, H264R_E_STATE, "frame_end: no open picture with this frame id");
    if (p->n_submitted != ctx->n_mbs) return fail(ctx, H264R_E_STATE, "frame_end: picture is missing macroblocks");
    CU(cudaSetDevice(ctx->device));
    // the picture's descriptor travels with its syntax (see PicList in h264r_kernels.cuh)
    PicDesc* h_desc = reinterpret_cast<PicDesc*>(ctx->h_ctrl);
    // h_ctrl[slot] is the source of an asynchronous copy: the one made for the slot's previous picture must have run
    // before the entry is rewritten (the copy stream may be parked behind a flush while the host runs ahead)
    CU(cudaEventSynchronize(ctx->ev_copied[p->slot]));
    fill_desc(ctx, *p, h_desc[p->slot]);
    CU(cudaMemcpyAsync(ctx->slot_ptr(p->slot) + ctx->off_desc, &h_desc[p->slot], sizeof(PicDesc), cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_copied[p->slot], ctx->copy_stream));
    p->seq = ++ctx->end_seq;
    p->ended = true;
    ctx->frame_state[frame_id] = FRAME_QUEUED;
    return H264R_OK;
}

int h264r_frame_abort(h264r_ctx* ctx, int frame_id)
{
    if (!ctx || frame_id < 0 || frame_id >= ctx->max_frames) return fail(ctx, H264R_E_INVAL, "frame_abort: bad frame id");
    for (size_t i = 0; i < ctx->pending.size(); i++) {
        Pending& p = ctx->pending[i];
        if (p.frame_id != frame_id || p.ended) continue;
        // chunks already on the copy stream land in the slot before anything a later picture uploads into it (one stream)
        ctx->free_slots.push_back(p.slot);
        ctx->pending.erase(ctx->pending.begin() + (long)i);
        ctx->frame_state[frame_id] = FRAME_FREE;
        return H264R_OK;
    }
    return fail(ctx, H264R_E_STATE, "frame_abort: no open picture with this frame id");
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
    const int NL = ctx->n_lanes;
    const bool pdl = ctx->pdl && !ctx->profiling;      // per-launch events sit between the kernels anyway
    // Dependency waves, and the LANE of every picture.  A lane is a CUDA stream that carries whole prediction chains:
    // a picture goes to the lane of its references (stream order is then its dependency), a picture without
    // references to the lane its frame slot used before (stream order protects the slot's old readers) or, for a
    // fresh slot, to the next lane round-robin.  Independent GOPs / streams therefore run as independent pipelines:
    // the latency-bound wavefront kernels of one lane overlap the throughput-bound kernels of the others.
    int n_waves = 0;
    std::vector<uint32_t> foreign(run.size(), 0);      // lanes (bit mask) a picture must wait for besides its own
    // Write after read: a picture that OVERWRITES a frame slot which pictures of this flush still read as a reference
    // (small pools: P1 = f1 <- f0, P2 = f2 <- f1, P3 = f0 <- f2; two streams sharing a pool) goes to a later wave than
    // the last of those readers, and waits for the lanes that carry readers (of this or of earlier flushes).  `run` is in
    // frame_begin order, which is the order that gives "the picture in slot f" its meaning.
    std::vector<int> reader_wave(ctx->max_frames, -1);      // highest wave of this flush that reads the slot's current picture
    for (size_t k = 0; k < run.size(); k++) {
        Pending* p = run[k];
        int wv = 0, lane = -1;
        for (int i = 0; i < p->pp.num_ref_idx_l0; i++) {
            int r = p->pp.ref_list0[i];
            if (ctx->frame_state[r] == FRAME_QUEUED && ctx->frame_wave[r] >= 0 && ctx->frame_wave[r] + 1 > wv) wv = ctx->frame_wave[r] + 1;
            int rl = ctx->frame_lane[r];
            if (rl >= 0) { if (lane < 0) lane = rl; else if (rl != lane) foreign[k] |= 1u << rl; }
        }
        if (reader_wave[p->frame_id] + 1 > wv) wv = reader_wave[p->frame_id] + 1;
        int old = ctx->frame_lane[p->frame_id];
        if (lane < 0) lane = old >= 0 ? old : (ctx->next_lane++ % NL);
        else if (old >= 0 && old != lane) foreign[k] |= 1u << old;
        foreign[k] |= ctx->frame_reader_lanes[p->frame_id] & ~(1u << lane);
        ctx->frame_reader_lanes[p->frame_id] = 0;       // the slot holds a new picture from here on
        reader_wave[p->frame_id] = -1;
        for (int i = 0; i < p->pp.num_ref_idx_l0; i++) {
            int r = p->pp.ref_list0[i];
            if (wv > reader_wave[r]) reader_wave[r] = wv;
            ctx->frame_reader_lanes[r] |= 1u << lane;
        }
        p->wave = wv;
        p->lane = lane;
        ctx->frame_wave[p->frame_id] = wv;
        ctx->frame_lane[p->frame_id] = lane;
        if (wv + 1 > n_waves) n_waves = wv + 1;
    }
    const int n = (int)run.size();
    const DescTable d_desc{ctx->d_syntax + ctx->off_desc, ctx->slot_bytes};
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->main_dirty) {
        // something was queued on the main stream that the lanes must not overtake (read-backs, map resets)
        CU(cudaEventRecord(ctx->ev_main, ctx->stream));
        for (int l = 0; l < NL; l++) CU(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_main, 0));
        ctx->main_dirty = false;
    }
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    ctx->kev_used = 0;
    ctx->kev_family.clear();
    cudaStream_t st = nullptr;       // stream of the lane being issued
    cudaStream_t prof_st = nullptr;  // stream of the launch being timed (null: the lane's)
    auto prof_begin = [&](int family) {
        if (!ctx->profiling) return;
        if ((int)ctx->kev.size() < ctx->kev_used + 2) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            ctx->kev.push_back(a); ctx->kev.push_back(b);
        }
        ctx->kev_family.push_back(family);
        cudaEventRecord(ctx->kev[ctx->kev_used], prof_st ? prof_st : st);
    };
    auto prof_end = [&]() {
        if (!ctx->profiling) return;
        cudaEventRecord(ctx->kev[ctx->kev_used + 1], prof_st ? prof_st : st);
        ctx->kev_used += 2;
    };
    // hand-over words of one wavefront launch: a zeroed region of the ring (a region comes round again only after
    // far more pictures than can be in flight, see sync_ints)
    const size_t ring = ctx->sync_ints;
    // Regions are clean when handed out: the whole ring is zeroed at creation and what a flush used is zeroed again
    // behind its kernels (end of this function) -- no memset between two kernels of a lane, which would undo their
    // programmatic overlap (see launch_pdl).
    std::vector<std::pair<size_t, size_t>> sync_used;
    auto sync_region = [&](size_t ints) -> int* {
        if (ctx->sync_pos + ints > ring) ctx->sync_pos = 0;
        int* r = ctx->d_sync + ctx->sync_pos;
        if (!sync_used.empty() && sync_used.back().first + sync_used.back().second == ctx->sync_pos) sync_used.back().second += ints;
        else sync_used.push_back({ctx->sync_pos, ints});
        ctx->sync_pos += ints;
        return r;
    };
    // runs f(list) over the pictures of (wave, lane) selected by pred, in chunks that fit the kernel parameters
    PicList list, list_frames;       // arena slots of a chunk, and the frame ids of the same pictures
    int cur_wave = 0, cur_lane = 0, cur_wave_lo = -1;      // cur_wave_lo >= 0: waves cur_wave_lo .. cur_wave
    auto for_chunks = [&](auto pred, auto f) {
        list.n = 0;
        for (int i = 0; i < n; i++) {
            if (run[i]->wave > cur_wave || run[i]->wave < (cur_wave_lo >= 0 ? cur_wave_lo : cur_wave) || run[i]->lane != cur_lane || !pred(*run[i])) continue;
            list_frames.slot[list.n] = run[i]->frame_id;
            list.slot[list.n++] = run[i]->slot;
            if (list.n == PIC_LIST_MAX) { list_frames.n = list.n; f(list); list.n = 0; }
        }
        if (list.n) { list_frames.n = list.n; f(list); }
    };
    std::vector<uint8_t> lane_used(NL, 0), lane_late_waited(NL, 0);
    bool dig_used = false, late_scans = false;       // late_scans: the structures of the waves behind wave 0 come with ev_scan[5]
    auto sparse = [&](const Pending& p) { return ctx->sparse_intra && p.pp.slice_type == H264R_SLICE_P && p.n_intra > 0 && p.n_intra * 8 < ctx->n_mbs; };
    // Pictures whose upload has landed already (all of them in a resident batch): their index structures in one go, ahead
    // of every wave, and no per-wave wait later -- stream operations between the kernels of two waves would serialise
    // what launch_pdl lets overlap.
    {
        PicList EL;
        EL.n = 0;
        bool any = false;
        auto launch = [&]() {
            if (!EL.n) return;
            k_blk_scan<<<EL.n, 1024, 0, ctx->prep_stream>>>(d_desc, EL, ctx->n_mbs);
            ctx->last_launches++;
            if (ctx->any_stream) {
                k_stream_expand<<<dim3((unsigned)((ctx->n_mbs + 255) / 256), (unsigned)EL.n), 256, 0, ctx->prep_stream>>>(d_desc, EL, ctx->n_mbs);
                ctx->last_launches++;
            }
            any = true;
            EL.n = 0;
        };
        // uploads complete in submission order (one copy stream): a binary search over that order finds how many have landed
        std::vector<Pending*> by_seq(run.begin(), run.end());
        std::sort(by_seq.begin(), by_seq.end(), [](const Pending* x, const Pending* y) { return x->seq < y->seq; });
        int lo = 0, hi = n;        // pictures [0, lo) have landed, [hi, n) have not
        bool first_probe = true;
        while (lo < hi) {
            const int mid = first_probe ? hi - 1 : (lo + hi) / 2;       // first probe: the last picture (resident batches end here)
            first_probe = false;
            if (cudaEventQuery(ctx->ev_copied[by_seq[mid]->slot]) == cudaSuccess) lo = mid + 1; else hi = mid;
        }
        for (int i = 0; i < n; i++) by_seq[i]->early = i < lo;
        // the pictures of wave 0 first: the lanes start as soon as THEIR structures are made, and the scans of the later waves
        // (an expansion of the whole picture in the stream form) run beside wave 0
        for (int pass = 0; pass < 2; pass++) {
            any = false;
            for (int i = 0; i < n; i++) {
                Pending* p = run[i];
                if (p->early && (p->wave == 0) == (pass == 0) && (p->layout == 2 || sparse(*p))) {
                    EL.slot[EL.n++] = p->slot;
                    if (EL.n == PIC_LIST_MAX) launch();
                }
            }
            launch();
            (void)cudaGetLastError();       // cudaErrorNotReady of the queries is not an error
            if (any) {
                CU(cudaEventRecord(ctx->ev_scan[pass ? 5 : 0], ctx->prep_stream));
                if (pass == 0) for (int l = 0; l < NL; l++) CU(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_scan[0], 0));
                else late_scans = true;
            }
        }
    }
    const int dig_every = ctx->digest_every;
    for (int wv = 0; wv < n_waves; wv++) {
        for (int lane = 0; lane < NL; lane++) {
            cur_wave = wv; cur_lane = lane;
            st = ctx->lane_stream[lane];
            // uploads run on the copy stream in frame_end order: the lane waits for the last-ended picture of this
            // (wave, lane), so the upload of later pictures overlaps the kernels of earlier ones
            Pending* last = nullptr;       // last-ended picture of this (wave, lane) whose upload is still on its way
            bool any_pic = false;
            uint32_t wait_lanes = 0;
            for (int i = 0; i < n; i++)
                if (run[i]->wave == wv && run[i]->lane == lane) {
                    any_pic = true;
                    if (!run[i]->early && (!last || run[i]->seq > last->seq)) last = run[i];
                    wait_lanes |= foreign[i];
                }
            if (!any_pic) continue;
            lane_used[lane] = 1;
            if (wv > 0 && late_scans && !lane_late_waited[lane]) {
                CU(cudaStreamWaitEvent(st, ctx->ev_scan[5], 0));
                lane_late_waited[lane] = 1;
            }
            // block / vector offsets of the pictures that arrived packed: on the prep stream, behind the uploads and
            // ahead of the lanes (in a resident batch the scans of all waves are done while wave 0 computes)
            bool scanned = false;
            for_chunks([&](const Pending& p) { return !p.early && (p.layout == 2 || sparse(p)); }, [&](const PicList& L) {
                if (!scanned) cudaStreamWaitEvent(ctx->prep_stream, ctx->ev_copied[last->slot], 0);
                scanned = true;
                k_blk_scan<<<L.n, 1024, 0, ctx->prep_stream>>>(d_desc, L, ctx->n_mbs);
                ctx->last_launches++;
                if (ctx->any_stream) {
                    k_stream_expand<<<dim3((unsigned)((ctx->n_mbs + 255) / 256), (unsigned)L.n), 256, 0, ctx->prep_stream>>>(d_desc, L, ctx->n_mbs);
                    ctx->last_launches++;
                }
            });
            if (scanned) {
                CU(cudaEventRecord(ctx->ev_scan[1 + (lane & 3)], ctx->prep_stream));
                CU(cudaStreamWaitEvent(st, ctx->ev_scan[1 + (lane & 3)], 0));
            } else if (last) CU(cudaStreamWaitEvent(st, ctx->ev_copied[last->slot], 0));
            // a frame slot whose previous picture is still being digested (dig stream) must not be overwritten
            {
                uint64_t need = 0;
                for (int i = 0; i < n; i++)
                    if (run[i]->wave == wv && run[i]->lane == lane) need = std::max(need, ctx->frame_dig_seq[run[i]->frame_id]);
                if (need > ctx->lane_dig_waited[lane]) {
                    CU(cudaStreamWaitEvent(st, ctx->ev_dig[need & 63], 0));
                    ctx->lane_dig_waited[lane] = need;
                }
                // ... nor one that is still being copied to the host (h264r_read_frame_async)
                uint64_t need_out = 0;
                for (int i = 0; i < n; i++)
                    if (run[i]->wave == wv && run[i]->lane == lane) need_out = std::max(need_out, ctx->frame_out_seq[run[i]->frame_id]);
                if (need_out > ctx->lane_out_waited[lane]) {
                    // the ring holds the last 64 read-backs; an older one is covered by any later event of the same stream
                    const uint64_t ev = need_out + 64 <= ctx->out_seq ? ctx->out_seq : need_out;
                    CU(cudaStreamWaitEvent(st, ctx->ev_out[ev & 63], 0));
                    ctx->lane_out_waited[lane] = need_out;
                }
            }
            for (int l = 0; l < NL; l++)
                if ((wait_lanes >> l) & 1u) {           // references (or the frame slot's old readers) live on another lane
                    CU(cudaEventRecord(ctx->ev_lane[l], ctx->lane_stream[l]));
                    CU(cudaStreamWaitEvent(st, ctx->ev_lane[l], 0));
                }
            // lean pictures (picture buffers with partition-order vectors and compact levels, no weighting) and the others: see k_inter
            auto lean_pic = [&](const Pending& p) { return ctx->lean_inter && p.layout == 2 && p.mv_partition && p.level_bytes == H264R_LEVELS_COMPACT && !p.pp.weighted_pred; };
            for (int lean = 1; lean >= 0; lean--)
            for_chunks([&](const Pending& p) { return p.pp.slice_type == H264R_SLICE_P && p.n_intra < ctx->n_mbs && (int)lean_pic(p) == lean; }, [&](const PicList& L) {
                // grid.y = picture, grid.x strides over its macroblocks: inter_ctas_per_sm CTAs per SM (the register file holds
                // 8) times a few generations, i.e. every warp reconstructs a handful of macroblocks (its next header and
                // offsets are always in flight) and late CTAs even out the tail.
                int want = (ctx->n_mbs + KINTER_WARPS - 1) / KINTER_WARPS;
                int cap = (sm_count * ctx->inter_ctas_per_sm * (4 / KINTER_WARPS) * ctx->inter_generations + L.n - 1) / L.n;      // inter_ctas_per_sm counts four warps
                dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)L.n);
                prof_begin(0);
                const bool tma = ctx->d_tmaps != nullptr && ctx->tma_old;
                const dim3 blk(KINTER_WARPS * 32);
                if (lean && compat && !tma && ctx->inter_group) {
                    // group kernel: one warp per CTA, GRP_MBS macroblocks per step, inter_group_gens CTA generations
                    const int groups = (ctx->n_mbs + GRP_MBS - 1) / GRP_MBS;
                    const int gcap = (sm_count * H264R_GRP_MIN_CTAS * ctx->inter_group_gens + L.n - 1) / L.n;
                    const dim3 ggrid((unsigned)(groups < gcap ? groups : gcap), (unsigned)L.n);
                    if (ctx->group_tma && ctx->d_tmaps) launch_pdl(k_inter_grp<true, true>, ggrid, dim3(32), st, pdl, d_desc, L, ctx->g);
                    else launch_pdl(k_inter_grp<true, false>, ggrid, dim3(32), st, pdl, d_desc, L, ctx->g);
                } else if (lean) {
                    if (compat) { if (tma) launch_pdl(k_inter<true, true, true>, grid, blk, st, pdl, d_desc, L, ctx->g); else launch_pdl(k_inter<true, false, true>, grid, blk, st, pdl, d_desc, L, ctx->g); }
                    else if (tma) launch_pdl(k_inter<false, true, true>, grid, blk, st, pdl, d_desc, L, ctx->g);
                    else if (ctx->inter_parts) {
                        // spec mode: three kernels that fit the instruction cache instead of one that does not (see PART in h264r_mc.cuh)
                        launch_pdl(k_inter<false, false, true, PART_LUMA4>, grid, blk, st, pdl, d_desc, L, ctx->g);
                        launch_pdl(k_inter<false, false, true, PART_LUMA8>, grid, blk, st, pdl, d_desc, L, ctx->g);
                        launch_pdl(k_inter<false, false, true, PART_CHROMA>, grid, blk, st, pdl, d_desc, L, ctx->g);
                        ctx->last_launches += 2;
                    } else launch_pdl(k_inter<false, false, true>, grid, blk, st, pdl, d_desc, L, ctx->g);
                } else {
                    if (compat) { if (tma) launch_pdl(k_inter<true, true, false>, grid, blk, st, pdl, d_desc, L, ctx->g); else launch_pdl(k_inter<true, false, false>, grid, blk, st, pdl, d_desc, L, ctx->g); }
                    else { if (tma) launch_pdl(k_inter<false, true, false>, grid, blk, st, pdl, d_desc, L, ctx->g); else launch_pdl(k_inter<false, false, false>, grid, blk, st, pdl, d_desc, L, ctx->g); }
                }
                prof_end();
                ctx->last_launches++;
            });
            // intra macroblocks.  Few and scattered (P pictures): one launch, a warp ticket per macroblock of the lists
            // k_blk_scan made.  Otherwise the residual pre-pass and the row wavefront.
            {
                SparseList SL;
                SL.n = 0; SL.cum[0] = 0;
                auto launch = [&]() {
                    if (!SL.n) return;
                    const int total = SL.cum[SL.n];
                    int blocks = std::min((total + ROW_WARPS - 1) / ROW_WARPS, sm_count * list_ctas_per_sm(compat));      // persistent warps
                    int* sync = sync_region(1);       // the ticket
                    ctx->sparse_stamp++;
                    prof_begin(1);
                    if (compat) launch_pdl(k_intra_list<true>, dim3(blocks), dim3(ROW_WARPS * 32), st, pdl, d_desc, SL, ctx->g, sync, ctx->sparse_stamp);
                    else launch_pdl(k_intra_list<false>, dim3(blocks), dim3(ROW_WARPS * 32), st, pdl, d_desc, SL, ctx->g, sync, ctx->sparse_stamp);
                    prof_end();
                    ctx->last_launches++;
                    SL.n = 0;
                };
                for (int i = 0; i < n; i++) {
                    if (run[i]->wave != wv || run[i]->lane != lane || !sparse(*run[i])) continue;
                    SL.slot[SL.n] = run[i]->slot;
                    SL.cum[SL.n + 1] = SL.cum[SL.n] + run[i]->n_intra;
                    if (++SL.n == SPARSE_PICS) launch();
                }
                launch();
            }
            for_chunks([&](const Pending& p) { return p.n_intra > 0 && !sparse(p); }, [&](const PicList& L) {
                int want = (ctx->n_mbs + INTER_WARPS * 32 - 1) / (INTER_WARPS * 32);       // 32 macroblocks per warp
                dim3 grid((unsigned)want, (unsigned)L.n);
                prof_begin(1);
                if (compat) k_residual_pics<true><<<grid, INTER_WARPS * 32, 0, st>>>(d_desc, L, ctx->g);
                else k_residual_pics<false><<<grid, INTER_WARPS * 32, 0, st>>>(d_desc, L, ctx->g);
                prof_end();
                ctx->last_launches++;
            });
            for_chunks([&](const Pending& p) { return p.n_intra > 0 && !sparse(p); }, [&](const PicList& L) {
                int rows = L.n * ctx->g.h_mbs, blocks = (rows + ROW_WARPS - 1) / ROW_WARPS;
                int* sync = sync_region((size_t)rows * 2 + 1);    // two progress words per row (one per warp of its pair), then the ticket
                prof_begin(1);
                if (compat) k_intra_rows<true><<<blocks, ROW_WARPS * 64, 0, st>>>(d_desc, L, ctx->g, sync, sync + rows * 2);
                else k_intra_rows<false><<<blocks, ROW_WARPS * 64, 0, st>>>(d_desc, L, ctx->g, sync, sync + rows * 2);
                prof_end();
                ctx->last_launches++;
            });
            for_chunks([&](const Pending& p) { return p.deblock_on; }, [&](const PicList& L) {
                {
                    int want = (ctx->n_mbs + 3) / 4;
                    int cap = (sm_count * 16 * 4 + L.n - 1) / L.n;
                    prof_begin(2);
                    k_deblock_prep<<<dim3((unsigned)(want < cap ? want : cap), (unsigned)L.n), 128, 0, st>>>(d_desc, L, ctx->g);
                    prof_end();
                    ctx->last_launches++;
                }
                int rows = L.n * ctx->g.h_mbs, blocks = (rows + ROW_WARPS - 1) / ROW_WARPS;
                int* sync = sync_region(1);       // the ticket
                prof_begin(2);
                k_deblock_rows<<<blocks, ROW_WARPS * 32, 0, st>>>(d_desc, L, ctx->g, ++ctx->mbox_stamp, sync);
                prof_end();
                ctx->last_launches++;
            });
            // border replication of the finished pictures.  REF_COMPAT: the warps of the border macroblocks wrote it
            // (phase_pad_mb_luma), nothing to launch.
            if (!compat)
                for_chunks([&](const Pending&) { return true; }, [&](const PicList& L) {
                    int rows = ctx->g.h_mbs * 16 + 2 * PAD_Y + 2 * (ctx->g.h_mbs * 8 + 2 * PAD_YC);
                    long warps = (long)L.n * rows;
                    long cap = (long)sm_count * 16;
                    int blocks = (int)((warps + 3) / 4 < cap ? (warps + 3) / 4 : cap);
                    prof_begin(3);
                    k_pad<<<blocks, 128, 0, st>>>(d_desc, L, ctx->g, 1, nullptr);
                    prof_end();
                    ctx->last_launches++;
                });
            // K5 as part of the reconstruction, on the dig stream: nothing in the lane waits for it
            if ((ctx->flags & H264R_DIGESTS) && (wv % dig_every == dig_every - 1 || wv == n_waves - 1)) {
                cur_wave_lo = wv - wv % dig_every;
                CU(cudaEventRecord(ctx->ev_wave[lane], st));
                CU(cudaStreamWaitEvent(ctx->dig_stream, ctx->ev_wave[lane], 0));
                prof_st = ctx->dig_stream;
                for_chunks([&](const Pending&) { return true; }, [&](const PicList&) {
                    const PicList& F = list_frames;      // frame ids of this chunk
                    k_zero_digests<<<1, 256, 0, ctx->dig_stream>>>(F, ctx->d_digests);
                    prof_begin(4);
                    k_digest<<<dim3(DIGEST_CTAS, F.n), 256, 0, ctx->dig_stream>>>(digest_pool(ctx), F, ctx->g, ctx->d_digests);
                    prof_end();
                    ctx->last_launches += 2;
                    ctx->dig_seq++;
                    cudaEventRecord(ctx->ev_dig[ctx->dig_seq & 63], ctx->dig_stream);
                    for (int i = 0; i < F.n; i++) { ctx->digest_valid[F.slot[i]] = 1; ctx->frame_dig_seq[F.slot[i]] = ctx->dig_seq; }
                });
                prof_st = nullptr;
                cur_wave_lo = -1;
                dig_used = true;
            }
        }
    }
    // join: the main stream (read-backs, timing, slot reuse) continues when every lane has finished this flush
    for (int l = 0; l < NL; l++)
        if (lane_used[l]) {
            CU(cudaEventRecord(ctx->ev_lane[l], ctx->lane_stream[l]));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_lane[l], 0));
        }
    if (dig_used) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_dig[ctx->dig_seq & 63], 0));      // the dig stream runs in order: its last launch
    for (auto& u : sync_used) CU(cudaMemsetAsync(ctx->d_sync + u.first, 0, u.second * sizeof(int), ctx->stream));
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaGetLastError());
    ctx->timing_valid = true;
    ctx->flush_seq++;
    CU(cudaEventRecord(ctx->ev_flush[ctx->flush_seq & 63], ctx->stream));
    std::vector<Pending> keep;
    for (auto& p : ctx->pending) {
        if (p.ended) {
            ctx->frame_state[p.frame_id] = FRAME_DONE;
            ctx->frame_wave[p.frame_id] = -1;
            ctx->slot_flush_seq[p.slot] = ctx->flush_seq;
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
    CU(cudaStreamSynchronize(ctx->copy_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return H264R_OK;
}

int h264r_set_lanes(h264r_ctx* ctx, int lanes)
{
    if (!ctx || lanes < 1) return fail(ctx, H264R_E_INVAL, "set_lanes: bad argument");
    int r = h264r_sync(ctx);
    if (r != H264R_OK) return r;
    for (int l = 0; l < ctx->n_lanes_created; l++) CU(cudaStreamSynchronize(ctx->lane_stream[l]));
    ctx->n_lanes = lanes < ctx->n_lanes_created ? lanes : ctx->n_lanes_created;
    std::fill(ctx->frame_lane.begin(), ctx->frame_lane.end(), -1);      // everything is complete: no stream order to preserve
    std::fill(ctx->frame_reader_lanes.begin(), ctx->frame_reader_lanes.end(), 0u);
    ctx->next_lane = 0;
    return H264R_OK;
}

int h264r_wait_uploads(h264r_ctx* ctx)
{
    if (!ctx) return H264R_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->copy_stream));
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

int h264r_frame_layout(h264r_ctx* ctx, h264r_frame_layout_t* out)
{
    if (!ctx || !out) return fail(ctx, H264R_E_INVAL, "frame_layout: bad argument");
    out->frame_bytes = ctx->frame_bytes;
    out->y_offset = (size_t)PAD_Y * ctx->g.pitch_y + PAD_X;
    out->u_offset = ctx->y_bytes + (size_t)PAD_YC * ctx->g.pitch_c + PAD_XC;
    out->v_offset = out->u_offset + ctx->c_bytes;
    out->y_pitch = ctx->g.pitch_y;
    out->c_pitch = ctx->g.pitch_c;
    out->width = ctx->g.w_mbs * 16;
    out->height = ctx->g.h_mbs * 16;
    return H264R_OK;
}

int h264r_read_frame_async(h264r_ctx* ctx, int frame_id, void* dst)
{
    if (!dst) return fail(ctx, H264R_E_INVAL, "read_frame_async: bad argument");
    int r = require_done(ctx, frame_id, "read_frame_async");       // flushes when the frame is still queued: its kernels are issued
    if (r != H264R_OK) return r;
    CU(cudaSetDevice(ctx->device));
    if (ctx->out_waited_flush != ctx->flush_seq) {
        // everything issued so far (ev_flush joins the lanes of the last flush) before the copy; later flushes run beside it
        CU(cudaStreamWaitEvent(ctx->out_stream, ctx->ev_flush[ctx->flush_seq & 63], 0));
        ctx->out_waited_flush = ctx->flush_seq;
    }
    CU(cudaMemcpyAsync(dst, ctx->d_frames + (size_t)frame_id * ctx->frame_bytes, ctx->frame_bytes, cudaMemcpyDeviceToHost, ctx->out_stream));
    ctx->out_seq++;
    CU(cudaEventRecord(ctx->ev_out[ctx->out_seq & 63], ctx->out_stream));
    ctx->frame_out_seq[frame_id] = ctx->out_seq;
    return H264R_OK;
}

int h264r_wait_readbacks(h264r_ctx* ctx)
{
    if (!ctx) return H264R_E_INVAL;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->out_stream));
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
    // K5 for the frames whose digest was not produced with their reconstruction (H264R_DIGESTS)
    PicList list;
    list.n = 0;
    auto run = [&]() {
        if (!list.n) return;
        k_zero_digests<<<1, 256, 0, ctx->stream>>>(list, ctx->d_digests);
        k_digest<<<dim3(DIGEST_CTAS, list.n), 256, 0, ctx->stream>>>(digest_pool(ctx), list, ctx->g, ctx->d_digests);
        list.n = 0;
    };
    for (int i = 0; i < n; i++) {
        if (ctx->digest_valid[frame_ids[i]]) continue;
        ctx->digest_valid[frame_ids[i]] = 1;
        list.slot[list.n++] = frame_ids[i];
        if (list.n == PIC_LIST_MAX) run();
    }
    run();
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(ctx->h_digests.data(), ctx->d_digests, (size_t)ctx->max_frames * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) memcpy(out + (size_t)i * 16, &ctx->h_digests[(size_t)frame_ids[i] * 4], 16);
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

int h264r_device_status(h264r_ctx* ctx, uint32_t* flags)
{
    if (!ctx || !flags) return H264R_E_INVAL;
    int r = h264r_flush(ctx);
    if (r != H264R_OK) return r;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->prep_stream));
    uint32_t v = 0;
    CU(cudaMemcpyAsync(&v, ctx->d_exc_counter + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_exc_counter + 1, 0, 4, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *flags = v;
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
    pd.level_fmt = H264R_LEVELS_INT16;
    // dense layout; the context's table covers n_mbs of a picture, larger calls get their own
    uint32_t* d_ref = nullptr;
    if (n_mbs <= ctx->n_mbs) { pd.blk_mask = ctx->d_dense_mask; pd.blk_off = ctx->d_dense_off; }
    else {
        std::vector<uint32_t> mo((size_t)2 * n_mbs);
        for (int a = 0; a < n_mbs; a++) { mo[a] = 0xFFFFFFu; mo[(size_t)n_mbs + a] = 24u * (uint32_t)a; }
        if (cudaMalloc(&d_ref, mo.size() * 4) != cudaSuccess) { cudaFree(d); return fail(ctx, H264R_E_NOMEM, "residual_mbs: out of device memory"); }
        cudaMemcpyAsync(d_ref, mo.data(), mo.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        pd.blk_mask = d_ref; pd.blk_off = d_ref + n_mbs;
    }
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
    cudaFree(d_ref);
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
