#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdint>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
template <int BYTES>
__global__ void probe(const __grid_constant__ CUtensorMap tensor_map, int x, int y, uint8_t* out)
{
    __shared__ alignas(128) uint8_t win[BYTES];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&win, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(win));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BYTES; i += 32) out[i] = win[i];
}
#include <cstdlib>
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BYTES>
void run(EncodeTiled enc, uint8_t* d, int pitch, int rows, CUtensorMapDataType dt, int esz, int bw, int bh, const char* name)
{
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)(pitch / esz), (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    uint8_t* dout; cudaMalloc(&dout, BYTES);
    probe<BYTES><<<1, 32>>>(m, (getenv("X") ? atoi(getenv("X")) : 16) / esz, 5, dout);
    cudaError_t e = cudaDeviceSynchronize();
    uint8_t o[4]; cudaMemcpy(o, dout, 4, cudaMemcpyDeviceToHost);
    printf("%s: encode %d kernel: %s  first bytes %d %d (expect %d %d)\n", name, (int)r, cudaGetErrorString(e), o[0], o[1], (5 * 7 + 16) & 255, (5 * 7 + 17) & 255);
    if (e != cudaSuccess) exit(0);
}
int main()
{
    const int pitch = 512, rows = 100;
    std::vector<uint8_t> h(pitch * rows);
    for (int r = 0; r < rows; r++) for (int c = 0; c < pitch; c++) h[r * pitch + c] = (uint8_t)(r * 7 + c);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    EncodeTiled enc = (EncodeTiled)fn;
    int which = getenv("W") ? atoi(getenv("W")) : 0;
    if (which == 0) run<64 * 8 * 4>(enc, d, pitch, rows, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, 64, 8, "int32 64x8");
    if (which == 1) run<64 * 8>(enc, d, pitch, rows, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 64, 8, "u8 64x8");
    if (which == 2) run<32 * 22>(enc, d, pitch, rows, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 32, 22, "u8 32x22");
    if (which == 3) run<16 * 14>(enc, d, pitch, rows, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 16, 14, "u8 16x14");
    if (which == 4) run<128 * 4>(enc, d, pitch, rows, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 128, 4, "u8 128x4");
    return 0;
}
