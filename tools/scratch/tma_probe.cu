// development probe: one warp loads a 32x22 u8 box with TMA and prints it
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, const void* gmap, int x, int y, uint8_t* out, int* status)
{
    const void* tmap = MODE == 0 ? (const void*)&pmap : gmap;
    __shared__ alignas(128) uint8_t win[32 * 22];
    __shared__ unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(32 * 22) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(win)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t done = 0; int spin = 0;
    while (!done && spin < (1 << 20)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        spin++;
    }
    if (threadIdx.x == 0) { status[0] = done; status[1] = spin; }
    for (int i = threadIdx.x; i < 32 * 22; i += 32) out[i] = win[i];
}
int main()
{
    const int pitch = 512, rows = 100;
    std::vector<uint8_t> h(pitch * rows);
    for (int r = 0; r < rows; r++) for (int c = 0; c < pitch; c++) h[r * pitch + c] = (uint8_t)(r * 7 + c);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    printf("entry point: %d %p %d\n", (int)e, fn, (int)qr);
    CUtensorMap m;
    cuuint64_t dims[2] = {pitch, rows}; cuuint64_t strides[1] = {pitch}; cuuint32_t box[2] = {32, 22}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    void* dm; cudaMalloc(&dm, 128); cudaMemcpy(dm, &m, 128, cudaMemcpyHostToDevice);
    uint8_t* dout; int* dst; cudaMalloc(&dout, 32 * 22); cudaMalloc(&dst, 8);
    if (getenv("MODE") && atoi(getenv("MODE")) == 1) probe<1><<<1, 32>>>(m, dm, 13, 5, dout, dst); else probe<0><<<1, 32>>>(m, dm, 13, 5, dout, dst);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    int st[2]; uint8_t o[32 * 22]; cudaMemcpy(st, dst, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, dout, sizeof o, cudaMemcpyDeviceToHost);
    printf("done=%d spins=%d first row: ", st[0], st[1]);
    for (int i = 0; i < 8; i++) printf("%d ", o[i]);
    printf(" expect %d %d ... row1: %d expect %d\n", (5 * 7 + 13) & 255, (5 * 7 + 14) & 255, o[32], (6 * 7 + 13) & 255);
    return 0;
}
