#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void bulk(const uint8_t* src, uint8_t* out)
{
    __shared__ alignas(128) uint8_t win[1024];
    __shared__ unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(win)), "l"(src), "r"(1024), "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t done = 0; int spin = 0;
    while (!done && spin < (1 << 20)) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        spin++;
    }
    for (int i = threadIdx.x; i < 1024; i += 32) out[i] = win[i];
}
__global__ void nothing(uint8_t* out) { out[threadIdx.x] = 7; }
int main()
{
    std::vector<uint8_t> h(4096);
    for (int i = 0; i < 4096; i++) h[i] = (uint8_t)(i * 3);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    uint8_t* dout; cudaMalloc(&dout, 1024);
    nothing<<<1, 32>>>(dout);
    printf("plain kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    bulk<<<1, 32>>>(d, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("bulk kernel: %s\n", cudaGetErrorString(e));
    uint8_t o[1024]; cudaMemcpy(o, dout, sizeof o, cudaMemcpyDeviceToHost);
    printf("%d %d %d expect 0 3 6\n", o[0], o[1], o[2]);
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
    printf("%s cc %d.%d driver %d runtime %d\n", p.name, p.major, p.minor, drv, rt);
    return 0;
}
