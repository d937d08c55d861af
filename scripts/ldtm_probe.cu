// Timing probe: tcgen05.ld (TMEM -> registers) throughput per SM for 4 / 8 / 16 warps, x8 and x32 shapes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__global__ void __launch_bounds__(512) probe(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t taddr = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t a = taddr + ((i * X) & 255);
        if (X == 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
            if ((i & 3) == 3) asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc ^= r[0] ^ r[7];
        } else {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(a) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc ^= r[0] ^ r[31];
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    const long long t1 = clock64();
    if (acc == 0x12345678u) *sink = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem_slot) : "memory");
}
int main() {
    long long* d; uint32_t* sink;
    cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 4096;
    for (int X : {8, 32})
        for (int warps : {1, 4, 8, 16}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (X == 8) probe<8><<<148, warps * 32>>>(iters, d, sink); else probe<32><<<148, warps * 32>>>(iters, d, sink);
            }
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<long long> c(148);
            cudaMemcpy(c.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)warps * iters * 32 * X * 4;
            printf("tcgen05.ld.32x32b.x%-2d %2d warps: %.1f clk per ld per warp, %.1f B/clk/SM  (%s)\n", X, warps, (double)c[0] / iters, bytes / c[0], cudaGetErrorString(e));
        }
    return 0;
}
