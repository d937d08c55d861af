// Probe of the tcgen05 conventions the tensor-core conv kernel relies on (no GPU in the build container):
//   * shared-memory matrix descriptor for K-major, no-swizzle operands: core matrix = 8 rows x 16 B, rows of a core
//     matrix 16 B apart, 8-row groups SBO apart, the two K halves of a K=16 step LBO apart; start address only 16-B aligned
//   * instruction descriptor bit layout for kind::f16 (f16 x f16 -> f32)
//   * tcgen05.alloc / mma / commit / ld round trip and the lane<->thread mapping of tcgen05.ld.32x32b
// One MMA chain: D[128 x N] = sum over `ksteps` K=16 steps of A_k[128x16] * B_k[N x16]^T, checked on the host.
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ int g_flag;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;                // layout_type (bits 61-63) = 0: no swizzle
}

template <int N>
__global__ void __launch_bounds__(192) probe(const __half* A, const __half* B, float* D, int ksteps, uint32_t a_lbo,
                                             uint32_t a_sbo, uint32_t a_kstep_bytes, uint32_t a_shift, uint32_t b_lbo,
                                             uint32_t b_sbo, uint32_t b_kstep_bytes, int a_bytes, int b_bytes,
                                             uint32_t idesc) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;
    unsigned char* sB = smem + ((a_bytes + 127) / 128) * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + ((b_bytes + 127) / 128) * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_bytes / 2; i += blockDim.x) reinterpret_cast<__half*>(sA)[i] = A[i];
    for (int i = tid; i < b_bytes / 2; i += blockDim.x) reinterpret_cast<__half*>(sB)[i] = B[i];
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes -> async proxy (tensor core) reads
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s32(tmem_slot)), "r"(N < 32 ? 32 : N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 32) {  // one thread issues every MMA
        for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_desc(s32(sA) + a_shift + k * a_kstep_bytes, a_lbo, a_sbo);
            const uint64_t bd = make_desc(s32(sB) + k * b_kstep_bytes, b_lbo, b_sbo);
            const uint32_t acc = k > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(bar)) : "memory");
    }
    if (warp >= 2) {  // 4 epilogue warps: warp w may only touch TMEM lanes 32*(w%4) .. +31
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
            if (++spins > (1u << 22)) { g_flag = 1; break; }
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int q = warp & 3, lane = tid & 31;
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t r[8];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int j = 0; j < 8; ++j) D[(q * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
}

// element (row, k) of step s lives at: s*kstep + (k/8)*lbo + (row/8)*sbo + (row%8)*16 + (k%8)*2   [+ shift for A]
static size_t off(int s, int row, int k, uint32_t kstep, uint32_t lbo, uint32_t sbo) {
    return (size_t)s * kstep + (size_t)(k / 8) * lbo + (size_t)(row / 8) * sbo + (row % 8) * 16 + (k % 8) * 2;
}

template <int N>
int run(const char* name, int ksteps, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t a_shift, uint32_t b_lbo,
        uint32_t b_sbo, uint32_t b_kstep, bool bf16_idesc_should_fail) {
    const int M = 128;
    size_t a_bytes = off(ksteps - 1, M - 1, 15, a_kstep, a_lbo, a_sbo) + 2 + a_shift, b_bytes = off(ksteps - 1, N - 1, 15, b_kstep, b_lbo, b_sbo) + 2;
    a_bytes = (a_bytes + 255) / 256 * 256; b_bytes = (b_bytes + 255) / 256 * 256;
    std::vector<__half> hA(a_bytes / 2, __float2half(7.0f)), hB(b_bytes / 2, __float2half(-3.0f));  // filler must not matter
    std::vector<float> fa((size_t)ksteps * M * 16), fb((size_t)ksteps * N * 16), ref((size_t)M * N, 0.f);
    srand(1);
    for (int s = 0; s < ksteps; ++s) {
        for (int m = 0; m < M; ++m) for (int k = 0; k < 16; ++k) {
            float v = (float)((rand() % 17) - 8) / 8.f; fa[(s * M + m) * 16 + k] = v;
            hA[(off(s, m, k, a_kstep, a_lbo, a_sbo) + a_shift) / 2] = __float2half(v);
        }
        for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) {
            float v = (float)((rand() % 13) - 6) / 4.f; fb[(s * N + n) * 16 + k] = v;
            hB[off(s, n, k, b_kstep, b_lbo, b_sbo) / 2] = __float2half(v);
        }
    }
    // overlapping layouts (the conv kernel's shifted views) write the same bytes twice; re-read the image for the reference
    for (int s = 0; s < ksteps; ++s) {
        for (int m = 0; m < M; ++m) for (int k = 0; k < 16; ++k) fa[(s * M + m) * 16 + k] = __half2float(hA[(off(s, m, k, a_kstep, a_lbo, a_sbo) + a_shift) / 2]);
        for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) fb[(s * N + n) * 16 + k] = __half2float(hB[off(s, n, k, b_kstep, b_lbo, b_sbo) / 2]);
    }
    for (int s = 0; s < ksteps; ++s) for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float acc = 0; for (int k = 0; k < 16; ++k) acc += fa[(s * M + m) * 16 + k] * fb[(s * N + n) * 16 + k];
        ref[m * N + n] += acc;
    }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, a_bytes); cudaMalloc(&dB, b_bytes); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), a_bytes, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), b_bytes, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * N * 4);
    // instruction descriptor: c_format F32 (1<<4), a/b format F16 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (bf16_idesc_should_fail) idesc |= (1u << 7) | (1u << 10);
    size_t smem = ((a_bytes + 127) / 128) * 128 + ((b_bytes + 127) / 128) * 128 + 64;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int zero = 0; cudaMemcpyToSymbol(g_flag, &zero, 4);
    probe<N><<<1, 192, smem>>>(dA, dB, dD, ksteps, a_lbo, a_sbo, a_kstep, a_shift, b_lbo, b_sbo, b_kstep, (int)a_bytes, (int)b_bytes, idesc);
    cudaError_t e = cudaDeviceSynchronize();
    int flag = 0; cudaMemcpyFromSymbol(&flag, g_flag, 4);
    std::vector<float> hD(M * N);
    cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs(hD[i] - ref[i]); if (!(d <= 1e-3)) { if (bad < 4) printf("   [%d,%d] got %f want %f\n", i / N, i % N, hD[i], ref[i]); ++bad; } if (d > maxerr) maxerr = d; }
    printf("%-46s N=%3d ksteps=%d: %s, timeout=%d, mismatches %d/%d maxerr %.3g -> %s\n", name, N, ksteps, cudaGetErrorString(e), flag, bad, M * N, maxerr,
           (e == cudaSuccess && !flag && (bad == 0) != bf16_idesc_should_fail) ? "OK" : "FAILED");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return e == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    // dense canonical layout: k halves 128*16 B apart (A) / N*16 B apart (B), 8-row groups 128 B apart
    if (which == 0 || which == 1) if (run<32>("canonical dense, N=32", 1, 2048, 128, 4096, 0, 512, 128, 1024, false)) return 1;
    if (which == 0 || which == 2) if (run<32>("3 accumulated k-steps", 3, 2048, 128, 4096, 0, 512, 128, 1024, false)) return 1;
    // the conv kernel's A operand: a [plane][pixel][8 ch] image; a k-step = (tap, chunk) pair: second half LBO bytes away
    // (here: next tap row, 184 pixels further), m-tiles start at arbitrary 16-byte aligned pixel offsets
    if (which == 0 || which == 3) if (run<32>("shifted view: start +124 px, LBO = 184 px", 2, 184 * 16, 128, 2 * 184 * 16, 124 * 16, 512, 128, 1024, false)) return 1;
    if (which == 0 || which == 4) if (run<96>("N=96 (conv1: 3 taps x 32 filters)", 2, 2048, 128, 4096, 16, 96 * 16, 128, 2 * 96 * 16, false)) return 1;
    if (which == 0 || which == 5) if (run<32>("bf16 idesc on f16 data must NOT match", 1, 2048, 128, 4096, 0, 512, 128, 1024, true)) return 1;
    return 0;
}
