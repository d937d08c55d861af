// Timing probe for small-N tcgen05.mma on B200: what bounds a 128 x N x 16 f16 MMA when N is 16..64?
// (data is arbitrary -- only issue order, operand source and accumulator dependencies vary).  One CTA per SM, one elected
// thread issues `iters` MMAs, tcgen05.commit -> mbarrier, clock64 around issue+completion.  Prints cycles per MMA.
//   modes: see the table in main().
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
#define MMA_SS(QUAL)                                                                                              \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16" QUAL         \
                 " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),                                                         \
                 "l"(ad), "l"(bd), "r"(idesc), "r"(1u)                                                            \
                 : "memory")
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(""); }
__device__ __forceinline__ void mma_ss_fill(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::fill"); }
__device__ __forceinline__ void mma_ss_use(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::use"); }
__device__ __forceinline__ void mma_ss_last(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::lastuse"); }
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t bd, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(ta), "l"(bd), "r"(idesc), "r"(1u)
                 : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t ta, uint64_t sd) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(ta), "l"(sd) : "memory");
}

__global__ void __launch_bounds__(128) probe(int mode, int iters, int N, long long* out_cycles, int* out_flag) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                 // 96 KB of operand bytes
    unsigned char* sB = smem + 96 * 1024;     // 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 128 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 1024 / 4; i += blockDim.x) {  // small finite fp16 values
        uint32_t h = (0x2c00u + ((i * 2654435761u) >> 23)) & 0x3fffu;
        reinterpret_cast<uint32_t*>(smem)[i] = h | (h << 16);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    // zero the TMEM A region (columns 256..511) and accumulators so no NaN/Inf garbage is multiplied
    {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr + c), "r"(0x2c002c00u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    long long cycles = 0;
    int flag = 0;
    if (warp == 0) {
        // the whole warp runs the issue loops convergently (uniform datapath); one elected lane issues
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = s32(sA), b0 = s32(sB);
        const uint32_t b_lbo = N * 16;  // B: [2 k-halves][N][8] fp16
        uint64_t av[8], bv[15], asw[8], ac[8];
        uint32_t dv[6], tav[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            av[i] = mk_desc(a0 + i * 4096, 2048, 128, 0);            // dense canonical views
            ac[i] = mk_desc(a0 + i * 1984 + (i >= 6 ? 49152 : 0), 184 * 16, 128, 0);  // conv-like views (LBO = one padded row)
            asw[i] = mk_desc(a0 + (i & 3) * 16384 + (i >> 2) * 32, 16, 1024, 2);
            tav[i] = tmem + 256 + i * 8;
        }
#pragma unroll
        for (int i = 0; i < 15; ++i) bv[i] = mk_desc(b0 + i * 1024, b_lbo, 128, 0);
#pragma unroll
        for (int i = 0; i < 6; ++i) dv[i] = tmem + (uint32_t)i * N;
        const long long t0 = clock64();
        if (mode == 0) {
            for (int i = 0; i < iters; i += 30) {
#pragma unroll
                for (int u = 0; u < 30; ++u) if (leader) mma_ss(dv[0], av[0], bv[0], idesc);
            }
        } else if (mode == 1) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ss(dv[0], av[u & 7], bv[0], idesc);
            }
        } else if (mode == 2) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ss(dv[u % 6], av[u & 7], bv[0], idesc);
            }
        } else if (mode == 3) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ss(dv[u % 6], ac[u % 6], bv[0], idesc);
            }
        } else if (mode == 4) {
            for (int i = 0; i < iters; i += 30) {
#pragma unroll
                for (int u = 0; u < 30; ++u) if (leader) mma_ss(dv[u % 5], av[u / 10], bv[u % 10], idesc);
            }
        } else if (mode == 5) {
            for (int i = 0; i < iters; i += 30) {
#pragma unroll
                for (int u = 0; u < 30; ++u) if (leader) {
                    if (u % 10 == 0) mma_ss_fill(dv[u % 5], av[u / 10], bv[u % 10], idesc);
                    else if (u % 10 == 9) mma_ss_last(dv[u % 5], av[u / 10], bv[u % 10], idesc);
                    else mma_ss_use(dv[u % 5], av[u / 10], bv[u % 10], idesc);
                }
            }
        } else if (mode == 6) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ts(dv[0], tav[u & 7], bv[0], idesc);
            }
        } else if (mode == 7) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ts(dv[u % 6], tav[u & 7], bv[0], idesc);
            }
        } else if (mode == 8) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) cp_128x256b(tav[u & 7], av[u & 7]);
            }
        } else if (mode == 9) {  // sliding-window mix: per input row 4 cp + 30 TS MMAs on 5 accumulators
            for (int i = 0; i < iters; i += 30) {
#pragma unroll
                for (int c = 0; c < 4; ++c) if (leader) cp_128x256b(tav[c + ((i / 30) & 1) * 4], av[c]);
#pragma unroll
                for (int u = 0; u < 30; ++u) if (leader) mma_ts(dv[u % 5], tav[(u & 3) + ((i / 30) & 1) * 4], bv[u % 15], idesc);
            }
        } else if (mode == 10) {
            for (int i = 0; i < iters; i += 24) {
#pragma unroll
                for (int u = 0; u < 24; ++u) if (leader) mma_ss(dv[u % 6], asw[u & 7], bv[0], idesc);
            }
        } else if (mode == 11) {  // tile-inner order: (hh x6 tiles, hl x6, lh x6)
            for (int i = 0; i < iters; i += 18) {
#pragma unroll
                for (int term = 0; term < 3; ++term)
#pragma unroll
                    for (int t = 0; t < 6; ++t)
                        if (leader) mma_ss(dv[t], term == 2 ? ac[6 + (t & 1)] : ac[t], bv[term == 1 ? 8 : 0], idesc);
            }
        } else if (mode == 12) {  // the conv kernel's order today: per tile (hh, hl, lh) on one accumulator
            for (int i = 0; i < iters; i += 18) {
#pragma unroll
                for (int t = 0; t < 6; ++t)
#pragma unroll
                    for (int term = 0; term < 3; ++term)
                        if (leader) mma_ss(dv[t], term == 2 ? ac[6 + (t & 1)] : ac[t], bv[term == 1 ? 8 : 0], idesc);
            }
        }
        if (mode >= 100) {
#define GROUPS(NG)                                                                                          \
    for (int i = 0; i < iters; i += 30) {                                                                   \
        _Pragma("unroll") for (int u = 0; u < 30; ++u) {                                                    \
            if (leader) {                                                                                   \
                constexpr int n = NG;                                                                       \
                const int g = u / n, j = u % n;                                                             \
                if (n == 1) mma_ss(dv[u % 5], av[g & 7], bv[u % 10], idesc);                                \
                else if (j == 0) mma_ss_fill(dv[u % 5], av[g & 7], bv[u % 10], idesc);                      \
                else if (j == n - 1) mma_ss_last(dv[u % 5], av[g & 7], bv[u % 10], idesc);                  \
                else mma_ss_use(dv[u % 5], av[g & 7], bv[u % 10], idesc);                                   \
            }                                                                                               \
        }                                                                                                   \
    }
            switch (mode - 100) {
                case 1: { GROUPS(1) } break;
                case 2: { GROUPS(2) } break;
                case 3: { GROUPS(3) } break;
                case 5: { GROUPS(5) } break;
                case 6: { GROUPS(6) } break;
                case 10: { GROUPS(10) } break;
                case 15: { GROUPS(15) } break;
                case 30: { GROUPS(30) } break;
            }
        }
        __syncwarp();
        if (leader) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(bar)) : "memory");
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
            if (++spins > (1u << 24)) { flag = 1; break; }
        }
        cycles = clock64() - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (cycles != 0) out_cycles[blockIdx.x] = cycles;
    if (flag) *out_flag = 1;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    struct Case { int mode, N; const char* what; };
    const Case cases[] = {
        {0, 32, "SS N=32, one accumulator, one A view"},
        {1, 32, "SS N=32, one accumulator, A view varies"},
        {2, 32, "SS N=32, 6 rotating accumulators, A varies"},
        {3, 32, "SS N=32, 6 rot. acc., conv-like A view (LBO = 184 px row)"},
        {12, 32, "SS N=32, conv kernel order today (per tile hh,hl,lh)"},
        {11, 32, "SS N=32, tile-inner order (hh x6, hl x6, lh x6)"},
        {4, 32, "SS N=32, same A x10 on 5 accumulators, no hint"},
        {5, 32, "SS N=32, same A x10, collector::a fill/use/lastuse"},
        {6, 32, "TS N=32 (A in TMEM), one accumulator"},
        {7, 32, "TS N=32, 6 rotating accumulators"},
        {8, 32, "tcgen05.cp 128x256b alone (per cp)"},
        {9, 32, "sliding window mix: 4 cp + 30 TS MMAs per row (per MMA)"},
        {10, 32, "SS N=32, SW128 K-major A, 6 rot. acc."},
        {1, 16, "SS N=16 chain"}, {1, 48, "SS N=48 chain"}, {1, 64, "SS N=64 chain"}, {2, 64, "SS N=64 6 rot"},
        {1, 96, "SS N=96 chain"}, {1, 128, "SS N=128 chain"}, {1, 256, "SS N=256 chain"},
        {7, 64, "TS N=64 6 rot"}, {6, 128, "TS N=128 chain"}, {6, 256, "TS N=256 chain"},
        {5, 64, "SS N=64 collector x10"}, {4, 64, "SS N=64 same A x10 no hint"},
        {101, 32, "group 1 (no reuse)"}, {102, 32, "collector groups of 2"}, {103, 32, "collector groups of 3"}, {105, 32, "collector groups of 5"},
        {106, 32, "collector groups of 6"}, {110, 32, "collector groups of 10"}, {115, 32, "collector groups of 15"}, {130, 32, "collector groups of 30"},
        {102, 64, "N=64 groups of 2"}, {105, 64, "N=64 groups of 5"}, {110, 64, "N=64 groups of 10"}, {105, 96, "N=96 groups of 5"}, {103, 96, "N=96 groups of 3"},
    };
    long long* d_cycles; int* d_flag;
    cudaMalloc(&d_cycles, 148 * 8); cudaMalloc(&d_flag, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 129 * 1024);
    const int iters = 5400;  // multiple of 30 and 18
    int idx = 0;
    for (const Case& c : cases) {
        if (only >= 0 && idx++ != only) continue;
        if (c.N * 6 > 256 && (c.mode == 2 || c.mode == 7)) continue;
        cudaMemset(d_flag, 0, 4);
        for (int rep = 0; rep < 2; ++rep) probe<<<148, 128, 129 * 1024>>>(c.mode, iters, c.N, d_cycles, d_flag);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> cyc(148);
        int flag = 0;
        cudaMemcpy(cyc.data(), d_cycles, 148 * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
        long long mn = cyc[0], mx = cyc[0];
        for (long long v : cyc) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
        printf("%-62s N=%3d: %7.1f clk/op (min CTA) %7.1f (max CTA)  floor N/2=%d  %s%s\n", c.what, c.N, (double)mn / iters,
               (double)mx / iters, c.N / 2, cudaGetErrorString(e), flag ? " TIMEOUT" : "");
        fflush(stdout);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
