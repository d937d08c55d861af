// Timing probe: the sliding-window kernel's MMA issue pattern (30 collector-chained 128x32x16 MMAs per input row) in
// isolation and with the kernel's other activities switched on one by one: tcgen05.commit per row, epilogue-style
// tcgen05.ld traffic from 16 warps, bulk copies into the operand stages.  Prints clocks per MMA.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
#define MMA_SS(QUAL)                                                                                              \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16" QUAL         \
                 " [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory")
__device__ __forceinline__ void mma_plain(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(""); }
__device__ __forceinline__ void mma_fill(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::fill"); }
__device__ __forceinline__ void mma_use(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::use"); }
__device__ __forceinline__ void mma_last(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) { MMA_SS(".collector::a::lastuse"); }

// flags: 1 = commit twice per row, 2 = 16 warps of tcgen05.ld, 4 = bulk copies into the stages, 8 = no collector hints,
//        16 = one commit per row only
__global__ void __launch_bounds__(576) probe(int flags, int rows, const uint4* gsrc, long long* out_cycles, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                  // 8 stages x 16 KB
    unsigned char* sB = smem + 128 * 1024;     // 20 KB of weight blocks (20 x 1 KB)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 150 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    volatile int* done = reinterpret_cast<volatile int*>(bars + 10);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 150 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (0x2c00u + ((i * 2654435761u) >> 23)) & 0x3fffu;
        reinterpret_cast<uint32_t*>(smem)[i] = h | (h << 16);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (tid == 0) {
        for (int b = 0; b < 4; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bars[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        *done = 0;
    }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(s32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp < 16) {  // fill TMEM with finite values
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
        for (int c = 0; c < 128; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr + c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    if (warp == 17) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
        const uint32_t idesc = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = s32(sA), b0 = s32(sB);
        if (flags & 256) rows = rows;  // (placeholder)
        uint64_t bd[20];
#pragma unroll
        for (int i = 0; i < 20; ++i) bd[i] = mk_desc(b0 + ((flags & 256) ? (i % 10) : i) * 1024, 512, 128);
        const long long t0 = clock64();
        for (int r = 0; r < rows; ++r) {
            const uint32_t sbase = a0 + ((flags & 512) ? 0u : (uint32_t)(r & 7) * 16384);
            uint32_t dc[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) dc[i] = (flags & 64) ? tmem + (uint32_t)(i * 32) : tmem + (uint32_t)(((r - i) & 15) * 32);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ah = (flags & 128) ? mk_desc(sbase + ks * 8192, 2048, 128) : mk_desc(sbase + ks * 8192, 4096, 128);
                const uint64_t al = (flags & 128) ? mk_desc(sbase + ks * 8192 + 4096, 2048, 128) : mk_desc(sbase + ks * 8192 + 2048, 4096, 128);
                if (leader) {
                    if (flags & 8) {
#pragma unroll
                        for (int i = 0; i < 5; ++i) { mma_plain(dc[i], ah, bd[ks * 10 + 2 * i], idesc); mma_plain(dc[i], ah, bd[ks * 10 + 2 * i + 1], idesc); }
#pragma unroll
                        for (int i = 0; i < 5; ++i) mma_plain(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                    } else if (flags & 32) {  // tap-inner order: consecutive MMAs never share an accumulator
#pragma unroll
                        for (int i = 0; i < 5; ++i) { if (i == 0) mma_fill(dc[i], ah, bd[ks * 10 + 2 * i], idesc); else mma_use(dc[i], ah, bd[ks * 10 + 2 * i], idesc); }
#pragma unroll
                        for (int i = 0; i < 5; ++i) { if (i == 4) mma_last(dc[i], ah, bd[ks * 10 + 2 * i + 1], idesc); else mma_use(dc[i], ah, bd[ks * 10 + 2 * i + 1], idesc); }
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (i == 0) mma_fill(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                            else if (i == 4) mma_last(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                            else mma_use(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (i == 0) mma_fill(dc[i], ah, bd[ks * 10 + 2 * i], idesc); else mma_use(dc[i], ah, bd[ks * 10 + 2 * i], idesc);
                            if (i == 4) mma_last(dc[i], ah, bd[ks * 10 + 2 * i + 1], idesc); else mma_use(dc[i], ah, bd[ks * 10 + 2 * i + 1], idesc);
                        }
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (i == 0) mma_fill(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                            else if (i == 4) mma_last(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                            else mma_use(dc[i], al, bd[ks * 10 + 2 * i], idesc);
                        }
                    }
                }
            }
            if (leader && (flags & 1)) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(&bars[1])) : "memory");
                if (!(flags & 16)) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(&bars[2])) : "memory");
            }
        }
        if (leader) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s32(&bars[0])) : "memory");
            uint32_t ok = 0, spins = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok) : "r"(s32(&bars[0])), "r"(0) : "memory");
                if (++spins > (1u << 24)) break;
            }
            out_cycles[blockIdx.x] = clock64() - t0;
            *done = 1;
        }
        __syncwarp();
    } else if (warp == 16) {
        if ((flags & 4) && lane == 0) {  // 8 bulk copies of 2 KB per "row", as the producer does (no one waits on them)
            int r = 0;
            while (!*done && r < rows * 4) {
                const uint32_t dst = s32(sA) + (uint32_t)(r & 7) * 16384;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bars[3])), "r"(16384) : "memory");
                for (int q = 0; q < 8; ++q)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst + q * 2048),
                                 "l"(gsrc + ((r * 8 + q) & 4095) * 128), "r"(2048), "r"(s32(&bars[3])) : "memory");
                uint32_t ok = 0, spins = 0;
                while (!ok && spins < (1u << 20)) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                 : "=r"(ok) : "r"(s32(&bars[3])), "r"(r & 1) : "memory");
                    ++spins;
                }
                ++r;
            }
        }
    } else if (flags & 2) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        int it = 0;
        while (!*done && it < rows * 64) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + ((it * 8) & 511)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            acc ^= v[0] ^ v[7];
            // ~60 cycles of epilogue-like math between loads
#pragma unroll
            for (int k = 0; k < 16; ++k) acc = acc * 1664525u + 1013904223u;
            ++it;
        }
        if (acc == 0x12345u) *sink = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 17) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

int main() {
    long long* d_cycles; uint32_t* sink; uint4* gsrc;
    cudaMalloc(&d_cycles, 148 * 8); cudaMalloc(&sink, 4); cudaMalloc(&gsrc, 4096 * 2048); cudaMemset(gsrc, 0x2c, 4096 * 2048);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 151 * 1024);
    const int rows = 300;
    struct C { int flags; const char* what; } cases[] = {
        {0, "collector chain, no commits"}, {64, "  D ring of 5 fixed slots"}, {128, "  dense A (LBO 2048)"}, {256, "  10 distinct B blocks"}, {512, "  one A stage"}, {64|128|256|512, "  all four"}, {32, "collector chain, tap-inner order"}, {32 | 1 | 2 | 4, "tap-inner + commits + ld + copies"}, {8, "no collector hints, no commits"}, {1, "+ 2 commits per row"}, {1 | 16, "+ 1 commit per row"},
        {1 | 2, "+ 2 commits + 16 warps of tcgen05.ld"}, {1 | 4, "+ 2 commits + bulk copies into the stages"},
        {1 | 2 | 4, "+ commits + tcgen05.ld + bulk copies"}, {2, "tcgen05.ld only (no commits)"}, {4, "bulk copies only (no commits)"}};
    for (auto c : cases) {
        for (int rep = 0; rep < 2; ++rep) probe<<<148, 576, 151 * 1024>>>(c.flags, rows, gsrc, d_cycles, sink);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<long long> cyc(148);
        cudaMemcpy(cyc.data(), d_cycles, 148 * 8, cudaMemcpyDeviceToHost);
        long long mn = cyc[0], mx = cyc[0];
        for (long long v : cyc) { mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
        printf("%-48s: %6.1f clk/MMA (min CTA) %6.1f (max CTA)  [%s]\n", c.what, (double)mn / (rows * 30), (double)mx / (rows * 30), cudaGetErrorString(e));
        fflush(stdout);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
