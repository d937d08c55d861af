// Tensor-core (tcgen05 / TMEM / TMA) implementations of the fused periodic-pad + Conv2D layer, and the data movers that let
// a whole U-Net run on them.
//
// Common ground.  With pixels as the M dimension a 'valid' conv is
//   D[m, co] = sum_{i,j,c} A[m + shift(i,j), c] * W[i,j,c,co]
// and the example nets have tiny N (6..32 filters).  Precision: operands are fp16 hi/lo splits of fp32 values
// (x = hi + lo, 22 significant bits); three MMAs per K step (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM -> ~1e-6
// relative error, inside the 1e-4 / 50-step gate.  Activations live in HBM already split and channel-blocked ("P layout"):
// [n][plane = 2*c8 + {hi,lo}][H + zero rows][Wp][8] fp16 with the periodic longitude halo (wpad columns each side)
// materialised by the PRODUCER's epilogue and zero rows stored beyond the poles, so a consumer stages rows with plain bulk /
// tensor-map copies and needs no wrap arithmetic.  A-operand views are K-major, no-swizzle UMMA descriptors into that
// image (8 pixels x 16 B core matrices; validated by scripts/umma_probe.cu on B200).
//
// Two kernels:
//  * conv_sw_kernel  (mode 1, default): sliding window over row strips -- one staged input row feeds all vertical taps
//    with the same A operand (collector reuse), one TMEM accumulator per output row in flight.  See the block comment
//    above the kernel; DESIGN.md 4.1 has the measurements that led to it.
//  * conv_tc_kernel  (mode 0, DLWP_TC_KERNEL=flat or geometries the sliding window rejects): flattened (row, column)
//    tiles, vertical taps folded into K (a tap is a +i*dil*Wp*16-byte offset of the A view), horizontal taps folded into N
//    (N = (j, co), shifted sum in the epilogue) or K; 192+ threads: warp 0 bulk-copy producer, warp 1 MMA issuer, 16
//    epilogue warps; smem stages full/empty, double-buffered TMEM accumulator sets.
//  * p_ew_kernel: MaxPooling2D(2) / UpSampling2D(2) / channel-window copies on P images.
#define DLWP_CONV_TU  // mbarrier helpers
#include <cuda_fp16.h>
#include <cuda_runtime.h>
namespace dlwp {
__device__ int g_tc_flags = 0;  // bit 0: mbarrier wait timed out
}
#define g_device_flags g_tc_flags
#include "internal.h"
#undef g_device_flags

#include <stdlib.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "conv_tc.h"

namespace dlwp {

constexpr int TC_SETS = 4;          // epilogue warp sets (each = 4 warps, one per TMEM lane quadrant)
constexpr int TC_THREADS = 64 + TC_SETS * 128;  // warp 0 producer, warp 1 MMA issuer, then the epilogue sets
constexpr int TC_HPAD = 12;        // zero rows stored above and below every P-layout plane (>= pad + rows per tile)
constexpr int TC_MAX_KSTEPS = 32;

struct TcParams {
    int N, H, W, Wp;          // source image; Wp = W + 2*wpad_in
    int R_out, Rin, MT, S;    // rows per tile, staged rows, M-tiles per tile, M-tile stride (128 - (KW-1)*D)
    int KW, D, pad_t;
    int Cout, NCOLS, CBLK, CSTRIDE;  // filters, MMA N, 8-filter blocks, TMEM columns per horizontal tap of a block
    int G, KS, NS;            // channel groups per tile, K steps per group, smem stages
    int NACC;                 // TMEM accumulator sets: 2 (MMA of tile t+1 overlaps the epilogue of tile t) or 1
    int planes_per_group;
    int tiles_per_sample, total_tiles;
    int row0, row1;           // output rows [row0, row1) this launch computes (latitude band); tiles start at row0
    uint32_t stage_bytes, stage_stride, plane_bytes, b_bytes;  // b_bytes: one (hi or lo) weight image
    int XL;                   // lanes a pixel reaches to its right: (KW-1)*D
    uint32_t idesc;
    int act;
    const float* bias;
    const __half* bimg;       // [hi image | lo image], each [G*KS*2 units][NCOLS][8]
    const __half* xp;         // source activation, P layout with TC_HPAD zero rows above/below each plane
    int planes_in;            // real planes of the source (the last channel group may be partial)
    float* y32; long long ys_n, ys_c, ys_h;
    __half* yp; int Wp_out, wpad_out, planes_out;
    TcKStep kst[TC_MAX_KSTEPS];  // [G][KS]
};

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // sm_100 descriptor version; layout_type 0 = no swizzle
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

struct alignas(16) Half8 {
    __half2 a, b, c, d;
};

// ===================================================================================================================
template <int KW>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stages = smem_raw;
    unsigned char* b_hi = stages + (size_t)p.NS * p.stage_stride;
    unsigned char* b_lo = b_hi + p.b_bytes;
    float* xch = reinterpret_cast<float*>(b_lo + p.b_bytes);  // mailbox [set][item of the set][4 quadrants][XL][(KW-1)*8]
    const int XQ = p.XL * (KW - 1) * 8;                       // floats one quadrant publishes per item
    const int items_per_set = (p.MT * p.CBLK + TC_SETS - 1) / TC_SETS;
    float* sbias = xch + (size_t)TC_SETS * items_per_set * 4 * XQ;  // [CBLK*8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + p.CBLK * 8);
    uint64_t* full = bars;             // [NS]
    uint64_t* empty = bars + 8;        // [NS]
    uint64_t* acc_full = bars + 16;    // [2]
    uint64_t* acc_empty = bars + 18;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ACC_COLS = p.MT * p.NCOLS;  // columns of one accumulator set (NACC * ACC_COLS <= 512)

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    for (uint32_t i = tid; i < 2 * p.b_bytes / 16; i += TC_THREADS)  // weight images -> smem (hi then lo, contiguous)
        reinterpret_cast<uint4*>(b_hi)[i] = reinterpret_cast<const uint4*>(p.bimg)[i];
    // stages start as zeros: planes of a partial last channel group are never loaded and meet zero weights (0 * 0, not
    // 0 * stale NaN)
    for (uint32_t i = tid; i < (uint32_t)p.NS * p.stage_stride / 16; i += TC_THREADS)
        reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < p.CBLK * 8; i += TC_THREADS) sbias[i] = (p.bias != nullptr && i < p.Cout) ? p.bias[i] : 0.f;
    fence_proxy_async();  // generic-proxy writes -> visible to the async proxy (bulk copies, tensor core)
    if (tid == 0) {
        for (int s = 0; s < p.NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], TC_SETS * 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int Halloc = p.H + 2 * TC_HPAD;

    if (warp == 0) {
        // =============================== producer: one bulk copy per plane ===============================
        // A tile's Rin rows of one plane are contiguous in the P layout (full-width rows, zero rows stored beyond the
        // poles), so a stage is planes_per_group copies of plane_bytes each -- large requests, not 16-byte tensor rows.
        if (lane == 0) {
            int idx = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n = tile / p.tiles_per_sample;
                const int y0 = p.row0 + (tile % p.tiles_per_sample) * p.R_out;
                for (int g = 0; g < p.G; ++g, ++idx) {
                    const int s = idx % p.NS;
                    const int plane0 = g * p.planes_per_group;
                    const int np = min(p.planes_per_group, p.planes_in - plane0);
                    mbar_wait(&empty[s], ((idx / p.NS) & 1) ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)np * p.plane_bytes);
                    unsigned char* dst = stages + (size_t)s * p.stage_stride;
                    for (int q = 0; q < np; ++q) {
                        const __half* src = p.xp + ((((size_t)n * p.planes_in + plane0 + q) * Halloc) +
                                                    (size_t)(y0 - p.pad_t + TC_HPAD)) * p.Wp * 8;
                        bulk_load(dst + (size_t)q * p.plane_bytes, src, p.plane_bytes, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // The whole warp runs the loop convergently (all values are warp-uniform, so the descriptor arithmetic stays on
        // the uniform datapath); one elected lane issues the tcgen05 instructions.  The MMAs are small (N = 32..96), so
        // the issue rate matters: per K step the descriptors cost two integer adds on precomputed 32-bit low words.
        const bool leader = elect_one();
        int idx = 0, it = 0;
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, sm_100 descriptor version
        const uint32_t b_lbo_field = ((p.NCOLS * 16u) >> 4) << 16;
        const uint32_t bhi16 = smem_u32(b_hi) >> 4, blo16 = smem_u32(b_lo) >> 4;
        const uint32_t plane16 = p.plane_bytes >> 4;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int ab = it % p.NACC;
            mbar_wait(&acc_empty[ab], ((it / p.NACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc_base = tmem + ab * ACC_COLS;
            for (int g = 0; g < p.G; ++g, ++idx) {
                const int s = idx % p.NS;
                mbar_wait(&full[s], (idx / p.NS) & 1);
                tc_fence_after();
                const uint32_t sbase16 = smem_u32(stages + (size_t)s * p.stage_stride) >> 4;
                for (int t = 0; t < p.MT; ++t) {
                    const uint32_t d = acc_base + t * p.NCOLS;
                    const uint32_t row16 = sbase16 + (uint32_t)(t * p.S);
                    for (int ks = 0; ks < p.KS; ++ks) {
                        const TcKStep k = p.kst[g * p.KS + ks];
                        const uint32_t a_lo = ((k.a_lbo >> 4) << 16) | (row16 + (k.a_off >> 4));
                        const uint32_t boff16 = (uint32_t)((g * p.KS + ks) * 2 * p.NCOLS);
                        const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | a_lo;
                        const uint64_t ad_lo = ((uint64_t)desc_hi << 32) | (a_lo + plane16);
                        const uint64_t bd_hi = ((uint64_t)desc_hi << 32) | (b_lbo_field | (bhi16 + boff16));
                        const uint64_t bd_lo = ((uint64_t)desc_hi << 32) | (b_lbo_field | (blo16 + boff16));
                        if (leader) {
                            umma_f16(d, ad_hi, bd_hi, p.idesc, (g | ks) != 0);
                            umma_f16(d, ad_hi, bd_lo, p.idesc, 1u);
                            umma_f16(d, ad_lo, bd_hi, p.idesc, 1u);
                        }
                    }
                }
                __syncwarp();
                if (leader) umma_commit(&empty[s]);  // stage s may be refilled once these MMAs have read it
            }
            if (leader) umma_commit(&acc_full[ab]);
        }
    } else {
        // =============================== epilogue (TC_SETS sets x 4 quadrant warps) ===============================
        // Work items (M tile t, filter block cb) are dealt round-robin to the sets; inside a set, the warp with hardware
        // index w owns TMEM lanes 32*(w%4)..+31.  out[p] = sum_j D[p + j*dil][(j, co)]: taps from the same quadrant come by
        // warp shuffle, the XL lanes that spill into the next quadrant through a shared-memory mailbox.  Two passes per tile
        // so that a set synchronises twice per tile instead of once per item: (1) every item's boundary taps -> mailbox,
        // barrier, (2) the sums, barrier.
        const int q = warp & 3, set = (warp - 2) >> 2;
        const int XL = p.XL;
        float* xset = xch + (size_t)set * items_per_set * 4 * XQ;
        const int nitems = p.MT * p.CBLK;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int ab = it % p.NACC;
            const int n = tile / p.tiles_per_sample;
            const int y0 = p.row0 + (tile % p.tiles_per_sample) * p.R_out;
            mbar_wait(&acc_full[ab], (it / p.NACC) & 1);
            tc_fence_after();
            // ---- pass 1: publish the taps the previous quadrant will need --------------------------------------------
            int li = 0;
            if (KW > 1)
            for (int item = set; item < nitems; item += TC_SETS, ++li) {
                const int t = item / p.CBLK, cb = item - t * p.CBLK;
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + ab * ACC_COLS + t * p.NCOLS + cb * KW * p.CSTRIDE;
                float d[KW > 1 ? KW - 1 : 1][8];
#pragma unroll
                for (int j = 1; j < KW; ++j) tmem_ld8(taddr + j * p.CSTRIDE, d[j - 1]);
                tmem_ld_wait();
                if (lane < XL) {
                    float* xb = xset + ((size_t)li * 4 + q) * XQ + lane * (KW - 1) * 8;
#pragma unroll
                    for (int j = 0; j < KW - 1; ++j) {
                        *reinterpret_cast<float4*>(xb + j * 8) = make_float4(d[j][0], d[j][1], d[j][2], d[j][3]);
                        *reinterpret_cast<float4*>(xb + j * 8 + 4) = make_float4(d[j][4], d[j][5], d[j][6], d[j][7]);
                    }
                }
            }
            if (KW > 1) named_bar_sync(1 + set, 128);
            // ---- pass 2: shifted sums, bias, activation, stores ---------------------------------------------------------
            li = 0;
            for (int item = set; item < nitems; item += TC_SETS, ++li) {
                const int t = item / p.CBLK, cb = item - t * p.CBLK;
                const int ml = q * 32 + lane;           // row of the M tile
                const int pos = t * p.S + ml;           // flattened (row, padded column) position in the tile
                const int r = pos / p.Wp, xq = pos - r * p.Wp;
                const int y = y0 + r;
                const bool valid = (ml < p.S) && (r < p.R_out) && (y < p.row1) && (xq < p.W);
                float d[KW][8];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + ab * ACC_COLS + t * p.NCOLS + cb * KW * p.CSTRIDE;
#pragma unroll
                for (int j = 0; j < KW; ++j) tmem_ld8(taddr + j * p.CSTRIDE, d[j]);
                tmem_ld_wait();
                const float* xn = xset + ((size_t)li * 4 + ((q + 1) & 3)) * XQ;
                float o[8];
#pragma unroll
                for (int ci = 0; ci < 8; ++ci) o[ci] = d[0][ci] + sbias[cb * 8 + ci];
#pragma unroll
                for (int j = 1; j < KW; ++j) {
                    const int sh = j * p.D;
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) d[j][ci] = __shfl_down_sync(0xffffffffu, d[j][ci], sh);
                    if (lane + sh >= 32) {  // the tap lives in the next quadrant's first lanes
                        const float* m = xn + (min(lane + sh - 32, XL - 1) * (KW - 1) + (j - 1)) * 8;
                        const float4 m0 = *reinterpret_cast<const float4*>(m), m1 = *reinterpret_cast<const float4*>(m + 4);
                        d[j][0] = m0.x; d[j][1] = m0.y; d[j][2] = m0.z; d[j][3] = m0.w;
                        d[j][4] = m1.x; d[j][5] = m1.y; d[j][6] = m1.z; d[j][7] = m1.w;
                    }
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) o[ci] += d[j][ci];
                }
                if (valid) {
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) o[ci] = (cb * 8 + ci < p.Cout) ? apply_act(o[ci], p.act) : 0.f;
                    if (p.y32 != nullptr) {
                        float* yb = p.y32 + (long long)n * p.ys_n + (long long)(cb * 8) * p.ys_c + (long long)y * p.ys_h + xq;
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci)
                            if (cb * 8 + ci < p.Cout) yb[(long long)ci * p.ys_c] = o[ci];
                    }
                    if (p.yp != nullptr) {
                        __half h[8], l[8];
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) {
                            h[ci] = __float2half_rn(o[ci]);
                            l[ci] = __float2half_rn(o[ci] - __half2float(h[ci]));
                        }
                        float amax = 0.f;
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) amax = fmaxf(amax, fabsf(o[ci]));
                        if (!(amax <= 65504.f)) atomicOr(&g_tc_flags, 2);  // outside the fp16 split's range (or NaN)
                        Half8 vh, vl;
                        vh.a = __halves2half2(h[0], h[1]); vh.b = __halves2half2(h[2], h[3]);
                        vh.c = __halves2half2(h[4], h[5]); vh.d = __halves2half2(h[6], h[7]);
                        vl.a = __halves2half2(l[0], l[1]); vl.b = __halves2half2(l[2], l[3]);
                        vl.c = __halves2half2(l[4], l[5]); vl.d = __halves2half2(l[6], l[7]);
                        const int Hout = p.H + 2 * TC_HPAD;
                        Half8* row_hi = reinterpret_cast<Half8*>(p.yp) +
                                        (((size_t)n * p.planes_out + cb * 2) * Hout + y + TC_HPAD) * p.Wp_out + xq + p.wpad_out;
                        Half8* row_lo = row_hi + (size_t)Hout * p.Wp_out;
                        row_hi[0] = vh;
                        row_lo[0] = vl;
                        if (xq < p.wpad_out) {            // periodic longitude halo of the NEXT layer, right side
                            row_hi[p.W] = vh;
                            row_lo[p.W] = vl;
                        }
                        if (xq >= p.W - p.wpad_out) {     // ... and left side
                            row_hi[-p.W] = vh;
                            row_lo[-p.W] = vl;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);
            if (KW > 1) named_bar_sync(1 + set, 128);  // the mailbox may be rewritten (next tile) only after every reader is done
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// fp32 (N,C,H,W) -> P layout with the periodic halo; one thread per (n, c8, y, padded x)
__global__ void __launch_bounds__(256) pack_state_kernel(const float* __restrict__ x, __half* __restrict__ yp, int N, int C,
                                                         int H, int W, int wpad, long long xs_n, long long xs_c,
                                                         long long xs_h, int row0, int rows) {
    const int Wp = W + 2 * wpad, C8 = (C + 7) / 8;
    const long long total = (long long)N * C8 * rows * Wp;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(idx % Wp);
        long long t = idx / Wp;
        const int y = row0 + (int)(t % rows);
        t /= rows;
        const int c8 = (int)(t % C8);
        const int n = (int)(t / C8);
        const int gx = wrap_index(xq - wpad, W);
        __half h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            const float v = c < C ? x[(long long)n * xs_n + (long long)c * xs_c + (long long)y * xs_h + gx] : 0.f;
            h[e] = __float2half_rn(v);
            l[e] = __float2half_rn(v - __half2float(h[e]));
            if (__hisinf(h[e]) || __hisnan(h[e])) atomicOr(&g_tc_flags, 2);  // |x| > 65504: outside the fp16 split's range
        }
        Half8 vh, vl;
        vh.a = __halves2half2(h[0], h[1]); vh.b = __halves2half2(h[2], h[3]);
        vh.c = __halves2half2(h[4], h[5]); vh.d = __halves2half2(h[6], h[7]);
        vl.a = __halves2half2(l[0], l[1]); vl.b = __halves2half2(l[2], l[3]);
        vl.c = __halves2half2(l[4], l[5]); vl.d = __halves2half2(l[6], l[7]);
        Half8* out = reinterpret_cast<Half8*>(yp);
        const long long Ha = H + 2 * TC_HPAD;
        out[(((long long)n * 2 * C8 + 2 * c8) * Ha + y + TC_HPAD) * Wp + xq] = vh;
        out[(((long long)n * 2 * C8 + 2 * c8 + 1) * Ha + y + TC_HPAD) * Wp + xq] = vl;
    }
}

// ===================================================================================================================
// Sliding-window kernel (mode 1).
//
// Measured on B200 (scripts/mma_rate_probe.cu, profiles/r01_mma_rate_probe.txt): a 128 x N x 16 f16 MMA with both operands
// in shared memory takes max(N/2, 32 + N/4) clocks -- for the N = 32..96 of these layers it is bound by the 4 KB A read,
// not by the tensor pipe.  With .collector::a::fill/use/lastuse consecutive MMAs that share their A operand run at
// ~19 clk (N = 32; tensor floor 16).  So the tile is laid out to make A sharing the common case:
//   * an M tile is 128 consecutive pixels of ONE padded row ("strip"; lane = x - x0).  The remainder strips of two
//     samples share a tile (lanes 0-63 / 64-127) when they fit, so W = 180 costs 1.5 tiles per row, not 2;
//   * a CTA walks a strip top to bottom.  Each padded input row is staged once (one bulk copy per plane and segment
//     into a ring of row stages) and feeds ALL vertical taps: row rp updates the accumulators of output rows
//     rp - i*dil, i = 0..KH-1, with the same A view and different weight blocks -> per K step the hi view is read once
//     for 2*KH MMAs (hi*hi, hi*lo) and the lo view once for KH MMAs;
//   * TMEM holds a ring of NACC accumulators (one output row x 128 lanes x NCOLS columns each); a row is committed to
//     the epilogue when its last tap has been issued.  Four epilogue warp sets take rows round robin.
// Horizontal taps live in N (few filters: the epilogue's shifted sum, as in the flattened kernel) or in K (A views
// shifted by j*dil pixels inside the staged row; N = filters).
// ===================================================================================================================
constexpr int SW_MAX_STAGES = 12;
constexpr int SW_MAX_ACC = 16;

struct SwParams {
    int N, H, W, Wp;
    int D, pad_t;
    int S, nfull, rem, pair;          // strips: valid outputs per full strip, full strips per row, remainder, pairing
    int units_per_group, nbands, RB, total_units;
    int row0, row1;
    int Cout, NCOLS, CBLK, CSTRIDE, XL;
    int KS, NS, NACC;
    int planes_in;
    uint32_t rowpitch, stage_stride, b_unit16, b_bytes;  // b_unit16: one (k step, tap, hi|lo) weight block in 16-byte units
    uint32_t idesc;
    int act;
    const float* bias;
    const __half* bimg;
    const __half* xp;
    int in_plane0, in_planes_total, out_plane0;  // channel windows of the source / destination P images
    int f32_C, f32_rawpitch, f32_segstride;  // F32IN: source channels, bytes of a raw channel row (W * 4), of a sample's rows
    int use_tma;  // 1: the producer stages rows with tensor-map loads (one per row and unit) instead of per-plane bulk copies
    int debug;  // DLWP_SW_DEBUG: 1 = epilogue only waits/arrives, 2 = issuer only commits (bottleneck triage; wrong results)
    float* y32; long long ys_n, ys_c, ys_h;
    __half* yp; int Wp_out, wpad_out, planes_out;
    TcKStep kst[TC_MAX_KSTEPS];
};

struct SwUnit {
    int n0, n1;      // samples of the two segments (n1 = -1: none)
    int x0;          // first padded column of the strip (both segments of a paired tile start at the same column)
    int nva, nvb;    // valid output lanes per segment
    int ya, yb;      // output rows [ya, yb)
    int paired;      // lanes 64.. belong to segment b
};

__host__ __device__ __forceinline__ bool sw_decode(const SwParams& p, int u, SwUnit& U) {
    const int band = u % p.nbands, su = u / p.nbands;
    const int g = su / p.units_per_group, k = su - g * p.units_per_group;
    U.ya = p.row0 + band * p.RB;
    U.yb = p.row1 < U.ya + p.RB ? p.row1 : U.ya + p.RB;
    U.n1 = -1; U.nvb = 0; U.paired = 0;
    if (!p.pair) {
        U.n0 = g; U.x0 = k * p.S; U.nva = k < p.nfull ? p.S : p.rem;
    } else if (k < 2 * p.nfull) {
        const int which = k / p.nfull;
        U.n0 = 2 * g + which; U.x0 = (k - which * p.nfull) * p.S; U.nva = p.S;
    } else {
        U.paired = 1;
        U.n0 = 2 * g; U.x0 = p.nfull * p.S; U.nva = p.rem;
        if (2 * g + 1 < p.N) { U.n1 = 2 * g + 1; U.nvb = p.rem; }
    }
    return U.n0 < p.N && U.ya < U.yb;
}

template <int COLL>  // 0: no collector hint, 1: fill, 2: use, 3: lastuse
__device__ __forceinline__ void umma_f16_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if constexpr (COLL == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (COLL == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else if constexpr (COLL == 3)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else
        umma_f16(tmem_d, adesc, bdesc, idesc, acc);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void p_unpack8(const uint4& h, const uint4& l, float (&v)[8]) {
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(hh[k]), b = __half22float2(ll[k]);
        v[2 * k] = a.x + b.x;      // exact: hi and lo are an exact split of an fp32 value
        v[2 * k + 1] = a.y + b.y;
    }
}
__device__ __forceinline__ void p_pack8(const float (&v)[8], uint4& h, uint4& l) {
    uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
    uint32_t* lp = reinterpret_cast<uint32_t*>(&l);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 hh = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        const float2 f = __half22float2(hh);
        hp[k] = *reinterpret_cast<const uint32_t*>(&hh);
        lp[k] = pack_half2(v[2 * k] - f.x, v[2 * k + 1] - f.y);
    }
}


// KH: kernel height (vertical taps), KW: horizontal taps summed by the epilogue (1 = folded into K), NC: filters per
// 8-filter block that exist (6: the single packed block of a 6-filter layer, else 8)
//
// ST: compile-time copy of the per-layer constants (0 / -1 = take the value from SwParams at run time).  The generic
// instance serves any layer; the benchmark nets' layers get instances with everything folded, which matters because the
// single MMA-issuing warp (and, for many filters, the epilogue's instruction issue) is the kernel's critical path.
template <int NCOLS_, int KS_, int D_, int CBLK_, int ACT_, int OUT_, int FULL_ = 0, int F32IN_ = 0>
struct SwStatic {
    static constexpr int NCOLS = NCOLS_, KS = KS_, D = D_, CBLK = CBLK_, ACT = ACT_, OUT = OUT_;  // OUT: 1 = P, 2 = fp32, 3 = both
    static constexpr int FULL = FULL_;  // 1: Cout == CBLK * NC, no partial filter block
    // 1: the source is the fp32 (N,C,H,W) state itself (first layer of a rollout): the producer stages raw fp32 rows and
    // two converter warps build the hi/lo A layout (periodic wrap and pole rows included) -- no P image of the state, no
    // pack kernel, no feedback copy written by the last layer
    static constexpr int F32IN = F32IN_;
};
constexpr int SW_CONV_WARPS = 2;   // converter warps of the F32IN variant
constexpr int SW_RAW_STAGES = 4;   // raw fp32 row stages
using SwGeneric = SwStatic<0, 0, 0, 0, -1, 0>;

template <int KH, int KW, int NC, class ST>
__global__ void __launch_bounds__(TC_THREADS + (ST::F32IN ? SW_CONV_WARPS * 32 : 0), 1)
conv_sw_kernel(const SwParams p, const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_pair) {
    constexpr int NTHREADS = TC_THREADS + (ST::F32IN ? SW_CONV_WARPS * 32 : 0);
    const int NCOLS = ST::NCOLS ? ST::NCOLS : p.NCOLS;
    const int KS = ST::KS ? ST::KS : p.KS;
    const int D = ST::D ? ST::D : p.D;
    const int CBLK = ST::CBLK ? ST::CBLK : p.CBLK;
    const int NACC = ST::NCOLS ? (512 / (ST::NCOLS ? ST::NCOLS : 1) > SW_MAX_ACC ? SW_MAX_ACC : 512 / (ST::NCOLS ? ST::NCOLS : 1)) : p.NACC;
    const int act = ST::ACT >= 0 ? ST::ACT : p.act;
    const bool has_yp = ST::OUT ? (ST::OUT & 1) != 0 : p.yp != nullptr;
    const bool has_y32 = ST::OUT ? (ST::OUT & 2) != 0 : p.y32 != nullptr;
    constexpr int CSTRIDE = NC;  // 6: the single packed block of a 6-filter layer, else 8
    const uint32_t unit16 = (uint32_t)(2 * NCOLS);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stages = smem_raw;
    unsigned char* bsm = stages + (size_t)p.NS * p.stage_stride;
    float* xch = reinterpret_cast<float*>(bsm + p.b_bytes);  // mailbox [set][parity][block][quadrant][XL][(KW-1)*8]
    const int XLc = (KW - 1) * D;
    const int XQ = XLc * (KW - 1) * 8;
    float* sbias = xch + (size_t)TC_SETS * 2 * CBLK * 4 * XQ;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + CBLK * 8);
    uint64_t* full = bars;                                  // [NS]       bulk copies of a row landed
    uint64_t* empty = bars + SW_MAX_STAGES;                 // [NS]       the row's MMAs have read the stage
    uint64_t* acc_full = bars + 2 * SW_MAX_STAGES;          // [NACC]     output row complete in TMEM
    uint64_t* acc_empty = acc_full + SW_MAX_ACC;            // [NACC]     epilogue done with the accumulator
    uint64_t* raw_full = acc_empty + SW_MAX_ACC;            // [SW_RAW_STAGES]  F32IN: raw fp32 row landed
    uint64_t* raw_empty = raw_full + SW_RAW_STAGES;         // [SW_RAW_STAGES]  F32IN: converters done with the raw row
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + SW_RAW_STAGES);
    unsigned char* raw = reinterpret_cast<unsigned char*>(tmem_slot + 4);  // F32IN: [stage][segment][channel][W] fp32
    raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);

    // Warp roles: warps 0-15 epilogue (set = warp / 4, TMEM lane quadrant = warp % 4), warp 16 producer, warp 17 MMA
    // issuer.  The two single-warp roles sit on the highest warp ids: the issuer's serial instruction stream is the
    // critical path (ncu: with it on warp 1 the epilogue warps waited on acc_full 36 % of the time while the issuer never
    // waited on a barrier), and the scheduler favours higher warp ids among eligible warps.
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int W_PROD = TC_SETS * 4, W_MMA = TC_SETS * 4 + 1;
    const int SPAN = (KH - 1) * D;

    for (uint32_t i = tid; i < p.b_bytes / 16; i += NTHREADS)
        reinterpret_cast<uint4*>(bsm)[i] = reinterpret_cast<const uint4*>(p.bimg)[i];
    for (uint32_t i = tid; i < (uint32_t)p.NS * p.stage_stride / 16; i += NTHREADS)  // lanes past a row's end stay finite
        reinterpret_cast<uint4*>(stages)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < CBLK * 8; i += NTHREADS) sbias[i] = (p.bias != nullptr && i < p.Cout) ? p.bias[i] : 0.f;
    fence_proxy_async();
    if (tid == 0) {
        // F32IN: a stage is filled by the converter threads (generic stores), not by a bulk copy
        for (int s = 0; s < p.NS; ++s) { mbar_init(&full[s], ST::F32IN ? SW_CONV_WARPS * 32 : 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        if (ST::F32IN)
            for (int s = 0; s < SW_RAW_STAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], SW_CONV_WARPS * 32); }
        fence_mbar_init();
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int Halloc = p.H + 2 * TC_HPAD;

    if (ST::F32IN && warp == W_PROD) {
        // =============================== producer, fp32-state flavour ===================================================
        // One tiled load per input row and sample: box (W floats, 1 row, C channels) of the (N*C, H, W) fp32 state; rows
        // beyond the poles are out of bounds -> zero filled (ZeroPadding2D).  map_full is the fp32 map here.
        if (lane == 0) {
            prefetch_tensormap(&map_full);
            int rs = 0;
            uint32_t rph = 0;
            const uint32_t seg_bytes = (uint32_t)p.f32_C * (uint32_t)p.f32_rawpitch, seg_stride = (uint32_t)p.f32_segstride;
            SwUnit U;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                if (!sw_decode(p, u, U)) continue;
                const int nrows = U.yb - U.ya + SPAN;
                const int nseg = U.n1 >= 0 ? 2 : 1;
                int row = U.ya - p.pad_t;  // unpadded source row, may lie beyond the poles
                for (int r = 0; r < nrows; ++r, ++row) {
                    mbar_wait_relaxed(&raw_empty[rs], rph ^ 1);
                    mbar_expect_tx(&raw_full[rs], seg_bytes * nseg);
                    unsigned char* dst = raw + (size_t)rs * 2 * seg_stride;
                    tma_load_3d(dst, &map_full, &raw_full[rs], 0, row, U.n0 * p.f32_C);
                    if (nseg == 2) tma_load_3d(dst + seg_stride, &map_full, &raw_full[rs], 0, row, U.n1 * p.f32_C);
                    if (++rs == SW_RAW_STAGES) { rs = 0; rph ^= 1; }
                }
            }
        }
    } else if (ST::F32IN && warp >= W_MMA + 1) {
        // =============================== converters (F32IN): raw fp32 row -> hi/lo A layout ===============================
        // Thread t owns lanes t and t + 64 of the 128-lane tile (in a paired tile: the same column of the two samples).
        // Lane l is padded column x0 + l = source column (x0 + l - wpad) mod W: the periodic wrap is index arithmetic here.
        const int t = tid - (W_MMA + 1) * 32;
        const int wpad = (p.Wp - p.W) / 2;
        const uint32_t seg_stride = (uint32_t)p.f32_segstride;
        int s = 0, rs = 0;
        uint32_t ph = 0, rph = 0;
        SwUnit U;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            if (!sw_decode(p, u, U)) continue;
            const int nrows = U.yb - U.ya + SPAN;
            int col[2], seg[2];
            bool live[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int ml = t + 64 * k;
                seg[k] = (U.paired && ml >= 64) ? 1 : 0;
                const int l = ml - 64 * seg[k];
                const int xpad = U.x0 + l;
                live[k] = xpad < p.Wp && (seg[k] ? U.n1 : U.n0) >= 0;
                int c = xpad - wpad;
                if (c < 0) c += p.W;
                if (c >= p.W) c -= p.W;
                col[k] = live[k] ? c : 0;
            }
            for (int r = 0; r < nrows; ++r) {
                mbar_wait_relaxed(&raw_full[rs], rph);
                mbar_wait_relaxed(&empty[s], ph ^ 1);
                const unsigned char* rbase = raw + (size_t)rs * 2 * seg_stride;
                unsigned char* sbase = stages + (size_t)s * p.stage_stride;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float* src = reinterpret_cast<const float*>(rbase + (size_t)seg[k] * seg_stride) + col[k];
                    float v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        v[c] = (live[k] && c < p.f32_C) ? src[(size_t)c * (p.f32_rawpitch >> 2)] : 0.f;
                    float amax = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) amax = fmaxf(amax, fabsf(v[c]));
                    if (!(amax <= 65504.f)) atomicOr(&g_tc_flags, 2);  // outside the fp16 split's range (or NaN)
                    uint4 vh, vl;
                    p_pack8(v, vh, vl);
                    uint4* dst = reinterpret_cast<uint4*>(sbase) + (t + 64 * k);   // plane 0 (hi), lane
                    dst[0] = vh;
                    dst[p.rowpitch >> 4] = vl;                                      // plane 1 (lo)
                }
                fence_proxy_async();            // generic-proxy stores -> visible to the tensor core's async-proxy reads
                mbar_arrive(&full[s]);
                mbar_arrive(&raw_empty[rs]);
                if (++s == p.NS) { s = 0; ph ^= 1; }
                if (++rs == SW_RAW_STAGES) { rs = 0; rph ^= 1; }
            }
        }
    } else if (warp == W_PROD && p.use_tma) {
        // =============================== producer, tensor-map flavour ===================================================
        // One tiled load per input row and unit brings all planes of the strip: box (128 px, 1 row, planes) for a full
        // strip, (64 px, 1 row, 2 samples, planes) for a paired remainder strip -- the box order makes the stage layout
        // [plane][segment][pixel][8 ch] either way.  Columns past the padded row end and the absent partner sample of an odd
        // batch are out of bounds: zero filled (and counted by complete_tx).  ~15 instructions per row instead of ~10 per
        // plane: with per-plane bulk copies the producer's own instruction stream capped the load path at 4 TB/s.
        if (lane == 0) {
            prefetch_tensormap(&map_full);
            prefetch_tensormap(&map_pair);
            int s = 0;
            uint32_t ph = 0;
            const uint32_t row_bytes = (uint32_t)p.planes_in * p.rowpitch;
            SwUnit U;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                if (!sw_decode(p, u, U)) continue;
                const int nrows = U.yb - U.ya + SPAN;
                int row = U.ya - p.pad_t + TC_HPAD;
                const int plane = U.n0 * p.in_planes_total + p.in_plane0;
                for (int r = 0; r < nrows; ++r, ++row) {
                    mbar_wait_relaxed(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], row_bytes);
                    unsigned char* dst = stages + (size_t)s * p.stage_stride;
                    if (!U.paired) tma_load_3d(dst, &map_full, &full[s], U.x0 * 2, row, plane);
                    else tma_load_4d(dst, &map_pair, &full[s], U.x0 * 2, row, U.n0, p.in_plane0);
                    if (++s == p.NS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == W_PROD) {
        // =============================== producer: one bulk copy per (plane, segment) and input row =======================
        int s = 0;
        uint32_t ph = 0;  // parity of the stage ring's current lap
        SwUnit U;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            if (!sw_decode(p, u, U)) continue;
            const int nseg = U.n1 >= 0 ? 2 : 1;
            const int avail = p.Wp - U.x0;                                   // pixels left in the padded row
            const uint32_t lenA = (uint32_t)min(U.paired ? 64 : (int)(p.rowpitch >> 4), avail) * 16u;
            const uint32_t lenB = (uint32_t)min((int)(p.rowpitch >> 4) - 64, avail) * 16u;
            const uint32_t row_bytes = (uint32_t)p.planes_in * (lenA + (nseg == 2 ? lenB : 0u));
            const int ncp = p.planes_in * nseg;
            // this lane's copies: c = lane and c = lane + 32 (planes_in <= 32, two segments)
            const __half* src[2];
            uint32_t dsto[2], len[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = lane + 32 * k;
                const int seg = c >= p.planes_in ? 1 : 0;
                const int q = c - seg * p.planes_in;
                const int n = seg ? U.n1 : U.n0;
                src[k] = p.xp + ((((size_t)max(n, 0) * p.in_planes_total + p.in_plane0 + q) * Halloc +
                                  (size_t)(U.ya - p.pad_t + TC_HPAD)) * p.Wp + U.x0) * 8;
                dsto[k] = (uint32_t)q * p.rowpitch + (uint32_t)seg * 1024u;
                len[k] = seg ? lenB : lenA;
            }
            const size_t row_halfs = (size_t)p.Wp * 8;
            const int nrows = U.yb - U.ya + SPAN;
            for (int r = 0; r < nrows; ++r) {
                if (lane == 0) {
                    mbar_wait_relaxed(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], row_bytes);
                }
                __syncwarp();
                unsigned char* dst0 = stages + (size_t)s * p.stage_stride;
                if (lane < ncp) bulk_load(dst0 + dsto[0], src[0], len[0], &full[s]);
                if (lane + 32 < ncp) bulk_load(dst0 + dsto[1], src[1], len[1], &full[s]);
                src[0] += row_halfs;
                src[1] += row_halfs;
                if (++s == p.NS) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == W_MMA) {
        // =============================== MMA issuer (warp-convergent, one elected lane issues) ===========================
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, sm_100 descriptor version
        const uint32_t b_lbo_field = ((NCOLS * 16u) >> 4) << 16;
        const uint32_t b16 = smem_u32(bsm) >> 4;
        const uint32_t pitch16 = p.rowpitch >> 4;
        const uint32_t stages16 = smem_u32(stages) >> 4, stride16 = p.stage_stride >> 4;
                int s = 0;
        uint32_t ph = 0;
        int sl0 = 0;        // accumulator slot of the output row that starts at the current input row
        uint32_t aph = 0;   // parity of the accumulator ring's current lap
        SwUnit U;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            if (!sw_decode(p, u, U)) continue;
            const int nout = U.yb - U.ya;
            const int nrows = nout + SPAN;
            int sl = sl0;           // slot of (virtual) output row r; rows r >= nout are never started
            uint32_t ap = aph;
            uint32_t dh[(KH - 1) * (ST::D ? ST::D : 1) + 1];  // dh[k]: accumulator columns of output row r - k
#pragma unroll
            for (int k = 0; k <= (KH - 1) * (ST::D ? ST::D : 1); ++k) dh[k] = tmem;
            uint32_t dcur = tmem + (uint32_t)(sl0 * NCOLS);
            for (int r = 0; r < nrows; ++r) {
                if (r < nout) mbar_wait(&acc_empty[sl], ap ^ 1);  // output row r starts accumulating: slot must be drained
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sbase16 = stages16 + (uint32_t)s * stride16;
                uint32_t dcol[KH];
                if constexpr (ST::D != 0) {  // rolling register file of the live rows' accumulator columns: no index math
#pragma unroll
                    for (int k = (KH - 1) * (ST::D ? ST::D : 1); k > 0; --k) dh[k] = dh[k - 1];
                    dh[0] = dcur;
#pragma unroll
                    for (int i = 0; i < KH; ++i) dcol[i] = dh[i * (ST::D ? ST::D : 1)];
                } else {
#pragma unroll
                    for (int i = 0; i < KH; ++i) {
                        int si = sl - i * D;
                        if (si < 0) si += NACC;
                        dcol[i] = tmem + (uint32_t)(si * NCOLS);
                    }
                }
                const bool interior = (r >= SPAN) && (r < nout);
                if (p.debug & 2) {
                } else if (interior) {
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint32_t a_lo32 = p.kst[ks].a_off + sbase16;  // (LBO field | offset) precomputed on the host
                        const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | a_lo32;
                        const uint64_t ad_lo = ((uint64_t)desc_hi << 32) | (a_lo32 + pitch16);
                        const uint32_t bks = b_lbo_field | (b16 + (uint32_t)(ks * KH) * 2u * unit16);
#pragma unroll
                        for (int i = 0; i < KH; ++i) {
                            const uint64_t bh = ((uint64_t)desc_hi << 32) | (bks + (uint32_t)(2 * i) * unit16);
                            const uint64_t bl = ((uint64_t)desc_hi << 32) | (bks + (uint32_t)(2 * i + 1) * unit16);
                            if (leader) {
                                if (i == 0) umma_f16_c<1>(dcol[i], ad_hi, bh, p.idesc, ks != 0);
                                else umma_f16_c<2>(dcol[i], ad_hi, bh, p.idesc, 1u);
                                if (i == KH - 1) umma_f16_c<3>(dcol[i], ad_hi, bl, p.idesc, 1u);
                                else umma_f16_c<2>(dcol[i], ad_hi, bl, p.idesc, 1u);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < KH; ++i) {
                            const uint64_t bh = ((uint64_t)desc_hi << 32) | (bks + (uint32_t)(2 * i) * unit16);
                            if (leader) {
                                if (i == 0) umma_f16_c<1>(dcol[i], ad_lo, bh, p.idesc, 1u);
                                else if (i == KH - 1) umma_f16_c<3>(dcol[i], ad_lo, bh, p.idesc, 1u);
                                else umma_f16_c<2>(dcol[i], ad_lo, bh, p.idesc, 1u);
                            }
                        }
                    }
                } else {
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint32_t a_lo32 = p.kst[ks].a_off + sbase16;
                        const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | a_lo32;
                        const uint64_t ad_lo = ((uint64_t)desc_hi << 32) | (a_lo32 + pitch16);
                        const uint32_t bks = b_lbo_field | (b16 + (uint32_t)(ks * KH) * 2u * unit16);
#pragma unroll
                        for (int i = 0; i < KH; ++i) {
                            const int yo = r - i * D;  // output row (relative to the unit) this tap contributes to
                            if (yo < 0 || yo >= nout) continue;
                            const uint64_t bh = ((uint64_t)desc_hi << 32) | (bks + (uint32_t)(2 * i) * unit16);
                            const uint64_t bl = ((uint64_t)desc_hi << 32) | (bks + (uint32_t)(2 * i + 1) * unit16);
                            if (leader) {
                                umma_f16(dcol[i], ad_hi, bh, p.idesc, (i | ks) != 0);
                                umma_f16(dcol[i], ad_hi, bl, p.idesc, 1u);
                                umma_f16(dcol[i], ad_lo, bh, p.idesc, 1u);
                            }
                        }
                    }
                }
                __syncwarp();
                if (leader) {
                    umma_commit(&empty[s]);
                    if (r >= SPAN) {  // the last tap of output row r - SPAN has been issued
                        int sd = sl - SPAN;
                        if (sd < 0) sd += NACC;
                        umma_commit(&acc_full[sd]);
                    }
                }
                if (++s == p.NS) { s = 0; ph ^= 1; }
                dcur += (uint32_t)NCOLS;
                if (++sl == NACC) { sl = 0; ap ^= 1; dcur = tmem; }
            }
            // the next unit's first output row follows this unit's last one in the accumulator ring
            sl0 += nout;
            while (sl0 >= NACC) { sl0 -= NACC; aph ^= 1; }
        }
    } else {
        // =============================== epilogue: 4 sets x 4 quadrant warps, rows dealt round robin ===================
        // Output rows are numbered G = 0, 1, 2, ... across the units of this CTA; set s takes G = s, s + 4, ... so its
        // accumulator slot advances by 4 (mod NACC) per row, whatever the unit boundaries are.
        const int q = warp & 3, set = warp >> 2;
        const int XL = XLc;
        float* xset = xch + (size_t)set * 2 * CBLK * 4 * XQ;
        const int Hout = p.H + 2 * TC_HPAD;
        const size_t plane_stride = (size_t)Hout * p.Wp_out;  // uint4 units
        int slot = set;
        uint32_t aph = 0;
        while (slot >= NACC) { slot -= NACC; aph ^= 1; }
        int g = 0, lrow = 0;
        SwUnit U;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            if (!sw_decode(p, u, U)) continue;
            const int ml = q * 32 + lane;
            const int seg = (U.paired && ml >= 64) ? 1 : 0;
            const int l = ml - seg * 64;
            const int n = seg ? U.n1 : U.n0;
            const int x = U.x0 + l;
            const bool lane_ok = (n >= 0) && (l < (seg ? U.nvb : U.nva)) && (x < p.W);
            const bool halo_r = x < p.wpad_out, halo_l = x >= p.W - p.wpad_out;
            float* y32n = has_y32 ? p.y32 + (long long)max(n, 0) * p.ys_n + x : nullptr;
            uint4* ypn = has_yp
                             ? reinterpret_cast<uint4*>(p.yp) + (((size_t)max(n, 0) * p.planes_out + p.out_plane0) * Hout + TC_HPAD) * p.Wp_out + x + p.wpad_out
                             : nullptr;
            for (int y = U.ya + ((set - g) & 3); y < U.yb; y += TC_SETS, ++lrow) {
                mbar_wait_relaxed(&acc_full[slot], aph);
                tc_fence_after();
                if (p.debug & 1) {
                    tc_fence_before();
                    mbar_arrive(&acc_empty[slot]);
                    slot += TC_SETS;
                    if (slot >= NACC) { slot -= NACC; aph ^= 1; }
                    continue;
                }
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * NCOLS);
                float* mb = xset + (size_t)(lrow & 1) * CBLK * 4 * XQ;
                if (KW > 1) {
                    // ---- pass 1: the taps that the previous quadrant's last lanes need -> mailbox ----------------------
                    for (int cb = 0; cb < CBLK; ++cb) {
                        float d[KW > 1 ? KW - 1 : 1][8];
#pragma unroll
                        for (int j = 1; j < KW; ++j) tmem_ld8(tbase + (cb * KW + j) * CSTRIDE, d[j - 1]);
                        tmem_ld_wait();
                        if (lane < XL) {
                            float* xb = mb + ((size_t)cb * 4 + q) * XQ + lane * (KW - 1) * 8;
#pragma unroll
                            for (int j = 0; j < KW - 1; ++j) {
                                *reinterpret_cast<float4*>(xb + j * 8) = make_float4(d[j][0], d[j][1], d[j][2], d[j][3]);
                                *reinterpret_cast<float4*>(xb + j * 8 + 4) = make_float4(d[j][4], d[j][5], d[j][6], d[j][7]);
                            }
                        }
                    }
                    named_bar_sync(1 + set, 128);
                }
                // ---- pass 2: shifted sums, bias, activation, stores ---------------------------------------------------
                float* y32c = has_y32 ? y32n + (long long)y * p.ys_h : nullptr;
                uint4* row_hi = has_yp ? ypn + (size_t)y * p.Wp_out : nullptr;
#pragma unroll(ST::CBLK ? ST::CBLK : 1)
                for (int cb = 0; cb < CBLK; ++cb) {
                    float d[KW][8];
#pragma unroll
                    for (int j = 0; j < KW; ++j) tmem_ld8(tbase + (cb * KW + j) * CSTRIDE, d[j]);
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + cb * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + cb * 8 + 4);
                    tmem_ld_wait();
                    float o[8];
                    o[0] = d[0][0] + b0.x; o[1] = d[0][1] + b0.y; o[2] = d[0][2] + b0.z; o[3] = d[0][3] + b0.w;
                    o[4] = d[0][4] + b1.x; o[5] = d[0][5] + b1.y;
                    o[6] = NC > 6 ? d[0][6] + b1.z : 0.f; o[7] = NC > 6 ? d[0][7] + b1.w : 0.f;
                    if (KW > 1) {
                        const float* xn = mb + ((size_t)cb * 4 + ((q + 1) & 3)) * XQ;
#pragma unroll
                        for (int j = 1; j < KW; ++j) {
                            const int sh = j * D;
                            float v[8];
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) v[ci] = __shfl_down_sync(0xffffffffu, d[j][ci], sh);
                            if (lane + sh >= 32) {  // the tap lives in the next quadrant's first lanes
                                const float* m = xn + ((lane + sh - 32) * (KW - 1) + (j - 1)) * 8;
                                const float4 m0 = *reinterpret_cast<const float4*>(m);
                                const float4 m1 = *reinterpret_cast<const float4*>(m + 4);
                                v[0] = m0.x; v[1] = m0.y; v[2] = m0.z; v[3] = m0.w; v[4] = m1.x; v[5] = m1.y;
                                v[6] = m1.z; v[7] = m1.w;
                            }
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) o[ci] += v[ci];
                        }
                    }
                    if (lane_ok) {
                        if (act == DLWP_ACT_TANH) {
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) o[ci] = tanh_accurate(o[ci]);
                        } else if (act == DLWP_ACT_RELU) {
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) o[ci] = fmaxf(o[ci], 0.f);
                        }
                        const int nreal = ST::FULL ? NC : p.Cout - cb * 8;  // filters of this block that exist (>= NC: all)
                        if (nreal < NC) {
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci)
                                if (ci >= nreal) o[ci] = 0.f;
                        }
                        if (has_y32) {
                            float* yb = y32c + (long long)(cb * 8) * p.ys_c;
                            if (nreal >= NC) {
#pragma unroll
                                for (int ci = 0; ci < NC; ++ci) __stcs(yb + (long long)ci * p.ys_c, o[ci]);
                            } else {
#pragma unroll
                                for (int ci = 0; ci < NC; ++ci)
                                    if (ci < nreal) __stcs(yb + (long long)ci * p.ys_c, o[ci]);
                            }
                        }
                        if (has_yp) {
                            float amax = 0.f;
                            if (act != DLWP_ACT_TANH) {
                                amax = fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3])));
                                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o[4]), fabsf(o[5])), fmaxf(fabsf(o[6]), fabsf(o[7]))));
                            }
                            // outside the fp16 split's range (or NaN); tanh output cannot be (a NaN input was flagged upstream)
                            if (act != DLWP_ACT_TANH && !(amax <= 65504.f)) atomicOr(&g_tc_flags, 2);
                            uint4 vh, vl;
                            {
                                const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
                                const __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
                                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2),
                                             f3 = __half22float2(h3);
                                vh.x = *reinterpret_cast<const uint32_t*>(&h0); vh.y = *reinterpret_cast<const uint32_t*>(&h1);
                                vh.z = *reinterpret_cast<const uint32_t*>(&h2); vh.w = *reinterpret_cast<const uint32_t*>(&h3);
                                vl.x = pack_half2(o[0] - f0.x, o[1] - f0.y); vl.y = pack_half2(o[2] - f1.x, o[3] - f1.y);
                                vl.z = pack_half2(o[4] - f2.x, o[5] - f2.y); vl.w = pack_half2(o[6] - f3.x, o[7] - f3.y);
                            }
                            uint4* row_lo = row_hi + plane_stride;
                            __stcs(row_hi, vh);
                            __stcs(row_lo, vl);
                            if (halo_r) {            // periodic longitude halo of the NEXT layer, right side
                                __stcs(row_hi + p.W, vh);
                                __stcs(row_lo + p.W, vl);
                            }
                            if (halo_l) {            // ... and left side
                                __stcs(row_hi - p.W, vh);
                                __stcs(row_lo - p.W, vl);
                            }
                        }
                    }
                    if (has_yp) row_hi += 2 * plane_stride;
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[slot]);
                slot += TC_SETS;
                if (slot >= NACC) { slot -= NACC; aph ^= 1; }
            }
            g += U.yb - U.ya;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

// ===================================================================================================================
// Data movers on P images: MaxPooling2D(2) / UpSampling2D(2) / channel-window copy (skip connections of the U-Net).
// One thread per destination pixel and 8-channel chunk (a hi and a lo 16-byte vector); the destination's periodic halo
// columns are written with the interior.  HBM-bound byte movers: 32 B read (128 B for the pooling) + 32 B written per
// thread, coalesced along x.
// ===================================================================================================================
template <int KIND>  // 0 copy, 1 maxpool 2x2 stride 2 (floor), 2 nearest upsample x2
__global__ void __launch_bounds__(256) p_ew_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int C8, int Hs,
                                                  int Ws, int wpad_s, int src_plane0, int src_planes_total, int Hd, int Wd,
                                                  int wpad_d, int dst_plane0, int dst_planes_total, int row0, int rows) {
    const int Wps = Ws + 2 * wpad_s, Wpd = Wd + 2 * wpad_d;
    const long long Has = Hs + 2 * TC_HPAD, Had = Hd + 2 * TC_HPAD;
    const long long total = (long long)N * C8 * rows * Wd;  // destination rows [row0, row0 + rows): latitude-band window
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wd);
        long long t = idx / Wd;
        const int y = row0 + (int)(t % rows);
        t /= rows;
        const int c8 = (int)(t % C8);
        const int n = (int)(t / C8);
        const uint4* sh = src + (((long long)n * src_planes_total + src_plane0 + 2 * c8) * Has + TC_HPAD) * Wps + wpad_s;
        const uint4* sl = sh + Has * Wps;
        uint4 vh, vl;
        if (KIND == 0) {
            vh = sh[(long long)y * Wps + x];
            vl = sl[(long long)y * Wps + x];
        } else if (KIND == 2) {
            vh = sh[(long long)(y >> 1) * Wps + (x >> 1)];
            vl = sl[(long long)(y >> 1) * Wps + (x >> 1)];
        } else {
            float m[8];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long long o = (long long)(2 * y + dy) * Wps + 2 * x + dx;
                    float v[8];
                    p_unpack8(sh[o], sl[o], v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) m[e] = (dy | dx) ? fmaxf(m[e], v[e]) : v[e];
                }
            p_pack8(m, vh, vl);
        }
        uint4* dh = dst + (((long long)n * dst_planes_total + dst_plane0 + 2 * c8) * Had + TC_HPAD + y) * Wpd + wpad_d + x;
        uint4* dl = dh + Had * Wpd;
        dh[0] = vh;
        dl[0] = vl;
        if (x < wpad_d) { dh[Wd] = vh; dl[Wd] = vl; }
        if (x >= Wd - wpad_d) { dh[-Wd] = vh; dl[-Wd] = vl; }
    }
}

int tc_ew_launch(int kind, const __half* src, __half* dst, int N, int planes, int Hs, int Ws, int wpad_s, int src_plane0,
                 int src_planes_total, int wpad_d, int dst_plane0, int dst_planes_total, cudaStream_t stream, int row_begin,
                 int row_end) {
    const int C8 = planes / 2;
    int Hd = Hs, Wd = Ws;
    if (kind == DLWP_OP_MAXPOOL) { Hd = Hs / 2; Wd = Ws / 2; }
    if (kind == DLWP_OP_UPSAMPLE) { Hd = Hs * 2; Wd = Ws * 2; }
    const int row0 = (row_begin == 0 && row_end == 0) ? 0 : row_begin;
    const int rows = ((row_begin == 0 && row_end == 0) ? Hd : row_end) - row0;
    const long long total = (long long)N * C8 * rows * Wd;
    if (total <= 0) return 0;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    if (kind == DLWP_OP_MAXPOOL)
        p_ew_kernel<1><<<blocks, 256, 0, stream>>>(s4, d4, N, C8, Hs, Ws, wpad_s, src_plane0, src_planes_total, Hd, Wd, wpad_d, dst_plane0, dst_planes_total, row0, rows);
    else if (kind == DLWP_OP_UPSAMPLE)
        p_ew_kernel<2><<<blocks, 256, 0, stream>>>(s4, d4, N, C8, Hs, Ws, wpad_s, src_plane0, src_planes_total, Hd, Wd, wpad_d, dst_plane0, dst_planes_total, row0, rows);
    else
        p_ew_kernel<0><<<blocks, 256, 0, stream>>>(s4, d4, N, C8, Hs, Ws, wpad_s, src_plane0, src_planes_total, Hd, Wd, wpad_d, dst_plane0, dst_planes_total, row0, rows);
    return after_launch("p_ew_kernel");
}

// ===================================================================================================================
// Host side
// ===================================================================================================================
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static int g_tc_sms = 0;

bool tc_geometry_ok(const DlwpConvDesc& d) {
    if (d.rowwise || d.pre_op || d.dil_h != d.dil_w) return false;
    if (d.pad_mode_h != DLWP_PAD_ZERO || d.pad_mode_w != DLWP_PAD_PERIODIC) return false;
    if (d.kw != 3 && d.kw != 5) return false;
    const int halo_w = d.dil_w * (d.kw - 1), halo_h = d.dil_h * (d.kh - 1);
    if (d.pad_l != halo_w / 2 || d.pad_r != halo_w / 2 || (halo_w & 1)) return false;  // 'same' width
    if (d.pad_t + d.pad_b != halo_h) return false;                                      // 'same' height
    if (d.W < halo_w || d.W + halo_w > 1024) return false;                              // rows are staged whole
    if ((d.kw - 1) * d.dil_w > 8) return false;
    TcLayer L;
    return tc_plan_layer(d, &L) == 0;
}

// Sliding-window schedule of one layer (mode 1); -1 if the geometry does not fit (TMEM columns, shared memory, K steps).
static int sw_plan_layer(const DlwpConvDesc& d, TcLayer* L) {
    memset(L, 0, sizeof(*L));
    if (d.kh != d.kw || (d.kh != 3 && d.kh != 5)) return -1;
    const int halo_w = d.dil_w * (d.kw - 1), span = d.dil_h * (d.kh - 1);
    L->mode = 1;
    L->wpad = halo_w / 2;
    L->Wp = d.W + halo_w;
    L->C8 = cdiv(d.Cin, 8);
    L->planes = 2 * L->C8;
    if (L->planes > 32) return -1;  // the producer issues at most two copies per lane and row
    L->CBLK = cdiv(d.Cout, 8);
    // Many filters: N = filters is wide enough, fold the horizontal taps into K (shifted A views, no shifted sum in the
    // epilogue).  Few filters: fold them into N so that the MMA N reaches 32..96.
    L->taps_in_k = d.Cout > 16 ? 1 : 0;
    // ... unless there are many input channels and still few enough filters for the accumulator ring: with taps in K the
    // K loop is kw times longer (kw * C8 / 2 steps of N = Cout MMAs), and the single issuing warp is the bottleneck
    if (L->taps_in_k && L->C8 >= 8 && 512 / (cdiv(cdiv(d.Cout, 8) * d.kw * 8, 16) * 16) >= span + 2) L->taps_in_k = 0;
    const char* env_m = getenv("DLWP_TC_TAPS_IN_K");
    if (env_m) L->taps_in_k = atoi(env_m) ? 1 : 0;
    L->kw_eff = L->taps_in_k ? 1 : d.kw;
    L->CSTRIDE = (!L->taps_in_k && d.Cout == 6) ? 6 : 8;  // 5 taps x 6 filters pack into 32 columns
    L->NCOLS = cdiv(L->CBLK * L->kw_eff * L->CSTRIDE + (8 - L->CSTRIDE), 16) * 16;
    if (L->NCOLS > 256) return -1;
    L->NACC = std::min(SW_MAX_ACC, 512 / L->NCOLS);
    if (L->NACC < span + 2) return -1;  // rows in flight (span + 1) plus one being drained by the epilogue
    // Either way a strip yields 128 - halo_w outputs: with the taps in N the last halo_w lanes only feed the shifted sum,
    // with the taps in K the shifted A views of the last valid lane end at pixel 127 -- so the staged row is exactly 128
    // pixels per plane and one tensor-map box (<= 256 eight-byte elements) can bring it.
    L->XLK = 0;
    L->S = 128 - halo_w;
    L->nfull = d.W / L->S;
    L->rem = d.W - L->nfull * L->S;
    L->pair = (L->rem > 0 && L->rem + halo_w <= 64) ? 1 : 0;
    const int units = L->C8 * (L->taps_in_k ? d.kw : 1);
    L->KS = cdiv(units, 2);
    if (L->KS > TC_MAX_KSTEPS) return -1;
    L->G = 1; L->cpg = L->C8; L->MT = 1; L->R_out = 1; L->Rin = 1 + span;
    L->rowpitch = (uint32_t)(((128 + L->XLK) * 16 + 127) / 128 * 128);
    L->plane_bytes = L->rowpitch;
    L->stage_bytes = L->stage_stride = (uint32_t)L->planes * L->rowpitch;
    L->b_bytes = (uint32_t)(L->KS * d.kh * 2) * (uint32_t)(2 * L->NCOLS * 16);
    const size_t mailbox = L->kw_eff > 1 ? (size_t)TC_SETS * 2 * L->CBLK * 4 * halo_w * (L->kw_eff - 1) * 8 * 4 : 0;
    const size_t fixed = (size_t)L->b_bytes + mailbox + (size_t)L->CBLK * 32 + (2 * SW_MAX_STAGES + 2 * SW_MAX_ACC + 2 * SW_RAW_STAGES) * 8 + 128 + 1024;
    const size_t budget = 227 * 1024;
    if (fixed + L->stage_stride > budget) return -1;
    L->NS = (int)std::min<size_t>(SW_MAX_STAGES, (budget - fixed) / L->stage_stride);
    L->smem = (size_t)L->NS * L->stage_stride + fixed;
    return 0;
}

int tc_plan_layer(const DlwpConvDesc& d, TcLayer* L) {
    const char* env_k = getenv("DLWP_TC_KERNEL");  // "flat": the flattened-tile kernel only; default: sliding window first
    if (!(env_k && !strcmp(env_k, "flat")) && sw_plan_layer(d, L) == 0) return 0;
    memset(L, 0, sizeof(*L));
    const int halo_w = d.dil_w * (d.kw - 1), halo_h = d.dil_h * (d.kh - 1);
    L->wpad = halo_w / 2;
    L->Wp = d.W + halo_w;
    L->C8 = cdiv(d.Cin, 8);
    L->planes = 2 * L->C8;
    L->CBLK = cdiv(d.Cout, 8);
    L->CSTRIDE = d.Cout < 8 ? d.Cout : 8;  // a single partial block packs its taps tightly (5 taps x 6 filters -> 32 cols)
    // Few input channels (K small): fold the horizontal taps into K as well (K = (i, j, c), N = filters only).  The A view
    // of tap (i, j) is just +(i*dil*Wp + j*dil)*16 bytes, the epilogue needs no shifted sum, and N shrinks by kw.
    L->taps_in_k = (L->C8 * d.kh * d.kw <= 12) ? 1 : 0;
    const char* env_m = getenv("DLWP_TC_TAPS_IN_K");
    if (env_m) L->taps_in_k = atoi(env_m) ? 1 : 0;
    L->kw_eff = L->taps_in_k ? 1 : d.kw;
    L->NCOLS = cdiv(L->CBLK * L->kw_eff * L->CSTRIDE + (8 - L->CSTRIDE), 16) * 16;  // x8 loads of the last tap stay inside
    if (L->NCOLS > 256) return -1;
    L->S = L->taps_in_k ? 128 : 128 - halo_w;
    // units = (chunk, tap) in order; groups of chunks so that a group has an even number of units when possible
    const int smem_budget = 224 * 1024;
    const size_t bias_bars = (size_t)cdiv(d.Cout, 8) * 32 + 512;
    int best_r = 0, best_cpg = 0, best_ns = 0, best_mt = 0, best_nacc = 2;
    const char* env_r = getenv("DLWP_TC_ROUT");
    // big-K layers (>= 2 channel groups per tile) run best with small tiles: more tiles to balance, shorter TMA latency
    const int r_max = env_r ? atoi(env_r) : (L->C8 > 2 ? 4 : 8);
    // prefer two accumulator sets (the MMAs of tile t+1 overlap the epilogue of tile t); wide rows may only fit one
    for (int nacc = 2; nacc >= 1 && !best_r; --nacc)
    for (int r_out = r_max; r_out >= 1; --r_out) {
        const int rin = r_out + halo_h;
        const int mt = cdiv(r_out * L->Wp, L->S);
        if (mt * L->NCOLS * nacc > 512) continue;  // the accumulator sets must fit the 512 TMEM columns
        const size_t fixed = (size_t)TC_SETS * cdiv(mt * L->CBLK, TC_SETS) * 4 * halo_w * (L->kw_eff - 1) * 8 * 4 + bias_bars;
        for (int cpg = std::min(L->C8, 4); cpg >= 1; --cpg) {
            const int G = cdiv(L->C8, cpg);
            const int units = cpg * d.kh * (L->taps_in_k ? d.kw : 1);
            const int KS = cdiv(units, 2);
            if (G * KS > TC_MAX_KSTEPS) continue;
            const size_t stage = (size_t)2 * cpg * rin * L->Wp * 16;
            const size_t bimg = (size_t)G * KS * 2 * L->NCOLS * 16;
            const size_t stride = (stage + 127) / 128 * 128;
            const int ns = G >= 2 ? 2 : (stride * 3 + 2 * bimg + fixed <= (size_t)smem_budget ? 3 : 2);
            if (stride * ns + 2 * bimg + fixed > (size_t)smem_budget) continue;
            if (stage > (1u << 20)) continue;
            best_r = r_out; best_cpg = cpg; best_ns = ns; best_mt = mt; best_nacc = nacc;
            break;
        }
        if (best_r) break;
    }
    if (!best_r) return -1;
    L->R_out = best_r;
    L->Rin = best_r + halo_h;
    L->MT = best_mt;
    L->cpg = best_cpg;
    L->G = cdiv(L->C8, best_cpg);
    L->KS = cdiv(best_cpg * d.kh * (L->taps_in_k ? d.kw : 1), 2);
    L->NS = best_ns;
    L->NACC = best_nacc;
    L->stage_bytes = (uint32_t)(2 * best_cpg * L->Rin * L->Wp * 16);
    L->stage_stride = (L->stage_bytes + 127) / 128 * 128;
    L->plane_bytes = (uint32_t)(L->Rin * L->Wp * 16);
    L->b_bytes = (uint32_t)(L->G * L->KS * 2 * L->NCOLS * 16);
    L->smem = (size_t)L->NS * L->stage_stride + 2 * (size_t)L->b_bytes +
              (size_t)TC_SETS * cdiv(L->MT * L->CBLK, TC_SETS) * 4 * halo_w * (L->kw_eff - 1) * 8 * 4 + bias_bars + 1024;
    return 0;
}

// Weight image: for every group g and K step ks two units of [NCOLS][8] fp16; a unit = (chunk, vertical tap i) holding
// w[i][j][c8*8+e][co] at column n = (cb*KW + j)*8 + ci, co = cb*8 + ci.  Odd unit counts pair the last unit with a zero unit
// that re-reads the previous tap's A rows.
// Sliding-window weight image: for every K step ks and vertical tap i a hi block then a lo block, each
// [2 units][NCOLS][8] fp16 (K-major, LBO = NCOLS*16 B between the two 8-channel units).  A unit is (8-channel chunk c8)
// with all horizontal taps in N, column n = (cb*KW + j)*CSTRIDE + ci, or (c8, horizontal tap j) with N = filters.
static int sw_pack_weights(const DlwpConvDesc& d, const TcLayer& L, const float* w_host, std::vector<__half>* img,
                           TcKStep* kst_out) {
    struct U { int c8, j; bool zero; };
    std::vector<U> units;
    for (int c8 = 0; c8 < L.C8; ++c8) {
        if (L.taps_in_k) for (int j = 0; j < d.kw; ++j) units.push_back({c8, j, false});
        else units.push_back({c8, -1, false});
    }
    if (units.size() & 1) units.push_back({units.back().c8, units.back().j, true});
    auto aoff = [&](const U& u) {
        // the zero unit (zero weights) views the lo plane of the last real unit: finite data, positive LBO
        return (long long)(2 * u.c8 + (u.zero ? 1 : 0)) * L.rowpitch + (long long)(u.j > 0 ? u.j : 0) * d.dil_w * 16;
    };
    const size_t block = (size_t)2 * L.NCOLS * 8;  // fp16 elements of one (ks, i, hi|lo) block
    img->assign((size_t)L.KS * d.kh * 2 * block, __float2half(0.f));
    for (int ks = 0; ks < L.KS; ++ks) {
        const U& u0 = units[2 * ks];
        const U& u1 = units[2 * ks + 1];
        const long long lbo = aoff(u1) - aoff(u0);
        if (lbo <= 0 || (lbo >> 4) > 0x3FFF) return -1;
        // the issuer adds the stage base (in 16-byte units) to this: low word of the A descriptor, LBO field included
        kst_out[ks].a_off = (uint32_t)((((uint32_t)lbo >> 4) << 16) | ((uint32_t)aoff(u0) >> 4));
        kst_out[ks].a_lbo = (uint32_t)lbo;
        for (int i = 0; i < d.kh; ++i)
            for (int half = 0; half < 2; ++half) {
                const U& u = half ? u1 : u0;
                if (u.zero) continue;
                const size_t bh = ((size_t)(ks * d.kh + i) * 2 + 0) * block + (size_t)half * L.NCOLS * 8;
                const size_t bl = ((size_t)(ks * d.kh + i) * 2 + 1) * block + (size_t)half * L.NCOLS * 8;
                for (int cb = 0; cb < L.CBLK; ++cb)
                    for (int j = (u.j >= 0 ? u.j : 0); j < (u.j >= 0 ? u.j + 1 : d.kw); ++j)
                        for (int ci = 0; ci < L.CSTRIDE; ++ci) {
                            const int co = cb * 8 + ci;
                            if (co >= d.Cout) continue;
                            const int ncol = (cb * L.kw_eff + (u.j >= 0 ? 0 : j)) * L.CSTRIDE + ci;
                            for (int e = 0; e < 8; ++e) {
                                const int c = u.c8 * 8 + e;
                                if (c >= d.Cin) continue;
                                const float v = w_host[(((size_t)i * d.kw + j) * d.Cin + c) * d.Cout + co];
                                const __half h = __float2half_rn(v);
                                (*img)[bh + (size_t)ncol * 8 + e] = h;
                                (*img)[bl + (size_t)ncol * 8 + e] = __float2half_rn(v - __half2float(h));
                            }
                        }
            }
    }
    return 0;
}

int tc_pack_weights(const DlwpConvDesc& d, const TcLayer& L, const float* w_host, std::vector<__half>* img,
                    TcKStep* kst_out) {
    if (L.mode == 1) return sw_pack_weights(d, L, w_host, img, kst_out);
    const size_t per_img = (size_t)L.b_bytes / 2;
    img->assign(2 * per_img, __float2half(0.f));
    for (int g = 0; g < L.G; ++g) {
        struct U { int chunk_in_group, chunk, i, j; bool zero; };  // j = -1: all horizontal taps live in N
        std::vector<U> units;
        for (int cg = 0; cg < L.cpg; ++cg)
            for (int i = 0; i < d.kh; ++i) {
                if (L.taps_in_k)
                    for (int j = 0; j < d.kw; ++j) units.push_back({cg, g * L.cpg + cg, i, j, false});
                else
                    units.push_back({cg, g * L.cpg + cg, i, -1, false});
            }
        if (units.size() & 1) {  // [.., u(n-2), u(n-1)] -> [.., u(n-2), ZERO(view of u(n-2)), u(n-1)]  (needs >= 2 units)
            U last = units.back();
            units.pop_back();
            U z = units.empty() ? last : units.back();
            z.zero = true;
            if (units.empty()) { units.push_back(last); units.push_back(z); }  // single unit: (u, ZERO view of u) lbo = 0 ...
            else { units.push_back(z); units.push_back(last); }
        }
        for (int ks = 0; ks < L.KS; ++ks) {
            const U& u0 = units[2 * ks];
            const U& u1 = units[2 * ks + 1];
            auto aoff = [&](const U& u) {
                return (long long)(2 * u.chunk_in_group) * L.plane_bytes + (long long)u.i * d.dil_h * L.Wp * 16 +
                       (long long)(u.j > 0 ? u.j : 0) * d.dil_w * 16;
            };
            const long long lbo = aoff(u1) - aoff(u0);
            if (lbo < 0 || (lbo >> 4) > 0x3FFF) return -1;
            kst_out[g * L.KS + ks].a_off = (uint32_t)aoff(u0);
            kst_out[g * L.KS + ks].a_lbo = (uint32_t)lbo;
            for (int half = 0; half < 2; ++half) {
                const U& u = half ? u1 : u0;
                if (u.zero) continue;
                const size_t ubase = ((size_t)(g * L.KS + ks) * 2 + half) * L.NCOLS * 8;
                for (int cb = 0; cb < L.CBLK; ++cb)
                    for (int j = (u.j >= 0 ? u.j : 0); j < (u.j >= 0 ? u.j + 1 : d.kw); ++j)
                        for (int ci = 0; ci < 8; ++ci) {
                            const int co = cb * 8 + ci;
                            const int ncol = (cb * L.kw_eff + (u.j >= 0 ? 0 : j)) * L.CSTRIDE + ci;
                            if (ci >= L.CSTRIDE) continue;
                            if (co >= d.Cout) continue;
                            for (int e = 0; e < 8; ++e) {
                                const int c = u.chunk * 8 + e;
                                if (c >= d.Cin) continue;
                                const float v = w_host[(((size_t)u.i * d.kw + j) * d.Cin + c) * d.Cout + co];
                                const __half h = __float2half_rn(v);
                                (*img)[ubase + (size_t)ncol * 8 + e] = h;
                                (*img)[per_img + ubase + (size_t)ncol * 8 + e] = __float2half_rn(v - __half2float(h));
                            }
                        }
            }
        }
    }
    return 0;
}

static CUtensorMap g_sw_map_full, g_sw_map_pair;  // the maps of the launch being issued (copied into the kernel parameters)

template <int KH, int KW, int NC, class ST = SwGeneric>
static void sw_launch_one(const SwParams& p, int grid, size_t smem, cudaStream_t stream) {
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(conv_sw_kernel<KH, KW, NC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    constexpr int threads = TC_THREADS + (ST::F32IN ? SW_CONV_WARPS * 32 : 0);
    conv_sw_kernel<KH, KW, NC, ST><<<grid, threads, smem, stream>>>(p, g_sw_map_full, g_sw_map_pair);
}

bool tc_f32in_ok(const DlwpConvDesc& d, const TcLayer& L) {
    return L.mode == 1 && d.kh == 3 && L.kw_eff == 1 && L.NCOLS == 32 && L.KS == 2 && d.dil_w == 2 && L.CBLK == 4 &&
           d.Cout == 32 && d.act == DLWP_ACT_TANH && d.Cin <= 8 && d.W <= 256 && d.W % 4 == 0 && d.x_stride_h == d.W &&
           d.x_stride_c == (long long)d.H * d.W && d.x_stride_n == (long long)d.Cin * d.H * d.W;
}

// The scheduling units of one launch: (strip or paired remainder strips) x (latitude band), see sw_decode.  Host-only
// arithmetic (also driven by the CPU tests through dlwp_debug_sw_cover); returns the number of sample groups.
static int sw_unit_geometry(const DlwpConvDesc& d, const TcLayer& L, int sms, SwParams* pp) {
    SwParams& p = *pp;
    p.N = d.N; p.H = d.H; p.W = d.W; p.Wp = L.Wp;
    p.D = d.dil_w; p.pad_t = d.pad_t;
    p.S = L.S; p.nfull = L.nfull; p.rem = L.rem; p.pair = L.pair;
    const bool all_rows = d.row_begin == 0 && d.row_end == 0;
    p.row0 = all_rows ? 0 : d.row_begin;
    p.row1 = all_rows ? d.H : d.row_end;
    const int rows = p.row1 - p.row0, span = d.dil_h * (d.kh - 1);
    const int groups = L.pair ? cdiv(d.N, 2) : d.N;
    p.units_per_group = L.pair ? 2 * L.nfull + 1 : L.nfull + (L.rem > 0 ? 1 : 0);
    if (rows <= 0) {  // empty latitude window
        p.RB = 1; p.nbands = 1; p.total_units = 0;
        return groups;
    }
    // latitude bands per strip: enough units to balance the SMs, few enough that the (KH-1)*dil halo rows re-read at
    // every band edge stay a small fraction
    int best_nb = 1;
    double best_score = -1.0;
    const char* env_nb = getenv("DLWP_TC_BANDS");
    if (sms < 1) sms = 148;
    for (int nb = 1; nb <= 16 && nb <= rows; ++nb) {
        const int rb = cdiv(rows, nb);
        if (cdiv(rows, rb) != nb) continue;
        const long long total = (long long)groups * p.units_per_group * nb;
        const double balance = (double)total / (double)(cdiv((int)total, sms) * (long long)sms);
        const double reread = (double)rows / (double)(rows + span * nb);
        const double score = balance * (0.5 + 0.5 * reread);
        if (score > best_score + 1e-9) { best_score = score; best_nb = nb; }
    }
    if (env_nb && atoi(env_nb) > 0) best_nb = std::min(atoi(env_nb), rows);
    p.RB = cdiv(rows, best_nb);
    p.nbands = cdiv(rows, p.RB);
    p.total_units = groups * p.units_per_group * p.nbands;
    return groups;
}

static int sw_launch(const DlwpConvDesc& d, const TcLayer& L, const TcKStep* kst, const __half* xp, const __half* bimg,
                     const float* bias, float* y32, __half* yp, int wpad_out, int planes_out, cudaStream_t stream,
                     const TcWindow& win) {
    if (g_tc_sms == 0) {
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaGetDeviceProperties(&prop, dev);
        g_tc_sms = prop.multiProcessorCount;
    }
    if (!(d.row_begin == 0 && d.row_end == 0) && d.row_end <= d.row_begin) return 0;  // empty latitude window: nothing to do
    SwParams p;
    memset(&p, 0, sizeof(p));
    const int groups = sw_unit_geometry(d, L, g_tc_sms, &p);
    p.Cout = d.Cout; p.NCOLS = L.NCOLS; p.CBLK = L.CBLK; p.CSTRIDE = L.CSTRIDE; p.XL = (L.kw_eff - 1) * d.dil_w;
    p.KS = L.KS; p.NS = L.NS; p.NACC = L.NACC;
    p.planes_in = L.planes;
    p.rowpitch = L.rowpitch; p.stage_stride = L.stage_stride; p.b_unit16 = (uint32_t)(2 * L.NCOLS); p.b_bytes = L.b_bytes;
    p.idesc = (1u << 4) | ((uint32_t)(L.NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major
    p.act = d.act; p.bias = bias; p.bimg = bimg; p.xp = xp;
    {
        const char* env_dbg = getenv("DLWP_SW_DEBUG");
        p.debug = env_dbg ? atoi(env_dbg) : 0;
    }
    p.in_plane0 = win.in_plane0; p.in_planes_total = win.in_planes_total ? win.in_planes_total : L.planes;
    p.out_plane0 = win.out_plane0;
    p.y32 = y32; p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    p.yp = yp; p.wpad_out = wpad_out; p.Wp_out = d.W + 2 * wpad_out; p.planes_out = planes_out;
    for (int i = 0; i < L.KS; ++i) p.kst[i] = kst[i];
    const int grid = std::min(p.total_units, g_tc_sms);
    const int nc = L.CSTRIDE == 6 ? 6 : 8;
    size_t smem = L.smem;
    const bool f32in = win.x32 != nullptr;
    if (f32in) {  // fp32 state as the source: raw-row ring behind the regular carve-up, fp32 tensor map in map_full
        DLWP_REQUIRE(tc_f32in_ok(d, L) && yp != nullptr && y32 == nullptr, DLWP_ESHAPE, "layer cannot read the fp32 state directly");
        p.f32_C = d.Cin;
        p.f32_rawpitch = d.W * 4;
        p.f32_segstride = (d.Cin * d.W * 4 + 127) / 128 * 128;
        smem += (size_t)SW_RAW_STAGES * 2 * p.f32_segstride + 256;
        DLWP_REQUIRE(smem <= 227 * 1024, DLWP_ESHAPE, "no shared memory left for the raw fp32 row stages");
        const uint64_t dims[3] = {(uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.N * d.Cin};
        const uint64_t str[2] = {(uint64_t)d.W * 4, (uint64_t)d.H * d.W * 4};
        const uint32_t box[3] = {(uint32_t)d.W, 1, (uint32_t)d.Cin};
        int rc = encode_tensor_map_any(&g_sw_map_full, win.x32, 1, 3, dims, str, box);
        if (rc) return rc;
        g_sw_map_pair = g_sw_map_full;
    }
    // Tensor-map producer for layers whose staged row is exactly 128 pixels per plane (taps in N) and has several planes.
    p.use_tma = 0;
    if (!f32in && L.XLK == 0 && L.rowpitch == 2048 && !getenv("DLWP_SW_NO_TMA")) {
        const uint64_t Halloc = (uint64_t)d.H + 2 * TC_HPAD, row_b = (uint64_t)L.Wp * 16, plane_b = Halloc * row_b;
        const uint64_t dims3[3] = {(uint64_t)L.Wp * 2, Halloc, (uint64_t)d.N * p.in_planes_total};
        const uint64_t str3[2] = {row_b, plane_b};
        const uint32_t box3[3] = {256, 1, (uint32_t)L.planes};
        const uint64_t dims4[4] = {(uint64_t)L.Wp * 2, Halloc, (uint64_t)d.N, (uint64_t)p.in_planes_total};
        const uint64_t str4[3] = {row_b, plane_b * p.in_planes_total, plane_b};
        const uint32_t box4[4] = {128, 1, 2, (uint32_t)L.planes};
        if (encode_tensor_map_u64(&g_sw_map_full, xp, 3, dims3, str3, box3) == 0 &&
            (!L.pair || encode_tensor_map_u64(&g_sw_map_pair, xp, 4, dims4, str4, box4) == 0)) {
            if (!L.pair) g_sw_map_pair = g_sw_map_full;
            p.use_tma = 1;
        }
    }
    if (getenv("DLWP_TC_DEBUG"))
        fprintf(stderr, "sw_launch %d->%d k%d: units %d (groups %d x %d strips x %d bands of %d rows) grid %d NS %d NACC %d KS %d smem %zu tma %d\n",
                d.Cin, d.Cout, d.kh, p.total_units, groups, p.units_per_group, p.nbands, p.RB, grid, p.NS, p.NACC, p.KS, L.smem, p.use_tma);
    // fully folded instances for the benchmark nets' layers (SwStatic<NCOLS, KS, D, CBLK, ACT, OUT>)
    const int out_mode = (yp ? 1 : 0) | (y32 ? 2 : 0);
    const bool generic_only = getenv("DLWP_TC_GENERIC") != nullptr;
    if (f32in)                                                                // Net A conv1 reading the fp32 state itself
        sw_launch_one<3, 1, 8, SwStatic<32, 2, 2, 4, DLWP_ACT_TANH, 1, 1, 1>>(p, grid, smem, stream);
    else if (!generic_only && d.kh == 3 && L.kw_eff == 1 && L.NCOLS == 32 && L.KS == 2 && d.dil_w == 2 && L.CBLK == 4 &&
        d.Cout == 32 && d.act == DLWP_ACT_TANH && out_mode == 1)            // Net A conv1: 6 -> 32, 3x3 dilation 2, tanh, P-layout output
        sw_launch_one<3, 1, 8, SwStatic<32, 2, 2, 4, DLWP_ACT_TANH, 1, 1>>(p, grid, L.smem, stream);
    else if (!generic_only && d.kh == 5 && L.kw_eff == 5 && nc == 6 && L.NCOLS == 32 && L.KS == 2 && d.dil_w == 1 &&
             L.CBLK == 1 && d.act == DLWP_ACT_LINEAR && out_mode == 3)  // Net A conv2: 32 -> 6, 5x5, fp32 series + feedback
        sw_launch_one<5, 5, 6, SwStatic<32, 2, 1, 1, DLWP_ACT_LINEAR, 3, 1>>(p, grid, L.smem, stream);
    else if (!generic_only && d.kh == 5 && L.kw_eff == 5 && nc == 6 && L.NCOLS == 32 && L.KS == 2 && d.dil_w == 1 &&
             L.CBLK == 1 && d.act == DLWP_ACT_LINEAR && out_mode == 2)  // ... the same layer in a latitude band (fp32 only)
        sw_launch_one<5, 5, 6, SwStatic<32, 2, 1, 1, DLWP_ACT_LINEAR, 2, 1>>(p, grid, L.smem, stream);
    else if (d.kh == 3 && L.kw_eff == 1) sw_launch_one<3, 1, 8>(p, grid, L.smem, stream);
    else if (d.kh == 5 && L.kw_eff == 1) sw_launch_one<5, 1, 8>(p, grid, L.smem, stream);
    else if (d.kh == 3 && nc == 8) sw_launch_one<3, 3, 8>(p, grid, L.smem, stream);
    else if (d.kh == 3) sw_launch_one<3, 3, 6>(p, grid, L.smem, stream);
    else if (d.kh == 5 && nc == 8) sw_launch_one<5, 5, 8>(p, grid, L.smem, stream);
    else sw_launch_one<5, 5, 6>(p, grid, L.smem, stream);
    return after_launch("conv_sw_kernel");
}

int tc_launch(const DlwpConvDesc& d, const TcLayer& L, const TcKStep* kst, const __half* xp, const __half* bimg,
              const float* bias, float* y32, __half* yp, int wpad_out, int planes_out, cudaStream_t stream,
              const TcWindow& win) {
    static std::once_flag once;
    std::call_once(once, [] {
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaGetDeviceProperties(&prop, dev);
        g_tc_sms = prop.multiProcessorCount;
        cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        cudaFuncSetAttribute(conv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
        cudaFuncSetAttribute(conv_tc_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin);
    });
    if (L.mode == 1) return sw_launch(d, L, kst, xp, bimg, bias, y32, yp, wpad_out, planes_out, stream, win);
    DLWP_REQUIRE(win.in_plane0 == 0 && win.out_plane0 == 0 && (win.in_planes_total == 0 || win.in_planes_total == L.planes) &&
                     (yp == nullptr || planes_out == 2 * cdiv(d.Cout, 8)),
                 DLWP_ESHAPE, "the flattened-tile tensor-core kernel does not take channel windows");
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.N = d.N; p.H = d.H; p.W = d.W; p.Wp = L.Wp;
    p.R_out = L.R_out; p.Rin = L.Rin; p.MT = L.MT; p.S = L.S;
    p.KW = L.kw_eff; p.D = d.dil_w; p.pad_t = d.pad_t;
    p.Cout = d.Cout; p.NCOLS = L.NCOLS; p.CBLK = L.CBLK; p.CSTRIDE = L.CSTRIDE; p.XL = (L.kw_eff - 1) * d.dil_w;
    p.G = L.G; p.KS = L.KS; p.NS = L.NS; p.NACC = L.NACC; p.planes_per_group = 2 * L.cpg;
    const bool all_rows = d.row_begin == 0 && d.row_end == 0;
    p.row0 = all_rows ? 0 : d.row_begin;
    p.row1 = all_rows ? d.H : d.row_end;
    p.tiles_per_sample = cdiv(p.row1 - p.row0, L.R_out);
    p.total_tiles = p.tiles_per_sample * d.N;
    p.stage_bytes = L.stage_bytes; p.stage_stride = L.stage_stride; p.plane_bytes = L.plane_bytes; p.b_bytes = L.b_bytes;
    p.idesc = (1u << 4) | ((uint32_t)(L.NCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major
    p.act = d.act; p.bias = bias; p.bimg = bimg;
    p.y32 = y32; p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    p.yp = yp; p.wpad_out = wpad_out; p.Wp_out = d.W + 2 * wpad_out; p.planes_out = planes_out;
    for (int i = 0; i < L.G * L.KS; ++i) p.kst[i] = kst[i];
    p.xp = xp;
    p.planes_in = L.planes;
    const int grid = std::min(p.total_tiles, g_tc_sms);
    if (L.kw_eff == 1) conv_tc_kernel<1><<<grid, TC_THREADS, L.smem, stream>>>(p);
    else if (d.kw == 3) conv_tc_kernel<3><<<grid, TC_THREADS, L.smem, stream>>>(p);
    else conv_tc_kernel<5><<<grid, TC_THREADS, L.smem, stream>>>(p);
    return after_launch("conv_tc_kernel");
}

int tc_pack_state(const float* x, __half* xp, int N, int C, int H, int W, int wpad, long long xs_n, long long xs_c,
                  long long xs_h, cudaStream_t stream, int row0, int row1) {
    if (row0 == 0 && row1 == 0) row1 = H;
    const long long total = (long long)N * cdiv(C, 8) * (row1 - row0) * (W + 2 * wpad);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    pack_state_kernel<<<blocks, 256, 0, stream>>>(x, xp, N, C, H, W, wpad, xs_n, xs_c, xs_h, row0, row1 - row0);
    return after_launch("pack_state_kernel");
}

size_t tc_p_bytes(int N, int planes, int H, int Wp) { return (size_t)N * planes * (H + 2 * TC_HPAD) * Wp * 16; }

int tc_debug_flags() {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_tc_flags, sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(g_tc_flags, &zero, sizeof(int));
    return v;
}

// Stand-alone layer through the tensor-core path (tests / A-B timing): packs x and the weights into temporaries.
int conv2d_fwd_tc(const DlwpConvDesc& d, const float* x, const float* w_dev, const float* bias, float* y,
                  cudaStream_t stream) {
    DLWP_REQUIRE(tc_geometry_ok(d), DLWP_ESHAPE, "geometry not supported by the tensor-core conv kernel");
    TcLayer L;
    tc_plan_layer(d, &L);
    const size_t wn = (size_t)d.kh * d.kw * d.Cin * d.Cout;
    std::vector<float> w_host(wn);
    DLWP_CUDA_TRY(cudaMemcpyAsync(w_host.data(), w_dev, wn * 4, cudaMemcpyDeviceToHost, stream));
    DLWP_CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<__half> img;
    TcKStep kst[TC_MAX_KSTEPS];
    DLWP_REQUIRE(tc_pack_weights(d, L, w_host.data(), &img, kst) == 0, DLWP_ESHAPE, "weight packing failed");
    __half *bimg = nullptr, *xp = nullptr;
    const size_t xp_bytes = tc_p_bytes(d.N, L.planes, d.H, L.Wp);
    DLWP_CUDA_TRY(cudaMalloc(&bimg, img.size() * 2));
    DLWP_CUDA_TRY(cudaMalloc(&xp, xp_bytes));
    DLWP_CUDA_TRY(cudaMemsetAsync(xp, 0, xp_bytes, stream));
    DLWP_CUDA_TRY(cudaMemcpyAsync(bimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice, stream));
    int rc = tc_pack_state(x, xp, d.N, d.Cin, d.H, d.W, L.wpad, d.x_stride_n, d.x_stride_c, d.x_stride_h, stream);
    if (!rc) rc = tc_launch(d, L, kst, xp, bimg, bias, y, nullptr, 0, 0, stream);
    cudaStreamSynchronize(stream);
    cudaFree(bimg);
    cudaFree(xp);
    return rc;
}

}  // namespace dlwp

// ---- host-only test hooks (include/dlwp_b200.h) -------------------------------------------------------------------------
extern "C" int dlwp_debug_tc_plan(const DlwpConvDesc* desc, int32_t* out, int32_t n_out) {
    using namespace dlwp;
    DLWP_REQUIRE(desc && out && n_out >= 16, DLWP_EINVAL, "bad argument");
    TcLayer L;
    DLWP_REQUIRE(tc_geometry_ok(*desc) && tc_plan_layer(*desc, &L) == 0, DLWP_ESHAPE,
                 "geometry not supported by the tensor-core conv kernels");
    const int32_t v[16] = {L.mode, L.taps_in_k, L.NCOLS, L.NACC, L.KS, L.NS, L.S, L.nfull, L.rem, L.pair, (int32_t)L.smem,
                           (int32_t)L.b_bytes, L.CBLK, L.CSTRIDE, L.planes, (int32_t)L.rowpitch};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return 0;
}

extern "C" int64_t dlwp_debug_tc_pack(const DlwpConvDesc* desc, const float* kernel, uint16_t* image, int64_t image_cap,
                                      uint32_t* kstep_words, int32_t kstep_cap) {
    using namespace dlwp;
    if (!desc || !kernel || !image || !kstep_words) return DLWP_EINVAL;
    TcLayer L;
    if (!tc_geometry_ok(*desc) || tc_plan_layer(*desc, &L) != 0) return DLWP_ESHAPE;
    std::vector<__half> img;
    TcKStep kst[TC_MAX_KSTEPS];
    if (tc_pack_weights(*desc, L, kernel, &img, kst) != 0) return DLWP_ESHAPE;
    const int nks = L.mode == 1 ? L.KS : L.G * L.KS;
    if ((int64_t)img.size() > image_cap || 2 * nks > kstep_cap) return DLWP_EINVAL;
    memcpy(image, img.data(), img.size() * sizeof(__half));
    for (int i = 0; i < nks; ++i) {
        kstep_words[2 * i] = kst[i].a_off;
        kstep_words[2 * i + 1] = kst[i].a_lbo;
    }
    return (int64_t)img.size();
}

extern "C" int dlwp_debug_sw_cover(const DlwpConvDesc* desc, int32_t sms, int32_t* cover, int64_t cover_elems,
                                   int32_t* info, int32_t n_info) {
    using namespace dlwp;
    DLWP_REQUIRE(desc && cover && info && n_info >= 4, DLWP_EINVAL, "bad argument");
    DLWP_REQUIRE(cover_elems == (int64_t)desc->N * desc->H * desc->W, DLWP_EINVAL, "cover must hold N*H*W counters");
    TcLayer L;
    DLWP_REQUIRE(tc_geometry_ok(*desc) && tc_plan_layer(*desc, &L) == 0 && L.mode == 1, DLWP_ESHAPE,
                 "not a sliding-window layer");
    SwParams p;
    memset(&p, 0, sizeof(p));
    sw_unit_geometry(*desc, L, sms, &p);
    long long staged_rows = 0;
    int live_units = 0;
    SwUnit U;
    for (int u = 0; u < p.total_units; ++u) {
        if (!sw_decode(p, u, U)) continue;
        ++live_units;
        staged_rows += U.yb - U.ya + desc->dil_h * (desc->kh - 1);
        for (int ml = 0; ml < 128; ++ml) {  // the epilogue's lane -> (sample, column) map
            const int seg = (U.paired && ml >= 64) ? 1 : 0;
            const int l = ml - seg * 64;
            const int n = seg ? U.n1 : U.n0;
            const int x = U.x0 + l;
            if (n < 0 || l >= (seg ? U.nvb : U.nva) || x >= p.W) continue;
            for (int y = U.ya; y < U.yb; ++y) ++cover[((int64_t)n * desc->H + y) * desc->W + x];
        }
    }
    info[0] = p.total_units; info[1] = live_units; info[2] = p.nbands; info[3] = (int32_t)staged_rows;
    return 0;
}
