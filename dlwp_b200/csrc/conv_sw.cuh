// Device side of the tensor-core conv path: tcgen05 / TMEM / TMA wrappers and the sliding-window kernel template.
// Included by conv_tc.cu (generic instances + host code) and by the conv_sw_net_*.cu translation units that hold the
// fully folded instances of the example nets' layers (separate TUs so that nvcc compiles them in parallel).
//
// With pixels as the M dimension a 'valid' conv is
//   D[m, co] = sum_{i,j,c} A[m + shift(i,j), c] * W[i,j,c,co]
// and the example nets have tiny N (6..128 filters).  Precision: operands are fp16 hi/lo splits of power-of-two scaled
// fp32 values (x * 2^e = hi + lo, 22 significant bits; conv_tc.h explains the exponents); three MMAs per K step
// (hi*hi + hi*lo + lo*hi) accumulate in fp32 in TMEM -> ~1e-6 relative error, inside the 1e-4 / 50-step gate.
// Activations live in HBM already split and channel-blocked ("P layout"):
// [n][plane = 2*c8 + {hi,lo}][H + zero rows][Wp][8] fp16 with the periodic longitude halo (wpad columns each side)
// materialised by the PRODUCER's epilogue and zero rows stored beyond the poles, so a consumer stages rows with plain bulk /
// tensor-map copies and needs no wrap arithmetic.  A-operand views are K-major, no-swizzle UMMA descriptors into that
// image (8 pixels x 16 B core matrices; validated by scripts/umma_probe.cu on B200).
#pragma once
#define DLWP_CONV_TU  // mbarrier helpers of internal.h
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
// Device-side diagnostic flags: bit 0 = an mbarrier wait timed out, bit 1 = a value left the fp16 split's range (or is
// NaN / inf), bit 2 = an image's amax fell below 2^-6 in scaled units (precision underflow).  Without relocatable device
// code a __device__ variable cannot be shared between translation units, so every TU that instantiates kernels owns one
// word (DLWP_SW_TU_FLAGS names it) and exports a read-and-clear function that tc_debug_flags() ORs together.
#ifndef DLWP_SW_TU_FLAGS
#error "define DLWP_SW_TU_FLAGS (name of this translation unit's device flag word) before including conv_sw.cuh"
#endif
#define DLWP_SW_CAT2(a, b) a##b
#define DLWP_SW_CAT(a, b) DLWP_SW_CAT2(a, b)
#define DLWP_SW_TU_COUNTERS DLWP_SW_CAT(DLWP_SW_TU_FLAGS, _counters)
namespace dlwp {
__device__ int DLWP_SW_TU_FLAGS = 0;
// TcOptions::debug & 4: clock64 totals of the MMA-issuing warp, summed over CTAs:
// [0] waiting for an accumulator slot, [1] waiting for a staged row, [2] issuing MMAs + commits, [3] rows; [4..7] the same
// for the second layer of the fused kernel -- in conv_sw_kernel: [4] issuer waiting for the phantom rows' slots at segment
// starts, [5] issuer between segments, [6] / [7] epilogue warp 0 waiting for / working on rows; [8] issuing warp's clocks from kernel entry to its last commit, [9] the same
// span in globaltimer ns, [10] CTAs counted, [11] clocks from kernel entry to the first staged row's arrival
__device__ unsigned long long DLWP_SW_TU_COUNTERS[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
}
#define g_tc_counters DLWP_SW_TU_COUNTERS
#define g_device_flags DLWP_SW_TU_FLAGS
#define g_tc_flags DLWP_SW_TU_FLAGS
#include "internal.h"

#include <mutex>

#include "conv_tc.h"

namespace dlwp {

constexpr int TC_SETS = 4;          // epilogue warp sets (each = 4 warps, one per TMEM lane quadrant)
constexpr int TC_THREADS = 64 + TC_SETS * 128;  // epilogue sets, then the producer warp and the MMA issuer warp
constexpr int TC_FLAG_TIMEOUT = 1, TC_FLAG_RANGE = 2, TC_FLAG_UNDERFLOW = 4;

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
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
__device__ __forceinline__ void umma_f16_dyn(int coll, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t acc) {
    if (coll == 1) umma_f16_c<1>(tmem_d, adesc, bdesc, idesc, acc);
    else if (coll == 2) umma_f16_c<2>(tmem_d, adesc, bdesc, idesc, acc);
    else if (coll == 3) umma_f16_c<3>(tmem_d, adesc, bdesc, idesc, acc);
    else umma_f16(tmem_d, adesc, bdesc, idesc, acc);
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
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t tmem) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((e + 127) << 23); }  // |e| <= 126

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
// v (already scaled by the image's 2^e) -> hi / lo fp16 vectors
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
// bf16 mode (DlwpPlanOptions.precision = 1): one plane per 8-channel chunk, round-to-nearest-even bf16
__device__ __forceinline__ void p_pack8_bf16(const float (&v)[8], uint4& h) {
    uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        hp[k] = *reinterpret_cast<const uint32_t*>(&b);
    }
}
__device__ __forceinline__ void p_unpack8_bf16(const uint4& h, float (&v)[8]) {
    const uint32_t* hp = reinterpret_cast<const uint32_t*>(&h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[2 * k] = __uint_as_float(hp[k] << 16);
        v[2 * k + 1] = __uint_as_float(hp[k] & 0xffff0000u);
    }
}
__device__ __forceinline__ float amax8(const float (&o)[8]) {
    return fmaxf(fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3]))),
                 fmaxf(fmaxf(fabsf(o[4]), fabsf(o[5])), fmaxf(fabsf(o[6]), fabsf(o[7]))));
}
// max over the warp of a non-negative float -> one atomicMax on its bit pattern (order-preserving for x >= 0)
__device__ __forceinline__ void amax_publish(float* slot, float v, int lane) {
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
    if (lane == 0 && slot != nullptr && m != 0u) atomicMax(reinterpret_cast<unsigned*>(slot), m);
}

// The scale prologue every tensor-core kernel runs in one thread: reads the source's exponent and measured amax, decides
// the destination's exponent from the bound (dynamic destinations), publishes it, raises the range / underflow flags.
// Returns {2^-(e_in + e_w), 2^e_out}.
__device__ __forceinline__ float2 tc_scale_prologue(const TcScale& sc, int act, bool first_block) {
    const int e_in = sc.e_in ? *sc.e_in : sc.e_in_const;
    const float amax_in = sc.amax_in ? *sc.amax_in : 1.f;
    const float amax_chk = sc.amax_chk ? *sc.amax_chk : amax_in;
    int e_out = sc.e_out_const;
    if (sc.e_out) {
        float B = fmaf(sc.l1max, amax_in, sc.bmax);
        if (act == DLWP_ACT_TANH) B = fminf(B, 1.f);
        e_out = tc_exp_for_bound(B);
        if (!(B < 1e30f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
        if (first_block) *sc.e_out = e_out;
    }
    if (first_block && sc.amax_zero) *sc.amax_zero = 0.f;
    if (!(amax_in < 1e30f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
    // the source image's largest element sits below 2^-6: its small elements have lost relative precision
    if (amax_chk > 0.f && amax_chk * exp2i(e_in) < 0.015625f) atomicOr(&g_tc_flags, TC_FLAG_UNDERFLOW);
    return make_float2(exp2i(-(e_in + sc.e_w)), exp2i(e_out));
}

// ===================================================================================================================
// Sliding-window kernel.
//
// Measured on B200 (scripts/mma_rate_probe.cu, profiles/r01_mma_rate_probe.txt): a 128 x N x 16 f16 MMA with both operands
// in shared memory takes max(N/2, 32 + N/4) clocks -- for the N = 32..96 of these layers it is bound by the 4 KB A read,
// not by the tensor pipe.  With .collector::a::fill/use/lastuse consecutive MMAs that share their A operand run at
// ~19 clk (N = 32; tensor floor 16).  So the tile is laid out to make A sharing the common case:
//   * an M tile is 128 consecutive pixels of ONE padded row ("strip"; lane = x - x0).  The remainder strips of two
//     samples share a tile (lanes 0-63 / 64-127) when they fit, so W = 180 costs 1.5 tiles per row, not 2;
//   * a CTA walks a strip top to bottom.  Each padded input row is staged once (one tensor-map load per row and unit, or
//     one bulk copy per plane and segment, into a ring of row stages) and feeds ALL vertical taps: row rp updates the
//     accumulators of output rows rp - i*dil, i = 0..KH-1, with the same A view and different weight blocks -> per K step
//     the hi view is read once for 2*KH MMAs (hi*hi, hi*lo) and the lo view once for KH MMAs;
//   * TMEM holds a ring of NACC accumulators (one output row x 128 lanes x NCOLS columns each); a row is committed to
//     the epilogue when its last tap has been issued.  Four epilogue warp sets take rows round robin;
//   * the accumulators of the output rows one input row feeds (rp, rp - dil, ..., rp - (KH-1)*dil) sit in ADJACENT TMEM
//     columns (the ring is ordered by row within each residue class mod dil), and the weight blocks of the vertical taps
//     are stored side by side in the same order, so ONE MMA with N = taps * NCOLS updates all of them ("vertical taps
//     folded into N", up to N = 256): 6 MMAs of N = 160 per input row instead of 30 of N = 32 for a 5x5 layer.  A 32-column
//     MMA is bound by its 4 KB A-operand read (40 clk, ~19 with the collector); at N >= 128 the same A read is amortised
//     and the MMA runs at the tensor pipe's N/2 clk.  Runs are cut at the ring's end, at the band's first / last rows
//     (fewer valid taps) and where a new output row takes its first contribution (accumulate flag off).
// Horizontal taps live in N (few filters: the epilogue's shifted sum) or in K (A views shifted by j*dil pixels inside the
// staged row; N = filters).
// ===================================================================================================================
constexpr int SW_MAX_STAGES = 12;
constexpr int SW_MAX_ACC = 16;

struct SwParams {
    int N, H, W, Wp;
    int D, pad_t;
    int S, nfull, rem, pair;          // strips: valid outputs per full strip, full strips per row, remainder, pairing
    int units_per_group, nbands, RB, total_units;   // (strip x band) units: the fused pair kernel's schedule (sw_decode)
    int nstrips, total_rows;          // conv_sw_kernel: strips over all samples, nstrips * (row1 - row0) (sw_next)
    int row0, row1;
    int Cout, NCOLS, CBLK, CSTRIDE, XL;
    int KS, NS, NACC;
    int fold;                         // vertical taps one MMA covers (1 .. KH; fold * NCOLS <= 256)
    int planes_in;
    uint32_t rowpitch, stage_stride, b_unit16, b_bytes;  // b_unit16: one (k step, tap, hi|lo) weight block in 16-byte units
    uint32_t idesc;                   // instruction descriptor with the N field clear (set per MMA run)
    int act;
    const float* bias;
    const __half* bimg;
    const __half* xp;
    int in_plane0, in_planes_total, out_plane0;  // channel windows of the source / destination P images
    int bf16;     // 1: plain bf16 operands, one plane per 8-channel chunk, one MMA pass (BASELINE.json configs[2])
    int use_tma;  // 1: the producer stages rows with tensor-map loads (one per row and unit) instead of per-plane bulk copies
    int debug;    // TcOptions::debug
    float* y32; long long ys_n, ys_c, ys_h;
    __half* yp; int Wp_out, wpad_out, planes_out;
    // Latitude-band halo over peer memory (NVLink): output rows y < peer_up_end are ALSO stored into the upper neighbour's
    // copy of the destination image (same layout, same absolute rows -- they are its bottom halo rows), rows
    // y >= peer_down_begin into the lower neighbour's.  Null pointers: no neighbour / not a feedback layer.
    __half* yp_up; __half* yp_down; int peer_up_end, peer_down_begin;
    TcScale sc;
    TcKStep kst[TC_MAX_KSTEPS];
};

struct SwUnit {
    int n0, n1;      // samples of the two segments (n1 = -1: none)
    int x0;          // first padded column of the strip (both segments of a paired tile start at the same column)
    int nva, nvb;    // valid output lanes per segment
    int ya, yb;      // output rows [ya, yb)
    int paired;      // lanes 64.. belong to segment b
};

template <class P>
__host__ __device__ __forceinline__ bool sw_decode(const P& p, int u, SwUnit& U) {
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

// conv_sw_kernel's schedule: the output rows of all strips form one sequence (strip-major, top to bottom) that is cut
// into gridDim.x contiguous ranges of equal length; a CTA walks its range as segments ("units": one strip, rows [ya, yb)).
// Every segment costs (KH-1)*dil warm-up input rows, and a range touches at most rows/len + 2 strips, so the CTAs stay
// balanced to within a few rows -- against (strip x latitude band) units dealt round robin this removed ~9 % of the
// staged rows of the slowest CTA at batch 256 (profiles/r02_balanced_ranges.txt).  Strips: all full strips first
// (sample-major), then the (paired) remainder strips.
struct SwIter {
    long long g, g1;
};
template <class P>
__host__ __device__ __forceinline__ void sw_range(const P& p, int cta, int nctas, SwIter& it) {
    it.g = (long long)p.total_rows * cta / nctas;
    it.g1 = (long long)p.total_rows * (cta + 1) / nctas;
}
template <class P>
__host__ __device__ __forceinline__ bool sw_next(const P& p, SwIter& it, SwUnit& U) {
    if (it.g >= it.g1) return false;
    const int rows = p.row1 - p.row0;
    const int su = (int)(it.g / rows), y = (int)(it.g - (long long)su * rows);
    int n = rows - y;
    if ((long long)n > it.g1 - it.g) n = (int)(it.g1 - it.g);
    it.g += n;
    U.ya = p.row0 + y;
    U.yb = U.ya + n;
    U.n1 = -1; U.nvb = 0; U.paired = 0;
    const int nfs = p.N * p.nfull;           // full strips
    if (su < nfs) {
        U.n0 = su / p.nfull; U.x0 = (su - U.n0 * p.nfull) * p.S; U.nva = p.S;
    } else if (!p.pair) {
        U.n0 = su - nfs; U.x0 = p.nfull * p.S; U.nva = p.rem;
    } else {
        const int g = su - nfs;
        U.paired = 1;
        U.n0 = 2 * g; U.x0 = p.nfull * p.S; U.nva = p.rem;
        if (2 * g + 1 < p.N) { U.n1 = 2 * g + 1; U.nvb = p.rem; }
    }
    return true;
}

constexpr uint32_t SW_IDESC = (1u << 4) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major, M = 128; N per MMA
constexpr uint32_t SW_IDESC_BF16 = SW_IDESC | (1u << 7) | (1u << 10);     // A and B formats = bf16

// Position of a row in the accumulator ring.  Rows are numbered CTA-wide (phantom rows included); row G sits in residue
// class m = G % D at ring position q = (G / D) % RING, i.e. slot m * RING + q, and lap = (G / (D * RING)) & 1 is the phase
// of that slot's barriers.  Kept incrementally: no divisions on the issuing warp.
struct SwRing {
    int m = 0, q = 0;
    uint32_t lap = 0;
    __device__ __forceinline__ int slot(int RING) const { return m * RING + q; }
    __device__ __forceinline__ void advance(int D, int RING) {
        if (++m == D) {
            m = 0;
            if (++q == RING) { q = 0; lap ^= 1; }
        }
    }
    __device__ __forceinline__ void retreat(int D, int RING) {
        if (--m < 0) {
            m = D - 1;
            if (--q < 0) { q = RING - 1; lap ^= 1; }
        }
    }
};

// One input row's MMAs with everything but the stage address folded at compile time (POS = ring position of the window's
// lowest row).  Runs of vertical taps: cut every FOLD taps, at the ring's end, and before tap 0 in the very first MMA (the
// new row's accumulator is overwritten, the others accumulate).
template <int KH, int NCOLS, int KS, int RING, int FOLD, int POS, bool HINTS = true, int NPASS = 3>
__device__ __forceinline__ void sw_issue_row_pos(bool leader, uint32_t tres, uint32_t sbase16, uint32_t pitch16,
                                                 uint32_t desc_hi, uint32_t bbase, const TcKStep* kst) {
    constexpr uint32_t bblock16 = 2u * KH * NCOLS;
    constexpr uint32_t IDESC0 = NPASS == 1 ? SW_IDESC_BF16 : SW_IDESC;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const uint32_t a_lo32 = kst[ks].a_off + sbase16;
        const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | a_lo32;
        const uint64_t ad_lo = ((uint64_t)desc_hi << 32) | (a_lo32 + pitch16);
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
            const uint64_t ad = pass < 2 ? ad_hi : ad_lo;
            const uint32_t bb = bbase + (NPASS == 1 ? (uint32_t)ks : (uint32_t)(2 * ks + (pass == 1 ? 1 : 0))) * bblock16;
            const bool fresh = ks == 0 && pass == 0;
            int run_tb = KH - 1;
#pragma unroll
            for (int t = KH - 1; t >= 0; --t) {
                const int len = run_tb - t + 1;
                const bool wrap_next = (POS + (KH - 1 - t) + 1) % RING == 0;
                const bool end = t == 0 || len == FOLD || wrap_next || (fresh && t == 1);
                if (end) {
                    const uint32_t dcol = tres + (uint32_t)(((POS + (KH - 1 - run_tb)) % RING) * NCOLS);
                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (bb + (uint32_t)((KH - 1 - run_tb) * NCOLS));
                    const uint32_t idesc = IDESC0 | ((uint32_t)((len * NCOLS) >> 3) << 17);
                    const bool first = pass != 1 && run_tb == KH - 1, last = (NPASS == 1 || pass != 0) && t == 0;
                    const uint32_t acc = (fresh && run_tb == 0) ? 0u : 1u;
                    if (leader) {
                        // (HINTS off: two warps issue into the same tensor core -- another warp's MMA may sit between a
                        // fill and its use, so the collector cannot be relied on)
                        if (!HINTS) umma_f16_c<0>(dcol, ad, bd, idesc, acc);
                        else if (first && !last) umma_f16_c<1>(dcol, ad, bd, idesc, acc);
                        else if (!first && last) umma_f16_c<3>(dcol, ad, bd, idesc, acc);
                        else if (!first) umma_f16_c<2>(dcol, ad, bd, idesc, acc);
                        else umma_f16_c<0>(dcol, ad, bd, idesc, acc);
                    }
                    run_tb = t - 1;
                }
            }
        }
    }
}

template <int KH, class ST>
__device__ __forceinline__ void sw_issue_row_static(int pos_lo, bool leader, uint32_t tres, uint32_t sbase16,
                                                    uint32_t pitch16, uint32_t desc_hi, uint32_t bbase, const TcKStep* kst) {
    constexpr int NCOLS = ST::NCOLS ? ST::NCOLS : 16, D = ST::D ? ST::D : 1, KS = ST::KS ? ST::KS : 1;
    constexpr int NACC = (512 / NCOLS > SW_MAX_ACC ? SW_MAX_ACC : 512 / NCOLS) / D * D;
    constexpr int RING = NACC / D;
    constexpr int FOLD = ST::FOLD ? ST::FOLD : 1;
#define SW_POS_CASE(P_)                                                                                          \
    case P_:                                                                                                     \
        if constexpr (P_ < RING)                                                                                 \
            sw_issue_row_pos<KH, NCOLS, KS, RING, FOLD, P_, true, ST::BF16 ? 1 : 3>(leader, tres, sbase16, pitch16,  \
                                                                                     desc_hi, bbase, kst);         \
        break;
    switch (pos_lo) {
        SW_POS_CASE(0) SW_POS_CASE(1) SW_POS_CASE(2) SW_POS_CASE(3) SW_POS_CASE(4) SW_POS_CASE(5) SW_POS_CASE(6) SW_POS_CASE(7)
        SW_POS_CASE(8) SW_POS_CASE(9) SW_POS_CASE(10) SW_POS_CASE(11) SW_POS_CASE(12) SW_POS_CASE(13) SW_POS_CASE(14)
        SW_POS_CASE(15)
        default: break;
    }
#undef SW_POS_CASE
}

// KH: kernel height (vertical taps), KW: horizontal taps summed by the epilogue (1 = folded into K), NC: filters per
// 8-filter block that exist (6: the single packed block of a 6-filter layer, else 8)
//
// ST: compile-time copy of the per-layer constants (0 / -1 = take the value from SwParams at run time).  The generic
// instance serves any layer; the example nets' layers get instances with everything folded (conv_sw_net_*.cu), which
// matters because the single MMA-issuing warp (and, for many filters, the epilogue's instruction issue) is the kernel's
// critical path.
template <int NCOLS_, int KS_, int D_, int CBLK_, int ACT_, int OUT_, int FULL_ = 0, int FOLD_ = 0, int BF16_ = 0,
          int PEERS_ = 0>
struct SwStatic {
    static constexpr int BF16 = BF16_;  // 1: bf16 operands / images, one MMA pass
    // 1: the epilogue can also store rows into the latitude-band neighbours' images (SwParams::yp_up / yp_down).  A flavour of
    // its own because even the untaken branches cost the hottest epilogues 3-7 % (profiles/r02_variants_ab.txt); the
    // generic instances (NCOLS_ == 0) always carry the code.
    static constexpr int PEERS = (PEERS_ != 0 || NCOLS_ == 0) ? 1 : 0;
    static constexpr int NCOLS = NCOLS_, KS = KS_, D = D_, CBLK = CBLK_, ACT = ACT_, OUT = OUT_;  // OUT: 1 = P, 2 = fp32, 3 = both
    static constexpr int FULL = FULL_;  // 1: Cout == CBLK * NC, no partial filter block
    static constexpr int FOLD = FOLD_;  // vertical taps per MMA (0: SwParams::fold)
};
using SwGeneric = SwStatic<0, 0, 0, 0, -1, 0>;
using SwGenericBf16 = SwStatic<0, 0, 0, 0, -1, 0, 0, 0, 1>;

// accumulator ring: rows of one residue class mod D are neighbours; NACC is a multiple of D
__host__ __device__ __forceinline__ int sw_nacc(int ncols, int d) {
    int n = 512 / ncols;
    if (n > SW_MAX_ACC) n = SW_MAX_ACC;
    return n - n % d;
}
__device__ __forceinline__ int sw_slot(int G, int D, int R) { return (G % D) * R + (G / D) % R; }

template <int KH, int KW, int NC, class ST>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_sw_kernel(const SwParams p, const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_pair) {
    constexpr int NTHREADS = TC_THREADS;
    const int NCOLS = ST::NCOLS ? ST::NCOLS : p.NCOLS;
    const int KS = ST::KS ? ST::KS : p.KS;
    const int D = ST::D ? ST::D : p.D;
    const int CBLK = ST::CBLK ? ST::CBLK : p.CBLK;
    const int NACC = (ST::NCOLS && ST::D) ? sw_nacc(ST::NCOLS ? ST::NCOLS : 16, ST::D ? ST::D : 1) : p.NACC;
    const int RING = NACC / D;  // ring positions per residue class
    const int act = ST::ACT >= 0 ? ST::ACT : p.act;
    const bool has_yp = ST::OUT ? (ST::OUT & 1) != 0 : p.yp != nullptr;
    const bool has_y32 = ST::OUT ? (ST::OUT & 2) != 0 : p.y32 != nullptr;
    constexpr int CSTRIDE = NC;  // 6: the single packed block of a 6-filter layer, else 8
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + SW_MAX_ACC);
    float* sscale = reinterpret_cast<float*>(tmem_slot + 2);  // {2^-(e_in+e_w), 2^e_out}

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
        for (int s = 0; s < p.NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < NACC; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
        fence_mbar_init();
        const float2 sc = tc_scale_prologue(p.sc, act, blockIdx.x == 0);
        sscale[0] = sc.x;
        sscale[1] = sc.y;
    }
    if (warp == W_MMA) tmem_alloc512(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int Halloc = p.H + 2 * TC_HPAD;

    if (warp == W_PROD && p.use_tma) {
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
            SwIter it;
            sw_range(p, blockIdx.x, gridDim.x, it);
            while (sw_next(p, it, U)) {
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
        SwIter it;
        sw_range(p, blockIdx.x, gridDim.x, it);
        while (sw_next(p, it, U)) {
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
        // Every staged input row is issued the same way -- all KH taps, one new accumulator row started -- by giving each
        // unit SPAN "phantom" output rows above and below its band: they own ring slots like real rows, collect the taps
        // that would fall outside the band, and are released by the epilogue unread.  No boundary cases, so the row's MMA
        // runs depend only on the ring position of the window (a compile-time switch in the folded instances).
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, sm_100 descriptor version
        const uint32_t bunit16 = (uint32_t)(KH * NCOLS);                   // one 8-channel unit of a weight block: [tap][col] x 16 B
        const uint32_t bbase = (bunit16 << 16) | (smem_u32(bsm) >> 4);     // LBO field | block 0 address
        const uint32_t bblock16 = 2u * bunit16;                            // one (k step, hi|lo) block = two units
        const uint32_t pitch16 = p.rowpitch >> 4;
        const uint32_t stages16 = smem_u32(stages) >> 4, stride16 = p.stage_stride >> 4;
        int s = 0;
        uint32_t ph = 0;
        SwRing ring;        // position of the next row to start (CTA-wide numbering, phantoms included)
        SwUnit U;
        const bool timing = (p.debug & 4) != 0;
        long long t_acc = 0, t_full = 0, t_issue = 0, n_rows = 0, t_first = 0;
        const long long k_c0 = timing ? clock64() : 0;
        unsigned long long k_g0 = 0;
        if (timing) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(k_g0));
        SwIter it;
        sw_range(p, blockIdx.x, gridDim.x, it);
        long long t_ph = 0, t_unit = 0, c_prev = k_c0;
        while (sw_next(p, it, U)) {
            const int nrows = U.yb - U.ya + SPAN;
            const long long cu0 = timing ? clock64() : 0;
            for (int k = 0; k < SPAN; ++k) {        // phantom rows above the band: take their slots
                mbar_wait(&acc_empty[ring.slot(RING)], ring.lap ^ 1);
                ring.advance(D, RING);
            }
            if (timing) {
                const long long cu1 = clock64();
                t_ph += cu1 - cu0;
                t_unit += cu0 - c_prev;
            }
            for (int r = 0; r < nrows; ++r) {
                const long long c0 = timing ? clock64() : 0;
                mbar_wait(&acc_empty[ring.slot(RING)], ring.lap ^ 1);  // the row that starts here: its slot must be drained
                const long long c1 = timing ? clock64() : 0;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const long long c2 = timing ? clock64() : 0;
                const uint32_t sbase16 = stages16 + (uint32_t)s * stride16;
                int pos_lo = ring.q - (KH - 1);     // ring position of the window's lowest row (same residue class)
                if (pos_lo < 0) pos_lo += RING;
                const uint32_t tres = tmem + (uint32_t)(ring.m * RING * NCOLS);
                if (!(p.debug & 2)) {
                    if constexpr (ST::NCOLS != 0 && ST::D != 0 && ST::KS != 0) {
                        sw_issue_row_static<KH, ST>(pos_lo, leader, tres, sbase16, pitch16, desc_hi, bbase, p.kst);
                    } else {
                        const int FOLD = p.fold;
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint32_t a_lo32 = p.kst[ks].a_off + sbase16;  // (LBO field | offset) precomputed on the host
                            const uint64_t ad_hi = ((uint64_t)desc_hi << 32) | a_lo32;
                            const uint64_t ad_lo = ((uint64_t)desc_hi << 32) | (a_lo32 + pitch16);
                            const uint32_t bk = bbase + (uint32_t)(ST::BF16 ? ks : 2 * ks) * bblock16;
                            // pass 0: hi * hi, pass 1: hi * lo (same A view: collector), pass 2: lo * hi; bf16: one pass
#pragma unroll
                            for (int pass = 0; pass < (ST::BF16 ? 1 : 3); ++pass) {
                                const uint64_t ad = pass < 2 ? ad_hi : ad_lo;
                                const uint32_t bb = bk + (pass == 1 ? bblock16 : 0u);
                                const bool fresh = ks == 0 && pass == 0;  // tap 0 overwrites the new row's accumulator
                                int tb = KH - 1, pos = pos_lo;
                                while (tb >= 0) {
                                    int len = tb + 1 < FOLD ? tb + 1 : FOLD;
                                    if (len > RING - pos) len = RING - pos;          // stop at the ring's end
                                    if (fresh && len > tb && tb > 0) len = tb;        // ... and before tap 0 of a new row
                                    const uint32_t ncols = (uint32_t)(len * NCOLS);
                                    const uint64_t bd = ((uint64_t)desc_hi << 32) | (bb + (uint32_t)((KH - 1 - tb) * NCOLS));
                                    const bool first = pass != 1 && tb == KH - 1, last = (ST::BF16 || pass != 0) && len > tb;
                                    if (leader)
                                        umma_f16_dyn(first ? (last ? 0 : 1) : (last ? 3 : 2), tres + (uint32_t)(pos * NCOLS), ad, bd,
                                                     (ST::BF16 ? SW_IDESC_BF16 : SW_IDESC) | ((ncols >> 3) << 17),
                                                     (fresh && tb == 0) ? 0u : 1u);
                                    tb -= len;
                                    pos += len;
                                    if (pos == RING) pos = 0;
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (leader) {
                    umma_commit(&empty[s]);
                    umma_commit(&acc_full[ring.m * RING + pos_lo]);  // the window's lowest row has taken its last tap
                }
                if (++s == p.NS) { s = 0; ph ^= 1; }
                ring.advance(D, RING);
                if (timing) {
                    const long long c3 = clock64();
                    if (n_rows == 0) t_first = c2 - k_c0;
                    t_acc += c1 - c0; t_full += c2 - c1; t_issue += c3 - c2; ++n_rows;
                }
            }
            // phantom rows below the band (started by the last SPAN input rows, never completed): release them
            if (leader) {
                SwRing t = ring;
                for (int k = 0; k < SPAN; ++k) {
                    t.retreat(D, RING);
                    umma_commit(&acc_full[t.slot(RING)]);
                }
            }
            __syncwarp();
            if (timing) c_prev = clock64();
        }
        if (timing && leader) {
            atomicAdd(&g_tc_counters[4], (unsigned long long)t_ph);
            atomicAdd(&g_tc_counters[5], (unsigned long long)t_unit);
            atomicAdd(&g_tc_counters[0], (unsigned long long)t_acc);
            atomicAdd(&g_tc_counters[1], (unsigned long long)t_full);
            atomicAdd(&g_tc_counters[2], (unsigned long long)t_issue);
            atomicAdd(&g_tc_counters[3], (unsigned long long)n_rows);
            unsigned long long k_g1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(k_g1));
            atomicAdd(&g_tc_counters[8], (unsigned long long)(clock64() - k_c0));
            atomicAdd(&g_tc_counters[9], k_g1 - k_g0);
            atomicAdd(&g_tc_counters[10], 1ull);
            atomicAdd(&g_tc_counters[11], (unsigned long long)t_first);
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
        const float inv = sscale[0], sout = sscale[1];
        float amax_t = 0.f;   // max |output| this thread produced (true units)
        int g = 0, lrow = 0;
        const bool etiming = (p.debug & 4) != 0 && warp == 0;
        long long e_wait = 0, e_busy = 0;
        SwUnit U;
        SwIter it;
        sw_range(p, blockIdx.x, gridDim.x, it);
        while (sw_next(p, it, U)) {
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
            // rows of the unit in the CTA-wide numbering: SPAN phantom rows, the band's rows, SPAN phantom rows; set s takes
            // the rows with G % 4 == s
            const int nnum = U.yb - U.ya + 2 * SPAN;
            for (int k = (set - g) & 3; k < nnum; k += TC_SETS) {
                const int G = g + k;
                const int slot = sw_slot(G, D, RING);
                const long long ce0 = etiming ? clock64() : 0;
                mbar_wait_relaxed(&acc_full[slot], (uint32_t)((G / NACC) & 1));
                tc_fence_after();
                const long long ce1 = etiming ? clock64() : 0;
                e_wait += ce1 - ce0;
                const int y = U.ya + k - SPAN;
                if (y < U.ya || y >= U.yb || (p.debug & 1)) {   // phantom row: release the slot unread
                    tc_fence_before();
                    mbar_arrive(&acc_empty[slot]);
                    continue;
                }
                ++lrow;
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
                // ---- pass 2: shifted sums, scale, bias, activation, stores --------------------------------------------
                float* y32c = has_y32 ? y32n + (long long)y * p.ys_h : nullptr;
                uint4* row_hi = has_yp ? ypn + (size_t)y * p.Wp_out : nullptr;
                // element distance from this row of the local image to the same row of a neighbour's image (0: none)
                long long peer_d[2] = {0, 0};
                if constexpr (ST::PEERS != 0) {
                    if (has_yp && p.yp_up != nullptr && y < p.peer_up_end)
                        peer_d[0] = reinterpret_cast<uint4*>(p.yp_up) - reinterpret_cast<uint4*>(p.yp);
                    if (has_yp && p.yp_down != nullptr && y >= p.peer_down_begin)
                        peer_d[1] = reinterpret_cast<uint4*>(p.yp_down) - reinterpret_cast<uint4*>(p.yp);
                }
#pragma unroll(ST::CBLK ? ST::CBLK : 1)
                for (int cb = 0; cb < CBLK; ++cb) {
                    float d[KW][8];
#pragma unroll
                    for (int j = 0; j < KW; ++j) tmem_ld8(tbase + (cb * KW + j) * CSTRIDE, d[j]);
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + cb * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + cb * 8 + 4);
                    tmem_ld_wait();
                    float o[8];   // the accumulator holds sum * 2^(e_in + e_w)
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) o[ci] = ci < NC ? d[0][ci] : 0.f;
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
                    o[0] = fmaf(o[0], inv, b0.x); o[1] = fmaf(o[1], inv, b0.y); o[2] = fmaf(o[2], inv, b0.z);
                    o[3] = fmaf(o[3], inv, b0.w); o[4] = fmaf(o[4], inv, b1.x); o[5] = fmaf(o[5], inv, b1.y);
                    if (NC > 6) { o[6] = fmaf(o[6], inv, b1.z); o[7] = fmaf(o[7], inv, b1.w); }
                    if (lane_ok) {
                        if (act == DLWP_ACT_TANH) {
#pragma unroll
                            for (int ci = 0; ci + 1 < NC; ci += 2) tanh_accurate2(o[ci], o[ci + 1]);
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
                        const float am = amax8(o);
                        amax_t = fmaxf(amax_t, am);
                        if (has_yp && ST::BF16) {
                            if (!(am <= 3.0e38f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);    // NaN / inf
                            uint4 vb;
                            p_pack8_bf16(o, vb);
                            __stcs(row_hi, vb);
                            if (halo_r) __stcs(row_hi + p.W, vb);
                            if (halo_l) __stcs(row_hi - p.W, vb);
                            if constexpr (ST::PEERS != 0) {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    if (peer_d[k] != 0) {          // the neighbour's halo rows, over NVLink
                                        uint4* pr = row_hi + peer_d[k];
                                        __stcs(pr, vb);
                                        if (halo_r) __stcs(pr + p.W, vb);
                                        if (halo_l) __stcs(pr - p.W, vb);
                                    }
                            }
                        } else if (has_yp) {
                            // the bound behind 2^e_out makes this unreachable for finite data; NaN / inf land here
                            if (!(am * sout <= 65504.f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
                            float v[8];
#pragma unroll
                            for (int ci = 0; ci < 8; ++ci) v[ci] = o[ci] * sout;
                            uint4 vh, vl;
                            p_pack8(v, vh, vl);
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
                            if constexpr (ST::PEERS != 0) {
#pragma unroll
                                for (int k = 0; k < 2; ++k)
                                    if (peer_d[k] != 0) {          // the neighbour's halo rows, over NVLink
                                        uint4* ph = row_hi + peer_d[k];
                                        uint4* plo = row_lo + peer_d[k];
                                        __stcs(ph, vh);
                                        __stcs(plo, vl);
                                        if (halo_r) { __stcs(ph + p.W, vh); __stcs(plo + p.W, vl); }
                                        if (halo_l) { __stcs(ph - p.W, vh); __stcs(plo - p.W, vl); }
                                    }
                            }
                        }
                    }
                    if (has_yp) row_hi += (ST::BF16 ? 1 : 2) * plane_stride;
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[slot]);
                if (etiming) e_busy += clock64() - ce1;
            }
            g += nnum;
        }
        if (etiming && lane == 0) {
            atomicAdd(&g_tc_counters[6], (unsigned long long)e_wait);
            atomicAdd(&g_tc_counters[7], (unsigned long long)e_busy);
        }
        amax_publish(p.sc.amax_out, amax_t, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc512(tmem);
}

// ---- launch plumbing shared by the translation units that instantiate the kernel -------------------------------------
typedef void (*SwLaunchFn)(const SwParams& p, const CUtensorMap& map_full, const CUtensorMap& map_pair, int grid,
                           size_t smem, cudaStream_t stream);

template <int KH, int KW, int NC, class ST>
static void sw_launch_one(const SwParams& p, const CUtensorMap& map_full, const CUtensorMap& map_pair, int grid,
                          size_t smem, cudaStream_t stream) {
    static std::atomic<unsigned long long> done{0};
    if (first_use_on_device(done))
        cudaFuncSetAttribute(conv_sw_kernel<KH, KW, NC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    conv_sw_kernel<KH, KW, NC, ST><<<grid, TC_THREADS, smem, stream>>>(p, map_full, map_pair);
}

// Loads the kernel's module and opts in to its shared memory WITHOUT launching (cudaFuncGetAttributes).  With CUDA's lazy
// module loading the first launch of an instance can synchronise with kernels already running on the device -- fatal when
// that running kernel is a latitude-band halo wait spinning for a signal this very host thread has yet to enqueue
// (several bands in one process); dlwp_plan_halo_enable preloads what its rollout will launch.
template <int KH, int KW, int NC, class ST>
static void sw_preload_one() {
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, conv_sw_kernel<KH, KW, NC, ST>);
    cudaFuncSetAttribute(conv_sw_kernel<KH, KW, NC, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}
typedef void (*SwPreloadFn)();

// A folded instance: the layer constants it was compiled for and its launcher.
struct SwFolded {
    int KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL;   // the kernel folds min(KH, 256 / NCOLS) vertical taps per MMA
    SwLaunchFn fn;
    const char* what;
    int BF16;   // 1: the bf16 flavour of the instance
    int PEERS;  // 1: the flavour whose epilogue also stores into the latitude-band neighbours' images
    SwPreloadFn preload;
};
#define SW_FOLDED_ENTRY(KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL, WHAT) \
    {KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL,                          \
     &sw_launch_one<KH, KWE, NC, SwStatic<NCOLS, KS, D, CBLK, ACT, OUT, FULL, (256 / NCOLS < KH ? 256 / NCOLS : KH)>>, WHAT, 0, 0, \
     &sw_preload_one<KH, KWE, NC, SwStatic<NCOLS, KS, D, CBLK, ACT, OUT, FULL, (256 / NCOLS < KH ? 256 / NCOLS : KH)>>}
// BF16 / PEERS flavours: SW_FOLDED_ENTRY_X(..., WHAT, bf16, peers)
#define SW_FOLDED_ENTRY_X(KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL, WHAT, BF, PE)                                       \
    {KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL,                                                                          \
     &sw_launch_one<KH, KWE, NC, SwStatic<NCOLS, KS, D, CBLK, ACT, OUT, FULL, (256 / NCOLS < KH ? 256 / NCOLS : KH), BF, PE>>, \
     WHAT, BF, PE,                                                                                                             \
     &sw_preload_one<KH, KWE, NC, SwStatic<NCOLS, KS, D, CBLK, ACT, OUT, FULL, (256 / NCOLS < KH ? 256 / NCOLS : KH), BF, PE>>}
#define SW_FOLDED_ENTRY_BF16(KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL, WHAT) \
    SW_FOLDED_ENTRY_X(KH, KWE, NC, NCOLS, KS, D, CBLK, ACT, OUT, FULL, WHAT, 1, 0)

static inline int sw_tu_flags_read_clear() {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, DLWP_SW_TU_FLAGS, sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(DLWP_SW_TU_FLAGS, &zero, sizeof(int));
    return v;
}
static inline void sw_tu_counters_read_clear(unsigned long long* acc12) {
    unsigned long long v[12], zero[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(v, DLWP_SW_TU_COUNTERS, sizeof(v)) != cudaSuccess) return;
    cudaMemcpyToSymbol(DLWP_SW_TU_COUNTERS, zero, sizeof(zero));
    for (int i = 0; i < 12; ++i) acc12[i] += v[i];
}
void sw_counters_net_a(unsigned long long* acc8);
void sw_counters_net_b(unsigned long long* acc8);
void sw_counters_net_basic(unsigned long long* acc8);
void sw_counters_fused(unsigned long long* acc8);
int sw_flags_net_a();
int sw_flags_net_b();
int sw_flags_net_basic();
int sw_flags_fused();
const SwFolded* sw_folded_net_a(int* n);      // conv_sw_net_a.cu: the 2-layer benchmark net (BASELINE.json configs[0-1])
const SwFolded* sw_folded_net_b(int* n);      // conv_sw_net_b.cu: skip U-Net of examples/train_functional.py:248-275
const SwFolded* sw_folded_net_basic(int* n);  // conv_sw_net_basic.cu: examples/train.py:159-219 / train_functional.py:222-245
const SwFolded* sw_folded_bf16(int* n);       // conv_sw_bf16.cu: Net B's layers in plain bf16 (BASELINE.json configs[2])
int sw_flags_bf16();
void sw_counters_bf16(unsigned long long* acc12);
// generic bf16 instances (conv_sw_bf16.cu); false: no instance for this (KH, KW_eff, NC)
bool sw_launch_generic_bf16(int kh, int kw_eff, int nc, const SwParams& p, const CUtensorMap& map_full,
                            const CUtensorMap& map_pair, int grid, size_t smem, cudaStream_t stream);
void sw_preload_generic_bf16();

}  // namespace dlwp
