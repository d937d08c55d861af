// Two conv layers in ONE kernel: [pad + Conv2D(C1 filters, tanh)] -> [pad + Conv2D] with the C1-channel intermediate
// kept on the SM.  For the 2-layer benchmark net (BASELINE.json configs[0-1]) the 32-channel hi/lo intermediate is
// 2 x 537 MB of the step's 1.45 GB of HBM traffic when the layers are separate kernels; fused, a rollout step reads the
// state image (32 B/pixel) and writes the fp32 series slot (24 B) and the next state image (32 B): 88 B/pixel instead of
// 346, and the step becomes bound by the tensor pipe (3 fp16 MMAs per K step for fp32-level accuracy).
//
// Structure (one persistent CTA per SM, 736 threads):
//   warp 20     producer: bulk copies of state rows (P image, 2 planes x 132 pixels) into a ring of row stages
//   warp 21     MMA issuer of layer 1: sliding window over state rows -> ring of accumulators in TMEM columns
//               [0, NACC1*NCOLS1)
//   warp 22     MMA issuer of layer 2: sliding window over intermediate rows (shared memory) -> ring in the remaining
//               columns.  Two issuing warps because a row's barrier waits and commits (~250 clk) sit on the issuing thread's
//               critical path: with one warp for both layers the tensor pipe idled through them (measured: 2260 clk per row
//               pair for 1035 clk of MMAs, profiles/r02_fused_triage.txt).  Every wait points down the row order
//               (state row -> conv1 -> epilogue 1 -> conv2 -> epilogue 2), so the blocking waits cannot deadlock.
//   warps 0-11  epilogue 1 (3 sets x 4 TMEM lane quadrants): conv1 accumulator row -> scale, bias, tanh -> scaled fp16 hi/lo
//               split -> the intermediate row stage in shared memory, laid out exactly like a staged P-image row
//               ([plane][128 pixels][8 ch]), i.e. directly usable as conv2's A operand.  Rows beyond the poles are the
//               ZeroPadding2D zeros of the second layer: written as zeros, whatever conv1 computed there.
//   warps 12-19 epilogue 2 (2 sets): conv2 accumulator row -> horizontal shifted sum, bias, activation -> fp32 series slot
//               and the next iteration's state image (periodic halo columns included), as in conv_sw_kernel.
// Strips: a unit is 128 intermediate pixels wide -> 128 - halo2 valid outputs, fed by 128 + halo1 state pixels; the
// remainder strips of two samples share a tile when rem + halo2 + halo1 <= 64.  Both layers use the phantom-row ring
// and the vertical-taps-in-N MMAs of conv_sw.cuh.
#define DLWP_SW_TU_FLAGS g_tc_flags_fused
#include "conv_sw.cuh"

#include <algorithm>
#include <cstring>

namespace dlwp {

constexpr int PF_EP1_SETS = 3, PF_EP2_SETS = 2;
constexpr int PF_THREADS = (PF_EP1_SETS + PF_EP2_SETS) * 128 + 96;
constexpr int PF_W_EP2 = PF_EP1_SETS * 4;
constexpr int PF_W_PROD = (PF_EP1_SETS + PF_EP2_SETS) * 4;
constexpr int PF_W_MMA = PF_W_PROD + 1;    // layer-1 issuer (allocates TMEM)
constexpr int PF_W_MMA2 = PF_W_PROD + 2;   // layer-2 issuer
constexpr int PF_NS = 8;    // state row stages
constexpr int PF_NI = 6;    // intermediate row stages
constexpr int PF_MAXACC = 16;

struct PairParams {
    int N, H, W;
    int Wp;                   // padded width of the state image: W + 2 * (wpad1 + wpad2)
    int pad_t1, pad_t2;
    int S, nfull, rem, pair;  // strips of the SECOND layer's output row (sw_decode)
    int units_per_group, nbands, RB, total_units;
    int row0, row1;
    int Cout2;
    uint32_t spitch, sstage;  // plane pitch / stage stride of a staged state row (bytes)
    uint32_t b1_bytes, b2_bytes;
    const float *bias1, *bias2;
    const __half *bimg1, *bimg2;
    const __half* xp;
    int in_planes_total, in_plane0;
    float* y32; long long ys_n, ys_c, ys_h;
    __half* yp; int Wp_out, wpad_out, planes_out, out_plane0;
    TcScale sc1, sc2;
    int* done_counter;        // device word, zero between launches: the last CTA checks the intermediate's amax
    float* zero_at_end;       // amax word the last CTA resets (it may be one this launch's prologues still read)
    int debug;
    TcKStep kst1[8], kst2[8];
};

// compile-time description of the pair
template <int KH1_, int D1_, int KS1_, int NCOLS1_, int CBLK1_, int PLANES1_,
          int KH2_, int KW2_, int NC2_, int D2_, int KS2_, int NCOLS2_, int CBLK2_, int ACT2_, int OUT2_, int FULL2_>
struct PairCfg {
    static constexpr int KH1 = KH1_, D1 = D1_, KS1 = KS1_, NCOLS1 = NCOLS1_, CBLK1 = CBLK1_, PLANES1 = PLANES1_;
    static constexpr int KH2 = KH2_, KW2 = KW2_, NC2 = NC2_, D2 = D2_, KS2 = KS2_, NCOLS2 = NCOLS2_, CBLK2 = CBLK2_;
    static constexpr int ACT2 = ACT2_, OUT2 = OUT2_, FULL2 = FULL2_;
    static constexpr int SPAN1 = (KH1 - 1) * D1, SPAN2 = (KH2 - 1) * D2;
    static constexpr int NACC1 = 8 / D1 * D1, NACC2 = 8 / D2 * D2;   // accumulator rings (rows in flight per layer)
    static constexpr int RING1 = NACC1 / D1, RING2 = NACC2 / D2;
    static constexpr int FOLD1 = 256 / NCOLS1 < KH1 ? 256 / NCOLS1 : KH1, FOLD2 = 256 / NCOLS2 < KH2 ? 256 / NCOLS2 : KH2;
    static constexpr int C2BASE = NACC1 * NCOLS1;                    // first TMEM column of the second layer's ring
    static constexpr int PLANES2 = 2 * CBLK1;                        // planes of the intermediate (hi, lo per 8 channels)
    static constexpr uint32_t ISTAGE = PLANES2 * 2048u;              // bytes of one intermediate row stage
    static_assert(NACC1 * NCOLS1 + NACC2 * NCOLS2 <= 512, "accumulator rings exceed TMEM");
    static_assert(RING1 >= KH1 + 1 && RING2 >= KH2 + 1, "ring too short for the tap window");
};

template <int D, int RING>
__device__ __forceinline__ int pf_slot(int G) { return (G % D) * RING + (G / D) % RING; }

// advance to the next live unit of this CTA; returns false when the list is exhausted
__device__ __forceinline__ bool pf_next_unit(const PairParams& p, int& u, SwUnit& U) {
    while (u < p.total_units) {
        if (sw_decode(p, u, U)) return true;
        u += gridDim.x;
    }
    return false;
}

template <class C>
__global__ void __launch_bounds__(PF_THREADS, 1) conv_pair_kernel(const PairParams p) {
    constexpr int KH1 = C::KH1, D1 = C::D1, KH2 = C::KH2, KW = C::KW2, NC = C::NC2, D2 = C::D2;
    constexpr int SPAN1 = C::SPAN1, SPAN2 = C::SPAN2, CBLK1 = C::CBLK1, CBLK2 = C::CBLK2;
    constexpr int RING1 = C::RING1, RING2 = C::RING2, NACC1 = C::NACC1, NACC2 = C::NACC2;
    constexpr int CSTRIDE = NC;
    constexpr int XL = (KW - 1) * D2, XQ = XL * (KW - 1) * 8;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sst = smem_raw;                                   // state row stages
    unsigned char* ist = sst + (size_t)PF_NS * p.sstage;             // intermediate row stages
    unsigned char* b1sm = ist + (size_t)PF_NI * C::ISTAGE;
    unsigned char* b2sm = b1sm + p.b1_bytes;
    float* xch = reinterpret_cast<float*>(b2sm + p.b2_bytes);        // epilogue-2 mailbox [set][parity][block][quadrant][XQ]
    float* sbias1 = xch + (size_t)PF_EP2_SETS * 2 * CBLK2 * 4 * XQ;
    float* sbias2 = sbias1 + CBLK1 * 8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sbias2 + CBLK2 * 8);
    uint64_t* s_full = bars;
    uint64_t* s_empty = s_full + PF_NS;
    uint64_t* i_full = s_empty + PF_NS;
    uint64_t* i_empty = i_full + PF_NI;
    uint64_t* a1_full = i_empty + PF_NI;
    uint64_t* a1_empty = a1_full + PF_MAXACC;
    uint64_t* a2_full = a1_empty + PF_MAXACC;
    uint64_t* a2_empty = a2_full + PF_MAXACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_empty + PF_MAXACC);
    float* sscale = reinterpret_cast<float*>(tmem_slot + 2);         // {inv1, sout1, inv2, sout2}

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (uint32_t i = tid; i < p.b1_bytes / 16; i += PF_THREADS)
        reinterpret_cast<uint4*>(b1sm)[i] = reinterpret_cast<const uint4*>(p.bimg1)[i];
    for (uint32_t i = tid; i < p.b2_bytes / 16; i += PF_THREADS)
        reinterpret_cast<uint4*>(b2sm)[i] = reinterpret_cast<const uint4*>(p.bimg2)[i];
    for (uint32_t i = tid; i < ((uint32_t)PF_NS * p.sstage + PF_NI * C::ISTAGE) / 16; i += PF_THREADS)  // stay finite
        reinterpret_cast<uint4*>(sst)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < CBLK1 * 8; i += PF_THREADS) sbias1[i] = p.bias1 != nullptr ? p.bias1[i] : 0.f;
    for (int i = tid; i < CBLK2 * 8; i += PF_THREADS) sbias2[i] = (p.bias2 != nullptr && i < p.Cout2) ? p.bias2[i] : 0.f;
    fence_proxy_async();
    if (tid == 0) {
        for (int s = 0; s < PF_NS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < PF_NI; ++s) { mbar_init(&i_full[s], 128); mbar_init(&i_empty[s], 1); }
        for (int a = 0; a < PF_MAXACC; ++a) {
            mbar_init(&a1_full[a], 1); mbar_init(&a1_empty[a], 128);
            mbar_init(&a2_full[a], 1); mbar_init(&a2_empty[a], 128);
        }
        fence_mbar_init();
        const float2 s1 = tc_scale_prologue(p.sc1, DLWP_ACT_TANH, blockIdx.x == 0);
        const float2 s2 = tc_scale_prologue(p.sc2, C::ACT2, blockIdx.x == 0);
        sscale[0] = s1.x; sscale[1] = s1.y; sscale[2] = s2.x; sscale[3] = s2.y;
    }
    if (warp == PF_W_MMA) tmem_alloc512(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int Halloc = p.H + 2 * TC_HPAD;

    if (warp == PF_W_PROD) {
        // =============================== producer: state rows, one bulk copy per (plane, segment) ==========================
        int s = 0;
        uint32_t ph = 0;
        SwUnit U;
        int u = blockIdx.x;
        while (pf_next_unit(p, u, U)) {
            const int nseg = U.n1 >= 0 ? 2 : 1;
            const int avail = p.Wp - U.x0;
            const int want = U.paired ? 64 : 128 + (KH1 - 1) * D1;   // state pixels per segment (KW1 == KH1)
            const uint32_t len = (uint32_t)min(want, avail) * 16u;
            const int ncp = C::PLANES1 * nseg;
            const uint32_t row_bytes = (uint32_t)ncp * len;
            const int seg = lane / C::PLANES1, q = lane - seg * C::PLANES1;
            const int n = seg ? U.n1 : U.n0;
            const int row0 = U.ya - p.pad_t2 - p.pad_t1 + TC_HPAD;   // first state row of the unit (physical, zero rows incl.)
            const __half* src = p.xp + ((((size_t)max(n, 0) * p.in_planes_total + p.in_plane0 + q) * Halloc + (size_t)row0) * p.Wp + U.x0) * 8;
            const uint32_t dsto = (uint32_t)q * p.spitch + (uint32_t)seg * 1024u;
            const size_t row_halfs = (size_t)p.Wp * 8;
            const int nrows = U.yb - U.ya + SPAN2 + SPAN1;
            for (int r = 0; r < nrows; ++r) {
                if (lane == 0) {
                    mbar_wait_relaxed(&s_empty[s], ph ^ 1);
                    mbar_expect_tx(&s_full[s], row_bytes);
                }
                __syncwarp();
                if (lane < ncp) bulk_load(sst + (size_t)s * p.sstage + dsto, src, len, &s_full[s]);
                src += row_halfs;
                if (++s == PF_NS) { s = 0; ph ^= 1; }
            }
            u += gridDim.x;
        }
    } else if (warp == PF_W_MMA) {
        // =============================== MMA issuer, layer 1 ===============================================================
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t bbase1 = ((uint32_t)(KH1 * C::NCOLS1) << 16) | (smem_u32(b1sm) >> 4);
        const uint32_t sst16 = smem_u32(sst) >> 4, sstride16 = p.sstage >> 4, spitch16 = p.spitch >> 4;
        const bool timing = (p.debug & 4) != 0;
        long long tw0 = 0, tw1 = 0, tw2 = 0, nr = 0;
        SwRing ring1;
        int s1 = 0;
        uint32_t sph = 0;
        SwUnit U;
        int u = blockIdx.x;
        while (pf_next_unit(p, u, U)) {
            const int T1 = U.yb - U.ya + SPAN2 + SPAN1;
            for (int k = 0; k < SPAN1; ++k) {      // phantom rows above the unit's first intermediate row
                mbar_wait(&a1_empty[ring1.slot(RING1)], ring1.lap ^ 1);
                ring1.advance(D1, RING1);
            }
            for (int r = 0; r < T1; ++r) {
                const long long c0 = timing ? clock64() : 0;
                mbar_wait(&a1_empty[ring1.slot(RING1)], ring1.lap ^ 1);
                const long long c1t = timing ? clock64() : 0;
                mbar_wait(&s_full[s1], sph);
                tc_fence_after();
                const long long c2t = timing ? clock64() : 0;
                int pos_lo = ring1.q - (KH1 - 1);
                if (pos_lo < 0) pos_lo += RING1;
                const uint32_t tres = tmem + (uint32_t)(ring1.m * RING1 * C::NCOLS1);
                if (!(p.debug & 2)) {
#define PF_POS1(P_)                                                                                                            \
    case P_:                                                                                                                   \
        if constexpr (P_ < RING1)                                                                                              \
            sw_issue_row_pos<KH1, C::NCOLS1, C::KS1, RING1, C::FOLD1, P_, false>(leader, tres, sst16 + (uint32_t)s1 * sstride16, \
                                                                                 spitch16, desc_hi, bbase1, p.kst1);           \
        break;
                    switch (pos_lo) { PF_POS1(0) PF_POS1(1) PF_POS1(2) PF_POS1(3) PF_POS1(4) PF_POS1(5) PF_POS1(6) PF_POS1(7) default: break; }
#undef PF_POS1
                }
                __syncwarp();
                if (leader) {
                    umma_commit(&s_empty[s1]);
                    umma_commit(&a1_full[ring1.m * RING1 + pos_lo]);
                }
                if (++s1 == PF_NS) { s1 = 0; sph ^= 1; }
                ring1.advance(D1, RING1);
                if (timing) { const long long c3t = clock64(); tw0 += c1t - c0; tw1 += c2t - c1t; tw2 += c3t - c2t; ++nr; }
            }
            if (leader) {                              // release the phantom rows below the unit
                SwRing t = ring1;
                for (int k = 0; k < SPAN1; ++k) { t.retreat(D1, RING1); umma_commit(&a1_full[t.slot(RING1)]); }
            }
            __syncwarp();
            u += gridDim.x;
        }
        if (timing && leader) {   // [0..2] layer 1: wait slot / wait stage / issue, [3] rows
            atomicAdd(&g_tc_counters[0], (unsigned long long)tw0);
            atomicAdd(&g_tc_counters[1], (unsigned long long)tw1);
            atomicAdd(&g_tc_counters[2], (unsigned long long)tw2);
            atomicAdd(&g_tc_counters[3], (unsigned long long)nr);
        }
    } else if (warp == PF_W_MMA2) {
        // =============================== MMA issuer, layer 2 ===============================================================
        const bool leader = elect_one();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t bbase2 = ((uint32_t)(KH2 * C::NCOLS2) << 16) | (smem_u32(b2sm) >> 4);
        const uint32_t ist16 = smem_u32(ist) >> 4;
        const bool timing = (p.debug & 4) != 0;
        long long tw0 = 0, tw1 = 0, tw2 = 0, nr = 0;
        SwRing ring2;
        int i2 = 0;
        uint32_t iph = 0;
        SwUnit U;
        int u = blockIdx.x;
        while (pf_next_unit(p, u, U)) {
            const int NI2 = U.yb - U.ya + SPAN2;
            for (int k = 0; k < SPAN2; ++k) {
                mbar_wait(&a2_empty[ring2.slot(RING2)], ring2.lap ^ 1);
                ring2.advance(D2, RING2);
            }
            for (int j = 0; j < NI2; ++j) {
                const long long c0 = timing ? clock64() : 0;
                mbar_wait(&a2_empty[ring2.slot(RING2)], ring2.lap ^ 1);
                const long long c1t = timing ? clock64() : 0;
                mbar_wait(&i_full[i2], iph);
                tc_fence_after();
                const long long c2t = timing ? clock64() : 0;
                int pos_lo = ring2.q - (KH2 - 1);
                if (pos_lo < 0) pos_lo += RING2;
                const uint32_t tres = tmem + (uint32_t)(C::C2BASE + ring2.m * RING2 * C::NCOLS2);
                if (!(p.debug & 2)) {
#define PF_POS2(P_)                                                                                                                  \
    case P_:                                                                                                                         \
        if constexpr (P_ < RING2)                                                                                                    \
            sw_issue_row_pos<KH2, C::NCOLS2, C::KS2, RING2, C::FOLD2, P_, false>(leader, tres, ist16 + (uint32_t)i2 * (C::ISTAGE >> 4), \
                                                                                 2048u >> 4, desc_hi, bbase2, p.kst2);               \
        break;
                    switch (pos_lo) { PF_POS2(0) PF_POS2(1) PF_POS2(2) PF_POS2(3) PF_POS2(4) PF_POS2(5) PF_POS2(6) PF_POS2(7) default: break; }
#undef PF_POS2
                }
                __syncwarp();
                if (leader) {
                    umma_commit(&i_empty[i2]);
                    umma_commit(&a2_full[ring2.m * RING2 + pos_lo]);
                }
                if (++i2 == PF_NI) { i2 = 0; iph ^= 1; }
                ring2.advance(D2, RING2);
                if (timing) { const long long c3t = clock64(); tw0 += c1t - c0; tw1 += c2t - c1t; tw2 += c3t - c2t; ++nr; }
            }
            if (leader) {
                SwRing t = ring2;
                for (int k = 0; k < SPAN2; ++k) { t.retreat(D2, RING2); umma_commit(&a2_full[t.slot(RING2)]); }
            }
            __syncwarp();
            u += gridDim.x;
        }
        if (timing && leader) {   // [4..6] layer 2: wait slot / wait intermediate row / issue, [7] rows
            atomicAdd(&g_tc_counters[4], (unsigned long long)tw0);
            atomicAdd(&g_tc_counters[5], (unsigned long long)tw1);
            atomicAdd(&g_tc_counters[6], (unsigned long long)tw2);
            atomicAdd(&g_tc_counters[7], (unsigned long long)nr);
        }
    } else if (warp < PF_W_EP2) {
        // =============================== epilogue 1: conv1 accumulator row -> intermediate A-operand row ===================
        const int q = warp & 3, set = warp >> 2;
        const float inv = sscale[0], sout = sscale[1];
        float amax_t = 0.f;
        int g = 0, gi = 0;     // CTA-wide counters: numbered conv1 rows (phantoms included), intermediate rows
        SwUnit U;
        int u = blockIdx.x;
        while (pf_next_unit(p, u, U)) {
            const int ni = U.yb - U.ya + SPAN2;
            const int nnum = ni + 2 * SPAN1;
            int k = (set - g % PF_EP1_SETS + PF_EP1_SETS) % PF_EP1_SETS;
            for (; k < nnum; k += PF_EP1_SETS) {
                const int G = g + k;
                const int slot = pf_slot<D1, RING1>(G);
                mbar_wait_relaxed(&a1_full[slot], (uint32_t)((G / NACC1) & 1));
                tc_fence_after();
                const int j = k - SPAN1;
                if (j < 0 || j >= ni || (p.debug & 1)) {
                    if ((p.debug & 1) && j >= 0 && j < ni) {   // triage: hand the (zero) stage on without touching it
                        const int gj = gi + j, islot = gj % PF_NI;
                        mbar_wait_relaxed(&i_empty[islot], (uint32_t)(((gj / PF_NI) & 1) ^ 1));
                        mbar_arrive(&i_full[islot]);
                    }
                    tc_fence_before();
                    mbar_arrive(&a1_empty[slot]);
                    continue;
                }
                const int gj = gi + j, islot = gj % PF_NI;
                mbar_wait_relaxed(&i_empty[islot], (uint32_t)(((gj / PF_NI) & 1) ^ 1));
                const int yi = U.ya - p.pad_t2 + j;          // true row of the intermediate
                const bool row_ok = yi >= 0 && yi < p.H;
                uint4* dst = reinterpret_cast<uint4*>(ist + (size_t)islot * C::ISTAGE) + (q * 32 + lane);
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * C::NCOLS1);
                // the whole accumulator row into registers, then the TMEM slot goes straight back to the issuer: the tanh /
                // split / store work below no longer holds a ring slot (the ring has one spare row per residue class)
                float acc[CBLK1][8];
#pragma unroll
                for (int cb = 0; cb < CBLK1; ++cb) tmem_ld8(tbase + cb * 8, acc[cb]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&a1_empty[slot]);
#pragma unroll
                for (int cb = 0; cb < CBLK1; ++cb) {
                    uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
                    if (row_ok) {
                        float o[8];
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) o[ci] = acc[cb][ci];
                        const float4 b0 = *reinterpret_cast<const float4*>(sbias1 + cb * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(sbias1 + cb * 8 + 4);
                        o[0] = fmaf(o[0], inv, b0.x); o[1] = fmaf(o[1], inv, b0.y); o[2] = fmaf(o[2], inv, b0.z);
                        o[3] = fmaf(o[3], inv, b0.w); o[4] = fmaf(o[4], inv, b1.x); o[5] = fmaf(o[5], inv, b1.y);
                        o[6] = fmaf(o[6], inv, b1.z); o[7] = fmaf(o[7], inv, b1.w);
#pragma unroll
                        for (int ci = 0; ci < 8; ci += 2) tanh_accurate2(o[ci], o[ci + 1]);
                        amax_t = fmaxf(amax_t, amax8(o));
#pragma unroll
                        for (int ci = 0; ci < 8; ++ci) o[ci] *= sout;
                        p_pack8(o, vh, vl);
                    }
                    dst[(2 * cb) * 128] = vh;                 // plane pitch 2048 B = 128 x 16 B
                    dst[(2 * cb + 1) * 128] = vl;
                }
                fence_proxy_async();                          // generic stores -> visible to the tensor core's reads
                mbar_arrive(&i_full[islot]);
            }
            g += nnum;
            gi += ni;
            u += gridDim.x;
        }
        amax_publish(p.sc1.amax_out, amax_t, lane);
    } else {
        // =============================== epilogue 2: conv2 accumulator row -> fp32 series + next state image ================
        const int q = warp & 3, set = (warp - PF_W_EP2) >> 2;
        constexpr bool has_yp = (C::OUT2 & 1) != 0, has_y32 = (C::OUT2 & 2) != 0;
        float* xset = xch + (size_t)set * 2 * CBLK2 * 4 * XQ;
        const int Hout = p.H + 2 * TC_HPAD;
        const size_t plane_stride = (size_t)Hout * p.Wp_out;
        const float inv = sscale[2], sout = sscale[3];
        float amax_t = 0.f;
        int g = 0, lrow = 0;
        SwUnit U;
        int u = blockIdx.x;
        while (pf_next_unit(p, u, U)) {
            const int ml = q * 32 + lane;
            const int seg = (U.paired && ml >= 64) ? 1 : 0;
            const int l = ml - seg * 64;
            const int n = seg ? U.n1 : U.n0;
            const int x = U.x0 + l;
            const bool lane_ok = (n >= 0) && (l < (seg ? U.nvb : U.nva)) && (x < p.W);
            const bool halo_r = x < p.wpad_out, halo_l = x >= p.W - p.wpad_out;
            float* y32n = has_y32 ? p.y32 + (long long)max(n, 0) * p.ys_n + x : nullptr;
            uint4* ypn = has_yp ? reinterpret_cast<uint4*>(p.yp) +
                                      (((size_t)max(n, 0) * p.planes_out + p.out_plane0) * Hout + TC_HPAD) * p.Wp_out + x + p.wpad_out
                                : nullptr;
            const int nnum = U.yb - U.ya + 2 * SPAN2;
            for (int k = (set - g % PF_EP2_SETS + PF_EP2_SETS) % PF_EP2_SETS; k < nnum; k += PF_EP2_SETS) {
                const int G = g + k;
                const int slot = pf_slot<D2, RING2>(G);
                mbar_wait_relaxed(&a2_full[slot], (uint32_t)((G / NACC2) & 1));
                tc_fence_after();
                const int y = U.ya + k - SPAN2;
                if (y < U.ya || y >= U.yb || (p.debug & 1)) {
                    tc_fence_before();
                    mbar_arrive(&a2_empty[slot]);
                    continue;
                }
                ++lrow;
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(C::C2BASE + slot * C::NCOLS2);
                float* mb = xset + (size_t)(lrow & 1) * CBLK2 * 4 * XQ;
                if (KW > 1) {
                    for (int cb = 0; cb < CBLK2; ++cb) {
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
                float* y32c = has_y32 ? y32n + (long long)y * p.ys_h : nullptr;
                uint4* row_hi = has_yp ? ypn + (size_t)y * p.Wp_out : nullptr;
#pragma unroll
                for (int cb = 0; cb < CBLK2; ++cb) {
                    float d[KW][8];
#pragma unroll
                    for (int j = 0; j < KW; ++j) tmem_ld8(tbase + (cb * KW + j) * CSTRIDE, d[j]);
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias2 + cb * 8);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias2 + cb * 8 + 4);
                    tmem_ld_wait();
                    float o[8];
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) o[ci] = ci < NC ? d[0][ci] : 0.f;
                    if (KW > 1) {
                        const float* xn = mb + ((size_t)cb * 4 + ((q + 1) & 3)) * XQ;
#pragma unroll
                        for (int j = 1; j < KW; ++j) {
                            const int sh = j * D2;
                            float v[8];
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) v[ci] = __shfl_down_sync(0xffffffffu, d[j][ci], sh);
                            if (lane + sh >= 32) {
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
                        if (C::ACT2 == DLWP_ACT_TANH) {
#pragma unroll
                            for (int ci = 0; ci + 1 < NC; ci += 2) tanh_accurate2(o[ci], o[ci + 1]);
                        } else if (C::ACT2 == DLWP_ACT_RELU) {
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci) o[ci] = fmaxf(o[ci], 0.f);
                        }
                        const int nreal = C::FULL2 ? NC : p.Cout2 - cb * 8;
                        if (nreal < NC) {
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci)
                                if (ci >= nreal) o[ci] = 0.f;
                        }
                        if (has_y32) {
                            float* yb = y32c + (long long)(cb * 8) * p.ys_c;
#pragma unroll
                            for (int ci = 0; ci < NC; ++ci)
                                if (ci < nreal) __stcs(yb + (long long)ci * p.ys_c, o[ci]);
                        }
                        const float am = amax8(o);
                        amax_t = fmaxf(amax_t, am);
                        if (has_yp) {
                            if (!(am * sout <= 65504.f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
                            float v[8];
#pragma unroll
                            for (int ci = 0; ci < 8; ++ci) v[ci] = o[ci] * sout;
                            uint4 vh, vl;
                            p_pack8(v, vh, vl);
                            uint4* row_lo = row_hi + plane_stride;
                            __stcs(row_hi, vh);
                            __stcs(row_lo, vl);
                            if (halo_r) { __stcs(row_hi + p.W, vh); __stcs(row_lo + p.W, vl); }
                            if (halo_l) { __stcs(row_hi - p.W, vh); __stcs(row_lo - p.W, vl); }
                        }
                    }
                    if (has_yp) row_hi += 2 * plane_stride;
                }
                tc_fence_before();
                mbar_arrive(&a2_empty[slot]);
            }
            g += nnum;
            u += gridDim.x;
        }
        amax_publish(p.sc2.amax_out, amax_t, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == PF_W_MMA) tmem_dealloc512(tmem);
    // the last CTA to finish checks the intermediate image's amax against its static scale (precision underflow)
    if (tid == 0 && p.done_counter != nullptr) {
        __threadfence();
        const int ticket = atomicAdd(p.done_counter, 1);
        if (ticket == (int)gridDim.x - 1) {
            __threadfence();
            const float a = p.sc1.amax_out ? *reinterpret_cast<volatile float*>(p.sc1.amax_out) : 1.f;
            if (a > 0.f && a * exp2i(TC_EXP_STATIC) < 0.015625f) atomicOr(&g_tc_flags, TC_FLAG_UNDERFLOW);
            if (p.zero_at_end) *p.zero_at_end = 0.f;
            *p.done_counter = 0;
        }
    }
}

// ===================================================================================================================
// Host side
// ===================================================================================================================
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Net A: conv1 6->32 3x3 dil 2 tanh (taps in K: KS 2, N 32) + conv2 32->6 5x5 linear (taps in N: 5 x 6 -> 32 columns)
using PairNetA3 = PairCfg<3, 2, 2, 32, 4, 2, 5, 5, 6, 1, 2, 32, 1, DLWP_ACT_LINEAR, 3, 1>;  // fp32 series + state image
using PairNetA2 = PairCfg<3, 2, 2, 32, 4, 2, 5, 5, 6, 1, 2, 32, 1, DLWP_ACT_LINEAR, 2, 1>;  // fp32 only

bool tc_pair_ok(const DlwpConvDesc& d1, const TcLayer& L1, const DlwpConvDesc& d2, const TcLayer& L2) {
    if (d1.H != d2.H || d1.W != d2.W || d1.Cout != d2.Cin) return false;
    // the instantiated pair (PairNetA*): everything the kernel folded at compile time must match
    if (!(d1.kh == 3 && d1.dil_h == 2 && d1.act == DLWP_ACT_TANH && L1.taps_in_k == 1 && L1.KS == 2 && L1.NCOLS == 32 &&
          L1.CBLK == 4 && L1.planes == 2 && d1.Cout == 32))
        return false;
    if (!(d2.kh == 5 && d2.dil_h == 1 && d2.act == DLWP_ACT_LINEAR && L2.taps_in_k == 0 && L2.CSTRIDE == 6 && L2.KS == 2 &&
          L2.NCOLS == 32 && L2.CBLK == 1 && d2.Cout == 6))
        return false;
    const int halo1 = d1.dil_w * (d1.kw - 1), halo2 = d2.dil_w * (d2.kw - 1);
    if (d1.W < halo1 + halo2) return false;
    return true;
}

uint32_t tc_pair_state_pitch(const DlwpConvDesc& d1) {
    // plane pitch of a staged state row: 128 + halo1 pixels of 16 B, rounded up to 128 B; the two segments of a paired
    // tile sit 1024 B apart inside it
    return (uint32_t)(((128 + d1.dil_w * (d1.kw - 1)) * 16 + 127) / 128 * 128);
}

template <class C>
static void pair_launch_one(const PairParams& p, int grid, size_t smem, cudaStream_t stream) {
    static std::atomic<unsigned long long> done{0};
    if (first_use_on_device(done))
        cudaFuncSetAttribute(conv_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    conv_pair_kernel<C><<<grid, PF_THREADS, smem, stream>>>(p);
}

static int g_pf_sms = 0;

int tc_pair_launch(const DlwpConvDesc& d1, const TcLayer& L1, const TcKStep* kst1, const __half* bimg1, const float* bias1,
                   const DlwpConvDesc& d2, const TcLayer& L2, const TcKStep* kst2, const __half* bimg2, const float* bias2,
                   const __half* xp, int in_planes_total, int in_plane0, float* y32, __half* yp, int wpad_out,
                   int planes_out, int out_plane0, const TcScale& sc1, const TcScale& sc2_in, int* done_counter,
                   cudaStream_t stream, const TcOptions& opt) {
    TcScale sc2 = sc2_in;
    float* zero_at_end = sc2.amax_zero;   // the state's amax word of this iteration: prologues of late CTAs still read it
    sc2.amax_zero = nullptr;
    if (g_pf_sms == 0) {
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaGetDeviceProperties(&prop, dev);
        g_pf_sms = prop.multiProcessorCount;
    }
    if (!(d2.row_begin == 0 && d2.row_end == 0) && d2.row_end <= d2.row_begin) return 0;
    using C = PairNetA3;
    const int halo1 = d1.dil_w * (d1.kw - 1), halo2 = d2.dil_w * (d2.kw - 1);
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.N = d2.N; p.H = d2.H; p.W = d2.W;
    p.Wp = d2.W + halo1 + halo2;
    p.pad_t1 = d1.pad_t; p.pad_t2 = d2.pad_t;
    p.S = 128 - halo2;
    p.nfull = d2.W / p.S;
    p.rem = d2.W - p.nfull * p.S;
    p.pair = (p.rem > 0 && p.rem + halo2 + halo1 <= 64) ? 1 : 0;
    const bool all_rows = d2.row_begin == 0 && d2.row_end == 0;
    p.row0 = all_rows ? 0 : d2.row_begin;
    p.row1 = all_rows ? d2.H : d2.row_end;
    const int rows = p.row1 - p.row0, span = C::SPAN1 + C::SPAN2;
    const int groups = p.pair ? cdiv(p.N, 2) : p.N;
    p.units_per_group = p.pair ? 2 * p.nfull + 1 : p.nfull + (p.rem > 0 ? 1 : 0);
    {   // latitude bands per strip: balance over the SMs against the span rows recomputed at every band edge
        int best_nb = 1;
        double best_score = -1.0;
        for (int nb = 1; nb <= 16 && nb <= rows; ++nb) {
            const int rb = cdiv(rows, nb);
            if (cdiv(rows, rb) != nb) continue;
            const long long total = (long long)groups * p.units_per_group * nb;
            const double balance = (double)total / (double)(cdiv((int)total, g_pf_sms) * (long long)g_pf_sms);
            const double useful = (double)rows / (double)(rows + span * nb);
            const double score = balance * useful;
            if (score > best_score + 1e-9) { best_score = score; best_nb = nb; }
        }
        if (opt.bands > 0) best_nb = std::min(opt.bands, rows);
        p.RB = cdiv(rows, best_nb);
        p.nbands = cdiv(rows, p.RB);
        p.total_units = groups * p.units_per_group * p.nbands;
    }
    p.Cout2 = d2.Cout;
    p.spitch = tc_pair_state_pitch(d1);
    p.sstage = (uint32_t)((C::PLANES1 * p.spitch + 127) / 128 * 128);
    p.b1_bytes = L1.b_bytes; p.b2_bytes = L2.b_bytes;
    p.bias1 = bias1; p.bias2 = bias2; p.bimg1 = bimg1; p.bimg2 = bimg2;
    p.xp = xp; p.in_planes_total = in_planes_total ? in_planes_total : C::PLANES1; p.in_plane0 = in_plane0;
    p.y32 = y32; p.ys_n = d2.y_stride_n; p.ys_c = d2.y_stride_c; p.ys_h = d2.y_stride_h;
    p.yp = yp; p.wpad_out = wpad_out; p.Wp_out = d2.W + 2 * wpad_out; p.planes_out = planes_out; p.out_plane0 = out_plane0;
    p.sc1 = sc1; p.sc2 = sc2;
    p.done_counter = done_counter;
    p.zero_at_end = zero_at_end;
    p.debug = opt.debug;
    for (int i = 0; i < L1.KS; ++i) p.kst1[i] = kst1[i];
    for (int i = 0; i < L2.KS; ++i) p.kst2[i] = kst2[i];
    const int grid = std::min(p.total_units, g_pf_sms);
    if (grid <= 0) return 0;
    constexpr int XQ = (C::KW2 - 1) * C::D2 * (C::KW2 - 1) * 8;
    const size_t smem = (size_t)PF_NS * p.sstage + (size_t)PF_NI * C::ISTAGE + p.b1_bytes + p.b2_bytes +
                        (size_t)PF_EP2_SETS * 2 * C::CBLK2 * 4 * XQ * 4 + (C::CBLK1 + C::CBLK2) * 32 +
                        (2 * PF_NS + 2 * PF_NI + 4 * PF_MAXACC) * 8 + 64 + 1024;
    DLWP_REQUIRE(smem <= 227 * 1024, DLWP_ESHAPE, "fused conv pair does not fit shared memory (%zu bytes)", smem);
    if (yp != nullptr) pair_launch_one<PairNetA3>(p, grid, smem, stream);
    else pair_launch_one<PairNetA2>(p, grid, smem, stream);
    return after_launch("conv_pair_kernel");
}

int sw_flags_fused() { return sw_tu_flags_read_clear(); }
void sw_counters_fused(unsigned long long* acc8) { sw_tu_counters_read_clear(acc8); }
}  // namespace dlwp
