// Tensor-core (tcgen05 / TMEM / TMA) implementation of the fused periodic-pad + Conv2D layer: host side (planner, weight
// packing, launch dispatch), the generic instances of the sliding-window kernel (conv_sw.cuh), and the small kernels around
// it that let a whole U-Net run as a chain of P-layout images:
//  * pack_state_kernel / amax_kernel: fp32 (N,C,H,W) state -> P image with a power-of-two exponent chosen from max|x|;
//  * p_ew_kernel: MaxPooling2D(2) / UpSampling2D(2) / channel-window copies on P images.
// The fully folded kernel instances of the example nets' layers live in conv_sw_net_{a,b,basic}.cu.
#define DLWP_SW_TU_FLAGS g_tc_flags_main
#include "conv_sw.cuh"

#include <stdlib.h>

#include <algorithm>
#include <cstring>
#include <vector>

namespace dlwp {

// ===================================================================================================================
// fp32 (N,C,H,W) -> P layout with the periodic halo
// ===================================================================================================================
// max|x| of a strided (N,C,H,W) tensor -> atomicMax into *amax; NaN / inf raise the range flag
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, float* amax, int N, int C, int H, int W,
                                                   long long xs_n, long long xs_c, long long xs_h, int row0) {
    const long long total = (long long)N * C * H * W;  // H = rows scanned, starting at row0
    float m = 0.f;
    bool bad = false;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(idx % W);
        long long t = idx / W;
        const int y = row0 + (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        const float v = fabsf(x[(long long)n * xs_n + (long long)c * xs_c + (long long)y * xs_h + xq]);
        if (!(v <= 3.0e38f)) bad = true;
        m = fmaxf(m, v);
    }
    if (bad) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
    amax_publish(amax, m, threadIdx.x & 31);
}

// dense tensors: a flat sweep with 16-byte loads, four in flight per thread (the strided kernel above reads the bench state at
// 0.8 TB/s; this one is bound by HBM)
__global__ void __launch_bounds__(256) amax_flat_kernel(const float4* __restrict__ x, float* amax, long long n4) {
    float m = 0.f;
    bool bad = false;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(x + i + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = fmaxf(fmaxf(fabsf(v[k].x), fabsf(v[k].y)), fmaxf(fabsf(v[k].z), fabsf(v[k].w)));
            // fmaxf drops NaNs: test the sum of the four
            if (!(fabsf(v[k].x) + fabsf(v[k].y) + fabsf(v[k].z) + fabsf(v[k].w) <= 3.0e38f)) bad = true;
            m = fmaxf(m, a);
        }
    }
    for (; i < n4; i += stride) {
        const float4 v = __ldg(x + i);
        if (!(fabsf(v.x) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w) <= 3.0e38f)) bad = true;
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (bad) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
    amax_publish(amax, m, threadIdx.x & 31);
}

// one thread per (n, c8, y, padded x)
__global__ void __launch_bounds__(256) pack_state_kernel(const float* __restrict__ x, __half* __restrict__ yp, int N, int C,
                                                         int H, int W, int wpad, long long xs_n, long long xs_c,
                                                         long long xs_h, int row0, int rows, TcPackScale ps) {
    __shared__ float s_scale;
    if (threadIdx.x == 0 && ps.bf16) s_scale = 1.f;
    if (threadIdx.x == 0 && !ps.bf16) {
        int e;
        if (ps.fresh) {
            const float a = *ps.amax;
            e = tc_exp_for_bound(a);
            if (!(a < 1e30f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
            if (blockIdx.x == 0) *ps.e = e;
        } else {
            e = *ps.e;
        }
        if (blockIdx.x == 0 && ps.amax_zero) *ps.amax_zero = 0.f;
        s_scale = exp2i(e);
    }
    __syncthreads();
    const float scale = s_scale;
    const int Wp = W + 2 * wpad, C8 = (C + 7) / 8;
    const long long total = (long long)N * C8 * rows * Wp;
    float am = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(idx % Wp);
        long long t = idx / Wp;
        const int y = row0 + (int)(t % rows);
        t /= rows;
        const int c8 = (int)(t % C8);
        const int n = (int)(t / C8);
        const int gx = wrap_index(xq - wpad, W);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            v[e] = c < C ? x[(long long)n * xs_n + (long long)c * xs_c + (long long)y * xs_h + gx] : 0.f;
        }
        const float a = amax8(v);
        am = fmaxf(am, a);
        uint4* out = reinterpret_cast<uint4*>(yp);
        const long long Ha = H + 2 * TC_HPAD;
        if (ps.bf16) {
            if (!(a <= 3.0e38f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
            uint4 vb;
            p_pack8_bf16(v, vb);
            out[(((long long)n * C8 + c8) * Ha + y + TC_HPAD) * Wp + xq] = vb;
            continue;
        }
        if (!(a * scale <= 65504.f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);  // rows joining an image with a smaller bound; NaN
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= scale;
        uint4 vh, vl;
        p_pack8(v, vh, vl);
        out[(((long long)n * 2 * C8 + 2 * c8) * Ha + y + TC_HPAD) * Wp + xq] = vh;
        out[(((long long)n * 2 * C8 + 2 * c8 + 1) * Ha + y + TC_HPAD) * Wp + xq] = vl;
    }
    if (!ps.fresh && !ps.bf16) amax_publish(ps.amax, am, threadIdx.x & 31);
}

// Halo rows received from the latitude-band neighbours (two contiguous (N, C, rows, W) staging buffers) -> rows
// [row0[k], row0[k] + rows[k]) of a P image another kernel produced: the image keeps its exponent *ps.e, the packed rows'
// max|x| is max-ed into *ps.amax.  One launch for both neighbours.
struct HaloPackParams {
    const float* src[2];
    int row0[2], rows[2];
};
__global__ void __launch_bounds__(256) pack_halo_kernel(HaloPackParams hp, __half* __restrict__ yp, int N, int C, int H, int W,
                                                        int wpad, TcPackScale ps) {
    const float scale = ps.bf16 ? 1.f : exp2i(*ps.e);
    const int Wp = W + 2 * wpad, C8 = (C + 7) / 8;
    const long long per0 = (long long)N * C8 * hp.rows[0] * Wp;
    const long long total = per0 + (long long)N * C8 * hp.rows[1] * Wp;
    float am = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx >= per0 ? 1 : 0;
        long long t = idx - (k ? per0 : 0);
        const int xq = (int)(t % Wp);
        t /= Wp;
        const int yl = (int)(t % hp.rows[k]);
        t /= hp.rows[k];
        const int c8 = (int)(t % C8);
        const int n = (int)(t / C8);
        const int gx = wrap_index(xq - wpad, W);
        const float* s = hp.src[k] + ((long long)n * C * hp.rows[k] + yl) * W + gx;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c8 * 8 + e;
            v[e] = c < C ? s[(long long)c * hp.rows[k] * W] : 0.f;
        }
        const float a = amax8(v);
        am = fmaxf(am, a);
        uint4* out = reinterpret_cast<uint4*>(yp);
        const long long Ha = H + 2 * TC_HPAD;
        const int y = hp.row0[k] + yl;
        if (ps.bf16) {
            uint4 vb;
            p_pack8_bf16(v, vb);
            out[(((long long)n * C8 + c8) * Ha + y + TC_HPAD) * Wp + xq] = vb;
            continue;
        }
        if (!(a * scale <= 65504.f)) atomicOr(&g_tc_flags, TC_FLAG_RANGE);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= scale;
        uint4 vh, vl;
        p_pack8(v, vh, vl);
        out[(((long long)n * 2 * C8 + 2 * c8) * Ha + y + TC_HPAD) * Wp + xq] = vh;
        out[(((long long)n * 2 * C8 + 2 * c8 + 1) * Ha + y + TC_HPAD) * Wp + xq] = vl;
    }
    if (!ps.bf16) amax_publish(ps.amax, am, threadIdx.x & 31);
}

int tc_pack_halo(const float* const src[2], const int row0[2], const int rows[2], __half* xp, int N, int C, int H, int W,
                 int wpad, cudaStream_t stream, const TcPackScale& ps) {
    DLWP_REQUIRE(ps.bf16 || (ps.e && ps.amax), DLWP_EINVAL, "pack_halo needs the image's scale words");
    HaloPackParams hp;
    long long total = 0;
    for (int k = 0; k < 2; ++k) {
        hp.src[k] = src[k]; hp.row0[k] = row0[k]; hp.rows[k] = src[k] ? rows[k] : 0;
        total += (long long)N * ((C + 7) / 8) * hp.rows[k] * (W + 2 * wpad);
    }
    if (total <= 0) return 0;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    pack_halo_kernel<<<blocks, 256, 0, stream>>>(hp, xp, N, C, H, W, wpad, ps);
    return after_launch("pack_halo_kernel");
}

int tc_pack_state(const float* x, __half* xp, int N, int C, int H, int W, int wpad, long long xs_n, long long xs_c,
                  long long xs_h, cudaStream_t stream, int row0, int row1, const TcPackScale& ps) {
    if (row0 == 0 && row1 == 0) row1 = H;
    DLWP_REQUIRE(ps.bf16 || (ps.e && ps.amax), DLWP_EINVAL, "pack_state needs the image's scale words");
    if (ps.fresh && !ps.bf16) {
        const int a0 = (ps.amax_row0 == 0 && ps.amax_row1 == 0) ? 0 : ps.amax_row0;
        const int a1 = (ps.amax_row0 == 0 && ps.amax_row1 == 0) ? H : ps.amax_row1;
        const long long all = (long long)N * C * (a1 - a0) * W;
        const bool dense = a0 == 0 && a1 == H && xs_h == W && xs_c == (long long)H * W && xs_n == (long long)C * H * W &&
                           all % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
        if (dense) {
            const int ablocks = (int)std::max<long long>(1, std::min<long long>((all / 4 + 1023) / 1024, 148LL * 8));
            amax_flat_kernel<<<ablocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), ps.amax, all / 4);
        } else {
            const int ablocks = (int)std::max<long long>(1, std::min<long long>((all + 1023) / 1024, 148LL * 8));
            amax_kernel<<<ablocks, 256, 0, stream>>>(x, ps.amax, N, C, a1 - a0, W, xs_n, xs_c, xs_h, a0);
        }
        int rc = after_launch("amax_kernel");
        if (rc) return rc;
        if (ps.after_amax && (rc = ps.after_amax(ps.after_amax_ctx, ps.amax, stream))) return rc;
    }
    const long long total = (long long)N * ((C + 7) / 8) * (row1 - row0) * (W + 2 * wpad);
    if (total <= 0) return 0;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    pack_state_kernel<<<blocks, 256, 0, stream>>>(x, xp, N, C, H, W, wpad, xs_n, xs_c, xs_h, row0, row1 - row0, ps);
    return after_launch("pack_state_kernel");
}

// ===================================================================================================================
// Data movers on P images: MaxPooling2D(2) / UpSampling2D(2) / channel-window copy (skip connections of the U-Net).
// One thread per destination pixel and 8-channel chunk (a hi and a lo 16-byte vector); the destination's periodic halo
// columns are written with the interior.  HBM-bound byte movers: 32 B read (128 B for the pooling) + 32 B written per
// thread, coalesced along x.  The values keep the source image's exponent (max of exact hi + lo sums is scale invariant).
// ===================================================================================================================
template <int KIND, int BF16>  // KIND: 0 copy, 1 maxpool 2x2 stride 2 (floor), 2 nearest upsample x2; BF16: one plane per chunk
__global__ void __launch_bounds__(256) p_ew_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int C8, int Hs,
                                                  int Ws, int wpad_s, int src_plane0, int src_planes_total, int Hd, int Wd,
                                                  int wpad_d, int dst_plane0, int dst_planes_total, int row0, int rows,
                                                  TcScale sc) {
    constexpr int PPC = BF16 ? 1 : 2;
    if (!BF16 && blockIdx.x == 0 && threadIdx.x == 0) {  // forward the exponent and the measured amax with the data
        if (sc.e_out) *sc.e_out = sc.e_in ? *sc.e_in : sc.e_in_const;
        if (sc.amax_zero) *sc.amax_zero = 0.f;
        if (sc.amax_out) {
            const float a = sc.amax_chk ? *sc.amax_chk : (sc.amax_in ? *sc.amax_in : 1.f);
            atomicMax(reinterpret_cast<unsigned*>(sc.amax_out), __float_as_uint(a));
        }
    }
    const int Wps = Ws + 2 * wpad_s, Wpd = Wd + 2 * wpad_d;
    const long long Has = Hs + 2 * TC_HPAD, Had = Hd + 2 * TC_HPAD;
    if (KIND == 2) {
        // UpSampling2D(2): one thread per destination COLUMN and source row -- the source vector is loaded once for the two
        // destination rows, and the stores of a warp stay contiguous (one thread per source pixel writing its 2 x 2 block
        // halves the store efficiency: 2x slower on the fp16 split, profiles/r02_net_b_per_op.jsonl)
        const int ys0 = row0 >> 1, ys1 = (row0 + rows + 1) >> 1;
        const long long totals = (long long)N * C8 * (ys1 - ys0) * Wd;
        for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < totals; idx += (long long)gridDim.x * blockDim.x) {
            const int x = (int)(idx % Wd);
            long long t = idx / Wd;
            const int ys = ys0 + (int)(t % (ys1 - ys0));
            t /= (ys1 - ys0);
            const int c8 = (int)(t % C8);
            const int n = (int)(t / C8);
            const uint4* sh = src + (((long long)n * src_planes_total + src_plane0 + PPC * c8) * Has + TC_HPAD + ys) * Wps + wpad_s + (x >> 1);
            const uint4 vh = __ldg(sh);
            uint4 vl = make_uint4(0, 0, 0, 0);
            if (!BF16) vl = __ldg(sh + Has * Wps);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int y = 2 * ys + dy;
                if (y < row0 || y >= row0 + rows) continue;
                uint4* dh = dst + (((long long)n * dst_planes_total + dst_plane0 + PPC * c8) * Had + TC_HPAD + y) * Wpd + wpad_d + x;
                uint4* dl = dh + Had * Wpd;
                dh[0] = vh;
                if (!BF16) dl[0] = vl;
                if (x < wpad_d) { dh[Wd] = vh; if (!BF16) dl[Wd] = vl; }
                if (x >= Wd - wpad_d) { dh[-Wd] = vh; if (!BF16) dl[-Wd] = vl; }
            }
        }
        return;
    }
    const long long total = (long long)N * C8 * rows * Wd;  // destination rows [row0, row0 + rows): latitude-band window
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wd);
        long long t = idx / Wd;
        const int y = row0 + (int)(t % rows);
        t /= rows;
        const int c8 = (int)(t % C8);
        const int n = (int)(t / C8);
        const uint4* sh = src + (((long long)n * src_planes_total + src_plane0 + PPC * c8) * Has + TC_HPAD) * Wps + wpad_s;
        const uint4* sl = sh + Has * Wps;   // fp16 split only
        uint4 vh, vl = make_uint4(0, 0, 0, 0);
        if (KIND == 0) {
            vh = sh[(long long)y * Wps + x];
            if (!BF16) vl = sl[(long long)y * Wps + x];
        } else if (KIND == 2) {
            vh = sh[(long long)(y >> 1) * Wps + (x >> 1)];
            if (!BF16) vl = sl[(long long)(y >> 1) * Wps + (x >> 1)];
        } else {
            float m[8];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const long long o = (long long)(2 * y + dy) * Wps + 2 * x + dx;
                    float v[8];
                    if (BF16) p_unpack8_bf16(sh[o], v);
                    else p_unpack8(sh[o], sl[o], v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) m[e] = (dy | dx) ? fmaxf(m[e], v[e]) : v[e];
                }
            if (BF16) p_pack8_bf16(m, vh);    // exact: the maximum of bf16 values is one of them
            else p_pack8(m, vh, vl);
        }
        uint4* dh = dst + (((long long)n * dst_planes_total + dst_plane0 + PPC * c8) * Had + TC_HPAD + y) * Wpd + wpad_d + x;
        uint4* dl = dh + Had * Wpd;
        dh[0] = vh;
        if (!BF16) dl[0] = vl;
        if (x < wpad_d) { dh[Wd] = vh; if (!BF16) dl[Wd] = vl; }
        if (x >= Wd - wpad_d) { dh[-Wd] = vh; if (!BF16) dl[-Wd] = vl; }
    }
}

int tc_ew_launch(int kind, const __half* src, __half* dst, int N, int planes, int Hs, int Ws, int wpad_s, int src_plane0,
                 int src_planes_total, int wpad_d, int dst_plane0, int dst_planes_total, cudaStream_t stream, int row_begin,
                 int row_end, const TcScale& sc, int bf16) {
    const int C8 = bf16 ? planes : planes / 2;
    int Hd = Hs, Wd = Ws;
    if (kind == DLWP_OP_MAXPOOL) { Hd = Hs / 2; Wd = Ws / 2; }
    if (kind == DLWP_OP_UPSAMPLE) { Hd = Hs * 2; Wd = Ws * 2; }
    const int row0 = (row_begin == 0 && row_end == 0) ? 0 : row_begin;
    const int rows = ((row_begin == 0 && row_end == 0) ? Hd : row_end) - row0;
    long long total = (long long)N * C8 * rows * Wd;
    if (total <= 0) return 0;
    if (kind == DLWP_OP_UPSAMPLE) total = (long long)N * C8 * (((row0 + rows + 1) >> 1) - (row0 >> 1)) * Wd;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#define DLWP_EW_LAUNCH(K_, B_)                                                                                              \
    p_ew_kernel<K_, B_><<<blocks, 256, 0, stream>>>(s4, d4, N, C8, Hs, Ws, wpad_s, src_plane0, src_planes_total, Hd, Wd, \
                                                    wpad_d, dst_plane0, dst_planes_total, row0, rows, sc)
    if (kind == DLWP_OP_MAXPOOL) { if (bf16) DLWP_EW_LAUNCH(1, 1); else DLWP_EW_LAUNCH(1, 0); }
    else if (kind == DLWP_OP_UPSAMPLE) { if (bf16) DLWP_EW_LAUNCH(2, 1); else DLWP_EW_LAUNCH(2, 0); }
    else { if (bf16) DLWP_EW_LAUNCH(0, 1); else DLWP_EW_LAUNCH(0, 0); }
#undef DLWP_EW_LAUNCH
    return after_launch("p_ew_kernel");
}

// ===================================================================================================================
// Host side
// ===================================================================================================================
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static int g_tc_sms = 0;

// Sliding-window schedule of one layer; -1 if the geometry does not fit (TMEM columns, shared memory, K steps).
int tc_plan_layer(const DlwpConvDesc& d, TcLayer* L, const TcOptions& opt) {
    memset(L, 0, sizeof(*L));
    if (d.kh != d.kw || (d.kh != 3 && d.kh != 5)) return -1;
    const int halo_w = d.dil_w * (d.kw - 1), span = d.dil_h * (d.kh - 1);
    L->wpad = halo_w / 2;
    L->Wp = d.W + halo_w;
    L->C8 = cdiv(d.Cin, 8);
    L->ppc = opt.bf16 ? 1 : 2;
    L->planes = L->ppc * L->C8;
    if (L->planes > 32) return -1;  // the bulk-copy producer issues at most two copies per lane and row
    L->CBLK = cdiv(d.Cout, 8);
    // Many filters: N = filters is wide enough, fold the horizontal taps into K (shifted A views, no shifted sum in the
    // epilogue).  Few filters: fold them into N so that the MMA N reaches 32..96.
    L->taps_in_k = d.Cout > 16 ? 1 : 0;
    // ... unless there are many input channels and still few enough filters for the accumulator ring: with taps in K the
    // K loop is kw times longer (kw * C8 / 2 steps of N = Cout MMAs), and the single issuing warp is the bottleneck
    if (L->taps_in_k && L->C8 >= 8 && 512 / (cdiv(cdiv(d.Cout, 8) * d.kw * 8, 16) * 16) >= span + 2) L->taps_in_k = 0;
    if (opt.taps_in_k >= 0) L->taps_in_k = opt.taps_in_k ? 1 : 0;
    L->kw_eff = L->taps_in_k ? 1 : d.kw;
    L->CSTRIDE = (!L->taps_in_k && d.Cout == 6) ? 6 : 8;  // 5 taps x 6 filters pack into 32 columns
    L->NCOLS = cdiv(L->CBLK * L->kw_eff * L->CSTRIDE + (8 - L->CSTRIDE), 16) * 16;
    if (L->NCOLS > 256) return -1;
    L->NACC = sw_nacc(L->NCOLS, d.dil_h);  // a multiple of the dilation: rows of one residue class are ring neighbours
    if (L->NACC < span + 2) return -1;  // rows in flight (span + 1) plus one being drained by the epilogue
    L->fold = std::min(d.kh, 256 / L->NCOLS);  // vertical taps one MMA covers
    // Either way a strip yields 128 - halo_w outputs: with the taps in N the last halo_w lanes only feed the shifted sum,
    // with the taps in K the shifted A views of the last valid lane end at pixel 127 -- so the staged row is exactly 128
    // pixels per plane and one tensor-map box (<= 256 eight-byte elements) can bring it.
    L->S = 128 - halo_w;
    L->nfull = d.W / L->S;
    L->rem = d.W - L->nfull * L->S;
    L->pair = (L->rem > 0 && L->rem + halo_w <= 64) ? 1 : 0;
    const int units = L->C8 * (L->taps_in_k ? d.kw : 1);
    L->KS = cdiv(units, 2);
    if (L->KS > TC_MAX_KSTEPS) return -1;
    L->rowpitch = 128 * 16;
    L->stage_stride = (uint32_t)L->planes * L->rowpitch;
    L->b_bytes = (uint32_t)(L->KS * d.kh * L->ppc) * (uint32_t)(2 * L->NCOLS * 16);
    const size_t mailbox = L->kw_eff > 1 ? (size_t)TC_SETS * 2 * L->CBLK * 4 * halo_w * (L->kw_eff - 1) * 8 * 4 : 0;
    const size_t fixed = (size_t)L->b_bytes + mailbox + (size_t)L->CBLK * 32 + (2 * SW_MAX_STAGES + 2 * SW_MAX_ACC) * 8 + 128 + 1024;
    const size_t budget = 227 * 1024;
    if (fixed + L->stage_stride > budget) return -1;
    L->NS = (int)std::min<size_t>(SW_MAX_STAGES, (budget - fixed) / L->stage_stride);
    L->smem = (size_t)L->NS * L->stage_stride + fixed;
    return 0;
}

bool tc_geometry_ok(const DlwpConvDesc& d, const TcOptions& opt) {
    if (d.rowwise || d.pre_op || d.dil_h != d.dil_w) return false;
    if (d.pad_mode_h != DLWP_PAD_ZERO || d.pad_mode_w != DLWP_PAD_PERIODIC) return false;
    if (d.kw != 3 && d.kw != 5) return false;
    const int halo_w = d.dil_w * (d.kw - 1), halo_h = d.dil_h * (d.kh - 1);
    if (d.pad_l != halo_w / 2 || d.pad_r != halo_w / 2 || (halo_w & 1)) return false;  // 'same' width
    if (d.pad_t + d.pad_b != halo_h) return false;                                      // 'same' height
    if (d.W < halo_w || d.W + halo_w > 1024) return false;                              // rows are staged whole
    if ((d.kw - 1) * d.dil_w > 8) return false;
    TcLayer L;
    return tc_plan_layer(d, &L, opt) == 0;
}

// Weight image: for every K step ks a hi block then a lo block, each [2 units][kh taps][NCOLS][8] fp16 (K-major,
// LBO = kh*NCOLS*16 B between the two 8-channel units).  The vertical taps sit side by side in DESCENDING order (tap kh-1
// first): input row r feeds output rows r - i*dil, whose accumulators are adjacent TMEM columns in ascending row order, so
// one MMA over a run of taps reads a contiguous column range of the block.  A unit is (8-channel chunk c8) with all
// horizontal taps in N, column n = (cb*KW + j)*CSTRIDE + ci, or (c8, horizontal tap j) with N = filters.  Odd unit counts pair the last
// unit with a zero unit that re-reads finite data.  The weights are scaled by 2^e_w (max|w| * 2^e_w in [2^13, 2^14))
// before they are split, so that a trained layer's small weights keep all 22 bits of the split.
int tc_pack_weights(const DlwpConvDesc& d, const TcLayer& L, const float* w_host, std::vector<__half>* img,
                    TcKStep* kst_out, TcWeightScale* ws) {
    struct U { int c8, j; bool zero; };
    std::vector<U> units;
    for (int c8 = 0; c8 < L.C8; ++c8) {
        if (L.taps_in_k) for (int j = 0; j < d.kw; ++j) units.push_back({c8, j, false});
        else units.push_back({c8, -1, false});
    }
    if (units.size() & 1) units.push_back({units.back().c8, units.back().j, true});
    const bool bf16 = L.ppc == 1;
    auto aoff = [&](const U& u) {
        // the zero unit (zero weights) views finite data at a positive LBO: the lo plane of the last real unit, or (bf16:
        // no lo plane) the same plane one pixel further (stages are zero-initialised past the row's end)
        if (bf16) return (long long)u.c8 * L.rowpitch + (long long)(u.j > 0 ? u.j : 0) * d.dil_w * 16 + (u.zero ? 16 : 0);
        return (long long)(2 * u.c8 + (u.zero ? 1 : 0)) * L.rowpitch + (long long)(u.j > 0 ? u.j : 0) * d.dil_w * 16;
    };
    // scale: exact power of two from max|w|; the output bound's coefficient max_co sum|w[..., co]| in true units
    const size_t wn = (size_t)d.kh * d.kw * d.Cin * d.Cout;
    float wmax = 0.f;
    std::vector<double> l1(d.Cout, 0.0);
    for (size_t i = 0; i < wn; ++i) {
        const float a = fabsf(w_host[i]);
        if (!(a <= 3.0e38f)) return -2;  // NaN / inf weights: not representable, the caller keeps the fp32 kernels
        wmax = std::max(wmax, a);
        l1[i % d.Cout] += a;
    }
    const int e_w = bf16 ? 0 : tc_exp_for_bound(wmax);   // bf16 has fp32's range: no scaling
    const float wscale = ldexpf(1.f, e_w);
    if (ws) {
        ws->e_w = e_w;
        double m = 0.0;
        for (double v : l1) m = std::max(m, v);
        ws->l1max = (float)(m * (1.0 + 1e-6));
    }
    const size_t unit = (size_t)d.kh * L.NCOLS * 8;  // fp16 elements of one 8-channel unit of a (ks, hi|lo) block
    const size_t block = 2 * unit;
    img->assign((size_t)L.KS * L.ppc * block, __float2half(0.f));
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
                const size_t bh = ((size_t)ks * L.ppc + 0) * block + (size_t)half * unit + (size_t)(d.kh - 1 - i) * L.NCOLS * 8;
                const size_t bl = ((size_t)ks * 2 + 1) * block + (size_t)half * unit + (size_t)(d.kh - 1 - i) * L.NCOLS * 8;
                for (int cb = 0; cb < L.CBLK; ++cb)
                    for (int j = (u.j >= 0 ? u.j : 0); j < (u.j >= 0 ? u.j + 1 : d.kw); ++j)
                        for (int ci = 0; ci < L.CSTRIDE; ++ci) {
                            const int co = cb * 8 + ci;
                            if (co >= d.Cout) continue;
                            const int ncol = (cb * L.kw_eff + (u.j >= 0 ? 0 : j)) * L.CSTRIDE + ci;
                            for (int e = 0; e < 8; ++e) {
                                const int c = u.c8 * 8 + e;
                                if (c >= d.Cin) continue;
                                const float v = w_host[(((size_t)i * d.kw + j) * d.Cin + c) * d.Cout + co] * wscale;
                                if (bf16) {   // the image is a bag of 16-bit words: store the bf16 bit pattern
                                    const __nv_bfloat16 b = __float2bfloat16_rn(v);
                                    memcpy(&(*img)[bh + (size_t)ncol * 8 + e], &b, 2);
                                    continue;
                                }
                                const __half h = __float2half_rn(v);
                                (*img)[bh + (size_t)ncol * 8 + e] = h;
                                (*img)[bl + (size_t)ncol * 8 + e] = __float2half_rn(v - __half2float(h));
                            }
                        }
            }
    }
    return 0;
}

// The schedule of one launch (sw_range / sw_next): nstrips strips x rows output rows, cut into `grid` contiguous ranges.
// min_rows (DlwpPlanOptions.tc_bands; default 8): a CTA is not started for fewer output rows than this -- every range pays
// (KH-1)*dil warm-up rows per strip it touches.  Host-only arithmetic (also driven by the CPU tests through
// dlwp_debug_sw_cover); returns the grid size.
static int sw_unit_geometry(const DlwpConvDesc& d, const TcLayer& L, int sms, int min_rows, SwParams* pp) {
    SwParams& p = *pp;
    p.N = d.N; p.H = d.H; p.W = d.W; p.Wp = L.Wp;
    p.D = d.dil_w; p.pad_t = d.pad_t;
    p.S = L.S; p.nfull = L.nfull; p.rem = L.rem; p.pair = L.pair;
    const bool all_rows = d.row_begin == 0 && d.row_end == 0;
    p.row0 = all_rows ? 0 : d.row_begin;
    p.row1 = all_rows ? d.H : d.row_end;
    const int rows = p.row1 - p.row0;
    p.units_per_group = 0; p.nbands = 1; p.RB = rows > 0 ? rows : 1; p.total_units = 0;
    p.nstrips = d.N * L.nfull + (L.rem > 0 ? (L.pair ? cdiv(d.N, 2) : d.N) : 0);
    if (rows <= 0) {  // empty latitude window
        p.total_rows = 0;
        return 0;
    }
    p.total_rows = p.nstrips * rows;
    if (sms < 1) sms = 148;
    if (min_rows < 1) min_rows = 8;
    return std::max(1, std::min(sms, p.total_rows / min_rows));
}

static const SwFolded* find_folded(const DlwpConvDesc& d, const TcLayer& L, int nc, int out_mode, int peers = 0) {
    const int full = (d.Cout == L.CBLK * nc) ? 1 : 0;
    const int bf16 = L.ppc == 1 ? 1 : 0;
    typedef const SwFolded* (*TableFn)(int*);
    const TableFn tables[] = {sw_folded_net_a, sw_folded_net_b, sw_folded_net_basic, sw_folded_bf16};
    for (TableFn t : tables) {
        int n = 0;
        const SwFolded* f = t(&n);
        for (int i = 0; i < n; ++i)
            if (f[i].BF16 == bf16 && f[i].PEERS == peers && f[i].KH == d.kh && f[i].KWE == L.kw_eff && f[i].NC == nc && f[i].NCOLS == L.NCOLS && f[i].KS == L.KS &&
                f[i].D == d.dil_w && f[i].CBLK == L.CBLK && f[i].ACT == d.act && f[i].OUT == out_mode && f[i].FULL == full)
                return &f[i];
    }
    return nullptr;
}

int tc_launch(const DlwpConvDesc& d, const TcLayer& L, const TcKStep* kst, const __half* xp, const __half* bimg,
              const float* bias, float* y32, __half* yp, int wpad_out, int planes_out, cudaStream_t stream,
              const TcWindow& win, const TcScale& sc, const TcOptions& opt, const TcPeers& peers) {
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
    const int grid = sw_unit_geometry(d, L, opt.max_ctas > 0 ? std::min(g_tc_sms, opt.max_ctas) : g_tc_sms, opt.bands, &p);
    p.Cout = d.Cout; p.NCOLS = L.NCOLS; p.CBLK = L.CBLK; p.CSTRIDE = L.CSTRIDE; p.XL = (L.kw_eff - 1) * d.dil_w;
    p.KS = L.KS; p.NS = L.NS; p.NACC = L.NACC; p.fold = L.fold;
    p.planes_in = L.planes;
    p.rowpitch = L.rowpitch; p.stage_stride = L.stage_stride; p.b_unit16 = (uint32_t)(2 * L.NCOLS); p.b_bytes = L.b_bytes;
    p.idesc = (1u << 4) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major, M = 128; N is set per MMA run
    p.act = d.act; p.bias = bias; p.bimg = bimg; p.xp = xp;
    p.debug = opt.debug;
    p.bf16 = L.ppc == 1 ? 1 : 0;
    p.in_plane0 = win.in_plane0; p.in_planes_total = win.in_planes_total ? win.in_planes_total : L.planes;
    p.out_plane0 = win.out_plane0;
    p.y32 = y32; p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    p.yp = yp; p.wpad_out = wpad_out; p.Wp_out = d.W + 2 * wpad_out; p.planes_out = planes_out;
    p.yp_up = yp ? peers.up : nullptr; p.yp_down = yp ? peers.down : nullptr;
    p.peer_up_end = peers.up_end; p.peer_down_begin = peers.down_begin;
    p.sc = sc;
    for (int i = 0; i < L.KS; ++i) p.kst[i] = kst[i];
    if (grid <= 0) return 0;
    const int nc = L.CSTRIDE == 6 ? 6 : 8;
    // Tensor-map producer: the staged row is exactly 128 pixels per plane, one box per row and unit.
    CUtensorMap map_full, map_pair;
    memset(&map_full, 0, sizeof(map_full));
    memset(&map_pair, 0, sizeof(map_pair));
    p.use_tma = 0;
    if (!opt.no_tma) {
        const uint64_t Halloc = (uint64_t)d.H + 2 * TC_HPAD, row_b = (uint64_t)L.Wp * 16, plane_b = Halloc * row_b;
        const uint64_t dims3[3] = {(uint64_t)L.Wp * 2, Halloc, (uint64_t)d.N * p.in_planes_total};
        const uint64_t str3[2] = {row_b, plane_b};
        const uint32_t box3[3] = {256, 1, (uint32_t)L.planes};
        const uint64_t dims4[4] = {(uint64_t)L.Wp * 2, Halloc, (uint64_t)d.N, (uint64_t)p.in_planes_total};
        const uint64_t str4[3] = {row_b, plane_b * p.in_planes_total, plane_b};
        const uint32_t box4[4] = {128, 1, 2, (uint32_t)L.planes};
        if (encode_tensor_map_u64(&map_full, xp, 3, dims3, str3, box3) == 0 &&
            (!L.pair || encode_tensor_map_u64(&map_pair, xp, 4, dims4, str4, box4) == 0)) {
            if (!L.pair) map_pair = map_full;
            p.use_tma = 1;
        }
    }
    const int out_mode = (yp ? 1 : 0) | (y32 ? 2 : 0);
    const int want_peers = (p.yp_up || p.yp_down) ? 1 : 0;
    const SwFolded* f = opt.generic ? nullptr : find_folded(d, L, nc, out_mode, want_peers);
    if (f) f->fn(p, map_full, map_pair, grid, L.smem, stream);
    else if (p.bf16) {
        DLWP_REQUIRE(sw_launch_generic_bf16(d.kh, L.kw_eff, nc, p, map_full, map_pair, grid, L.smem, stream), DLWP_ESHAPE,
                     "no bf16 kernel instance for this layer");
    } else if (d.kh == 3 && L.kw_eff == 1) sw_launch_one<3, 1, 8, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    else if (d.kh == 5 && L.kw_eff == 1) sw_launch_one<5, 1, 8, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    else if (d.kh == 3 && nc == 8) sw_launch_one<3, 3, 8, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    else if (d.kh == 3) sw_launch_one<3, 3, 6, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    else if (d.kh == 5 && nc == 8) sw_launch_one<5, 5, 8, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    else sw_launch_one<5, 5, 6, SwGeneric>(p, map_full, map_pair, grid, L.smem, stream);
    return after_launch("conv_sw_kernel");
}

// Load (without launching) the kernel instance tc_launch would pick for this layer: see sw_preload_one.
void tc_preload(const DlwpConvDesc& d, const TcLayer& L, int out_mode, int peers, const TcOptions& opt) {
    const int nc = L.CSTRIDE == 6 ? 6 : 8;
    const SwFolded* f = opt.generic ? nullptr : find_folded(d, L, nc, out_mode, peers);
    if (f) { f->preload(); return; }
    if (L.ppc == 1) { sw_preload_generic_bf16(); return; }
    sw_preload_one<3, 1, 8, SwGeneric>(); sw_preload_one<5, 1, 8, SwGeneric>(); sw_preload_one<3, 3, 8, SwGeneric>();
    sw_preload_one<3, 3, 6, SwGeneric>(); sw_preload_one<5, 5, 8, SwGeneric>(); sw_preload_one<5, 5, 6, SwGeneric>();
}

void tc_preload_aux() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, amax_kernel);
    cudaFuncGetAttributes(&a, amax_flat_kernel);
    cudaFuncGetAttributes(&a, pack_state_kernel);
    cudaFuncGetAttributes(&a, pack_halo_kernel);
    cudaFuncGetAttributes(&a, p_ew_kernel<0, 0>); cudaFuncGetAttributes(&a, p_ew_kernel<1, 0>); cudaFuncGetAttributes(&a, p_ew_kernel<2, 0>);
    cudaFuncGetAttributes(&a, p_ew_kernel<0, 1>); cudaFuncGetAttributes(&a, p_ew_kernel<1, 1>); cudaFuncGetAttributes(&a, p_ew_kernel<2, 1>);
}

size_t tc_p_bytes(int N, int planes, int H, int Wp) { return (size_t)N * planes * (H + 2 * TC_HPAD) * Wp * 16; }

int tc_debug_flags() {
    const int parts[] = {sw_tu_flags_read_clear(), sw_flags_net_a(), sw_flags_net_b(), sw_flags_net_basic(), sw_flags_fused(),
                         sw_flags_bf16()};
    int v = 0;
    for (int p : parts) {
        if (p < 0) return -1;
        v |= p;
    }
    return v;
}

// Stand-alone layer through the tensor-core path (tests / A-B timing): packs x and the weights into temporaries.
int conv2d_fwd_tc(const DlwpConvDesc& d, const float* x, const float* w_dev, const float* bias, float* y,
                  cudaStream_t stream) {
    DLWP_REQUIRE(tc_geometry_ok(d), DLWP_ESHAPE, "geometry not supported by the tensor-core conv kernel");
    TcLayer L;
    tc_plan_layer(d, &L);
    const size_t wn = (size_t)d.kh * d.kw * d.Cin * d.Cout;
    std::vector<float> w_host(wn), b_host(bias ? d.Cout : 0);
    DLWP_CUDA_TRY(cudaMemcpyAsync(w_host.data(), w_dev, wn * 4, cudaMemcpyDeviceToHost, stream));
    if (bias) DLWP_CUDA_TRY(cudaMemcpyAsync(b_host.data(), bias, (size_t)d.Cout * 4, cudaMemcpyDeviceToHost, stream));
    DLWP_CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<__half> img;
    TcKStep kst[TC_MAX_KSTEPS];
    TcWeightScale ws;
    DLWP_REQUIRE(tc_pack_weights(d, L, w_host.data(), &img, kst, &ws) == 0, DLWP_ESHAPE, "weight packing failed");
    __half *bimg = nullptr, *xp = nullptr;
    int* words = nullptr;  // [0] exponent of the packed input, [1] its amax (float)
    const size_t xp_bytes = tc_p_bytes(d.N, L.planes, d.H, L.Wp);
    DLWP_CUDA_TRY(cudaMalloc(&bimg, img.size() * 2));
    DLWP_CUDA_TRY(cudaMalloc(&xp, xp_bytes));
    DLWP_CUDA_TRY(cudaMalloc(&words, 16));
    DLWP_CUDA_TRY(cudaMemsetAsync(words, 0, 16, stream));
    DLWP_CUDA_TRY(cudaMemsetAsync(xp, 0, xp_bytes, stream));
    DLWP_CUDA_TRY(cudaMemcpyAsync(bimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice, stream));
    TcPackScale ps;
    ps.e = words; ps.amax = reinterpret_cast<float*>(words + 1); ps.fresh = 1;
    int rc = tc_pack_state(x, xp, d.N, d.Cin, d.H, d.W, L.wpad, d.x_stride_n, d.x_stride_c, d.x_stride_h, stream, 0, 0, ps);
    TcScale sc;
    sc.e_in = words; sc.amax_in = reinterpret_cast<float*>(words + 1); sc.amax_chk = sc.amax_in;
    sc.e_w = ws.e_w; sc.l1max = ws.l1max;
    for (float b : b_host) sc.bmax = std::max(sc.bmax, fabsf(b));
    if (!rc) rc = tc_launch(d, L, kst, xp, bimg, bias, y, nullptr, 0, 0, stream, TcWindow(), sc, TcOptions());
    cudaStreamSynchronize(stream);
    cudaFree(bimg);
    cudaFree(xp);
    cudaFree(words);
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
    const int32_t v[16] = {1, L.taps_in_k, L.NCOLS, L.NACC, L.KS, L.NS, L.S, L.nfull, L.rem, L.pair, (int32_t)L.smem,
                           (int32_t)L.b_bytes, L.CBLK, L.CSTRIDE, L.planes, (int32_t)L.rowpitch};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    return 0;
}

extern "C" int64_t dlwp_debug_tc_pack(const DlwpConvDesc* desc, const float* kernel, uint16_t* image, int64_t image_cap,
                                      uint32_t* kstep_words, int32_t kstep_cap, int32_t* weight_exponent, float* l1max) {
    using namespace dlwp;
    if (!desc || !kernel || !image || !kstep_words) return DLWP_EINVAL;
    TcLayer L;
    if (!tc_geometry_ok(*desc) || tc_plan_layer(*desc, &L) != 0) return DLWP_ESHAPE;
    std::vector<__half> img;
    TcKStep kst[TC_MAX_KSTEPS];
    TcWeightScale ws;
    if (tc_pack_weights(*desc, L, kernel, &img, kst, &ws) != 0) return DLWP_ESHAPE;
    if ((int64_t)img.size() > image_cap || 2 * L.KS > kstep_cap) return DLWP_EINVAL;
    memcpy(image, img.data(), img.size() * sizeof(__half));
    for (int i = 0; i < L.KS; ++i) {
        kstep_words[2 * i] = kst[i].a_off;
        kstep_words[2 * i + 1] = kst[i].a_lbo;
    }
    if (weight_exponent) *weight_exponent = ws.e_w;
    if (l1max) *l1max = ws.l1max;
    return (int64_t)img.size();
}

extern "C" int dlwp_debug_sw_cover(const DlwpConvDesc* desc, int32_t sms, int32_t* cover, int64_t cover_elems,
                                   int32_t* info, int32_t n_info) {
    using namespace dlwp;
    DLWP_REQUIRE(desc && cover && info && n_info >= 4, DLWP_EINVAL, "bad argument");
    DLWP_REQUIRE(cover_elems == (int64_t)desc->N * desc->H * desc->W, DLWP_EINVAL, "cover must hold N*H*W counters");
    TcLayer L;
    DLWP_REQUIRE(tc_geometry_ok(*desc) && tc_plan_layer(*desc, &L) == 0, DLWP_ESHAPE, "not a sliding-window layer");
    SwParams p;
    memset(&p, 0, sizeof(p));
    const int grid = sw_unit_geometry(*desc, L, sms, 0, &p);
    long long staged_rows = 0, worst = 0;
    int live_units = 0;
    SwUnit U;
    for (int cta = 0; cta < grid; ++cta) {
      SwIter it;
      sw_range(p, cta, grid, it);
      long long mine = 0;
      while (sw_next(p, it, U)) {
        ++live_units;
        mine += U.yb - U.ya + desc->dil_h * (desc->kh - 1);
        for (int ml = 0; ml < 128; ++ml) {  // the epilogue's lane -> (sample, column) map
            const int seg = (U.paired && ml >= 64) ? 1 : 0;
            const int l = ml - seg * 64;
            const int n = seg ? U.n1 : U.n0;
            const int x = U.x0 + l;
            if (n < 0 || l >= (seg ? U.nvb : U.nva) || x >= p.W) continue;
            for (int y = U.ya; y < U.yb; ++y) ++cover[((int64_t)n * desc->H + y) * desc->W + x];
        }
      }
      staged_rows += mine;
      worst = std::max(worst, mine);
    }
    info[0] = grid; info[1] = live_units; info[2] = (int32_t)worst; info[3] = (int32_t)staged_rows;
    return 0;
}

extern "C" int dlwp_debug_exp_for_bound(float bound) { return dlwp::tc_exp_for_bound(bound); }

// Name of the folded kernel instance this layer would launch with the given outputs (out_mode: 1 = P image, 2 = fp32,
// 3 = both), or NULL when it takes a generic instance.
extern "C" const char* dlwp_debug_tc_folded(const DlwpConvDesc* desc, int32_t out_mode) {
    using namespace dlwp;
    if (!desc) return nullptr;
    TcLayer L;
    if (!tc_geometry_ok(*desc) || tc_plan_layer(*desc, &L) != 0) return nullptr;
    const SwFolded* f = find_folded(*desc, L, L.CSTRIDE == 6 ? 6 : 8, out_mode);
    return f ? f->what : nullptr;
}

// TcOptions::debug & 4 (DlwpPlanOptions.tc_debug): clock64 totals of the MMA-issuing warps since the last call, summed over
// CTAs and kernels: out[0] waiting for accumulator slots, [1] waiting for staged rows, [2] issuing, [3] rows issued.
extern "C" int dlwp_debug_counters(int64_t* out, int32_t n) {
    using namespace dlwp;
    DLWP_REQUIRE(out && n >= 8, DLWP_EINVAL, "bad argument");
    unsigned long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaDeviceSynchronize();
    sw_tu_counters_read_clear(acc);
    sw_counters_net_a(acc);
    sw_counters_net_b(acc);
    sw_counters_net_basic(acc);
    sw_counters_fused(acc);
    sw_counters_bf16(acc);
    for (int i = 0; i < 12 && i < n; ++i) out[i] = (int64_t)acc[i];
    return 0;
}
