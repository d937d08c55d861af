// Stand-alone data-movement layers of the example nets: PeriodicPadding2D / ZeroPadding2D (DLWP/custom.py:191-214),
// MaxPooling2D(2), UpSampling2D(2), and a strided channel-block copy (materialised concatenate / slice).
// HBM-bound byte movers: coalesced along W, grid-stride, one thread per output element.
#include "internal.h"

#include <algorithm>

namespace dlwp {

struct EwParams {
    const float* x;
    float* y;
    int N, C, H, W;      // source dims
    int Ho, Wo;          // destination dims
    int pad_t, pad_l, mode_h, mode_w;
    int row0, rows;      // destination row window: rows [row0, row0 + rows)
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;
};

enum { EW_PAD = 0, EW_POOL = 1, EW_UP = 2, EW_COPY = 3 };

// Source index of padded coordinate g (already shifted by the leading pad) on an axis of length n; ok = false: a zero.
__device__ __forceinline__ int pad_source_index(int g, int n, int mode, bool& ok) {
    if (g >= 0 && g < n) return g;
    switch (mode) {
        case DLWP_PAD_PERIODIC: return wrap_index(g, n);
        case DLWP_PAD_EDGE: return g < 0 ? 0 : n - 1;
        case DLWP_PAD_REFLECT: return g < 0 ? -g : 2 * (n - 1) - g;        // pad <= n - 1 (checked on the host)
        case DLWP_PAD_SYMMETRIC: return g < 0 ? -g - 1 : 2 * n - 1 - g;    // pad <= n
        default: ok = false; return 0;
    }
}

template <int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(const EwParams p) {
    const long long total = (long long)p.N * p.C * p.rows * p.Wo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % p.Wo);
        long long t = idx / p.Wo;
        const int yo = p.row0 + (int)(t % p.rows);
        t /= p.rows;
        const int c = (int)(t % p.C);
        const int n = (int)(t / p.C);
        const float* xc = p.x + (long long)n * p.xs_n + (long long)c * p.xs_c;
        float v;
        if (OP == EW_PAD) {
            bool ok = true;
            const int gy = pad_source_index(yo - p.pad_t, p.H, p.mode_h, ok);
            const int gx = pad_source_index(xo - p.pad_l, p.W, p.mode_w, ok);
            v = ok ? xc[(long long)gy * p.xs_h + gx] : 0.f;
        } else if (OP == EW_POOL) {
            const float* q = xc + (long long)(2 * yo) * p.xs_h + 2 * xo;
            v = fmaxf(fmaxf(q[0], q[1]), fmaxf(q[p.xs_h], q[p.xs_h + 1]));
        } else if (OP == EW_UP) {
            v = xc[(long long)(yo >> 1) * p.xs_h + (xo >> 1)];
        } else {
            v = xc[(long long)yo * p.xs_h + xo];
        }
        p.y[(long long)n * p.ys_n + (long long)c * p.ys_c + (long long)yo * p.ys_h + xo] = v;
    }
}

template <int OP>
static int launch(EwParams p, cudaStream_t stream, const char* name, int row_begin = 0, int row_end = 0) {
    int rc = check_device();
    if (rc) return rc;
    DLWP_REQUIRE(p.x && p.y, DLWP_EINVAL, "null tensor pointer");
    DLWP_REQUIRE(p.N > 0 && p.C > 0 && p.H > 0 && p.W > 0 && p.Ho > 0 && p.Wo > 0, DLWP_ESHAPE,
                 "%s: non-positive dims", name);
    if (row_begin == 0 && row_end == 0) row_end = p.Ho;
    DLWP_REQUIRE(row_begin >= 0 && row_end <= p.Ho && row_begin < row_end, DLWP_ESHAPE, "%s: bad row window [%d,%d)",
                 name, row_begin, row_end);
    p.row0 = row_begin;
    p.rows = row_end - row_begin;
    const long long total = (long long)p.N * p.C * p.rows * p.Wo;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    elementwise_kernel<OP><<<blocks, 256, 0, stream>>>(p);
    return after_launch(name);
}

// (N, H, W, C) <-> (N, C, H, W).  One thread per element, indexed in the DESTINATION order (coalesced writes; the reads are
// strided by C or H*W -- served from L2 for the small C of these nets).
__global__ void __launch_bounds__(256) layout_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C, int H,
                                                    int W, int to_nchw) {
    const long long total = (long long)N * C * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        long long t = idx;
        int n, c, h, w;
        if (to_nchw) { w = (int)(t % W); t /= W; h = (int)(t % H); t /= H; c = (int)(t % C); n = (int)(t / C); }
        else { c = (int)(t % C); t /= C; w = (int)(t % W); t /= W; h = (int)(t % H); n = (int)(t / H); }
        const long long nchw = (((long long)n * C + c) * H + h) * W + w, nhwc = (((long long)n * H + h) * W + w) * C + c;
        y[to_nchw ? nchw : nhwc] = x[to_nchw ? nhwc : nchw];
    }
}

// SeriesDataGenerator batch assembly: a gather of (time, variable) planes.  One thread per 4 consecutive floats of a plane
// when H*W is a multiple of 4 (16-byte loads / stores), else per float.
template <int VEC>
__global__ void __launch_bounds__(256) gather_series_kernel(const float* __restrict__ data, const long long* __restrict__ samples,
                                                           const int* __restrict__ sel, float* __restrict__ out, int B, int T,
                                                           int V, int V_total, long long hw, int t_off, int out_c_per_t,
                                                           int out_c0) {
    const long long per = hw / VEC;
    const long long total = (long long)B * T * V * per;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long e = idx % per;
        long long r = idx / per;
        const int v = (int)(r % V);
        r /= V;
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        const float* src = data + ((samples[b] + t_off + t) * V_total + sel[v]) * hw + e * VEC;
        float* dst = out + (((long long)b * T + t) * out_c_per_t + out_c0 + v) * hw + e * VEC;
        if (VEC == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
        else *dst = __ldg(src);
    }
}

// Halo rows of a latitude-band rollout: up to two row blocks of a (N, C, H, W) tensor <-> two contiguous staging buffers
// (N, C, rows, W), ONE launch for both neighbours (plan.cu: halo_exchange).
struct HaloCopyParams {
    const float* src[2];
    float* dst[2];
    int rows[2];
    long long ss_n[2], ss_c[2], ds_n[2], ds_c[2];   // row stride is W on both sides
    int N, C, W;
};

__global__ void __launch_bounds__(256) halo_copy_kernel(const HaloCopyParams p) {
    const long long per0 = (long long)p.N * p.C * p.rows[0] * p.W;
    const long long total = per0 + (long long)p.N * p.C * p.rows[1] * p.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = idx >= per0 ? 1 : 0;
        long long t = idx - (k ? per0 : 0);
        const int x = (int)(t % p.W);
        t /= p.W;
        const int y = (int)(t % p.rows[k]);
        t /= p.rows[k];
        const int c = (int)(t % p.C);
        const int n = (int)(t / p.C);
        p.dst[k][(long long)n * p.ds_n[k] + (long long)c * p.ds_c[k] + (long long)y * p.W + x] =
            p.src[k][(long long)n * p.ss_n[k] + (long long)c * p.ss_c[k] + (long long)y * p.W + x];
    }
}

// segment k: rows[k] rows, src[k] / dst[k] point at the block's first row; strides in elements
int halo_copy(const float* const src[2], float* const dst[2], const int rows[2], const long long ss_n[2],
              const long long ss_c[2], const long long ds_n[2], const long long ds_c[2], int N, int C, int W,
              cudaStream_t stream) {
    HaloCopyParams p;
    long long total = 0;
    for (int k = 0; k < 2; ++k) {
        p.src[k] = src[k]; p.dst[k] = dst[k]; p.rows[k] = (src[k] && dst[k]) ? rows[k] : 0;
        p.ss_n[k] = ss_n[k]; p.ss_c[k] = ss_c[k]; p.ds_n[k] = ds_n[k]; p.ds_c[k] = ds_c[k];
        total += (long long)N * C * p.rows[k] * W;
    }
    p.N = N; p.C = C; p.W = W;
    if (total <= 0) return 0;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    halo_copy_kernel<<<blocks, 256, 0, stream>>>(p);
    return after_launch("halo_copy_kernel");
}

// ConvLSTM2D gate step (keras ConvLSTM2DCell.call).  One thread per (n, f, y, x): reads the four (eight) gate
// pre-activation planes and c_prev, writes c and h -- 10 floats read, 2 written per element, coalesced along W.
struct LstmParams {
    const float* z;
    const float* r;
    const float* c_prev;
    float* c_out;
    float* h_out;
    int N, F, H, W, row0, rows, act, ract;
    long long zs_n, rs_n, cs_n, hs_n;
};

__device__ __forceinline__ float recurrent_act(float v, int ract) {
    if (ract == DLWP_RACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return fminf(fmaxf(fmaf(0.2f, v, 0.5f), 0.f), 1.f);   // keras hard_sigmoid
}

__global__ void __launch_bounds__(256) convlstm_gate_kernel(const LstmParams p) {
    const long long hw = (long long)p.H * p.W;
    const long long total = (long long)p.N * p.F * p.rows * p.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.W);
        long long t = idx / p.W;
        const int y = p.row0 + (int)(t % p.rows);
        t /= p.rows;
        const int f = (int)(t % p.F);
        const int n = (int)(t / p.F);
        const long long px = (long long)y * p.W + x;
        const float* z = p.z + (long long)n * p.zs_n + (long long)f * hw + px;
        const long long g = (long long)p.F * hw;   // distance between gate blocks
        float zi = z[0], zf = z[g], zc = z[2 * g], zo = z[3 * g];
        if (p.r != nullptr) {
            const float* r = p.r + (long long)n * p.rs_n + (long long)f * hw + px;
            zi += r[0]; zf += r[g]; zc += r[2 * g]; zo += r[3 * g];
        }
        const float ig = recurrent_act(zi, p.ract), fg = recurrent_act(zf, p.ract), og = recurrent_act(zo, p.ract);
        const float cp = p.c_prev != nullptr ? p.c_prev[(long long)n * p.cs_n + (long long)f * hw + px] : 0.f;
        const float c = fmaf(fg, cp, ig * apply_act(zc, p.act));
        p.c_out[(long long)n * p.cs_n + (long long)f * hw + px] = c;
        p.h_out[(long long)n * p.hs_n + (long long)f * hw + px] = og * apply_act(c, p.act);
    }
}

}  // namespace dlwp

using namespace dlwp;

static int check_pad_args(int H, int W, int pad_t, int pad_b, int pad_l, int pad_r, int mode_h, int mode_w) {
    DLWP_REQUIRE(pad_t >= 0 && pad_b >= 0 && pad_l >= 0 && pad_r >= 0, DLWP_ESHAPE, "negative padding");
    DLWP_REQUIRE(mode_h >= DLWP_PAD_ZERO && mode_h <= DLWP_PAD_SYMMETRIC && mode_w >= DLWP_PAD_ZERO &&
                     mode_w <= DLWP_PAD_SYMMETRIC, DLWP_EINVAL, "bad pad mode");
    const int lim_h = mode_h == DLWP_PAD_REFLECT ? H - 1 : H, lim_w = mode_w == DLWP_PAD_REFLECT ? W - 1 : W;
    if (mode_h == DLWP_PAD_PERIODIC || mode_h == DLWP_PAD_REFLECT || mode_h == DLWP_PAD_SYMMETRIC)
        DLWP_REQUIRE(pad_t <= lim_h && pad_b <= lim_h, DLWP_ESHAPE, "periodic / mirrored pad > axis");
    if (mode_w == DLWP_PAD_PERIODIC || mode_w == DLWP_PAD_REFLECT || mode_w == DLWP_PAD_SYMMETRIC)
        DLWP_REQUIRE(pad_l <= lim_w && pad_r <= lim_w, DLWP_ESHAPE, "periodic / mirrored pad > axis");
    return 0;
}

extern "C" int dlwp_convlstm_gates(const float* z, const float* r, const float* c_prev, float* c_out, float* h_out,
                                   int32_t N, int32_t F, int32_t H, int32_t W, int64_t zs_n, int64_t rs_n, int64_t cs_n,
                                   int64_t hs_n, int32_t act, int32_t ract, int32_t row_begin, int32_t row_end,
                                   dlwp_stream_t stream) {
    int rc = check_device();
    if (rc) return rc;
    DLWP_REQUIRE(z && c_out && h_out, DLWP_EINVAL, "null tensor pointer");
    DLWP_REQUIRE(N > 0 && F > 0 && H > 0 && W > 0, DLWP_ESHAPE, "convlstm_gates: non-positive dims");
    DLWP_REQUIRE(act == DLWP_ACT_LINEAR || act == DLWP_ACT_TANH || act == DLWP_ACT_RELU, DLWP_EINVAL, "bad activation");
    DLWP_REQUIRE(ract == DLWP_RACT_HARD_SIGMOID || ract == DLWP_RACT_SIGMOID, DLWP_EINVAL, "bad recurrent activation");
    if (row_begin == 0 && row_end == 0) row_end = H;
    DLWP_REQUIRE(row_begin >= 0 && row_end <= H && row_begin < row_end, DLWP_ESHAPE, "convlstm_gates: bad row window");
    LstmParams p{z, r, r ? c_prev : nullptr, c_out, h_out, N, F, H, W, row_begin, row_end - row_begin, act, ract,
                 zs_n, rs_n, cs_n, hs_n};
    const long long total = (long long)N * F * p.rows * W;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    convlstm_gate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return after_launch("convlstm_gates");
}

extern "C" int dlwp_pad2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t pad_t,
                          int32_t pad_b, int32_t pad_l, int32_t pad_r, int32_t mode_h, int32_t mode_w, int64_t xs_n,
                          int64_t xs_c, int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h, dlwp_stream_t stream) {
    int rc = check_pad_args(H, W, pad_t, pad_b, pad_l, pad_r, mode_h, mode_w);
    if (rc) return rc;
    EwParams p{x, y, N, C, H, W, H + pad_t + pad_b, W + pad_l + pad_r, pad_t, pad_l, mode_h, mode_w, 0, 0,
               xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
    return launch<EW_PAD>(p, (cudaStream_t)stream, "pad2d");
}

extern "C" int dlwp_gather_series(const float* data, const int64_t* samples, const int32_t* sel, float* out, int32_t B,
                                  int32_t T, int32_t V, int32_t V_total, int32_t H, int32_t W, int32_t t_off,
                                  int32_t out_c_per_t, int32_t out_c0, dlwp_stream_t stream) {
    int rc = check_device();
    if (rc) return rc;
    DLWP_REQUIRE(data && samples && sel && out, DLWP_EINVAL, "null pointer");
    DLWP_REQUIRE(B > 0 && T > 0 && V > 0 && V_total > 0 && H > 0 && W > 0 && t_off >= 0, DLWP_ESHAPE, "bad dims");
    DLWP_REQUIRE(out_c0 >= 0 && out_c0 + V <= out_c_per_t, DLWP_ESHAPE, "channel window outside the destination");
    const long long hw = (long long)H * W;
    const bool vec = hw % 4 == 0 && (reinterpret_cast<uintptr_t>(data) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const long long total = (long long)B * T * V * (vec ? hw / 4 : hw);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    const long long* sm = reinterpret_cast<const long long*>(samples);
    if (vec) gather_series_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(data, sm, sel, out, B, T, V, V_total, hw, t_off, out_c_per_t, out_c0);
    else gather_series_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(data, sm, sel, out, B, T, V, V_total, hw, t_off, out_c_per_t, out_c0);
    return after_launch("gather_series_kernel");
}

extern "C" int dlwp_layout2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t to_nchw,
                             dlwp_stream_t stream) {
    int rc = check_device();
    if (rc) return rc;
    DLWP_REQUIRE(x && y && x != y, DLWP_EINVAL, "layout2d needs distinct tensors");
    DLWP_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, DLWP_ESHAPE, "layout2d: non-positive dims");
    const long long total = (long long)N * C * H * W;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    layout_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, N, C, H, W, to_nchw ? 1 : 0);
    return after_launch("layout_kernel");
}

extern "C" int dlwp_maxpool2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n,
                              int64_t xs_c, int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h,
                              dlwp_stream_t stream) {
    EwParams p{x, y, N, C, H, W, H / 2, W / 2, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
    return launch<EW_POOL>(p, (cudaStream_t)stream, "maxpool2d");
}

extern "C" int dlwp_upsample2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n,
                               int64_t xs_c, int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h,
                               dlwp_stream_t stream) {
    EwParams p{x, y, N, C, H, W, H * 2, W * 2, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
    return launch<EW_UP>(p, (cudaStream_t)stream, "upsample2d");
}

extern "C" int dlwp_copy4d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n,
                           int64_t xs_c, int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h,
                           dlwp_stream_t stream) {
    EwParams p{x, y, N, C, H, W, H, W, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
    return launch<EW_COPY>(p, (cudaStream_t)stream, "copy4d");
}

extern "C" int dlwp_rows_op(int32_t op, const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W,
                            int32_t pad_t, int32_t pad_b, int32_t pad_l, int32_t pad_r, int32_t mode_h, int32_t mode_w,
                            int64_t xs_n, int64_t xs_c, int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h,
                            int32_t row_begin, int32_t row_end, dlwp_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    switch (op) {
        case DLWP_OP_PAD: {
            int rc = check_pad_args(H, W, pad_t, pad_b, pad_l, pad_r, mode_h, mode_w);
            if (rc) return rc;
            EwParams p{x, y, N, C, H, W, H + pad_t + pad_b, W + pad_l + pad_r, pad_t, pad_l, mode_h, mode_w, 0, 0,
                       xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
            return launch<EW_PAD>(p, st, "pad2d", row_begin, row_end);
        }
        case DLWP_OP_MAXPOOL: {
            EwParams p{x, y, N, C, H, W, H / 2, W / 2, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
            return launch<EW_POOL>(p, st, "maxpool2d", row_begin, row_end);
        }
        case DLWP_OP_UPSAMPLE: {
            EwParams p{x, y, N, C, H, W, H * 2, W * 2, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
            return launch<EW_UP>(p, st, "upsample2d", row_begin, row_end);
        }
        case DLWP_OP_COPY: {
            EwParams p{x, y, N, C, H, W, H, W, 0, 0, 0, 0, 0, 0, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h};
            return launch<EW_COPY>(p, st, "copy4d", row_begin, row_end);
        }
        default: DLWP_REQUIRE(false, DLWP_EINVAL, "dlwp_rows_op: unknown op %d", op);
    }
}
