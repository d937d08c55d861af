// Backward kernels of the training path (BASELINE.json configs[4]; reference: keras.Model.fit_generator as driven by
// DLWP/model/models.py:216-228, :394-402 -- forward, sum_i w_i * MSE_i, backprop, Adam).
//
//   conv_bwd_input_kernel   dX = adjoint of [periodic/zero pad -> conv]: the adjoint of PeriodicPadding2D is a wrap-ADD (halo
//                           gradients fold back onto the columns they were copied from), of ZeroPadding2D a crop.
//   conv_bwd_weight_kernel  dW[i,j,c,o] += sum_{n,y,x} dY[n,o,y,x] * Xpad[n,c,y+d*i,x+d*j]: one CTA per (tap, c, 32-filter
//                           chunk, batch slice), 32 fp32 accumulators per thread over a coalesced pixel sweep, block
//                           reduction, one atomicAdd per weight.
//   elementwise adjoints    activation', MaxPooling2D(2) (route to the first arg-max), UpSampling2D(2) (sum of the 2x2 block),
//                           channel-window add (slice / concatenate), MSE gradient + loss / MAE reduction, Adam.
// fp32 FFMA kernels: correctness-first (gradient parity vs torch autograd), not yet tuned like the forward path.
#include "internal.h"

#include <algorithm>

namespace dlwp {

struct BwdParams {
    const float* x;    // forward input of the layer (bwd_weight) or unused
    const float* dy;   // gradient w.r.t. the layer output (already multiplied by activation')
    const float* w;    // Keras layout (kh,kw,Cin,Cout)
    float* dx;
    float* dw;
    float* db;
    int N, Cin, H, W, Cout, Ho, Wo, kh, kw, dh, dw_, pad_t, pad_l, mode_h, mode_w;
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;
    int nsplit;
};

// ---- dX ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_bwd_input_kernel(const BwdParams p) {
    const long long total = (long long)p.N * p.Cin * p.H * p.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.W);
        long long t = idx / p.W;
        const int y = (int)(t % p.H);
        t /= p.H;
        const int c = (int)(t % p.Cin);
        const int n = (int)(t / p.Cin);
        float acc = 0.f;
        for (int i = 0; i < p.kh; ++i) {
            // output rows yo whose tap i reads source row y:  (yo + dh*i - pad_t) == y  (mod H when periodic)
            const int yb = y + p.pad_t - p.dh * i;
            const int ky0 = p.mode_h == DLWP_PAD_PERIODIC ? -1 : 0, ky1 = p.mode_h == DLWP_PAD_PERIODIC ? 1 : 0;
            for (int ky = ky0; ky <= ky1; ++ky) {
                const int yo = yb + ky * p.H;
                if (yo < 0 || yo >= p.Ho) continue;
                for (int j = 0; j < p.kw; ++j) {
                    const int xb = x + p.pad_l - p.dw_ * j;
                    const int kx0 = p.mode_w == DLWP_PAD_PERIODIC ? -1 : 0, kx1 = p.mode_w == DLWP_PAD_PERIODIC ? 1 : 0;
                    for (int kx = kx0; kx <= kx1; ++kx) {
                        const int xo = xb + kx * p.W;
                        if (xo < 0 || xo >= p.Wo) continue;
                        const float* wt = p.w + ((long long)(i * p.kw + j) * p.Cin + c) * p.Cout;
                        const float* g = p.dy + (long long)n * p.ys_n + (long long)yo * p.ys_h + xo;
                        for (int o = 0; o < p.Cout; ++o) acc = fmaf(g[(long long)o * p.ys_c], __ldg(wt + o), acc);
                    }
                }
            }
        }
        float* dst = p.dx + (long long)n * p.xs_n + (long long)c * p.xs_c + (long long)y * p.xs_h + x;
        *dst += acc;
    }
}

// ---- dW, dB -------------------------------------------------------------------------------------------------------------
constexpr int BW_OCH = 32;  // filters per CTA
__global__ void __launch_bounds__(256) conv_bwd_weight_kernel(const BwdParams p) {
    // blockIdx.x -> (tap, c, filter chunk); blockIdx.y -> batch slice
    const int nchunks = (p.Cout + BW_OCH - 1) / BW_OCH;
    int b = blockIdx.x;
    const int oc = b % nchunks; b /= nchunks;
    const int c = b % p.Cin; b /= p.Cin;
    const int j = b % p.kw;
    const int i = b / p.kw;
    const int o0 = oc * BW_OCH;
    const int no = min(BW_OCH, p.Cout - o0);
    float acc[BW_OCH];
#pragma unroll
    for (int o = 0; o < BW_OCH; ++o) acc[o] = 0.f;
    const long long plane = (long long)p.Ho * p.Wo;
    const int n0 = (int)((long long)p.N * blockIdx.y / p.nsplit), n1 = (int)((long long)p.N * (blockIdx.y + 1) / p.nsplit);
    for (int n = n0; n < n1; ++n) {
        const float* xc = p.x + (long long)n * p.xs_n + (long long)c * p.xs_c;
        const float* gn = p.dy + (long long)n * p.ys_n + (long long)o0 * p.ys_c;
        for (long long q = threadIdx.x; q < plane; q += blockDim.x) {
            const int yo = (int)(q / p.Wo), xo = (int)(q - (long long)yo * p.Wo);
            int gy = yo + p.dh * i - p.pad_t, gx = xo + p.dw_ * j - p.pad_l;
            bool ok = true;
            if (p.mode_h == DLWP_PAD_PERIODIC) gy = wrap_index(gy, p.H);
            else ok = gy >= 0 && gy < p.H;
            if (p.mode_w == DLWP_PAD_PERIODIC) gx = wrap_index(gx, p.W);
            else ok = ok && gx >= 0 && gx < p.W;
            if (!ok) continue;
            const float xv = xc[(long long)gy * p.xs_h + gx];
            const float* g = gn + (long long)yo * p.ys_h + xo;
#pragma unroll
            for (int o = 0; o < BW_OCH; ++o)
                if (o < no) acc[o] = fmaf(xv, g[(long long)o * p.ys_c], acc[o]);
        }
    }
    __shared__ float red[8][BW_OCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 0; o < BW_OCH; ++o) {
        float v = acc[o];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[warp][o] = v;
    }
    __syncthreads();
    if (threadIdx.x < no) {
        float v = 0.f;
        for (int wq = 0; wq < 8; ++wq) v += red[wq][threadIdx.x];
        atomicAdd(p.dw + ((long long)(i * p.kw + j) * p.Cin + c) * p.Cout + o0 + threadIdx.x, v);
    }
}

// ---- dW, tiled ----------------------------------------------------------------------------------------------------------
// dW[i,j,c,o] = sum over pixels of X(c, y + d*i - pad_t, x + d*j - pad_l) * dY(o, y, x).  The kernel above re-reads all of dY
// for every (tap, channel) and does one FMA per load; it took 71 % of a training step (profiles/r02_launches_train.txt).
// Here a CTA owns 16 input channels x up to 64 filters for a share of the output-pixel tiles: per tile the (padded) X tile
// and the dY tile are staged in shared memory once; thread (channel, filter group) keeps K*K*OT accumulators in registers
// and slides a K x (d*(K-1)+1) register window along the row, so a pixel costs K + OT shared-memory loads for K*K*OT FMAs.
// Partial sums go to the gradient with one atomicAdd per weight and CTA.
constexpr int BWT_CC = 16, BWT_TH = 8, BWT_TW = 32;

template <int K, int D, int OT>
__global__ void __launch_bounds__(256) conv_bwd_weight_tiled_kernel(const BwdParams p, int OC, int tiles_x, int tiles_y) {
    constexpr int WW = D * (K - 1) + 1, XH = BWT_TH + D * (K - 1), XW = BWT_TW + D * (K - 1), XWP = XW | 1;
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                   // [BWT_CC][XH][XWP]
    float* ds = sm + BWT_CC * XH * XWP;               // [OC][BWT_TH * BWT_TW]
    const int ochunks = (p.Cout + OC - 1) / OC;
    const int c0 = (blockIdx.x / ochunks) * BWT_CC, o0 = (blockIdx.x % ochunks) * OC;
    const int cl = threadIdx.x % BWT_CC, og = threadIdx.x / BWT_CC;
    float acc[K * K][OT];
    float bsum[OT];      // dB: the channel-0 thread of every filter group of the FIRST channel chunk sums its dY values
#pragma unroll
    for (int q = 0; q < OT; ++q) bsum[q] = 0.f;
    const bool do_bias = p.db != nullptr && c0 == 0 && cl == 0;
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
        for (int q = 0; q < OT; ++q) acc[t][q] = 0.f;
    const long long ntiles = (long long)p.N * tiles_y * tiles_x;
    for (long long tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
        const int tx = (int)(tile % tiles_x);
        long long r = tile / tiles_x;
        const int ty = (int)(r % tiles_y), n = (int)(r / tiles_y);
        const int y0 = ty * BWT_TH, x0 = tx * BWT_TW;
        __syncthreads();
        for (int idx = threadIdx.x; idx < BWT_CC * XH * XW; idx += blockDim.x) {
            const int xx = idx % XW;
            int q = idx / XW;
            const int yy = q % XH, c = q / XH;
            int gy = y0 + yy - p.pad_t, gx = x0 + xx - p.pad_l;
            bool ok = c0 + c < p.Cin;
            if (p.mode_h == DLWP_PAD_PERIODIC) gy = wrap_index(gy, p.H);
            else ok = ok && gy >= 0 && gy < p.H;
            if (p.mode_w == DLWP_PAD_PERIODIC) gx = wrap_index(gx, p.W);
            else ok = ok && gx >= 0 && gx < p.W;
            xs[(c * XH + yy) * XWP + xx] =
                ok ? __ldg(p.x + (long long)n * p.xs_n + (long long)(c0 + c) * p.xs_c + (long long)gy * p.xs_h + gx) : 0.f;
        }
        for (int idx = threadIdx.x; idx < OC * BWT_TH * BWT_TW; idx += blockDim.x) {
            const int xx = idx % BWT_TW;
            int q = idx / BWT_TW;
            const int yy = q % BWT_TH, o = q / BWT_TH;
            const bool ok = o0 + o < p.Cout && y0 + yy < p.Ho && x0 + xx < p.Wo;
            ds[o * (BWT_TH * BWT_TW) + yy * BWT_TW + xx] =
                ok ? __ldg(p.dy + (long long)n * p.ys_n + (long long)(o0 + o) * p.ys_c + (long long)(y0 + yy) * p.ys_h + x0 + xx) : 0.f;
        }
        __syncthreads();
        if (og * OT < OC) {
            const float* xc = xs + cl * XH * XWP;
            const float* dq = ds + (og * OT) * (BWT_TH * BWT_TW);
#pragma unroll 1
            for (int yy = 0; yy < BWT_TH; ++yy) {
                float win[K][WW];
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int t = 0; t + 1 < WW; ++t) win[i][t + 1] = xc[(yy + i * D) * XWP + t];
#pragma unroll
                for (int xx = 0; xx < BWT_TW; ++xx) {
#pragma unroll
                    for (int i = 0; i < K; ++i) {
#pragma unroll
                        for (int t = 0; t + 1 < WW; ++t) win[i][t] = win[i][t + 1];
                        win[i][WW - 1] = xc[(yy + i * D) * XWP + xx + WW - 1];
                    }
                    float g[OT];
#pragma unroll
                    for (int q = 0; q < OT; ++q) g[q] = dq[q * (BWT_TH * BWT_TW) + yy * BWT_TW + xx];
                    if (do_bias) {
#pragma unroll
                        for (int q = 0; q < OT; ++q) bsum[q] += g[q];
                    }
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int j = 0; j < K; ++j)
#pragma unroll
                            for (int q = 0; q < OT; ++q) acc[i * K + j][q] = fmaf(win[i][j * D], g[q], acc[i * K + j][q]);
                }
            }
        }
    }
    if (og * OT < OC && c0 + cl < p.Cin) {
#pragma unroll
        for (int t = 0; t < K * K; ++t)
#pragma unroll
            for (int q = 0; q < OT; ++q) {
                const int o = o0 + og * OT + q;
                if (o < p.Cout) atomicAdd(p.dw + ((long long)t * p.Cin + c0 + cl) * p.Cout + o, acc[t][q]);
            }
    }
    if (do_bias && og * OT < OC) {
#pragma unroll
        for (int q = 0; q < OT; ++q)
            if (o0 + og * OT + q < p.Cout) atomicAdd(p.db + o0 + og * OT + q, bsum[q]);
    }
}

template <int K, int D, int OT>
static int launch_bwd_weight_tiled(const BwdParams& p, cudaStream_t stream) {
    constexpr int XH = BWT_TH + D * (K - 1), XW = BWT_TW + D * (K - 1), XWP = XW | 1;
    int OC = std::min(64, (p.Cout + OT - 1) / OT * OT);
    OC = std::min(OC, 16 * OT);                       // 256 threads = 16 channels x 16 filter groups
    const int threads = BWT_CC * (OC / OT);
    const size_t smem = sizeof(float) * ((size_t)BWT_CC * XH * XWP + (size_t)OC * BWT_TH * BWT_TW);
    static std::atomic<unsigned long long> done{0};
    if (first_use_on_device(done))
        cudaFuncSetAttribute(conv_bwd_weight_tiled_kernel<K, D, OT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int tiles_x = (p.Wo + BWT_TW - 1) / BWT_TW, tiles_y = (p.Ho + BWT_TH - 1) / BWT_TH;
    const int gx = ((p.Cin + BWT_CC - 1) / BWT_CC) * ((p.Cout + OC - 1) / OC);
    const long long ntiles = (long long)p.N * tiles_x * tiles_y;
    const int gy = (int)std::max<long long>(1, std::min<long long>(ntiles, (148 * 4 + gx - 1) / gx));
    conv_bwd_weight_tiled_kernel<K, D, OT><<<dim3(gx, gy), threads, smem, stream>>>(p, OC, tiles_x, tiles_y);
    return after_launch("conv_bwd_weight_tiled_kernel");
}

// W'(i, j, o, c) = W(kh-1-i, kw-1-j, c, o): the weights of the forward conv that computes dX from dY
__global__ void __launch_bounds__(256) flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int kh, int kw,
                                                            int Cin, int Cout) {
    const int total = kh * kw * Cin * Cout;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = idx % Cin;
        int r = idx / Cin;
        const int o = r % Cout;
        r /= Cout;
        const int j = r % kw, i = r / kw;
        wt[idx] = w[(((long long)(kh - 1 - i) * kw + (kw - 1 - j)) * Cin + c) * Cout + o];
    }
}

__global__ void __launch_bounds__(256) bias_grad_kernel(const float* dy, float* db, int N, int Cout, int Ho, int Wo,
                                                        long long ys_n, long long ys_c, long long ys_h) {
    const int o = blockIdx.x;
    float acc = 0.f;
    const long long plane = (long long)Ho * Wo;
    for (int n = blockIdx.y; n < N; n += gridDim.y)
        for (long long q = threadIdx.x; q < plane; q += blockDim.x)
            acc += dy[(long long)n * ys_n + (long long)o * ys_c + (q / Wo) * ys_h + (q % Wo)];
    __shared__ float red[8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int wq = 0; wq < 8; ++wq) v += red[wq];
        atomicAdd(db + o, v);
    }
}

// ---- elementwise adjoints --------------------------------------------------------------------------------------------------
struct EwBwd {
    const float* a;   // forward tensor (activation output / pooling input)
    const float* g;   // incoming gradient
    float* out;       // gradient to accumulate into (or modify in place)
    int N, C, H, W;   // iteration space
    long long as_n, as_c, as_h, gs_n, gs_c, gs_h, os_n, os_c, os_h;
    int act;
};
enum { BW_ACT = 0, BW_POOL = 1, BW_UP = 2, BW_ADD = 3 };

template <int OP>
__global__ void __launch_bounds__(256) ew_bwd_kernel(const EwBwd p) {
    const long long total = (long long)p.N * p.C * p.H * p.W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.W);
        long long t = idx / p.W;
        const int y = (int)(t % p.H);
        t /= p.H;
        const int c = (int)(t % p.C);
        const int n = (int)(t / p.C);
        if (OP == BW_ACT) {  // g (in place) *= act'(a), iteration space = the activation tensor
            float* g = p.out + (long long)n * p.os_n + (long long)c * p.os_c + (long long)y * p.os_h + x;
            const float a = p.a[(long long)n * p.as_n + (long long)c * p.as_c + (long long)y * p.as_h + x];
            if (p.act == DLWP_ACT_TANH) *g *= (1.f - a * a);
            else if (p.act == DLWP_ACT_RELU) *g = a > 0.f ? *g : 0.f;
        } else if (OP == BW_POOL) {  // iteration space = pooled tensor; a = pooling input, out = its gradient
            const float* q = p.a + (long long)n * p.as_n + (long long)c * p.as_c + (long long)(2 * y) * p.as_h + 2 * x;
            int by = 0, bx = 0;
            float best = q[0];
            if (q[1] > best) { best = q[1]; bx = 1; }
            if (q[p.as_h] > best) { best = q[p.as_h]; by = 1; bx = 0; }
            if (q[p.as_h + 1] > best) { by = 1; bx = 1; }
            const float g = p.g[(long long)n * p.gs_n + (long long)c * p.gs_c + (long long)y * p.gs_h + x];
            p.out[(long long)n * p.os_n + (long long)c * p.os_c + (long long)(2 * y + by) * p.os_h + 2 * x + bx] += g;
        } else if (OP == BW_UP) {  // iteration space = the low-resolution tensor
            const float* g = p.g + (long long)n * p.gs_n + (long long)c * p.gs_c + (long long)(2 * y) * p.gs_h + 2 * x;
            p.out[(long long)n * p.os_n + (long long)c * p.os_c + (long long)y * p.os_h + x] +=
                g[0] + g[1] + g[p.gs_h] + g[p.gs_h + 1];
        } else {  // BW_ADD
            p.out[(long long)n * p.os_n + (long long)c * p.os_c + (long long)y * p.os_h + x] +=
                p.g[(long long)n * p.gs_n + (long long)c * p.gs_c + (long long)y * p.gs_h + x];
        }
    }
}

// dL/dyhat = scale * (yhat - y); stats[0] += sum (yhat-y)^2, stats[1] += sum |yhat-y|
// `wmap` (optional, hw elements): latitude weights of DLWP.custom.latitude_weighted_loss (custom.py:956-991), which scales
// BOTH tensors before the base loss: loss = mean((w*yhat - w*y)^2), dL/dyhat = scale * w^2 * (yhat - y); MAE stays unweighted.
__global__ void __launch_bounds__(256) mse_grad_kernel(const float* yhat, const float* y, float* g, long long n,
                                                       float scale, float* stats, const float* wmap, long long hw) {
    float s2 = 0.f, s1 = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = yhat[i] - y[i];
        const float w = wmap ? wmap[i % hw] : 1.f;
        const float w2 = w * w;
        if (g) g[i] += scale * w2 * d;
        s2 += w2 * d * d;
        s1 += fabsf(d);
    }
    __shared__ float r2[8], r1[8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        s2 += __shfl_xor_sync(0xffffffffu, s2, s);
        s1 += __shfl_xor_sync(0xffffffffu, s1, s);
    }
    if ((threadIdx.x & 31) == 0) { r2[threadIdx.x >> 5] = s2; r1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int wq = 0; wq < 8; ++wq) { a += r2[wq]; b += r1[wq]; }
        atomicAdd(stats, a);
        atomicAdd(stats + 1, b);
    }
}

// DLWP.custom.anomaly_correlation_loss (custom.py:1036-1088).  Over the WHOLE batch tensor of one output, with the optional
// climatology mu (one sample's worth, broadcast over the batch):
//   a = sum (p-mu)(y-mu) / sqrt(sum (p-mu)^2 * sum (y-mu)^2),   loss = s * (m - a),
// m = mean (p-y)^2 ('mse'), mean |p-y| ('mae') or 0 (regularize_mean=None), s = +1 (reverse) / -1.  Two passes: the five
// sums (double atomics: a is a ratio of sums of ~1e7 terms), then
//   dloss/dp_i = s * ( dm/dp_i - [ (y_i-mu_i) / sqrt(Spp Syy) - a (p_i-mu_i) / Spp ] ).
// stats: Spy, Spp, Syy, sum (p-y)^2, sum |p-y|
__global__ void __launch_bounds__(256) acc_stats_kernel(const float* yhat, const float* y, const float* mean, long long per,
                                                        long long n, double* stats) {
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float mu = mean ? mean[i % per] : 0.f;
        const float p = yhat[i] - mu, t = y[i] - mu, d = yhat[i] - y[i];
        v[0] += p * t; v[1] += p * p; v[2] += t * t; v[3] += d * d; v[4] += fabsf(d);
    }
    __shared__ float red[5][8];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], s);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double a = 0.0;
        for (int wq = 0; wq < 8; ++wq) a += (double)red[threadIdx.x][wq];
        atomicAdd(stats + threadIdx.x, a);
    }
}

__global__ void __launch_bounds__(256) acc_grad_kernel(const float* yhat, const float* y, const float* mean, long long per,
                                                       float* g, long long n, const double* stats, float scale,
                                                       int regularize) {
    const double spp = stats[1], syy = stats[2];
    const float inv = (float)(1.0 / sqrt(spp * syy));
    const float a_over_spp = (float)(stats[0] / sqrt(spp * syy) / spp);
    const float inv_n = 1.f / (float)n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float mu = mean ? mean[i % per] : 0.f;
        const float p = yhat[i] - mu, t = y[i] - mu, d = yhat[i] - y[i];
        float dm = 0.f;
        if (regularize == 1) dm = 2.f * d * inv_n;
        else if (regularize == 2) dm = (d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f));
        g[i] += scale * (dm - (t * inv - a_over_spp * p));
    }
}

// Keras Adam (no amsgrad): lr_t = lr * sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t m/(sqrt v + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* w, const float* g, float* m, float* v, long long n, float lr_t,
                                                   float b1, float b2, float eps) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        w[i] -= lr_t * mi / (sqrtf(vi) + eps);
    }
}

// ---- launchers ----------------------------------------------------------------------------------------------------------
static inline int blocks_for(long long total) { return (int)std::min<long long>((total + 255) / 256, 148LL * 32); }

static BwdParams bwd_params(const DlwpConvDesc& d) {
    BwdParams p;
    memset(&p, 0, sizeof(p));
    p.N = d.N; p.Cin = d.Cin; p.H = d.H; p.W = d.W; p.Cout = d.Cout;
    p.Ho = d.H + d.pad_t + d.pad_b - d.dil_h * (d.kh - 1);
    p.Wo = d.W + d.pad_l + d.pad_r - d.dil_w * (d.kw - 1);
    p.kh = d.kh; p.kw = d.kw; p.dh = d.dil_h; p.dw_ = d.dil_w; p.pad_t = d.pad_t; p.pad_l = d.pad_l;
    p.mode_h = d.pad_mode_h; p.mode_w = d.pad_mode_w;
    p.xs_n = d.x_stride_n; p.xs_c = d.x_stride_c; p.xs_h = d.x_stride_h;
    p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    return p;
}

int conv2d_bwd_input(const DlwpConvDesc& d, const float* dy, const float* w, float* dx, cudaStream_t stream) {
    DLWP_REQUIRE(!d.rowwise && !d.pre_op, DLWP_ESHAPE, "backward is not implemented for row-connected / pre_op layers");
    BwdParams p = bwd_params(d);
    p.dy = dy; p.w = w; p.dx = dx;
    conv_bwd_input_kernel<<<blocks_for((long long)d.N * d.Cin * d.H * d.W), 256, 0, stream>>>(p);
    return after_launch("conv_bwd_input_kernel");
}

// dX through the FORWARD kernels: dX = conv(dY, W') with W' the flipped, channel-transposed weights and the paddings mirrored
// (adjoint of PeriodicPadding2D = periodic, of ZeroPadding2D + 'valid' = zero padding of dY).  OVERWRITES dx (dense
// (N, Cin, H, W) with the strides of d.x_stride_*); wt: scratch for kh*kw*Cin*Cout floats.  Returns -1 when the layer is not
// a stride-1 'same' convolution (then the caller keeps conv2d_bwd_input).
int conv2d_bwd_input_fwd(const DlwpConvDesc& d, const float* dy, const float* w, float* dx, float* wt, cudaStream_t stream) {
    if (d.rowwise || d.pre_op) return -1;
    if (d.pad_t + d.pad_b != d.dil_h * (d.kh - 1) || d.pad_l + d.pad_r != d.dil_w * (d.kw - 1)) return -1;
    const int total = d.kh * d.kw * d.Cin * d.Cout;
    flip_transpose_kernel<<<std::min((total + 255) / 256, 148 * 4), 256, 0, stream>>>(w, wt, d.kh, d.kw, d.Cin, d.Cout);
    int rc = after_launch("flip_transpose_kernel");
    if (rc) return rc;
    DlwpConvDesc b = d;
    b.Cin = d.Cout; b.Cout = d.Cin;
    b.pad_t = d.pad_b; b.pad_b = d.pad_t; b.pad_l = d.pad_r; b.pad_r = d.pad_l;
    b.act = DLWP_ACT_LINEAR;
    b.impl = DLWP_IMPL_AUTO;
    b.row_begin = b.row_end = 0;
    b.x_stride_n = d.y_stride_n; b.x_stride_c = d.y_stride_c; b.x_stride_h = d.y_stride_h;
    b.y_stride_n = d.x_stride_n; b.y_stride_c = d.x_stride_c; b.y_stride_h = d.x_stride_h;
    return conv2d_fwd(b, dy, wt, nullptr, dx, stream);
}

int conv2d_bwd_weight(const DlwpConvDesc& d, const float* x, const float* dy, float* dw, float* db, cudaStream_t stream) {
    DLWP_REQUIRE(!d.rowwise && !d.pre_op, DLWP_ESHAPE, "backward is not implemented for row-connected / pre_op layers");
    BwdParams p = bwd_params(d);
    p.x = x; p.dy = dy; p.dw = dw; p.db = db;
    int rc = -1;
    if (d.kh == d.kw && d.dil_h == d.dil_w) {          // the tiled kernel's instances: the example nets' layer shapes
        if (d.kh == 3 && d.dil_h == 1) rc = launch_bwd_weight_tiled<3, 1, 4>(p, stream);
        else if (d.kh == 3 && d.dil_h == 2) rc = launch_bwd_weight_tiled<3, 2, 4>(p, stream);
        else if (d.kh == 5 && d.dil_h == 1) rc = launch_bwd_weight_tiled<5, 1, 2>(p, stream);
    }
    if (rc == -1) {
        const int nchunks = (d.Cout + BW_OCH - 1) / BW_OCH;
        const int nb = d.kh * d.kw * d.Cin * nchunks;
        p.nsplit = std::max(1, std::min(d.N, (148 * 4 + nb - 1) / nb));
        conv_bwd_weight_kernel<<<dim3(nb, p.nsplit), 256, 0, stream>>>(p);
        rc = after_launch("conv_bwd_weight_kernel");
    } else {
        return rc;      // the tiled kernel sums dB itself
    }
    if (rc || !db) return rc;
    bias_grad_kernel<<<dim3(d.Cout, std::min(d.N, 16)), 256, 0, stream>>>(dy, db, d.N, d.Cout, p.Ho, p.Wo, p.ys_n, p.ys_c,
                                                                          p.ys_h);
    return after_launch("bias_grad_kernel");
}

static EwBwd ew(const float* a, const float* g, float* out, int N, int C, int H, int W, const long long* as,
                const long long* gs, const long long* os, int act) {
    EwBwd p;
    p.a = a; p.g = g; p.out = out; p.N = N; p.C = C; p.H = H; p.W = W; p.act = act;
    p.as_n = as[0]; p.as_c = as[1]; p.as_h = as[2];
    p.gs_n = gs[0]; p.gs_c = gs[1]; p.gs_h = gs[2];
    p.os_n = os[0]; p.os_c = os[1]; p.os_h = os[2];
    return p;
}

int act_bwd(const float* y, float* g, int act, int N, int C, int H, int W, const long long* ys, const long long* gs,
            cudaStream_t stream) {
    if (act == DLWP_ACT_LINEAR) return 0;
    EwBwd p = ew(y, nullptr, g, N, C, H, W, ys, gs, gs, act);
    ew_bwd_kernel<BW_ACT><<<blocks_for((long long)N * C * H * W), 256, 0, stream>>>(p);
    return after_launch("act_bwd");
}
int maxpool_bwd(const float* x, const float* g, float* dx, int N, int C, int Hp, int Wp, const long long* xs,
                const long long* gs, const long long* dxs, cudaStream_t stream) {
    EwBwd p = ew(x, g, dx, N, C, Hp, Wp, xs, gs, dxs, 0);
    ew_bwd_kernel<BW_POOL><<<blocks_for((long long)N * C * Hp * Wp), 256, 0, stream>>>(p);
    return after_launch("maxpool_bwd");
}
int upsample_bwd(const float* g, float* dx, int N, int C, int Hl, int Wl, const long long* gs, const long long* dxs,
                 cudaStream_t stream) {
    EwBwd p = ew(nullptr, g, dx, N, C, Hl, Wl, gs, gs, dxs, 0);
    ew_bwd_kernel<BW_UP><<<blocks_for((long long)N * C * Hl * Wl), 256, 0, stream>>>(p);
    return after_launch("upsample_bwd");
}
int add_bwd(const float* g, float* dx, int N, int C, int H, int W, const long long* gs, const long long* dxs,
            cudaStream_t stream) {
    EwBwd p = ew(nullptr, g, dx, N, C, H, W, gs, gs, dxs, 0);
    ew_bwd_kernel<BW_ADD><<<blocks_for((long long)N * C * H * W), 256, 0, stream>>>(p);
    return after_launch("add_bwd");
}
int acc_loss_grad(const float* yhat, const float* y, const float* mean, long long per, float* g, long long n, double* stats,
                  float scale, int regularize, cudaStream_t stream) {
    acc_stats_kernel<<<blocks_for(n), 256, 0, stream>>>(yhat, y, mean, per, n, stats);
    int rc = after_launch("acc_stats_kernel");
    if (rc || !g) return rc;
    acc_grad_kernel<<<blocks_for(n), 256, 0, stream>>>(yhat, y, mean, per, g, n, stats, scale, regularize);
    return after_launch("acc_grad_kernel");
}

int mse_grad(const float* yhat, const float* y, float* g, long long n, float scale, float* stats, cudaStream_t stream,
             const float* wmap, long long hw) {
    mse_grad_kernel<<<blocks_for(n), 256, 0, stream>>>(yhat, y, g, n, scale, stats, wmap, hw);
    return after_launch("mse_grad_kernel");
}
// keras regularizers.L1L2 on one weight tensor: g += l1 * sign(w) + 2 * l2 * w; penalty = sum(l1 |w| + l2 w^2) -> *stat
__global__ void __launch_bounds__(256) regularize_kernel(const float* __restrict__ w, float* __restrict__ g, long long n,
                                                         float l1, float l2, float* stat) {
    float pen = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = w[i];
        g[i] += l1 * (v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f)) + 2.f * l2 * v;
        pen += l1 * fabsf(v) + l2 * v * v;
    }
    for (int o = 16; o > 0; o >>= 1) pen += __shfl_down_sync(0xffffffffu, pen, o);
    if ((threadIdx.x & 31) == 0 && pen != 0.f) atomicAdd(stat, pen);
}
int regularize_grad(const float* w, float* g, long long n, float l1, float l2, float* stat, cudaStream_t stream) {
    regularize_kernel<<<blocks_for(n), 256, 0, stream>>>(w, g, n, l1, l2, stat);
    return after_launch("regularize_kernel");
}
int adam_step(float* w, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
              cudaStream_t stream) {
    adam_kernel<<<blocks_for(n), 256, 0, stream>>>(w, g, m, v, n, lr_t, b1, b2, eps);
    return after_launch("adam_kernel");
}

}  // namespace dlwp
