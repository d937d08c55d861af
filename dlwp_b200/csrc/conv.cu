// Fused [PeriodicPadding2D / ZeroPadding2D] -> Conv2D('valid', dilation) -> bias -> activation, fp32, channels_first.
//
// Replaces, per layer, the three Keras-internal steps the reference runs inside keras.Model.predict
// (DLWP/model/models.py:241, :412): PeriodicPadding2D.call's two concatenations (DLWP/custom.py:191-214), ZeroPadding2D,
// and Conv2D.  No padded copy is ever materialised: the wrap in longitude and the zero rows beyond the poles are
// resolved while the input tile is staged into shared memory.
//
// Kernels
//   conv_direct_kernel   one thread per output element; any kernel size / dilation / pad mode, RowConnected2D,
//                        pool-on-load and upsample-on-load.  The reference CUDA implementation and the fallback.
//   conv_ffma_kernel     register-tiled FFMA kernel.  A CTA owns TH output rows x TW columns x (NCG*COUT_T) filters of
//                        one sample; each thread owns ROWS x 4 pixels x COUT_T filters in registers.  The input tile
//                        (with halo) is streamed through shared memory in channel chunks, double buffered:
//                          STAGE_CPASYNC  elementwise cp.async with wrap / zero-fill index arithmetic (any geometry)
//                          STAGE_TMA      one cp.async.bulk.tensor box per chunk; rows beyond the poles arrive as TMA
//                                         out-of-bounds zeros, the longitude wrap is patched into the halo columns in
//                                         shared memory from values prefetched while the TMA is in flight.
#define DLWP_CONV_TU
#include <cuda_runtime.h>
namespace dlwp {
__device__ int g_device_flags = 0;  // bit 0: an mbarrier wait timed out (TMA never completed)
}
#include "internal.h"
#include "conv_tc.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace dlwp {

// =================================================================================================================
// Direct kernel
// =================================================================================================================
struct DirectParams {
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int N, Cin, Hs, Ws;  // source dims as stored
    int H, W;            // logical input dims after pre_op
    int Cout, Ho, Wo, kh, kw, dh, dw, pad_t, pad_l, mode_h, mode_w, act, pre_op, rowwise;
    int row0, rows;      // output row window
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;
};

__device__ __forceinline__ float load_logical(const DirectParams& p, const float* xc, int gy, int gx) {
    // (gy, gx) are valid logical coordinates
    if (p.pre_op == 1) {  // MaxPooling2D(2): logical (gy,gx) = max of the 2x2 source block
        const float* q = xc + (long long)(2 * gy) * p.xs_h + 2 * gx;
        return fmaxf(fmaxf(q[0], q[1]), fmaxf(q[p.xs_h], q[p.xs_h + 1]));
    }
    if (p.pre_op == 2) return xc[(long long)(gy >> 1) * p.xs_h + (gx >> 1)];  // UpSampling2D(2), nearest
    return xc[(long long)gy * p.xs_h + gx];
}

__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectParams p) {
    const long long total = (long long)p.N * p.Cout * p.rows * p.Wo;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xo = (int)(idx % p.Wo);
        long long t = idx / p.Wo;
        const int yo = p.row0 + (int)(t % p.rows);
        t /= p.rows;
        const int o = (int)(t % p.Cout);
        const int n = (int)(t / p.Cout);
        const float* w = p.w;
        float acc = 0.f;
        if (p.rowwise) {
            w += (long long)yo * p.kh * p.kw * p.Cin * p.Cout;
            if (p.bias) acc = p.bias[(long long)yo * p.Cout + o];
        } else if (p.bias) {
            acc = p.bias[o];
        }
        for (int i = 0; i < p.kh; ++i) {
            int gy = yo + p.dh * i - p.pad_t;
            if (p.mode_h == DLWP_PAD_PERIODIC) gy = wrap_index(gy, p.H);
            else if (gy < 0 || gy >= p.H) continue;
            for (int j = 0; j < p.kw; ++j) {
                int gx = xo + p.dw * j - p.pad_l;
                if (p.mode_w == DLWP_PAD_PERIODIC) gx = wrap_index(gx, p.W);
                else if (gx < 0 || gx >= p.W) continue;
                const float* wt = w + ((long long)(i * p.kw + j) * p.Cin) * p.Cout + o;
                const float* xb = p.x + (long long)n * p.xs_n;
                for (int c = 0; c < p.Cin; ++c)
                    acc = fmaf(load_logical(p, xb + (long long)c * p.xs_c, gy, gx), wt[(long long)c * p.Cout], acc);
            }
        }
        p.y[(long long)n * p.ys_n + (long long)o * p.ys_c + (long long)yo * p.ys_h + xo] = apply_act(acc, p.act);
    }
}

// =================================================================================================================
// RowConnected2D (DLWP/custom.py:695-896: a conv whose weights differ per output row) as a tiled kernel.
// The reference runs H_out separate K.conv2d calls on row slabs (custom.py:879-893).  Here one CTA owns ONE output row of a
// chunk of samples: that row's whole weight set (kh*kw*Cin*Cout floats: 38 KB for the 5x5, 32 -> 12 layer of
// examples/train_functional.py:191-196) is staged in shared memory once and every thread computes 4 adjacent pixels x all
// filters in registers -- per (channel, tap row) 4 + dw*(kw-1) input values feed kw*4*Cout FMAs, weights arrive as
// broadcast 16-byte shared-memory loads (~10 FMAs per memory instruction; the direct kernel does 1).
// =================================================================================================================
struct RowwiseParams {
    const float* x;
    const float* w;      // (H_out, kh, kw, Cin, Cout)
    const float* bias;   // (H_out, Cout) or null
    float* y;
    int N, Cin, H, W, Cout, Wo, pad_t, pad_l, dh, mode_h, mode_w, act, row0, xgroups;
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;
};

template <int KH, int KW, int DW, int COUT_T>
__global__ void __launch_bounds__(256) conv_rowwise_kernel(const RowwiseParams p) {
    extern __shared__ __align__(16) float ws[];          // [i][j][c][COUT_T]
    constexpr int NV = 4 + DW * (KW - 1);
    const int yo = p.row0 + blockIdx.x;
    const float* wrow = p.w + (long long)yo * KH * KW * p.Cin * p.Cout;
    for (int idx = threadIdx.x; idx < KH * KW * p.Cin * COUT_T; idx += blockDim.x) {
        const int o = idx % COUT_T, r = idx / COUT_T;     // r = (i*KW + j)*Cin + c
        ws[idx] = o < p.Cout ? wrow[(long long)r * p.Cout + o] : 0.f;
    }
    __syncthreads();
    const long long items = (long long)p.N * p.xgroups;
    for (long long it = blockIdx.y * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.y * blockDim.x) {
        const int xg = (int)(it % p.xgroups), n = (int)(it / p.xgroups);
        const int x0 = 4 * xg;
        float acc[4][COUT_T];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int o = 0; o < COUT_T; ++o) acc[px][o] = (p.bias && o < p.Cout) ? p.bias[(long long)yo * p.Cout + o] : 0.f;
        int gxs[NV];
        bool okx[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            int gx = x0 + t - p.pad_l;
            bool ok = true;
            if (p.mode_w == DLWP_PAD_PERIODIC) gx = wrap_index(gx, p.W);
            else if (gx < 0 || gx >= p.W) { ok = false; gx = 0; }
            gxs[t] = gx; okx[t] = ok;
        }
        const float* xn = p.x + (long long)n * p.xs_n;
#pragma unroll 1
        for (int c = 0; c < p.Cin; ++c) {
#pragma unroll
            for (int i = 0; i < KH; ++i) {
                int gy = yo + p.dh * i - p.pad_t;
                if (p.mode_h == DLWP_PAD_PERIODIC) gy = wrap_index(gy, p.H);
                else if (gy < 0 || gy >= p.H) continue;
                const float* xr = xn + (long long)c * p.xs_c + (long long)gy * p.xs_h;
                float v[NV];
#pragma unroll
                for (int t = 0; t < NV; ++t) v[t] = okx[t] ? __ldg(xr + gxs[t]) : 0.f;
#pragma unroll
                for (int j = 0; j < KW; ++j) {
                    const float4* w4 = reinterpret_cast<const float4*>(ws + ((long long)(i * KW + j) * p.Cin + c) * COUT_T);
#pragma unroll
                    for (int o4 = 0; o4 < COUT_T / 4; ++o4) {
                        const float4 wv = w4[o4];
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            const float xv = v[px + DW * j];
                            acc[px][4 * o4 + 0] = fmaf(xv, wv.x, acc[px][4 * o4 + 0]);
                            acc[px][4 * o4 + 1] = fmaf(xv, wv.y, acc[px][4 * o4 + 1]);
                            acc[px][4 * o4 + 2] = fmaf(xv, wv.z, acc[px][4 * o4 + 2]);
                            acc[px][4 * o4 + 3] = fmaf(xv, wv.w, acc[px][4 * o4 + 3]);
                        }
                    }
                }
            }
        }
        float* yn = p.y + (long long)n * p.ys_n + (long long)yo * p.ys_h;
#pragma unroll
        for (int o = 0; o < COUT_T; ++o)
            if (o < p.Cout) {
#pragma unroll
                for (int px = 0; px < 4; ++px)
                    if (x0 + px < p.Wo) yn[(long long)o * p.ys_c + x0 + px] = apply_act(acc[px][o], p.act);
            }
    }
}

template <int KH, int KW, int DW, int COUT_T>
static int launch_rowwise(const RowwiseParams& p, int rows, size_t smem, cudaStream_t stream) {
    static std::atomic<unsigned long long> done{0};
    if (first_use_on_device(done))
        cudaFuncSetAttribute(conv_rowwise_kernel<KH, KW, DW, COUT_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const long long items = (long long)p.N * p.xgroups;
    const int by = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, 64));
    conv_rowwise_kernel<KH, KW, DW, COUT_T><<<dim3(rows, by), 256, smem, stream>>>(p);
    return after_launch("conv_rowwise_kernel");
}

// 0: launched; -1: no instance for this geometry (the caller falls back to the direct kernel)
static int conv2d_rowwise(const DlwpConvDesc& d, const float* x, const float* w, const float* bias, float* y, int Wo,
                          int row0, int rows, cudaStream_t stream, int* rc_out) {
    if (d.pre_op || d.kh != d.kw || d.Cout > 16) return -1;
    const int cout_t = d.Cout <= 4 ? 4 : (d.Cout <= 8 ? 8 : (d.Cout <= 12 ? 12 : 16));
    const size_t smem = (size_t)d.kh * d.kw * d.Cin * cout_t * sizeof(float);
    if (smem > 200 * 1024) return -1;
    RowwiseParams p;
    p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.N = d.N; p.Cin = d.Cin; p.H = d.H; p.W = d.W; p.Cout = d.Cout; p.Wo = Wo; p.pad_t = d.pad_t; p.pad_l = d.pad_l;
    p.dh = d.dil_h; p.mode_h = d.pad_mode_h; p.mode_w = d.pad_mode_w; p.act = d.act; p.row0 = row0;
    p.xgroups = (Wo + 3) / 4;
    p.xs_n = d.x_stride_n; p.xs_c = d.x_stride_c; p.xs_h = d.x_stride_h;
    p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
#define DLWP_ROWWISE_CASE(K_, D_, C_) \
    if (d.kh == K_ && d.dil_w == D_ && cout_t == C_) { *rc_out = launch_rowwise<K_, K_, D_, C_>(p, rows, smem, stream); return 0; }
    DLWP_ROWWISE_CASE(5, 1, 12) DLWP_ROWWISE_CASE(5, 1, 8) DLWP_ROWWISE_CASE(5, 1, 16) DLWP_ROWWISE_CASE(5, 1, 4)
    DLWP_ROWWISE_CASE(3, 1, 12) DLWP_ROWWISE_CASE(3, 1, 8) DLWP_ROWWISE_CASE(3, 1, 16) DLWP_ROWWISE_CASE(3, 1, 4)
    DLWP_ROWWISE_CASE(3, 2, 12) DLWP_ROWWISE_CASE(3, 2, 8) DLWP_ROWWISE_CASE(3, 2, 16) DLWP_ROWWISE_CASE(3, 2, 4)
#undef DLWP_ROWWISE_CASE
    return -1;
}

// =================================================================================================================
// Register-tiled FFMA kernel
// =================================================================================================================
struct TileParams {
    const float* x;
    const float* w;
    const float* bias;
    float* y;
    int N, Cin, H, W, Cout, Ho, Wo;
    int row0, row1;    // output rows [row0, row1) computed by this launch (latitude band)
    int pad_t, pad_l, mode_h, mode_w, act;
    long long xs_n, xs_c, xs_h, ys_n, ys_c, ys_h;
    int TH, TW;        // output tile (TW multiple of 4, TH multiple of ROWS)
    int RIN, PITCH;    // staged input tile: RIN rows x PITCH floats (PITCH multiple of 4)
    int CC, nchunks;   // channels per chunk, number of chunks
    int NXG, NRG, NCG; // thread grid inside the CTA: x groups of 4 px, row groups of ROWS rows, filter groups
    int COUT_BP;       // padded filters per CTA in the smem weight tile (= NCG * roundup4(COUT_T))
    int tiles_x, tiles_y;
    int in_stage_floats, w_stage_floats;  // per stage, multiples of 32 floats (128 B)
    int fix_l, fix_r;  // STAGE_TMA: number of halo columns patched on the left / right edge tiles
};

constexpr int STAGE_CPASYNC = 0;
constexpr int STAGE_TMA = 1;
constexpr int MAX_FIX_PER_THREAD = 8;

// ---- packed fp32 math (sm_100: FFMA2) -----------------------------------------------------------------------------------
// fma.rn.f32x2 does two IEEE fp32 FMAs (same rounding as two FFMAs) on 64-bit register pairs.  It halves the FMA
// instruction count and the register-file reads per flop, which is what lets a register-tiled loop approach the fp32
// pipe's peak: with scalar FFMA the same loop is limited by operand-bank conflicts (measured ~30-50% of peak).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// LEAD: unused columns in front of the halo (STAGE_TMA only): the TMA box must start on a 16-byte boundary in global
// memory, so a tile whose halo is pad_l columns wide starts (4 - pad_l % 4) % 4 columns earlier.
//
// acc[r][px][cp] holds filters (2cp, 2cp+1) of output pixel (row r, column px) as one packed pair: the weight pairs come
// straight out of shared memory (16-byte loads of 4 consecutive filters), the input value is duplicated into both lanes.
template <int KH, int KW, int D, int ROWS, int COUT_T, int LEAD>
__device__ __forceinline__ void ffma_chunk(const float* __restrict__ in_s, const float* __restrict__ w_s,
                                           f32x2 (&acc)[ROWS][4][COUT_T / 2], const int cc, const int plane,
                                           const int pitch, const int w_plane, const int cout_bp) {
    static_assert(COUT_T % 2 == 0, "filters are processed in packed pairs");
    constexpr int COUT_LD = (COUT_T + 3) / 4 * 4;
    constexpr int NV = (LEAD + (KW - 1) * D + 4 + 3) / 4 * 4;  // input floats a thread reads per row
#pragma unroll 1
    for (int c = 0; c < cc; ++c) {
        const float* in_c = in_s + c * plane;
        const float* w_c = w_s + c * w_plane;
#pragma unroll
        for (int i = 0; i < KH; ++i) {
            float v[ROWS][NV];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float4* q = reinterpret_cast<const float4*>(in_c + (r + i * D) * pitch);
#pragma unroll
                for (int q4 = 0; q4 < NV / 4; ++q4) {
                    const float4 a = q[q4];
                    v[r][4 * q4 + 0] = a.x; v[r][4 * q4 + 1] = a.y; v[r][4 * q4 + 2] = a.z; v[r][4 * q4 + 3] = a.w;
                }
            }
#pragma unroll
            for (int j = 0; j < KW; ++j) {
                f32x2 wv[COUT_LD / 2];
                const ulonglong2* wq = reinterpret_cast<const ulonglong2*>(w_c + (i * KW + j) * cout_bp);
#pragma unroll
                for (int q4 = 0; q4 < COUT_LD / 4; ++q4) {
                    const ulonglong2 t = wq[q4];
                    wv[2 * q4 + 0] = t.x; wv[2 * q4 + 1] = t.y;
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float x = v[r][LEAD + px + j * D];
                        const f32x2 xx = pack2(x, x);
#pragma unroll
                        for (int cp = 0; cp < COUT_T / 2; ++cp) acc[r][px][cp] = ffma2(xx, wv[cp], acc[r][px][cp]);
                    }
            }
        }
    }
}

template <int KH, int KW, int D, int ROWS, int COUT_T, int STAGE>
__global__ void __launch_bounds__(384) conv_ffma_kernel(const TileParams p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int COUT_LD = (COUT_T + 3) / 4 * 4;
    constexpr int NAT_PAD = D * (KW - 1) / 2;                                  // the 'same'-style halo of this kernel
    constexpr int LEAD = (STAGE == STAGE_TMA) ? (4 - NAT_PAD % 4) % 4 : 0;     // host guarantees pad_l == NAT_PAD
    extern __shared__ __align__(128) float smem[];
    auto in_stage = [&](int s) { return smem + s * p.in_stage_floats; };
    auto w_stage = [&](int s) { return smem + 2 * p.in_stage_floats + s * p.w_stage_floats; };
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 2 * p.in_stage_floats + 2 * p.w_stage_floats);

    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int tile = blockIdx.x;
    const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
    const int x0 = tx * p.TW, y0 = p.row0 + ty * p.TH;
    const int cout0 = blockIdx.y * (p.NCG * COUT_T);
    const int n = blockIdx.z;

    // thread -> (filter group, row group, x group)
    const int xg = tid % p.NXG;
    const int rg = (tid / p.NXG) % p.NRG;
    const int cg = tid / (p.NXG * p.NRG);
    const bool active = cg < p.NCG;

    const int plane = p.RIN * p.PITCH;
    const int w_plane = KH * KW * p.COUT_BP;
    const float* xn = p.x + (long long)n * p.xs_n;

    // ---- producers ---------------------------------------------------------------------------------------------
    auto stage_weights = [&](int chunk, float* dst) {
        // smem layout [c][i][j][COUT_BP]; Keras layout in global: w[((i*KW + j)*Cin + c)*Cout + o]
        const int c0 = chunk * p.CC;
        const int total = p.CC * w_plane;
        for (int idx = tid; idx < total; idx += nthreads) {
            const int ob = idx % p.COUT_BP;
            const int tap = (idx / p.COUT_BP) % (KH * KW);
            const int c = idx / w_plane;
            const int g = ob / COUT_LD, ol = ob % COUT_LD;
            const int o = cout0 + g * COUT_T + ol;
            const bool valid = (ol < COUT_T) && (o < p.Cout) && (c0 + c < p.Cin);
            const float* src = valid ? p.w + ((long long)tap * p.Cin + (c0 + c)) * p.Cout + o : p.w;
            cp_async_4(smem_u32(dst + idx), src, valid);
        }
    };
    auto stage_input_cpasync = [&](int chunk, float* dst) {
        const int c0 = chunk * p.CC;
        const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
        for (int row = warp; row < p.CC * p.RIN; row += nwarps) {
            const int c = row / p.RIN, r = row % p.RIN;
            int gy = y0 + r - p.pad_t;
            bool vrow = (c0 + c) < p.Cin;
            if (p.mode_h == DLWP_PAD_PERIODIC) gy = wrap_index(gy, p.H);
            else vrow = vrow && gy >= 0 && gy < p.H;
            const float* src_row = xn + (long long)(c0 + c) * p.xs_c + (long long)gy * p.xs_h;
            float* dst_row = dst + row * p.PITCH;
            for (int k = lane; k < p.PITCH; k += 32) {
                int gx = x0 + k - p.pad_l;
                bool v = vrow;
                if (p.mode_w == DLWP_PAD_PERIODIC) gx = wrap_index(gx, p.W);
                else v = v && gx >= 0 && gx < p.W;
                cp_async_4(smem_u32(dst_row + k), v ? src_row + gx : p.x, v);
            }
        }
    };

    // ---- STAGE_TMA halo patch bookkeeping ------------------------------------------------------------------------
    // Columns whose source longitude lies outside [0, W) were zero-filled by the TMA box; with periodic padding they
    // are overwritten with the wrapped value.  nfix elements per chunk, spread over the threads, prefetched into
    // registers while the TMA is in flight.
    const int fix_l = (STAGE == STAGE_TMA && p.mode_w == DLWP_PAD_PERIODIC && tx == 0) ? p.fix_l : 0;
    const int fix_r = (STAGE == STAGE_TMA && p.mode_w == DLWP_PAD_PERIODIC && tx == p.tiles_x - 1) ? p.fix_r : 0;
    const int fix_cols = fix_l + fix_r;
    const int nfix = fix_cols * p.CC * p.RIN;
    float hv[MAX_FIX_PER_THREAD];
    const bool fix_prefetched = nfix <= MAX_FIX_PER_THREAD * nthreads;
    auto fix_decode = [&](int e, int& c, int& r, int& k) {
        const int fc = fix_cols > 0 ? fix_cols : 1;
        const int q = e % fc;
        const int row = e / fc;
        c = row / p.RIN;
        r = row % p.RIN;
        // left patch columns are k = LEAD .. LEAD+fix_l-1; right patch columns start where the source x reaches W
        k = LEAD + ((q < fix_l) ? q : (p.W + p.pad_l - x0) + (q - fix_l));
    };
    auto prefetch_fix = [&](int chunk) {
        if constexpr (STAGE == STAGE_TMA) {
        if (!fix_prefetched) return;
        const int c0 = chunk * p.CC;
#pragma unroll
        for (int u = 0; u < MAX_FIX_PER_THREAD; ++u) {
            const int e = tid + u * nthreads;
            hv[u] = 0.f;
            if (e < nfix) {
                int c, r, k;
                fix_decode(e, c, r, k);
                const int gy = y0 + r - p.pad_t;
                const int gx = wrap_index(x0 + k - LEAD - p.pad_l, p.W);
                if (gy >= 0 && gy < p.H && c0 + c < p.Cin)
                    hv[u] = __ldg(xn + (long long)(c0 + c) * p.xs_c + (long long)gy * p.xs_h + gx);
            }
        }
        }
    };
    auto apply_fix = [&](float* dst, int chunk) {
        if constexpr (STAGE == STAGE_TMA) {
            if (fix_prefetched) {
#pragma unroll
                for (int u = 0; u < MAX_FIX_PER_THREAD; ++u) {
                    const int e = tid + u * nthreads;
                    if (e < nfix) {
                        int c, r, k;
                        fix_decode(e, c, r, k);
                        dst[(c * p.RIN + r) * p.PITCH + k] = hv[u];
                    }
                }
            } else {  // tiny CTAs (few threads, many halo elements): patch straight from global after the TMA landed
                const int c0 = chunk * p.CC;
                for (int e = tid; e < nfix; e += nthreads) {
                    int c, r, k;
                    fix_decode(e, c, r, k);
                    const int gy = y0 + r - p.pad_t;
                    const int gx = wrap_index(x0 + k - LEAD - p.pad_l, p.W);
                    float val = 0.f;
                    if (gy >= 0 && gy < p.H && c0 + c < p.Cin)
                        val = __ldg(xn + (long long)(c0 + c) * p.xs_c + (long long)gy * p.xs_h + gx);
                    dst[(c * p.RIN + r) * p.PITCH + k] = val;
                }
            }
        }
    };
    auto issue_tma = [&](int chunk, int s) {
        if (STAGE == STAGE_TMA && tid == 0) {
            mbar_expect_tx(&mbar[s], (uint32_t)(p.CC * plane * sizeof(float)));
            tma_load_4d(in_stage(s), &tmap, &mbar[s], x0 - p.pad_l - LEAD, y0 - p.pad_t, chunk * p.CC, n);
        }
    };

    if (STAGE == STAGE_TMA) {
        if (tid == 0) {
            prefetch_tensormap(&tmap);
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }

    f32x2 acc[ROWS][4][COUT_T / 2];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int cp = 0; cp < COUT_T / 2; ++cp) acc[r][px][cp] = 0ull;

    // ---- pipeline: chunk k+1 is in flight while chunk k is consumed ------------------------------------------------
    if (STAGE == STAGE_TMA) issue_tma(0, 0);
    else stage_input_cpasync(0, in_stage(0));
    stage_weights(0, w_stage(0));
    cp_async_commit();

    for (int k = 0; k < p.nchunks; ++k) {
        const int s = k & 1;
        if (k + 1 < p.nchunks) {
            if (STAGE == STAGE_TMA) issue_tma(k + 1, s ^ 1);
            else stage_input_cpasync(k + 1, in_stage(s ^ 1));
            stage_weights(k + 1, w_stage(s ^ 1));
            cp_async_commit();
        }
        if (STAGE == STAGE_TMA) {
            prefetch_fix(k);
            mbar_wait(&mbar[s], (k >> 1) & 1);
            apply_fix(in_stage(s), k);
        }
        if (k + 1 < p.nchunks) cp_async_wait<1>();
        else cp_async_wait<0>();
        __syncthreads();

        if (active) {
            const int cc = min(p.CC, p.Cin - k * p.CC);
            ffma_chunk<KH, KW, D, ROWS, COUT_T, LEAD>(in_stage(s) + (rg * ROWS) * p.PITCH + 4 * xg,
                                                w_stage(s) + cg * COUT_LD, acc, cc, plane, p.PITCH, w_plane,
                                                p.COUT_BP);
        }
        if (STAGE == STAGE_TMA) fence_proxy_async();  // generic writes (halo patch) before the next TMA refill
        __syncthreads();
    }

    // ---- epilogue: bias + activation + vector store ----------------------------------------------------------------
    if (!active) return;
    const int xo = x0 + 4 * xg;
    if (xo >= p.Wo) return;
    const bool vec_ok = (xo + 3 < p.Wo) && ((p.ys_h & 3) == 0) && ((p.ys_c & 3) == 0) && ((p.ys_n & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int yo = y0 + rg * ROWS + r;
        if (yo >= p.row1) continue;
#pragma unroll
        for (int co = 0; co < COUT_T; ++co) {
            const int o = cout0 + cg * COUT_T + co;
            if (o >= p.Cout) continue;
            const float b = p.bias ? __ldg(p.bias + o) : 0.f;
            float* dst = p.y + (long long)n * p.ys_n + (long long)o * p.ys_c + (long long)yo * p.ys_h + xo;
            float a4[4];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                float lo, hi;
                unpack2(acc[r][px][co / 2], lo, hi);
                a4[px] = (co & 1) ? hi : lo;
            }
            float4 v;
            v.x = apply_act(a4[0] + b, p.act);
            v.y = apply_act(a4[1] + b, p.act);
            v.z = apply_act(a4[2] + b, p.act);
            v.w = apply_act(a4[3] + b, p.act);
            if (vec_ok) {
                *reinterpret_cast<float4*>(dst) = v;
            } else {
                dst[0] = v.x;
                if (xo + 1 < p.Wo) dst[1] = v.y;
                if (xo + 2 < p.Wo) dst[2] = v.z;
                if (xo + 3 < p.Wo) dst[3] = v.w;
            }
        }
    }
}

// =================================================================================================================
// Host side
// =================================================================================================================
typedef void (*ffma_kernel_t)(const TileParams, const CUtensorMap);

struct KernelEntry {
    int kh, kw, d, cout_t, stage;
    ffma_kernel_t fn;
};

#define DLWP_ENTRY(KH, KW, D, CT)                                                            \
    {KH, KW, D, CT, STAGE_CPASYNC, conv_ffma_kernel<KH, KW, D, 2, CT, STAGE_CPASYNC>},       \
    {KH, KW, D, CT, STAGE_TMA, conv_ffma_kernel<KH, KW, D, 2, CT, STAGE_TMA>}

static const KernelEntry g_kernels[] = {
    DLWP_ENTRY(3, 3, 1, 4), DLWP_ENTRY(3, 3, 1, 6), DLWP_ENTRY(3, 3, 1, 8),
    DLWP_ENTRY(3, 3, 2, 4), DLWP_ENTRY(3, 3, 2, 6), DLWP_ENTRY(3, 3, 2, 8),
    DLWP_ENTRY(5, 5, 1, 4), DLWP_ENTRY(5, 5, 1, 6), DLWP_ENTRY(5, 5, 1, 8),
};
constexpr int ROWS_PER_THREAD = 2;

static ffma_kernel_t find_kernel(int kh, int kw, int d, int cout_t, int stage) {
    for (const KernelEntry& e : g_kernels)
        if (e.kh == kh && e.kw == kw && e.d == d && e.cout_t == cout_t && e.stage == stage) return e.fn;
    return nullptr;
}

static int g_num_sms = 0;
static int g_smem_optin = 0;
static std::once_flag g_dev_once;
static int g_dev_status = 0;

int check_device() {
    std::call_once(g_dev_once, [] {
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
            set_error("no CUDA device available");
            g_dev_status = DLWP_EARCH;
            return;
        }
        if (prop.major != 10) {
            set_error("libdlwp_b200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
            g_dev_status = DLWP_EARCH;
            return;
        }
        g_num_sms = prop.multiProcessorCount;
        g_smem_optin = (int)prop.sharedMemPerBlockOptin;
    });
    static std::atomic<unsigned long long> optin_done{0};
    if (g_dev_status == 0 && first_use_on_device(optin_done))   // per device: a process may drive several GPUs
        for (const KernelEntry& e : g_kernels)
            cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin);
    return g_dev_status;
}

typedef CUresult (*encode_tiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tensor_map_4d(CUtensorMap* map, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4]) {
    static encode_tiled_t fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_t>(ptr);
    });
    DLWP_REQUIRE(fn != nullptr, DLWP_EARCH, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdims[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t gbox[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdims, gstr, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLWP_REQUIRE(r == CUDA_SUCCESS, DLWP_ESHAPE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// Generic tiled tensor map over 8-byte elements (the tensor-core path views a P image as 16-byte pixel chunks = 2 elements):
// rank <= 5, dims fastest first, strides_bytes[i] = stride of dimension i + 1.  Returns nonzero (and sets the error string)
// when the driver refuses the descriptor, so callers can fall back to plain bulk copies.
int encode_tensor_map_any(CUtensorMap* map, const void* base, int is_f32, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);
int encode_tensor_map_u64(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box) {
    return encode_tensor_map_any(map, base, 0, rank, dims, strides_bytes, box);
}
int encode_tensor_map_any(CUtensorMap* map, const void* base, int is_f32, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
    static encode_tiled_t fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_t>(ptr);
    });
    DLWP_REQUIRE(fn != nullptr, DLWP_EARCH, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdims[5] = {1, 1, 1, 1, 1}, gstr[4] = {0, 0, 0, 0};
    cuuint32_t gbox[5] = {1, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank,
                    const_cast<void*>(base), gdims, gstr, gbox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLWP_REQUIRE(r == CUDA_SUCCESS, DLWP_ESHAPE, "cuTensorMapEncodeTiled (%s, rank %d) failed with CUresult %d", is_f32 ? "f32" : "u64", rank, (int)r);
    return 0;
}

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct TileChoice {
    bool ok = false;
    int impl = DLWP_IMPL_DIRECT;
    int cout_t = 0;
    TileParams p;
    int threads = 0;
    size_t smem = 0;
    dim3 grid;
};

static TileChoice build_tiling(const DlwpConvDesc& d, bool use_tma) {
    TileChoice tc;
    memset(&tc.p, 0, sizeof(tc.p));
    if (d.rowwise || d.pre_op != 0 || d.dil_h != d.dil_w) return tc;
    int cout_t;
    if (d.Cout % 8 == 0) cout_t = 8;
    else if (d.Cout % 6 == 0) cout_t = 6;
    else if (d.Cout % 4 == 0 || d.Cout < 4) cout_t = 4;
    else cout_t = (d.Cout > 6) ? 8 : 6;
    if (!find_kernel(d.kh, d.kw, d.dil_h, cout_t, STAGE_CPASYNC)) return tc;

    const int Ho_full = d.H + d.pad_t + d.pad_b - d.dil_h * (d.kh - 1);
    const int Wo = d.W + d.pad_l + d.pad_r - d.dil_w * (d.kw - 1);
    if (Ho_full <= 0 || Wo <= 0) return tc;
    const int row0 = (d.row_begin == 0 && d.row_end == 0) ? 0 : d.row_begin;
    const int row1 = (d.row_begin == 0 && d.row_end == 0) ? Ho_full : d.row_end;
    const int Ho = row1 - row0;  // rows this launch computes (a latitude band, or everything)
    if (Ho <= 0) return tc;
    const int halo_w = d.dil_w * (d.kw - 1), halo_h = d.dil_h * (d.kh - 1);
    // STAGE_TMA: the box must start on a 16-byte boundary, the kernel is instantiated for the natural halo only
    const int nat_pad = halo_w / 2;
    if (use_tma && (d.pad_l != nat_pad || d.pad_mode_h != DLWP_PAD_ZERO || (d.x_stride_h % 4) || (d.x_stride_c % 4) ||
                    (d.x_stride_n % 4)))
        return tc;
    const int lead = use_tma ? (4 - nat_pad % 4) % 4 : 0;
    const int nv = round_up(lead + halo_w + 4, 4);  // floats a thread reads per input row

    TileParams& p = tc.p;
    // --- width: full rows when they fit a TMA box (<= 256 elements) ---
    p.tiles_x = ceil_div(Wo + nv - 4, 256 - 4);
    p.TW = round_up(ceil_div(Wo, p.tiles_x), 4);
    p.tiles_x = ceil_div(Wo, p.TW);
    p.NXG = p.TW / 4;
    p.PITCH = p.TW + nv - 4;

    // --- filters per CTA ---
    const int groups = ceil_div(d.Cout, cout_t);
    int ncg = std::min(groups, 4);
    ncg = std::max(1, std::min(ncg, groups));
    // --- rows: aim for ~256-384 threads per CTA ---
    int th = 2 * std::max(1, std::min(8, 384 / (p.NXG * ncg)));
    while (th > 2 && p.NXG * (th / 2) * ncg > 384) th -= 2;
    th = std::min(th, round_up(Ho, 2));
    {   // prefer a nearby TH that wastes fewer padded rows (91 rows: 16 -> 96, 14 -> 98, 12 -> 96)
        int best = th, best_waste = round_up(Ho, th) - Ho;
        for (int t = th; t >= std::max(2, th - 4); t -= 2) {
            const int waste = round_up(Ho, t) - Ho;
            if (waste < best_waste) { best = t; best_waste = waste; }
        }
        th = best;
    }
    th = std::max(2, round_up(th, 2));
    if (p.NXG * (th / 2) * ncg > 384) return tc;
    p.TH = th;
    p.NRG = th / ROWS_PER_THREAD;
    p.NCG = ncg;
    p.RIN = p.TH + halo_h;
    p.tiles_y = ceil_div(Ho, p.TH);
    const int cout_ld = round_up(cout_t, 4);
    p.COUT_BP = ncg * cout_ld;

    // --- channel chunk: two stages of (input tile + weights) should leave room for 2 CTAs per SM ---
    const int budget = 100 * 1024;
    int cc = std::min(d.Cin, 16);
    auto stage_bytes = [&](int c) {
        return (size_t)(round_up(c * p.RIN * p.PITCH, 32) + round_up(c * d.kh * d.kw * p.COUT_BP, 32)) * 4;
    };
    while (cc > 1 && 2 * stage_bytes(cc) > (size_t)budget) cc = (cc > 4) ? cc - 2 : cc - 1;
    cc = std::max(1, std::min(cc, d.Cin));
    p.CC = cc;
    p.nchunks = ceil_div(d.Cin, cc);
    p.in_stage_floats = round_up(cc * p.RIN * p.PITCH, 32);
    p.w_stage_floats = round_up(cc * d.kh * d.kw * p.COUT_BP, 32);
    tc.smem = (size_t)(2 * p.in_stage_floats + 2 * p.w_stage_floats) * 4 + 64;
    if (tc.smem > (size_t)g_smem_optin) return tc;

    p.N = d.N; p.Cin = d.Cin; p.H = d.H; p.W = d.W; p.Cout = d.Cout; p.Ho = Ho_full; p.Wo = Wo;
    p.row0 = row0; p.row1 = row1;
    p.pad_t = d.pad_t; p.pad_l = d.pad_l; p.mode_h = d.pad_mode_h; p.mode_w = d.pad_mode_w; p.act = d.act;
    p.xs_n = d.x_stride_n; p.xs_c = d.x_stride_c; p.xs_h = d.x_stride_h;
    p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    // halo columns patched under TMA staging: the pad_l columns left of x = 0 (first tile) and the columns whose
    // source x >= W that a valid output can touch (last tile)
    p.fix_l = d.pad_l;
    {
        const int x0_last = (p.tiles_x - 1) * p.TW;
        const int needed = Wo - x0_last + halo_w;  // halo-relative columns any valid output of the last tile touches
        p.fix_r = std::max(0, needed - (d.W + d.pad_l - x0_last));
    }
    tc.threads = round_up(p.NXG * p.NRG * p.NCG, 32);
    tc.grid = dim3(p.tiles_x * p.tiles_y, ceil_div(groups, ncg), d.N);
    tc.cout_t = cout_t;
    if (use_tma) {
        const bool fix_ok = d.pad_mode_w == DLWP_PAD_ZERO || (p.fix_l <= d.W && p.fix_r <= d.W);
        if (!fix_ok || p.PITCH > 256 || p.RIN > 256 || cc > 256) return tc;
    }
    tc.impl = use_tma ? DLWP_IMPL_FFMA_TMA : DLWP_IMPL_FFMA;
    tc.ok = true;
    return tc;
}

static TileChoice choose_tiling(const DlwpConvDesc& d, int want_impl) {
    TileChoice tc;
    const bool try_tma = want_impl == DLWP_IMPL_FFMA_TMA || want_impl == DLWP_IMPL_AUTO;
    if (try_tma) tc = build_tiling(d, true);
    if (!tc.ok && want_impl != DLWP_IMPL_FFMA_TMA) tc = build_tiling(d, false);
    return tc;
}

static int validate(const DlwpConvDesc& d) {
    DLWP_REQUIRE(d.N > 0 && d.Cin > 0 && d.H > 0 && d.W > 0 && d.Cout > 0, DLWP_ESHAPE, "non-positive tensor dims");
    DLWP_REQUIRE(d.kh > 0 && d.kw > 0 && d.dil_h > 0 && d.dil_w > 0, DLWP_ESHAPE, "bad kernel size / dilation");
    DLWP_REQUIRE(d.pad_t >= 0 && d.pad_b >= 0 && d.pad_l >= 0 && d.pad_r >= 0, DLWP_ESHAPE, "negative padding");
    DLWP_REQUIRE(d.pad_mode_h == DLWP_PAD_ZERO || d.pad_mode_h == DLWP_PAD_PERIODIC, DLWP_EINVAL, "bad pad_mode_h");
    DLWP_REQUIRE(d.pad_mode_w == DLWP_PAD_ZERO || d.pad_mode_w == DLWP_PAD_PERIODIC, DLWP_EINVAL, "bad pad_mode_w");
    DLWP_REQUIRE(d.act >= DLWP_ACT_LINEAR && d.act <= DLWP_ACT_RELU, DLWP_EINVAL, "bad activation %d", d.act);
    DLWP_REQUIRE(d.pre_op >= 0 && d.pre_op <= 2, DLWP_EINVAL, "bad pre_op %d", d.pre_op);
    DLWP_REQUIRE(d.impl >= DLWP_IMPL_AUTO && d.impl <= DLWP_IMPL_TC, DLWP_EINVAL, "bad impl %d", d.impl);
    return 0;
}

static void logical_dims(const DlwpConvDesc& d, int& H, int& W) {
    H = d.H; W = d.W;
    if (d.pre_op == 1) { H = d.H / 2; W = d.W / 2; }
    if (d.pre_op == 2) { H = d.H * 2; W = d.W * 2; }
}

const char* conv2d_impl_name(const DlwpConvDesc& d) {
    if (check_device() != 0) return "unavailable";
    TileChoice tc = choose_tiling(d, d.impl);
    if (!tc.ok) return "direct";
    return tc.impl == DLWP_IMPL_FFMA_TMA ? "ffma_tma" : "ffma";
}

int conv2d_fwd(const DlwpConvDesc& d, const float* x, const float* w, const float* bias, float* y,
               cudaStream_t stream) {
    int rc = check_device();
    if (rc) return rc;
    DLWP_REQUIRE(x && w && y, DLWP_EINVAL, "null tensor pointer");
    rc = validate(d);
    if (rc) return rc;
    int H, W;
    logical_dims(d, H, W);
    DLWP_REQUIRE(H > 0 && W > 0, DLWP_ESHAPE, "pre_op leaves an empty image");
    // PeriodicPadding2D with a pad larger than the axis silently mis-slices in the reference (SURVEY.md A.7 iii)
    if (d.pad_mode_h == DLWP_PAD_PERIODIC)
        DLWP_REQUIRE(d.pad_t <= H && d.pad_b <= H, DLWP_ESHAPE, "periodic pad larger than the latitude axis");
    if (d.pad_mode_w == DLWP_PAD_PERIODIC)
        DLWP_REQUIRE(d.pad_l <= W && d.pad_r <= W, DLWP_ESHAPE, "periodic pad larger than the longitude axis");
    const int Ho = H + d.pad_t + d.pad_b - d.dil_h * (d.kh - 1);
    const int Wo = W + d.pad_l + d.pad_r - d.dil_w * (d.kw - 1);
    DLWP_REQUIRE(Ho > 0 && Wo > 0, DLWP_ESHAPE, "convolution output would be empty (%d x %d)", Ho, Wo);
    const bool all_rows = d.row_begin == 0 && d.row_end == 0;
    DLWP_REQUIRE(all_rows || (d.row_begin >= 0 && d.row_end <= Ho && d.row_begin < d.row_end), DLWP_ESHAPE,
                 "row window [%d,%d) outside the %d output rows", d.row_begin, d.row_end, Ho);

    if (d.impl == DLWP_IMPL_TC) {
        return conv2d_fwd_tc(d, x, w, bias, y, stream);
    }
    TileChoice tc;
    if (d.impl != DLWP_IMPL_DIRECT) tc = choose_tiling(d, d.impl);
    if (d.impl == DLWP_IMPL_FFMA || d.impl == DLWP_IMPL_FFMA_TMA)
        DLWP_REQUIRE(tc.ok && tc.impl == d.impl, DLWP_ESHAPE, "requested impl %d cannot run this geometry", d.impl);

    if (tc.ok) {
        tc.p.x = x; tc.p.w = w; tc.p.bias = bias; tc.p.y = y;
        const int stage = tc.impl == DLWP_IMPL_FFMA_TMA ? STAGE_TMA : STAGE_CPASYNC;
        ffma_kernel_t fn = find_kernel(d.kh, d.kw, d.dil_h, tc.cout_t, stage);
        CUtensorMap map;
        memset(&map, 0, sizeof(map));
        if (stage == STAGE_TMA) {
            const uint64_t dims[4] = {(uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.Cin, (uint64_t)d.N};
            const uint64_t strides[3] = {(uint64_t)d.x_stride_h * 4, (uint64_t)d.x_stride_c * 4,
                                         (uint64_t)d.x_stride_n * 4};
            const uint32_t box[4] = {(uint32_t)tc.p.PITCH, (uint32_t)tc.p.RIN, (uint32_t)tc.p.CC, 1};
            if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
                encode_tensor_map_4d(&map, x, dims, strides, box) != 0) {
                DLWP_REQUIRE(d.impl != DLWP_IMPL_FFMA_TMA, DLWP_ESHAPE, "TMA descriptor could not be built");
                fn = find_kernel(d.kh, d.kw, d.dil_h, tc.cout_t, STAGE_CPASYNC);
            }
        }
        fn<<<tc.grid, tc.threads, tc.smem, stream>>>(tc.p, map);
        return after_launch("conv_ffma_kernel");
    }

    if (d.rowwise && d.impl != DLWP_IMPL_DIRECT) {   // RowConnected2D: the tiled per-row kernel when it has an instance
        int rc = 0;
        if (conv2d_rowwise(d, x, w, bias, y, Wo, all_rows ? 0 : d.row_begin, all_rows ? Ho : d.row_end - d.row_begin,
                           stream, &rc) == 0)
            return rc;
    }
    DirectParams p;
    p.x = x; p.w = w; p.bias = bias; p.y = y;
    p.N = d.N; p.Cin = d.Cin; p.Hs = d.H; p.Ws = d.W; p.H = H; p.W = W;
    p.Cout = d.Cout; p.Ho = Ho; p.Wo = Wo; p.kh = d.kh; p.kw = d.kw; p.dh = d.dil_h; p.dw = d.dil_w;
    p.pad_t = d.pad_t; p.pad_l = d.pad_l; p.mode_h = d.pad_mode_h; p.mode_w = d.pad_mode_w; p.act = d.act;
    p.pre_op = d.pre_op; p.rowwise = d.rowwise;
    p.xs_n = d.x_stride_n; p.xs_c = d.x_stride_c; p.xs_h = d.x_stride_h;
    p.ys_n = d.y_stride_n; p.ys_c = d.y_stride_c; p.ys_h = d.y_stride_h;
    p.row0 = all_rows ? 0 : d.row_begin;
    p.rows = all_rows ? Ho : d.row_end - d.row_begin;
    const long long total = (long long)d.N * d.Cout * p.rows * Wo;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)g_num_sms * 32);
    conv_direct_kernel<<<blocks, 256, 0, stream>>>(p);
    return after_launch("conv_direct_kernel");
}

}  // namespace dlwp

// Read-and-clear the device-side diagnostic flags (bit 0: a TMA mbarrier wait timed out). Synchronises the device.
extern "C" int dlwp_debug_flags(void) {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, dlwp::g_device_flags, sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyToSymbol(dlwp::g_device_flags, &zero, sizeof(int));
    const int t = dlwp::tc_debug_flags();
    return v | dlwp::plan_flags_read_clear() | (t > 0 ? (t << 1) : 0);
}
