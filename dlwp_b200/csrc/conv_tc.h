// Host-side interface of the tensor-core conv path (conv_tc.cu, conv_fused.cu); internal, not part of the C ABI.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/dlwp_b200.h"

namespace dlwp {

constexpr int TC_HPAD = 12;  // zero rows stored above and below every P-layout plane (>= pad + rows per tile)
constexpr int TC_MAX_KSTEPS = 32;

// ---- power-of-two scaling of the fp16 hi/lo split -------------------------------------------------------------------
// A P image stores x * 2^e as (hi, lo) fp16 pairs.  fp16 keeps 11 significant bits down to 2^-14 and a fixed 2^-24 step
// below, so the pair carries 22 bits of a value v = x * 2^e as long as |v| >= 2^-3, and an absolute error of 2^-25
// otherwise.  With max|v| placed in [2^13, 2^14) every element within 2^-16 of the image's maximum keeps 22 bits and the
// rest are off by < 2^-38 of that maximum: fp32-level accuracy at any input / weight magnitude.  The exponents are exact
// (powers of two), weights get theirs on the host when they are packed, activations on the device from a rigorous bound
// |out| <= l1max(W) * amax(in) + max|b| evaluated with the MEASURED amax of the layer's input (tanh outputs: |y| <= 1,
// static exponent TC_EXP_STATIC).  The consumer's epilogue multiplies the accumulator by 2^-(e_in + e_w).
constexpr int TC_EXP_STATIC = 14;
constexpr int TC_EXP_LIMIT = 60;

// exponent e with B * 2^e in [2^13, 2^14) for a bound B on |values|; host and device
__host__ __device__ inline int tc_exp_for_bound(float B) {
    if (!(B > 1e-30f)) return TC_EXP_LIMIT;   // all zeros: any exponent works
    if (!(B < 1e30f)) return -TC_EXP_LIMIT;   // inf / NaN: the caller raises the range flag
    int ex;
    (void)frexpf(B, &ex);                      // B = m * 2^ex, m in [0.5, 1)  ->  B < 2^ex
    const int e = TC_EXP_STATIC - ex;
    return e > TC_EXP_LIMIT ? TC_EXP_LIMIT : (e < -TC_EXP_LIMIT ? -TC_EXP_LIMIT : e);
}

// Scale bookkeeping of one launch.  Device words live in the plan (one exponent and two amax slots per P buffer).
struct TcScale {
    const int* e_in = nullptr;       // exponent of the source image (device word); null: e_in_const
    int e_in_const = 0;
    const float* amax_in = nullptr;  // bound on the source's |x| in true units: its measured amax (device word) when the
                                     // source is dynamic; null for a static (tanh) source: 1, so that the exponent this
                                     // launch derives is a function of the weights alone (identical on every rank of a
                                     // latitude-band run, whatever rows the rank measured)
    const float* amax_chk = nullptr; // measured max|x| of the source (device word), for the underflow check only
    int* e_out = nullptr;            // exponent of the destination image, decided by this launch (device word); null: static
    int e_out_const = TC_EXP_STATIC;
    float* amax_out = nullptr;       // max|y| of this launch's outputs is atomically max-ed into this word (may be null)
    float* amax_zero = nullptr;      // word to reset for the next production of the destination (may be null)
    int e_w = 0;                     // exponent the weight image was packed with
    float l1max = 0.f, bmax = 0.f;   // max over filters of sum|w|, max|bias| (true units): the output bound's coefficients
};

struct TcWeightScale {
    int e_w = 0;
    float l1max = 0.f;
};

// Tuning / triage switches, fixed when a plan is created (DlwpPlanOptions); never read from the environment here.
struct TcOptions {
    int generic = 0;     // 1: only the generic kernel instances (A/B against the folded ones)
    int bands = 0;       // latitude bands per strip (0: chosen per launch)
    int no_tma = 0;      // 1: per-plane bulk copies instead of tensor-map row loads
    int taps_in_k = -1;  // -1: planner's choice
    int debug = 0;       // 1: epilogue only waits/arrives, 2: issuer only commits (bottleneck triage; wrong results)
    int bf16 = 0;        // 1: plain bf16 images / weights, one MMA pass (DlwpPlanOptions.precision = 1)
    int max_ctas = 0;    // > 0: launch at most this many persistent CTAs (the latitude-band rollout leaves a few SMs to the
                         // halo exchange that runs beside the interior rows of the next iteration)
};

struct TcKStep {
    uint32_t a_off;  // byte offset (inside a stage) of the first 8-channel unit's hi plane, tap row included
    uint32_t a_lbo;  // byte distance to the second unit of this K=16 step
};

// Static schedule of one layer: how the P-layout input is tiled, staged and fed to the tensor core (conv_sw_kernel:
// M tile = 128 consecutive pixels of ONE row; every staged input row feeds all vertical taps with the same A operand,
// one TMEM accumulator per output row in flight).
struct TcLayer {
    int wpad, Wp;          // periodic halo columns per side of the INPUT image, padded width
    int C8, planes;        // 8-channel chunks of the input, planes = ppc * C8
    int ppc;               // planes per chunk: 2 (fp16 hi, lo) or 1 (bf16)
    int CBLK, CSTRIDE, NCOLS;  // 8-filter blocks, TMEM columns per horizontal tap, MMA N
    int S;                 // valid outputs per strip: 128 - (kw-1)*dil
    int taps_in_k, kw_eff; // horizontal taps folded into K (shifted A views) -> the epilogue sees a 1-tap layer
    int KS, NS;            // K=16 steps per input row, shared-memory row stages
    int NACC;              // TMEM accumulator ring (output rows in flight); a multiple of the dilation
    int fold;              // vertical taps covered by one MMA (N = fold * NCOLS <= 256)
    uint32_t stage_stride, b_bytes;
    size_t smem;
    int nfull, rem, pair;  // full strips per row, valid outputs of the remainder strip, remainder strips of 2 samples share a tile
    uint32_t rowpitch;     // bytes of one staged plane row
};

bool tc_geometry_ok(const DlwpConvDesc& d, const TcOptions& opt = TcOptions());
int tc_plan_layer(const DlwpConvDesc& d, TcLayer* L, const TcOptions& opt = TcOptions());
int tc_pack_weights(const DlwpConvDesc& d, const TcLayer& L, const float* w_host, std::vector<__half>* img,
                    TcKStep* kst_out, TcWeightScale* ws);
// Channel windows of P images (slice_layer / concatenate without copies): the conv reads planes
// [in_plane0, in_plane0 + L.planes) of a source image that has in_planes_total planes per sample (0 = L.planes) and writes
// planes [out_plane0, ...) of a destination image with planes_out planes per sample.
struct TcWindow {
    int in_plane0 = 0, in_planes_total = 0, out_plane0 = 0;
};
// Latitude-band neighbours' copies of the destination P image (peer memory): output rows below up_end are also stored into
// `up`, rows from down_begin on into `down` (conv_sw.cuh, SwParams::yp_up).
struct TcPeers {
    __half* up = nullptr;
    __half* down = nullptr;
    int up_end = 0, down_begin = 0;
};
int tc_launch(const DlwpConvDesc& d, const TcLayer& L, const TcKStep* kst, const __half* xp, const __half* bimg,
              const float* bias, float* y32, __half* yp, int wpad_out, int planes_out, cudaStream_t stream,
              const TcWindow& win, const TcScale& sc, const TcOptions& opt, const TcPeers& peers = TcPeers());
// Data movers on P images (the U-Net's MaxPooling2D(2), UpSampling2D(2) and skip-connection copies): `planes` planes
// starting at src_plane0 of the source -> planes starting at dst_plane0 of the destination; the destination's periodic
// halo (wpad_d columns per side) is written with the interior.  kind: DLWP_OP_COPY / DLWP_OP_MAXPOOL / DLWP_OP_UPSAMPLE.
// (Hs, Ws): source image size.  The values are moved bit for bit, so the destination inherits the source's exponent:
// sc.e_in -> sc.e_out and sc.amax_in -> sc.amax_out are forwarded by the kernel (null pointers: nothing to forward).
int tc_ew_launch(int kind, const __half* src, __half* dst, int N, int planes, int Hs, int Ws, int wpad_s, int src_plane0,
                 int src_planes_total, int wpad_d, int dst_plane0, int dst_planes_total, cudaStream_t stream,
                 int row_begin, int row_end, const TcScale& sc, int bf16 = 0);  // destination rows [row_begin, row_end); 0,0 = all
// fp32 (N,C,H,W) -> P image, rows [row0, row1) (0,0 = all).  fresh: measure max|x| over rows [amax_row0, amax_row1) into
// *amax (which must be zero), derive the image's exponent from it and publish it in *e; otherwise the image keeps the
// exponent in *e (rows that join an image another kernel produced, e.g. halo rows received from a neighbour) and the
// packed rows' max|x| is max-ed into *amax.
struct TcPackScale {
    int* e = nullptr;
    float* amax = nullptr;
    float* amax_zero = nullptr;
    int fresh = 1;
    int amax_row0 = 0, amax_row1 = 0;  // fresh: rows whose max|x| defines the exponent (0,0 = all H rows)
    int bf16 = 0;                      // 1: one bf16 plane per chunk, no exponent (e / amax unused)
    // fresh: called (host side) between the max|x| measurement and the packing kernel -- a latitude-band rank that holds
    // only its own rows of x0 reduces the word over the ranks here, so that every rank derives the same exponent
    int (*after_amax)(void* ctx, float* amax, cudaStream_t stream) = nullptr;
    void* after_amax_ctx = nullptr;
};
int tc_pack_state(const float* x, __half* xp, int N, int C, int H, int W, int wpad, long long xs_n, long long xs_c,
                  long long xs_h, cudaStream_t stream, int row0, int row1, const TcPackScale& ps);
// Latitude bands: halo rows received as fp32 (two contiguous (N, C, rows[k], W) staging buffers, null = none) -> rows
// [row0[k], row0[k] + rows[k]) of an existing P image (its exponent *ps.e is kept).  One launch for both neighbours.
int tc_pack_halo(const float* const src[2], const int row0[2], const int rows[2], __half* xp, int N, int C, int H, int W,
                 int wpad, cudaStream_t stream, const TcPackScale& ps);
// Two consecutive layers in one kernel (conv_fused.cu): layer 1's output never leaves the SM.  tc_pair_ok: the pair matches
// an instantiated configuration.  The state image (layer 1's source) must carry wpad1 + wpad2 halo columns, and layer 1's
// weights must be packed with TcLayer::rowpitch = tc_pair_state_pitch (the staged state row is wider than 128 pixels).
bool tc_pair_ok(const DlwpConvDesc& d1, const TcLayer& L1, const DlwpConvDesc& d2, const TcLayer& L2);
uint32_t tc_pair_state_pitch(const DlwpConvDesc& d1);
int tc_pair_launch(const DlwpConvDesc& d1, const TcLayer& L1, const TcKStep* kst1, const __half* bimg1, const float* bias1,
                   const DlwpConvDesc& d2, const TcLayer& L2, const TcKStep* kst2, const __half* bimg2, const float* bias2,
                   const __half* xp, int in_planes_total, int in_plane0, float* y32, __half* yp, int wpad_out,
                   int planes_out, int out_plane0, const TcScale& sc1, const TcScale& sc2, int* done_counter,
                   cudaStream_t stream, const TcOptions& opt);
void tc_preload_aux();   // the small kernels around the conv chain (packing, data movers)
void tc_preload(const DlwpConvDesc& d, const TcLayer& L, int out_mode, int peers, const TcOptions& opt);
int conv2d_fwd_tc(const DlwpConvDesc& d, const float* x, const float* w_dev, const float* bias, float* y,
                  cudaStream_t stream);
int tc_debug_flags();
// bytes of a P-layout buffer (planes of (H + zero rows) x Wp pixels x 8 fp16)
size_t tc_p_bytes(int N, int planes, int H, int Wp);

}  // namespace dlwp
