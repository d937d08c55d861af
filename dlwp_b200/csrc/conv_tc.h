// Host-side interface of the tensor-core conv path (conv_tc.cu); internal, not part of the C ABI.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/dlwp_b200.h"

namespace dlwp {

struct TcKStep {
    uint32_t a_off;  // byte offset (inside a stage) of the first 8-channel unit's hi plane, tap row included
    uint32_t a_lbo;  // byte distance to the second unit of this K=16 step
};

// Static schedule of one layer: how the P-layout input is tiled, staged and fed to the tensor core.
struct TcLayer {
    int wpad, Wp;          // periodic halo columns per side of the INPUT image, padded width
    int C8, planes;        // 8-channel chunks of the input, planes = 2 * C8 (hi, lo)
    int CBLK, CSTRIDE, NCOLS;  // 8-filter blocks, TMEM columns per horizontal tap, MMA N
    int S;                 // M-tile stride in pixels (128 - (kw-1)*dil, or 128 with taps_in_k)
    int taps_in_k, kw_eff; // horizontal taps folded into K (small Cin) -> the epilogue sees a 1-tap layer
    int R_out, Rin, MT;    // rows per tile, staged rows, M tiles per tile
    int cpg, G, KS, NS;    // chunks per group, groups per tile, K=16 steps per group, smem stages
    int NACC;              // TMEM accumulator sets (2 = double buffered)
    uint32_t stage_bytes, stage_stride, plane_bytes, b_bytes;
    size_t smem;
    // sliding-window kernel (mode 1, conv_sw_kernel): M tile = 128 consecutive pixels of ONE row; every staged input row
    // feeds all vertical taps (same A operand -> collector reuse), one TMEM accumulator per output row in flight
    int mode;              // 0 = flattened-tile kernel (conv_tc_kernel), 1 = sliding window
    int XLK;               // pixels an A view reaches to the right of its lane (horizontal taps folded into K)
    int nfull, rem, pair;  // full strips per row, valid outputs of the remainder strip, remainder strips of 2 samples share a tile
    uint32_t rowpitch;     // bytes of one staged plane row
};

bool tc_geometry_ok(const DlwpConvDesc& d);
int tc_plan_layer(const DlwpConvDesc& d, TcLayer* L);
int tc_pack_weights(const DlwpConvDesc& d, const TcLayer& L, const float* w_host, std::vector<__half>* img,
                    TcKStep* kst_out);
// Channel windows of P images (slice_layer / concatenate without copies): the conv reads planes
// [in_plane0, in_plane0 + L.planes) of a source image that has in_planes_total planes per sample (0 = L.planes) and writes
// planes [out_plane0, ...) of a destination image with planes_out planes per sample.
struct TcWindow {
    int in_plane0 = 0, in_planes_total = 0, out_plane0 = 0;
    // first layer of a rollout reading the fp32 (N,C,H,W) state directly (no P image of the state): dense source, see
    // tc_f32in_ok; xp is ignored when x32 is set
    const float* x32 = nullptr;
};
// The layer can take its input as fp32 (N,C,H,W) instead of a P image (the folded Net A first-layer instance only, for now).
bool tc_f32in_ok(const DlwpConvDesc& d, const TcLayer& L);
int tc_launch(const DlwpConvDesc& d, const TcLayer& L, const TcKStep* kst, const __half* xp, const __half* bimg,
              const float* bias, float* y32, __half* yp, int wpad_out, int planes_out, cudaStream_t stream,
              const TcWindow& win = TcWindow());
// Data movers on P images (the U-Net's MaxPooling2D(2), UpSampling2D(2) and skip-connection copies): `planes` planes
// starting at src_plane0 of the source -> planes starting at dst_plane0 of the destination; the destination's periodic
// halo (wpad_d columns per side) is written with the interior.  kind: DLWP_OP_COPY / DLWP_OP_MAXPOOL / DLWP_OP_UPSAMPLE.
// (Hs, Ws): source image size.
int tc_ew_launch(int kind, const __half* src, __half* dst, int N, int planes, int Hs, int Ws, int wpad_s, int src_plane0,
                 int src_planes_total, int wpad_d, int dst_plane0, int dst_planes_total, cudaStream_t stream,
                 int row_begin = 0, int row_end = 0);  // destination rows [row_begin, row_end); 0,0 = all
int tc_pack_state(const float* x, __half* xp, int N, int C, int H, int W, int wpad, long long xs_n, long long xs_c,
                  long long xs_h, cudaStream_t stream, int row0 = 0, int row1 = 0);
int conv2d_fwd_tc(const DlwpConvDesc& d, const float* x, const float* w_dev, const float* bias, float* y,
                  cudaStream_t stream);
int tc_debug_flags();
// bytes of a P-layout buffer (planes of (H + zero rows) x Wp pixels x 8 fp16)
size_t tc_p_bytes(int N, int planes, int H, int Wp);

}  // namespace dlwp
