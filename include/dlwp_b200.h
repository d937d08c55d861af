/*
 * dlwp_b200.h -- C ABI of libdlwp_b200.so, the B200-native implementation of DLWP's forecast-rollout hot path.
 *
 * The reference (jweyn/DLWP @ 3f32bfab) is pure Python on Keras/TensorFlow and has NO FFI: what "calls" the arithmetic
 * is `keras.Model.predict` (DLWP/model/models.py:241, :412) inside the Python rollout loops
 * (DLWP/model/models.py:247-301, :414-452).  This header is the boundary a maintainer would bind from Python
 * (ctypes) to replace exactly those calls.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; no torch / numpy types.  Device pointers are raw `float*` in the CUDA device's address
 *     space (e.g. torch.Tensor.data_ptr()); `dlwp_stream_t` is a `cudaStream_t` passed as an opaque pointer.
 *   - all tensors are fp32, channels_first: (N, C, H, W) with explicit element strides for N, C and H (W is
 *     contiguous), so channel slices (`slice_layer`, DLWP/custom.py:675-692) and channel concatenation
 *     (`keras.layers.concatenate(axis=1)`) are pointer arithmetic, not copies.
 *   - every function returns 0 on success, a negative DLWP_E* code for argument errors, or a positive `cudaError_t`;
 *     nothing throws; `dlwp_last_error_string()` returns a thread-local description of the last failure.
 *   - all device work is asynchronous on the stream passed in, unless the name ends in `_host`.
 *   - the caller owns every pointer it passes; a DlwpPlan owns its weights, workspace and CUDA graphs.
 */
#ifndef DLWP_B200_H
#define DLWP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLWP_B200_ABI_VERSION 3

typedef void* dlwp_stream_t; /* cudaStream_t */

enum {
    DLWP_OK = 0,
    DLWP_EINVAL = -1,   /* bad argument (null pointer, unknown enum, ...)            */
    DLWP_ESHAPE = -2,   /* inconsistent or unsupported tensor geometry               */
    DLWP_EARCH = -3,    /* no sm_100 device / kernel image missing                   */
    DLWP_ENOMEM = -4,   /* workspace allocation failed                               */
    DLWP_ESTATE = -5    /* call sequence error (e.g. weights not set before forward) */
};

/* Padding applied (logically) in front of a convolution.  PERIODIC = PeriodicPadding2D (DLWP/custom.py:191-214),
 * ZERO = keras ZeroPadding2D.  One mode per axis; the example nets use PERIODIC on W (longitude) and ZERO on H. */
enum { DLWP_PAD_ZERO = 0, DLWP_PAD_PERIODIC = 1,
       /* stand-alone padding ops only (dlwp_pad2d / DLWP_OP_PAD), never fused into a conv: */
       DLWP_PAD_EDGE = 2,        /* FillPadding2D (DLWP/custom.py:359-402): replicate the border row / column      */
       DLWP_PAD_REFLECT = 3,     /* TFPadding2D mode='REFLECT' (custom.py:581-586 -> tf.pad): mirror, border excluded */
       DLWP_PAD_SYMMETRIC = 4 }; /* TFPadding2D mode='SYMMETRIC': mirror, border included                            */

/* Conv2D `activation=` values used by the example nets (examples/train.py:159-219). */
enum { DLWP_ACT_LINEAR = 0, DLWP_ACT_TANH = 1, DLWP_ACT_RELU = 2 };

/* Kernel selection. AUTO picks the fastest applicable implementation; the others force one (tests, A/B timing). */
enum {
    DLWP_IMPL_AUTO = 0,
    DLWP_IMPL_DIRECT = 1,     /* one thread per output element, any geometry (reference CUDA kernel)              */
    DLWP_IMPL_FFMA = 2,       /* register-tiled fp32 FFMA kernel, input tiles staged with cp.async                */
    DLWP_IMPL_FFMA_TMA = 3,   /* same math, input tiles staged by TMA (cp.async.bulk.tensor) + in-smem wrap fix-up */
    DLWP_IMPL_TC = 4          /* tcgen05 tensor cores: scaled fp16 hi/lo split x3 (fp32-level accuracy), TMEM accumulators */
};

/* Geometry of one fused [periodic/zero pad] -> Conv2D('valid', stride 1, dilation) -> bias -> activation.
 *   out[n,o,y,x] = act( b[o] + sum_{c,i,j} w[i,j,c,o] * X(n, c, y + dil_h*i - pad_t, x + dil_w*j - pad_l) )
 * X is the (optionally 2x2-max-pooled or nearest-2x-upsampled) source, wrapped modulo the axis length where the axis
 * mode is PERIODIC and 0 outside the axis where it is ZERO.  H_out = H_in + pad_t + pad_b - dil_h*(kh-1), likewise W.
 * Weights use the Keras layout (kh, kw, Cin, Cout); `rowwise` selects RowConnected2D (DLWP/custom.py:695-896): weights
 * (H_out, kh, kw, Cin, Cout) and bias (H_out, Cout). */
typedef struct DlwpConvDesc {
    int32_t N, Cin, H, W;                     /* source dims as stored (before pre_op)                           */
    int32_t Cout, kh, kw, dil_h, dil_w;
    int32_t pad_t, pad_b, pad_l, pad_r;
    int32_t pad_mode_h, pad_mode_w;           /* DLWP_PAD_*                                                      */
    int32_t act;                              /* DLWP_ACT_*                                                      */
    int32_t pre_op;                           /* 0 none, 1 MaxPooling2D(2) on load, 2 UpSampling2D(2) on load    */
    int32_t rowwise;                          /* 1 = RowConnected2D                                              */
    int32_t impl;                             /* DLWP_IMPL_*                                                     */
    int32_t reserved;
    int32_t row_begin, row_end;               /* compute only output rows [row_begin, row_end); 0,0 = all rows   */
    int64_t x_stride_n, x_stride_c, x_stride_h;   /* element strides of the source                              */
    int64_t y_stride_n, y_stride_c, y_stride_h;   /* element strides of the destination                         */
} DlwpConvDesc;

/* ---- stand-alone operators (each replaces one Keras/DLWP.custom layer call inside keras.Model.predict) ---------- */

/* PeriodicPadding2D / ZeroPadding2D + Conv2D (+bias +activation): DLWP/custom.py:191-214 + keras Conv2D as built at
 * DLWP/model/models.py:97-103; RowConnected2D.call (custom.py:823-837) when desc->rowwise. x, w, bias, y: device. */
int dlwp_conv2d_fwd(const DlwpConvDesc* desc, const float* x, const float* w, const float* bias, float* y,
                    dlwp_stream_t stream);

/* Adjoints of dlwp_conv2d_fwd in fp32 (what Keras' backward pass computes inside fit_generator, DLWP/model/models.py:216-228):
 *   dlwp_conv2d_bwd_input   dx += sum over taps of dy * w, with the padding's adjoint (PeriodicPadding2D: wrap-ADD,
 *                           ZeroPadding2D: crop); dy is the gradient w.r.t. the pre-activation output, dx is ACCUMULATED
 *                           into (zero it first for a plain gradient);
 *   dlwp_conv2d_bwd_weight  dw (kh, kw, Cin, Cout) and db (Cout; may be NULL) are ACCUMULATED into (atomic adds).
 * Strides of x / dx come from desc->x_stride_*, of dy from desc->y_stride_*.  rowwise / pre_op descriptors are refused. */
int dlwp_conv2d_bwd_input(const DlwpConvDesc* desc, const float* dy, const float* w, float* dx, dlwp_stream_t stream);
int dlwp_conv2d_bwd_weight(const DlwpConvDesc* desc, const float* x, const float* dy, float* dw, float* db,
                           dlwp_stream_t stream);

/* Stand-alone PeriodicPadding2D.call / ZeroPadding2D.call (custom.py:191-214): y = pad(x). Strides in elements. */
int dlwp_pad2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t pad_t, int32_t pad_b,
               int32_t pad_l, int32_t pad_r, int32_t mode_h, int32_t mode_w, int64_t xs_n, int64_t xs_c, int64_t xs_h,
               int64_t ys_n, int64_t ys_c, int64_t ys_h, dlwp_stream_t stream);
/* Row-windowed variants of the data movers (op = DLWP_OP_PAD/MAXPOOL/UPSAMPLE/COPY): only destination rows
 * [row_begin, row_end) are written. Used by latitude-band plans. */
int dlwp_rows_op(int32_t op, const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t pad_t,
                 int32_t pad_b, int32_t pad_l, int32_t pad_r, int32_t mode_h, int32_t mode_w, int64_t xs_n, int64_t xs_c,
                 int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h, int32_t row_begin, int32_t row_end,
                 dlwp_stream_t stream);

/* (N, H, W, C) <-> (N, C, H, W), dense tensors: to_nchw != 0 reads NHWC and writes NCHW, else the reverse. */
int dlwp_layout2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t to_nchw,
                  dlwp_stream_t stream);

/* Batch assembly of the reference's SeriesDataGenerator.generate (DLWP/model/generators.py:529-605) on the device:
 *   out[b, t * out_c_per_t + out_c0 + v, :, :] = data[samples[b] + t_off + t, sel[v], :, :]     b < B, t < T, v < V
 * data: (n_times, V_total, H, W) dense, resident in HBM; samples (B) int64 and sel (V) int32: device arrays; out: (B, T *
 * out_c_per_t, H, W) dense.  Predictors: t_off = 0, T = input time steps; targets: t_off = input steps + interval - 1 (+
 * output steps * s for the s-th element of a sequence); the insolation channel is a second call with V_total = V = 1 and
 * out_c0 = the number of selected variables. */
int dlwp_gather_series(const float* data, const int64_t* samples, const int32_t* sel, float* out, int32_t B, int32_t T,
                       int32_t V, int32_t V_total, int32_t H, int32_t W, int32_t t_off, int32_t out_c_per_t,
                       int32_t out_c0, dlwp_stream_t stream);

/* keras MaxPooling2D(2) ('valid', floor) and UpSampling2D(2) (nearest), channels_first. */
int dlwp_maxpool2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n, int64_t xs_c,
                   int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h, dlwp_stream_t stream);
int dlwp_upsample2d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n, int64_t xs_c,
                    int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h, dlwp_stream_t stream);
/* Strided channel-block copy (materialised concatenate / slice when a zero-copy view is not possible). */
int dlwp_copy4d(const float* x, float* y, int32_t N, int32_t C, int32_t H, int32_t W, int64_t xs_n, int64_t xs_c,
                int64_t xs_h, int64_t ys_n, int64_t ys_c, int64_t ys_h, dlwp_stream_t stream);

/* One time step of keras ConvLSTM2D's cell after its two convolutions (keras ConvLSTM2DCell.call; the layer is built at
 * examples/train.py:144-157):  i = s(zi + ri), f = s(zf + rf), c = f * c_prev + i * act(zc + rc), o = s(zo + ro),
 * h = o * act(c), with s = hard_sigmoid (clip(0.2 x + 0.5, 0, 1)) or sigmoid.  z: (N, 4F, H, W) input-conv
 * pre-activations (bias included), gate blocks in keras order i, f, c, o; r: the recurrent conv's, or NULL (first step:
 * h_-1 = 0 and c_prev is not read, may be NULL).  c_out may alias c_prev.  Strides in elements; rows [row_begin, row_end)
 * only (0,0 = all). */
int dlwp_convlstm_gates(const float* z, const float* r, const float* c_prev, float* c_out, float* h_out, int32_t N,
                        int32_t F, int32_t H, int32_t W, int64_t zs_n, int64_t rs_n, int64_t cs_n, int64_t hs_n,
                        int32_t act, int32_t ract, int32_t row_begin, int32_t row_end, dlwp_stream_t stream);

/* ---- the network plan: what keras.Model.predict executes for one model application ------------------------------ */

enum { DLWP_BUF_INTERNAL = 0, DLWP_BUF_INPUT = 1, DLWP_BUF_OUTPUT = 2 };
enum { DLWP_OP_CONV = 0, DLWP_OP_PAD = 1, DLWP_OP_MAXPOOL = 2, DLWP_OP_UPSAMPLE = 3, DLWP_OP_COPY = 4,
       DLWP_OP_LSTM = 5 /* ConvLSTM2D gate step: see dlwp_convlstm_gates */,
       /* data_format='channels_last' models (DLWP/custom.py:205-213; the Keras default): the plan computes channels_first,
        * the caller-bound input / outputs are (N, H, W, C) memory -- one layout op at each end (whole buffers only) */
       DLWP_OP_TO_NCHW = 6, DLWP_OP_TO_NHWC = 7 };
enum { DLWP_RACT_HARD_SIGMOID = 0, DLWP_RACT_SIGMOID = 1 };   /* keras recurrent_activation */

/* A dense (N, C, H, W) activation. INTERNAL buffers are allocated by the plan; the INPUT buffer and the OUTPUT buffers
 * are bound to caller memory at run time (dense, contiguous). `output_index` orders the model outputs. */
typedef struct DlwpBufferDesc {
    int32_t kind;       /* DLWP_BUF_* */
    int32_t C, H, W;
    int32_t output_index;
    int32_t reserved;
} DlwpBufferDesc;

/* One op: dst[:, dst_c0:dst_c0+Cout'] = f(src[:, src_c0:src_c0+Cin']). Channel windows implement slice_layer and
 * concatenate(axis=1) without copies. For DLWP_OP_CONV `conv` carries geometry (strides/N/H/W are filled in by the
 * plan); `weight_id` indexes the plan's weight table so layers shared between unrolled applications
 * (examples/train_functional.py:278-281) share storage. */
typedef struct DlwpOpDesc {
    int32_t kind;              /* DLWP_OP_* */
    int32_t src, src_c0, src_c;
    int32_t dst, dst_c0;
    int32_t weight_id;         /* conv only, else -1 */
    int32_t pad_t, pad_b, pad_l, pad_r, pad_mode_h, pad_mode_w;   /* conv and pad */
    int32_t Cout, kh, kw, dil_h, dil_w, act, pre_op, rowwise, impl; /* conv only */
    int32_t row_begin, row_end;    /* destination rows this plan computes ([0,0) = all): latitude-band partitioning */
    /* DLWP_OP_LSTM: src = pre-activations, channels [i | f | c | o] (4*Cout) of the input conv (bias included), followed
     * by the same four blocks of the recurrent conv when src_c == 8*Cout (absent on the first time step: h_-1 = 0);
     * aux / aux_c0 = the cell-state buffer (Cout channels, updated in place; ignored as an input when src_c == 4*Cout);
     * dst / dst_c0 = where h_t goes; act = cell activation, act2 = DLWP_RACT_* */
    int32_t aux, aux_c0, act2;
} DlwpOpDesc;

typedef struct DlwpNetDesc {
    int32_t n_buffers;
    int32_t n_ops;
    int32_t n_weights;
    int32_t max_batch;               /* INTERNAL buffers are sized for this many samples */
    const DlwpBufferDesc* buffers;
    const DlwpOpDesc* ops;
} DlwpNetDesc;

typedef struct DlwpPlan DlwpPlan;

/* How a plan is built; fixed at creation (the library never reads the environment).  All zeros = defaults. */
typedef struct DlwpPlanOptions {
    int32_t math;          /* 0: tensor-core chain when the whole plan is eligible, else fp32 kernels; 1: fp32 FFMA kernels
                              (training needs them: fp32 intermediates)                                                  */
    int32_t fuse;          /* 0: one kernel per layer; 1: eligible conv -> conv pairs run as one kernel (conv_fused.cu; the
                              intermediate never leaves the SM).  Opt-in: measured slower than the two-kernel chain      */
    int32_t tc_generic;    /* 1: only the generic tensor-core kernel instances (A/B against the folded ones)             */
    int32_t tc_bands;      /* sliding-window kernel: fewest output rows a CTA is started for (0 = default 8); the fused
                              pair kernel: latitude bands per strip (0 = chosen per launch)                               */
    int32_t tc_no_tma;     /* 1: per-plane bulk copies instead of tensor-map row loads                                   */
    int32_t tc_taps_in_k;  /* -1: planner's choice; 0 / 1: horizontal taps in the MMA N / K dimension                     */
    int32_t tc_debug;      /* bottleneck triage: 1 = epilogue only waits/arrives, 2 = issuer only commits (both: WRONG
                              RESULTS); 4 = the issuing warp times itself (dlwp_debug_counters)                         */
    int32_t precision;     /* tensor-core chain: 0 = fp32-equivalent (fp16 hi/lo split, three MMA passes; the 1e-4 / 50-step
                              parity gate), 1 = plain bf16 activations and weights, one MMA pass, fp32 accumulation
                              (BASELINE.json configs[2]; parity ~1e-2 per application, reported, not gated)              */
    int32_t latband_spare_sms; /* dlwp_rollout_latband, NCCL exchange: > 0 runs the exchange on a second stream beside the
                              interior rows of the next iteration's first layer, which leaves this many SMs free for it;
                              0 (default) = exchange, then compute, on one stream (measured faster at 2 GPUs:
                              profiles/r02_latband_2gpu.txt)                                                             */
    int32_t reserved[7];
} DlwpPlanOptions;

/* Build the executable form of a Keras graph (replaces graph construction at DLWP/model/models.py:96-112, :349-373).
 * dlwp_plan_create = dlwp_plan_create_opts with options NULL (defaults). */
int dlwp_plan_create(const DlwpNetDesc* net, DlwpPlan** plan);
int dlwp_plan_create_opts(const DlwpNetDesc* net, const DlwpPlanOptions* options, DlwpPlan** plan);
void dlwp_plan_destroy(DlwpPlan* plan);

/* keras `layer.set_weights([kernel, bias])` / `get_weights()` (DLWP/custom.py:126,136; DLWP/util.py:180).
 * HOST pointers, Keras layouts: kernel (kh,kw,Cin,Cout) [rowwise: (H_out,kh,kw,Cin,Cout)], bias (Cout) [rowwise:
 * (H_out,Cout)], bias may be NULL (use_bias=False). Synchronous. */
int dlwp_plan_set_weights(DlwpPlan* plan, int32_t weight_id, const float* kernel, int64_t kernel_elems,
                          const float* bias, int64_t bias_elems);
int dlwp_plan_get_weights(DlwpPlan* plan, int32_t weight_id, float* kernel, int64_t kernel_elems, float* bias,
                          int64_t bias_elems);

/* One model application on N <= max_batch samples: keras.Model.predict (DLWP/model/models.py:241, :412).
 * x: device (N,C,H,W); outputs[k]: device, dense, one per model output. */
int dlwp_plan_forward(DlwpPlan* plan, int32_t N, const float* x, float* const* outputs, dlwp_stream_t stream);

/* The rollout loop of DLWPNeuralNet.predict_timeseries (models.py:277-293, non-step_sequence branch) and
 * DLWPFunctional.predict_timeseries (models.py:439-447): `iterations` model applications, each fed the LAST output of
 * the previous one, every output stored.  series: device (iterations * n_outputs, N, C, H, W); step t reads series
 * slot t*n_outputs-1 directly (no state copy).  Requires every output shape == input shape.
 * use_graph != 0 captures the whole rollout into one CUDA graph (cached per (N, iterations, x0, series)). */
int dlwp_rollout(DlwpPlan* plan, int32_t N, const float* x0, float* series, int32_t iterations, int32_t use_graph,
                 dlwp_stream_t stream);

/* DLWPNeuralNet.predict_timeseries(step_sequence=True) (DLWP/model/models.py:280-290) on the device: the state holds
 * time_dim slices of C / time_dim channels ((N, T, C', H, W) recurrent inputs are the same memory); after every application
 * the next input is the old one without its first slice followed by the first slice of the prediction.  series
 * (iterations, N, C, H, W) receives every whole prediction.  x0, series: device; the input is never modified. */
int dlwp_rollout_step_sequence(DlwpPlan* plan, int32_t N, const float* x0, float* series, int32_t iterations,
                               int32_t time_dim, dlwp_stream_t stream);

/* Same with HOST buffers (numpy in / numpy out, like the reference's API): H2D of x0, rollout, D2H of the series
 * pipelined behind the compute in groups of `d2h_group` iterations. Pinned host memory gives full PCIe speed. Blocking.
 * The plan keeps the device buffers for reuse. */
int dlwp_rollout_host(DlwpPlan* plan, int32_t N, const float* x0_host, float* series_host, int32_t iterations,
                      int32_t d2h_group);

/* ---- latitude-band rollout: one process per GPU, halo rows exchanged with ONE grouped NCCL SendRecv per iteration ----- */

/* This rank's share of the H latitude rows (dlwp_b200.parallel.BandPlanner computes it): it owns output rows
 * [band_lo, band_hi); before an iteration it needs `recv_top` rows above and `recv_bot` rows below its band from its
 * neighbours, and serves them `send_up` / `send_down` rows of its own band. The plan's ops must be row-windowed. */
typedef struct DlwpBandInfo {
    int32_t rank, world;
    int32_t band_lo, band_hi;
    int32_t recv_top, recv_bot, send_up, send_down;
} DlwpBandInfo;

/* NCCL is resolved at run time from `libnccl_path` (torch's bundled libnccl.so.2); the library has no link dependency on
 * it. Rank 0 creates the 128-byte id and distributes it (e.g. torch.distributed.broadcast); every rank then creates the
 * communicator. */
int dlwp_comm_unique_id(const char* libnccl_path, void* id128);
int dlwp_comm_create(const char* libnccl_path, int32_t rank, int32_t world, const void* id128, void** comm);
void dlwp_comm_destroy(void* comm);

/* dlwp_rollout for a latitude band: after every iteration but the last, the band rows of the last output that the
 * neighbours need are packed, exchanged (ncclGroupStart; ncclSend/ncclRecv up and down; ncclGroupEnd) and unpacked into
 * the halo rows of the same series slot, all on `stream`. use_graph captures kernels AND the NCCL group into one CUDA
 * graph. series holds full (N,C,H,W) slots of which the band rows are valid. */
int dlwp_rollout_latband(DlwpPlan* plan, void* comm, int32_t N, const float* x0, float* series, int32_t iterations,
                         const DlwpBandInfo* band, int32_t use_graph, dlwp_stream_t stream);

/* The same with host buffers (what LatBandEngine.predict_timeseries calls; the band counterpart of dlwp_rollout_host and of
 * the loops at DLWP/model/models.py:277-293): x0_host is the full (N,C,H,W) initial state -- only the rows this rank reads
 * (band + halo) are copied to the device when `comm` is given (tensor-core plans then agree on max|x0|, which fixes the
 * exponent of the x0 image, with one 4-byte ncclAllReduce; comm == NULL: the whole x0 goes up) -- and band_host receives
 * THIS rank's band of every state, (iterations * n_outputs, N, C, band_hi - band_lo, W) contiguous.  The device series stays inside
 * the plan; the band rows of finished groups of `d2h_group` steps (0 = iterations / 8) travel to the host as strided 2-D
 * copies on a second stream while the next steps compute.  Blocks until band_host is complete. */
int dlwp_rollout_latband_host(DlwpPlan* plan, void* comm, int32_t N, const float* x0_host, float* band_host,
                              int32_t iterations, const DlwpBandInfo* band, int32_t d2h_group);

/* Halo exchange over peer memory (NVLink) instead of NCCL: after dlwp_plan_halo_enable the input image is double-buffered
 * and the last conv of iteration t stores the rows its neighbours need straight into THEIR images from its epilogue; two
 * one-thread kernels per iteration count arrivals (no packing, no copy kernels, no collective).  Set-up, once per plan:
 * every rank enables, exports 3 x 64 bytes of CUDA IPC handles, the ranks swap them (any side channel), and each imports
 * its upper (which = 0) and lower (which = 1) neighbour's.  dlwp_rollout_latband then ignores `comm` (may be NULL).
 * dlwp_plan_halo_enable returns DLWP_ESTATE for plans it cannot serve (fp32 kernels, no feedback conv, a first layer whose
 * output exponent depends on the data): keep the NCCL path for those.  dlwp_plan_halo_connect links two plans of ONE
 * process by raw pointers (tests: several bands on one GPU). */
int dlwp_plan_halo_enable(DlwpPlan* plan);
int dlwp_plan_halo_export(DlwpPlan* plan, void* handles192);
int dlwp_plan_halo_import(DlwpPlan* plan, int32_t which, const void* handles192);
int dlwp_plan_halo_connect(DlwpPlan* plan, int32_t which, DlwpPlan* neighbour);


/* ---- training (BASELINE.json configs[4]): what keras fit_generator / train_on_batch does per batch ------------------- */

/* Forward through the plan, loss = sum_k loss_weights[k] * mean((yhat_k - y_k)^2) (Keras 'mse'), and -- if `backward` --
 * backprop into one flat gradient buffer (kernel then bias per weight id, Keras layouts; shared layers accumulate).
 * x, targets[k]: device, dense. losses[k] / maes[k] (host, may be NULL for maes) receive the unweighted per-output MSE / MAE.
 * Needs a plan built on the fp32 kernels (DlwpPlanOptions.math = 1). Blocking (returns the loss). */
int dlwp_train_step(DlwpPlan* plan, int32_t N, const float* x, const float* const* targets, const float* loss_weights,
                    int32_t backward, int32_t input_grad, float* losses, float* maes, dlwp_stream_t stream);
/* Optional (H, W) weight map of DLWP.custom.latitude_weighted_loss (custom.py:956-991): both tensors are multiplied by it
 * before the MSE. HOST pointer; NULL removes it. */
int dlwp_train_loss_weights(DlwpPlan* plan, const float* wmap_host, int64_t elems);
/* DLWP.custom.anomaly_correlation_loss (custom.py:1036-1088) instead of the MSE (kind 1; 0 restores the MSE): per output,
 * over the whole batch tensor, a = <p-mu, y-mu> / sqrt(|p-mu|^2 |y-mu|^2) and loss = m - a (reverse != 0) or a - m, with
 * m = the MSE (regularize 1), the MAE (2) or nothing (0; regularized losses are always reversed).  mean_host: optional
 * climatology mu, one sample of an output (C*H*W floats, HOST), NULL = 0.  dlwp_train_step then reports this loss in
 * losses[k].  Not combined with the latitude weight map. */
int dlwp_train_loss_kind(DlwpPlan* plan, int32_t kind, int32_t regularize, int32_t reverse, const float* mean_host,
                         int64_t mean_elems);
/* The flat gradient buffer (device) for a data-parallel all-reduce between dlwp_train_step and dlwp_train_adam, and the
 * gradient w.r.t. the input (valid when input_grad was set). */
int dlwp_train_buffers(DlwpPlan* plan, float** flat_grad, int64_t* elems, float** input_grad);
/* Data-parallel training (one process per GPU; the reference's multi-GPU mode is keras multi_gpu_model, models.py:104-109):
 * averages the flat gradient buffer over the ranks of `comm` (dlwp_comm_create) in place with one ncclAllReduce(ncclAvg),
 * enqueued on `stream` -- between dlwp_train_step and dlwp_train_regularize / dlwp_train_adam. */
int dlwp_train_allreduce(DlwpPlan* plan, void* comm, dlwp_stream_t stream);
int dlwp_train_weight_offsets(DlwpPlan* plan, int32_t weight_id, int64_t* kernel_off, int64_t* bias_off);
/* keras.regularizers.L1L2 (kernel_regularizer= / bias_regularizer=, examples/train.py:155): adds l1 * sign(w) + 2 * l2 * w
 * to the weight's slots of the flat gradient buffer (after dlwp_train_step and the all-reduce, before dlwp_train_adam) and
 * returns the penalty sum(l1 |w| + l2 w^2) that Keras adds to the reported loss. Blocking. */
int dlwp_train_regularize(DlwpPlan* plan, int32_t weight_id, float kernel_l1, float kernel_l2, float bias_l1,
                          float bias_l2, float* penalty, dlwp_stream_t stream);
/* Read (set = 0) or write (set = 1) the Adam moments and step count: HOST arrays of `elems` = flat gradient elements.
 * Lets the optimizer state survive a plan rebuild (larger batch). */
int dlwp_train_adam_state(DlwpPlan* plan, float* m_host, float* v_host, int64_t elems, int64_t* step, int32_t set);
/* Keras Adam update of every weight from the flat gradient buffer (lr_t = lr*sqrt(1-b2^t)/(1-b1^t), eps outside sqrt). */
int dlwp_train_adam(DlwpPlan* plan, float lr, float beta1, float beta2, float eps, dlwp_stream_t stream);

/* Time one op of the plan alone: `iters` launches bracketed by CUDA events on `stream` (after 2 warm-up launches), using
 * whatever the plan's buffers hold from the last forward / rollout. Blocking. Used by bench.py for the roofline figure. */
int dlwp_plan_profile_op(DlwpPlan* plan, int32_t N, int32_t op_index, int32_t iters, float* ms_per_launch,
                         dlwp_stream_t stream);
/* 1 if the plan runs as a tcgen05 tensor-core chain (DlwpPlanOptions.math = 0 and every op eligible). */
int dlwp_plan_uses_tensor_cores(DlwpPlan* plan);

/* Index of the first op of the conv -> conv pair that runs as one fused kernel (its partner is the next op), or -1.
 * dlwp_plan_profile_op reports 0 ms for that op and the fused kernel's time for the partner. */
int dlwp_plan_fused_pair(DlwpPlan* plan);

/* ---- introspection --------------------------------------------------------------------------------------------- */

const char* dlwp_last_error_string(void);
int dlwp_abi_version(void);
/* Number of kernels this library launched since load (all streams); bench.py reports the delta as gpu_launches. */
int64_t dlwp_kernel_launch_count(void);
/* Read-and-clear device-side diagnostic flags. Synchronises the device.
 *   1: an fp32-kernel TMA wait timed out   2: a tensor-core kernel's mbarrier wait timed out
 *   4: a value left the fp16 hi/lo split's range (NaN / inf data)   8: an image's amax fell below 2^-6 in scaled units
 *   (precision underflow of a statically scaled tanh image).  dlwp_b200.engine reruns on the fp32 kernels on 4 / 8. */
int dlwp_debug_flags(void);
/* Name of the implementation AUTO would choose for this descriptor ("direct", "ffma", "ffma_tma"). */
const char* dlwp_conv2d_impl_name(const DlwpConvDesc* desc);
/* Host-only test hooks of the tensor-core path (no GPU needed).
 * dlwp_debug_tc_plan: the schedule the planner picks for one layer. out[0..15] = mode (0 flattened tiles, 1 sliding window),
 *   taps_in_k, NCOLS, NACC, KS, NS, S, nfull, rem, pair, shared-memory bytes, weight-image bytes, CBLK, CSTRIDE, planes,
 *   row pitch.  Returns 0, or DLWP_ESHAPE when the layer cannot run on the tensor-core kernels.
 * dlwp_debug_tc_pack: the packed fp16 hi/lo weight image of that schedule (kernel in Keras layout (kh,kw,Cin,Cout), HOST
 *   pointer; the image holds w * 2^weight_exponent) and the low words of its A-operand descriptors per K step; returns
 *   the number of fp16 elements written.  l1max = max over filters of sum|w| (the output bound's coefficient). */
int dlwp_debug_tc_plan(const DlwpConvDesc* desc, int32_t* out, int32_t n_out);
/* dlwp_debug_sw_cover: walks the sliding-window kernel's scheduling units for this layer on `sms` SMs exactly as the
 *   device roles do and counts, per output pixel (N,H,W), how many units write it (must be 1 inside [row_begin,row_end),
 *   0 outside).  info[0..3] = CTAs, segments (strip x row range) over all CTAs, staged input rows of the busiest CTA,
 *   staged input rows over all CTAs. */
int dlwp_debug_sw_cover(const DlwpConvDesc* desc, int32_t sms, int32_t* cover, int64_t cover_elems, int32_t* info,
                        int32_t n_info);
int64_t dlwp_debug_tc_pack(const DlwpConvDesc* desc, const float* kernel, uint16_t* image, int64_t image_cap,
                           uint32_t* kstep_words, int32_t kstep_cap, int32_t* weight_exponent, float* l1max);
/* Description of the fully folded tensor-core kernel instance this layer launches when it writes a P image (out_mode 1),
 * fp32 (2) or both (3); NULL = generic instance (profiles/r02_folded_vs_generic.txt compares the two). */
const char* dlwp_debug_tc_folded(const DlwpConvDesc* desc, int32_t out_mode);
/* Plans created with tc_debug & 4 make the MMA-issuing warp time itself (clock64): totals since the last call, summed over
 * CTAs and launches: out[0] cycles waiting for accumulator slots, [1] waiting for staged rows, [2] issuing MMAs,
 * [3] rows issued; out[4..7] the same for the second layer of a fused pair; n >= 12 adds [8] clocks from kernel entry to
 * the issuing warp's exit, [9] that span in ns (globaltimer), [10] CTAs counted, [11] clocks until the first staged row
 * arrived.  Synchronises the device. */
int dlwp_debug_counters(int64_t* out, int32_t n);
/* The power-of-two exponent e the tensor-core path gives an image whose values are bounded by `bound`:
 * bound * 2^e in [2^13, 2^14) (conv_tc.h). */
int dlwp_debug_exp_for_bound(float bound);

#ifdef __cplusplus
}
#endif
#endif /* DLWP_B200_H */
