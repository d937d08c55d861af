// The network plan and the device-resident rollout engine.
//
// A DlwpPlan is the executable form of the Keras graph that DLWPNeuralNet.build_model / DLWPFunctional.build_model
// compile (DLWP/model/models.py:96-112, :349-373): a list of ops over (N,C,H,W) buffers with channel windows, so that
// slice_layer (DLWP/custom.py:675-692) and concatenate(axis=1) are views.  dlwp_rollout replaces the Python feedback
// loops of DLWP/model/models.py:277-293 and :439-447: the forecast state never leaves HBM -- iteration t reads slot
// t*n_out-1 of the output series directly and writes slots t*n_out .. t*n_out+n_out-1 -- and the whole loop can be
// replayed as one CUDA graph.
#include "internal.h"
#include "conv_tc.h"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace dlwp {

__device__ int g_device_flags_plan = 0;   // bit 0: a halo wait timed out (read by dlwp_debug_flags through plan_flags_read_clear)
int plan_flags_read_clear() {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_device_flags_plan, sizeof(int)) != cudaSuccess) return 0;
    if (v) cudaMemcpyToSymbol(g_device_flags_plan, &zero, sizeof(int));
    return v;
}

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

struct Buffer {
    DlwpBufferDesc d;
    float* ptr = nullptr;  // INTERNAL: owned; INPUT/OUTPUT: bound per call
    long long sample_elems() const { return (long long)d.C * d.H * d.W; }
    // tensor-core chain: the same activation in P layout (fp16 hi/lo planes, periodic halo materialised)
    __half* P = nullptr;
    int wpad = -1, planes = 0;
    // power-of-two exponent of the P image (conv_tc.h): static = TC_EXP_STATIC (every producer is a tanh conv or moves data
    // from such an image), else decided per production on the device and kept in DlwpPlan::d_exp[buffer]
    bool e_static = true;
};

struct Weight {
    float* k = nullptr;
    float* b = nullptr;
    long long k_elems = 0, b_elems = 0;
    bool has_bias = false, set = false;
    __half* bimg = nullptr;        // tensor-core chain: packed hi/lo weight images
    TcKStep kst[32];
    TcWeightScale ws;              // exponent the image was packed with, max_co sum|w| (the output bound's coefficient)
    float bmax = 0.f;              // max|bias|
};

struct GraphKey {
    int N, iterations;
    const float* x0;
    float* series;
    bool operator<(const GraphKey& o) const {
        return std::tie(N, iterations, x0, series) < std::tie(o.N, o.iterations, o.x0, o.series);
    }
};

}  // namespace dlwp

using namespace dlwp;

struct DlwpPlan {
    std::vector<Buffer> buffers;
    std::vector<DlwpOpDesc> ops;
    std::vector<Weight> weights;
    int max_batch = 0;
    int input_buf = -1;
    std::vector<int> outputs;  // buffer index per output_index
    // rollout state
    std::map<GraphKey, cudaGraphExec_t> graphs;
    float* d_x0 = nullptr;
    float* d_series = nullptr;
    long long d_x0_cap = 0, d_series_cap = 0;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> d2h_events;   // dlwp_rollout_latband_host: one per group of steps
    void* x0_amax_comm = nullptr;          // set while x0 holds only this rank's rows: max|x0| is reduced over these ranks
    std::vector<cudaEvent_t> events;
    // tensor-core chain mode (every op a tc-capable conv): per-op schedule + which packed buffer each conv writes
    bool tc = false;
    std::vector<TcLayer> tc_layers;
    std::vector<int> tc_pdst;      // buffer whose P image op i writes (-1: none)
    int tc_feedback_op = -1;       // op that also serves as the packer of the next iteration's input
    int tc_in_row0 = 0, tc_in_row1 = 0;  // rows of the input the (row-windowed) convs read: only those are packed
    // latitude band: rows [tc_band_row0, tc_band_row1) of the next input are re-packed by the feedback conv itself; only the
    // halo rows that arrive as fp32 from the neighbours ([tc_in_row0, band_row0) and [band_row1, tc_in_row1)) are packed
    int tc_band_row0 = 0, tc_band_row1 = 0;
    DlwpPlanOptions opt;           // as given to dlwp_plan_create_opts (zeros = defaults)
    TcOptions tc_opt;
    // scale words of the P images (conv_tc.h): exponent and amax per buffer, each double-buffered by production parity --
    // the fused kernel reads the state's exponent of iteration t while its first CTA already publishes the one of t + 1
    int* d_exp = nullptr;
    float* d_amax = nullptr;
    int tc_last_t = 0;             // iteration index of the last application (dlwp_plan_profile_op reuses its parity)
    // ops pair_first and pair_first + 1 run as ONE kernel (conv_fused.cu); -1: no fused pair
    int pair_first = -1;
    int* d_counter = nullptr;      // device word for the fused kernel's last-CTA check
    // latitude-band rollout: contiguous staging for the halo rows sent / received per iteration
    float* halo_stage[4] = {nullptr, nullptr, nullptr, nullptr};  // send_up, send_down, recv_top, recv_bot
    bool halo_in_p = false;
    // latitude-band halo over peer memory (dlwp_plan_halo_*): the input image is double-buffered by iteration parity, the
    // feedback conv of iteration t stores its band rows into P_in[(t+1)&1] and the rows the neighbours need into THEIR
    // P_in[(t+1)&1]; arrivals are counted in `flags` ([0] from the upper, [1] from the lower neighbour, [2] / [3] how many
    // this rank has consumed)
    bool p2p = false;
    __half* P_in[2] = {nullptr, nullptr};
    __half* peer_P[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [up | down][parity]
    int* flags = nullptr;
    int* peer_flags[2] = {nullptr, nullptr};
    std::vector<void*> ipc_opened;
    int cur_parity = 0;            // parity of the iteration being launched (selects the feedback conv's destination)
    bool p2p_active = false;       // inside dlwp_rollout_latband's peer-memory loop
    bool p2p_send = false;         // this iteration's feedback conv also writes the neighbours' halo rows
    int p2p_up_end = 0, p2p_down_begin = 0;
    float* seq_state[2] = {nullptr, nullptr};   // dlwp_rollout_step_sequence: ping-pong input states
    long long seq_cap = 0;        // the last exchange wrote the received rows into the input's P image (not the fp32 slot)
    long long halo_cap = 0;
    std::map<GraphKey, cudaGraphExec_t> band_graphs;
    // training state (lazily created by dlwp_train_step)
    bool train_ready = false;
    std::vector<float*> gbuf;        // gradient w.r.t. every buffer (max_batch samples)
    std::vector<float*> out_store;   // plan-owned model outputs during training
    float *flat_g = nullptr, *flat_m = nullptr, *flat_v = nullptr, *stats = nullptr;
    float *bwd_scratch = nullptr, *wt_scratch = nullptr;   // dX via the forward kernels: scratch gradient, flipped weights
    std::vector<char> g_touched;                           // per buffer: something has accumulated into its gradient this step
    long long flat_elems = 0, adam_t = 0;
    std::vector<long long> gk_off, gb_off;
    float* loss_wmap = nullptr;      // optional (H, W) latitude weights of the loss
    // DLWP.custom.anomaly_correlation_loss instead of the MSE (dlwp_train_loss_kind)
    int loss_kind = 0, acc_regularize = 1, acc_reverse = 1;
    float* acc_mean = nullptr;       // optional climatology, one sample of an output
    long long acc_mean_elems = 0;
    double* acc_stats = nullptr;     // 5 sums per output
};

namespace dlwp {

static void out_dims(const DlwpPlan* pl, const DlwpOpDesc& op, int& C, int& H, int& W) {
    const Buffer& s = pl->buffers[op.src];
    C = op.src_c; H = s.d.H; W = s.d.W;
    switch (op.kind) {
        case DLWP_OP_CONV: {
            int h = H, w = W;
            if (op.pre_op == 1) { h /= 2; w /= 2; }
            if (op.pre_op == 2) { h *= 2; w *= 2; }
            C = op.Cout;
            H = h + op.pad_t + op.pad_b - op.dil_h * (op.kh - 1);
            W = w + op.pad_l + op.pad_r - op.dil_w * (op.kw - 1);
            break;
        }
        case DLWP_OP_PAD: H += op.pad_t + op.pad_b; W += op.pad_l + op.pad_r; break;
        case DLWP_OP_MAXPOOL: H /= 2; W /= 2; break;
        case DLWP_OP_UPSAMPLE: H *= 2; W *= 2; break;
        case DLWP_OP_LSTM: C = op.Cout; break;
        default: break;
    }
}

static DlwpConvDesc conv_desc_of(const DlwpPlan* pl, const DlwpOpDesc& op, int N) {
    const Buffer& s = pl->buffers[op.src];
    const Buffer& t = pl->buffers[op.dst];
    DlwpConvDesc d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.Cin = op.src_c; d.H = s.d.H; d.W = s.d.W;
    d.Cout = op.Cout; d.kh = op.kh; d.kw = op.kw; d.dil_h = op.dil_h; d.dil_w = op.dil_w;
    d.pad_t = op.pad_t; d.pad_b = op.pad_b; d.pad_l = op.pad_l; d.pad_r = op.pad_r;
    d.pad_mode_h = op.pad_mode_h; d.pad_mode_w = op.pad_mode_w;
    d.act = op.act; d.pre_op = op.pre_op; d.rowwise = op.rowwise; d.impl = op.impl;
    d.row_begin = op.row_begin; d.row_end = op.row_end;
    d.x_stride_n = s.sample_elems(); d.x_stride_c = (long long)s.d.H * s.d.W; d.x_stride_h = s.d.W;
    d.y_stride_n = t.sample_elems(); d.y_stride_c = (long long)t.d.H * t.d.W; d.y_stride_h = t.d.W;
    return d;
}

// Decide whether the whole plan can run as a tensor-core chain and size its packed buffers.
//
// Chain = every activation lives in P layout (fp16 hi/lo planes, 8 channels per plane pair).  Convs run on the tcgen05
// kernels; MaxPooling2D(2), UpSampling2D(2) and the skip-connection copies of the U-Net run as data movers on P images;
// slice_layer / concatenate are plane windows (channel offsets must be multiples of 8).  Anything else (stand-alone
// padding, RowConnected2D, pool/upsample fused into a conv, a data mover that produces a model output, ...) keeps the
// whole plan on the fp32 kernels.
static int tc_setup(DlwpPlan* pl) {
    // Default: use the tensor-core chain whenever the whole plan is eligible.  math = 1 (or any conv op forced to an
    // FFMA/direct implementation) keeps the fp32 FFMA kernels.
    if (pl->opt.math == 1) return 0;
    for (const DlwpOpDesc& op : pl->ops)
        if (op.kind == DLWP_OP_CONV && op.impl != DLWP_IMPL_AUTO && op.impl != DLWP_IMPL_TC) return 0;
    const int nops = (int)pl->ops.size();
    const int nbuf = (int)pl->buffers.size();
    std::vector<TcLayer> layers(nops);
    std::vector<int> wpad(nbuf, -1);
    std::vector<char> is_src(nbuf, 0), written(nbuf, 0);
    auto chunk_window_ok = [](int c0, int c, int C) { return c0 % 8 == 0 && (c % 8 == 0 || c0 + c == C); };
    written[pl->input_buf] = 1;
    for (int i = 0; i < nops; ++i) {
        const DlwpOpDesc& op = pl->ops[i];
        const Buffer& s = pl->buffers[op.src];
        const Buffer& t = pl->buffers[op.dst];
        if (!written[op.src]) return 0;  // source must be the input or the result of an earlier op
        is_src[op.src] = 1;
        written[op.dst] = 1;
        if (op.kind == DLWP_OP_CONV) {
            if (op.src == pl->input_buf && (op.row_begin || op.row_end)) {
                const int lo = std::max(0, op.row_begin - op.pad_t);
                const int hi = std::min(s.d.H, op.row_end - op.pad_t + op.dil_h * (op.kh - 1));
                if (pl->tc_in_row1 == 0) { pl->tc_in_row0 = lo; pl->tc_in_row1 = hi; }
                else { pl->tc_in_row0 = std::min(pl->tc_in_row0, lo); pl->tc_in_row1 = std::max(pl->tc_in_row1, hi); }
            }
            if (!chunk_window_ok(op.src_c0, op.src_c, s.d.C) || !chunk_window_ok(op.dst_c0, op.Cout, t.d.C)) return 0;
            DlwpConvDesc d = conv_desc_of(pl, op, pl->max_batch);
            if (!tc_geometry_ok(d, pl->tc_opt) || tc_plan_layer(d, &layers[i], pl->tc_opt) != 0) return 0;
            if (wpad[op.src] >= 0 && wpad[op.src] != layers[i].wpad) return 0;  // all conv readers must want the same halo
            wpad[op.src] = layers[i].wpad;
        } else if (op.kind == DLWP_OP_MAXPOOL || op.kind == DLWP_OP_UPSAMPLE || op.kind == DLWP_OP_COPY) {
            if (op.src_c0 % 8 || op.src_c % 8 || op.dst_c0 % 8) return 0;
            if (op.kind == DLWP_OP_MAXPOOL && ((s.d.H | s.d.W) & 1)) return 0;
            // the data movers only write P images: a model output produced by one would never reach caller memory
            if (t.d.kind == DLWP_BUF_OUTPUT) return 0;
        } else {
            return 0;
        }
    }
    // conv -> conv fusion: op i's output is an internal buffer that only op i + 1 reads (whole), op i is the only reader
    // of its own source, and the pair matches an instantiated fused kernel
    if (pl->opt.fuse == 1 && !pl->tc_opt.bf16)
        for (int i = 0; i + 1 < nops && pl->pair_first < 0; ++i) {
            const DlwpOpDesc& a = pl->ops[i];
            const DlwpOpDesc& b = pl->ops[i + 1];
            if (a.kind != DLWP_OP_CONV || b.kind != DLWP_OP_CONV || b.src != a.dst) continue;
            const Buffer& mid = pl->buffers[a.dst];
            if (mid.d.kind != DLWP_BUF_INTERNAL || a.dst_c0 != 0 || a.Cout != mid.d.C || b.src_c0 != 0 || b.src_c != mid.d.C) continue;
            if (a.src_c0 != 0 || a.src_c != pl->buffers[a.src].d.C) continue;
            int mid_readers = 0, src_readers = 0, mid_writers = 0;
            for (const DlwpOpDesc& o : pl->ops) {
                mid_readers += o.src == a.dst;
                src_readers += o.src == a.src;
                mid_writers += o.dst == a.dst;
            }
            if (mid_readers != 1 || src_readers != 1 || mid_writers != 1) continue;
            DlwpConvDesc d1 = conv_desc_of(pl, a, pl->max_batch), d2 = conv_desc_of(pl, b, pl->max_batch);
            if (!tc_pair_ok(d1, layers[i], d2, layers[i + 1])) continue;
            pl->pair_first = i;
            layers[i].rowpitch = tc_pair_state_pitch(d1);       // the staged state row is 128 + halo1 pixels
            wpad[a.src] = layers[i].wpad + layers[i + 1].wpad;  // the state image carries both layers' halo columns
            is_src[a.dst] = 0;                                  // the intermediate never exists in HBM
        }
    for (int b = 0; b < nbuf; ++b)
        if (is_src[b]) {
            Buffer& B = pl->buffers[b];
            B.wpad = wpad[b] >= 0 ? wpad[b] : 0;  // read only by data movers: no halo needed
            B.planes = (pl->tc_opt.bf16 ? 1 : 2) * ((B.d.C + 7) / 8);
        }
    pl->tc_pdst.assign(nops, -1);
    for (int i = 0; i < nops; ++i)
        if (pl->buffers[pl->ops[i].dst].wpad >= 0) pl->tc_pdst[i] = pl->ops[i].dst;
    // feedback: the producer of the LAST output re-packs the next iteration's input when shapes allow -- and when nothing
    // from that op on still reads the input image (a single conv with Cin == Cout would overwrite rows and halo columns
    // that other CTAs of the same launch are reading)
    const int last_out = pl->outputs.back();
    const Buffer& in = pl->buffers[pl->input_buf];
    const Buffer& lo = pl->buffers[last_out];
    int last_writer = -1, n_writers = 0;
    for (int i = 0; i < nops; ++i)
        if (pl->ops[i].dst == last_out) { last_writer = i; ++n_writers; }
    bool input_read_late = false;
    for (int i = std::max(0, last_writer); i < nops; ++i)
        if (pl->ops[i].src == pl->input_buf) input_read_late = true;
    // (latitude bands: the conv re-packs its own band rows; the halo rows arrive as fp32 from the neighbours and are
    // packed after the exchange, see run_ops_tc)
    if (n_writers == 1 && !input_read_late && pl->ops[last_writer].kind == DLWP_OP_CONV && pl->tc_pdst[last_writer] < 0 &&
        pl->ops[last_writer].dst_c0 == 0 && pl->ops[last_writer].Cout == lo.d.C && lo.d.C == in.d.C && lo.d.H == in.d.H &&
        lo.d.W == in.d.W && in.wpad >= 0) {
        pl->tc_feedback_op = last_writer;
        pl->tc_pdst[last_writer] = pl->input_buf;
        pl->tc_band_row0 = pl->ops[last_writer].row_begin;
        pl->tc_band_row1 = pl->ops[last_writer].row_end;
    }
    // exponents: the input image and everything a linear / relu conv produces are dynamic (decided per production from
    // the measured amax of the producer's source); tanh outputs and data moved from them are static.  A dynamic image has
    // ONE exponent word, so it must have one producer.
    for (Buffer& b : pl->buffers) b.e_static = true;
    pl->buffers[pl->input_buf].e_static = false;
    for (bool changed = true; changed;) {
        changed = false;
        for (int i = 0; i < nops; ++i) {
            if (pl->tc_pdst[i] < 0 || pl->tc_pdst[i] == pl->input_buf) continue;
            const DlwpOpDesc& op = pl->ops[i];
            Buffer& t = pl->buffers[pl->tc_pdst[i]];
            const bool dyn = op.kind == DLWP_OP_CONV ? op.act != DLWP_ACT_TANH : !pl->buffers[op.src].e_static;
            if (dyn && t.e_static) { t.e_static = false; changed = true; }
        }
    }
    std::vector<int> producers(nbuf, 0);
    for (int i = 0; i < nops; ++i)
        if (pl->tc_pdst[i] >= 0 && pl->tc_pdst[i] != pl->input_buf) ++producers[pl->tc_pdst[i]];
    for (int b = 0; b < nbuf; ++b)
        if (!pl->tc_opt.bf16 && !pl->buffers[b].e_static && producers[b] > 1) return 0;   // (bf16 images carry no exponent)
    for (Buffer& b : pl->buffers)
        if (b.wpad >= 0) {
            const size_t bytes = tc_p_bytes(pl->max_batch, b.planes, b.d.H, b.d.W + 2 * b.wpad);
            if (cudaMalloc(&b.P, bytes) != cudaSuccess) return DLWP_ENOMEM;
            cudaMemset(b.P, 0, bytes);
        }
    if (cudaMalloc(&pl->d_counter, sizeof(int)) != cudaSuccess) return DLWP_ENOMEM;
    cudaMemset(pl->d_counter, 0, sizeof(int));
    if (cudaMalloc(&pl->d_exp, sizeof(int) * 2 * nbuf) != cudaSuccess) return DLWP_ENOMEM;
    if (cudaMalloc(&pl->d_amax, sizeof(float) * 2 * nbuf) != cudaSuccess) return DLWP_ENOMEM;
    cudaMemset(pl->d_exp, 0, sizeof(int) * 2 * nbuf);
    cudaMemset(pl->d_amax, 0, sizeof(float) * 2 * nbuf);
    pl->tc_layers = layers;
    pl->tc = true;
    return 0;
}

static int tc_pack_plan_weights(DlwpPlan* pl, int weight_id, const float* kernel_host, const float* bias_host) {
    for (size_t i = 0; i < pl->ops.size(); ++i) {
        const DlwpOpDesc& op = pl->ops[i];
        if (op.weight_id != weight_id) continue;
        Weight& w = pl->weights[weight_id];
        DlwpConvDesc d = conv_desc_of(pl, op, pl->max_batch);
        std::vector<__half> img;
        DLWP_REQUIRE(tc_pack_weights(d, pl->tc_layers[i], kernel_host, &img, w.kst, &w.ws) == 0,
                     DLWP_ESHAPE, "tensor-core weight packing failed for weight %d (non-finite weights?)", weight_id);
        w.bmax = 0.f;
        if (bias_host)
            for (long long k = 0; k < w.b_elems; ++k) w.bmax = std::max(w.bmax, fabsf(bias_host[k]));
        if (!w.bimg) DLWP_CUDA_TRY(cudaMalloc(&w.bimg, img.size() * sizeof(__half)));
        DLWP_CUDA_TRY(cudaMemcpy(w.bimg, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
        return 0;  // ops sharing a weight id share geometry and therefore the image
    }
    return 0;
}

// Scale words of one op at iteration t (conv_tc.h).  An image produced at iteration t is read at iteration t, except
// the input image: the feedback conv of iteration t produces the input of iteration t + 1 (production index t + 1).
static inline int PPC(const DlwpPlan* pl) { return pl->tc_opt.bf16 ? 1 : 2; }   // planes per 8-channel chunk

static TcScale scale_of(const DlwpPlan* pl, int i, int t) {
    const DlwpOpDesc& op = pl->ops[i];
    TcScale sc;
    if (pl->tc_opt.bf16) {   // plain bf16: no exponents, nothing measured
        sc.e_in_const = 0;
        sc.e_out_const = 0;
        return sc;
    }
    const Buffer& s = pl->buffers[op.src];
    sc.e_in_const = TC_EXP_STATIC;
    sc.amax_chk = pl->d_amax + 2 * op.src + (t & 1);
    if (!s.e_static) {  // a static (tanh) source bounds the output with |x| <= 1: exponents independent of the data
        sc.e_in = pl->d_exp + 2 * op.src + (t & 1);
        sc.amax_in = sc.amax_chk;
    }
    const int pd = pl->tc_pdst[i];
    if (pd >= 0) {
        const int prod = pd == pl->input_buf ? t + 1 : t;
        if (!pl->buffers[pd].e_static) sc.e_out = pl->d_exp + 2 * pd + (prod & 1);
        sc.amax_out = pl->d_amax + 2 * pd + (prod & 1);
        sc.amax_zero = pl->d_amax + 2 * pd + ((prod + 1) & 1);
    }
    if (op.kind == DLWP_OP_CONV) {
        const Weight& w = pl->weights[op.weight_id];
        sc.e_w = w.ws.e_w; sc.l1max = w.ws.l1max; sc.bmax = w.has_bias ? w.bmax : 0.f;
    }
    return sc;
}

// ops pair_first, pair_first + 1 as one kernel
static int run_pair_tc(DlwpPlan* pl, int N, cudaStream_t stream, int t, bool feedback_write) {
    const int i = pl->pair_first;
    const DlwpOpDesc& a = pl->ops[i];
    const DlwpOpDesc& b = pl->ops[i + 1];
    const Buffer& s = pl->buffers[a.src];
    const Buffer& out = pl->buffers[b.dst];
    const Weight& w1 = pl->weights[a.weight_id];
    const Weight& w2 = pl->weights[b.weight_id];
    DLWP_REQUIRE(w1.set && w1.bimg && w2.set && w2.bimg, DLWP_ESTATE, "weights of the fused pair were never set");
    int pd = pl->tc_pdst[i + 1];
    if (pd == pl->input_buf && i + 1 == pl->tc_feedback_op && !feedback_write) pd = -1;
    TcScale sc1 = scale_of(pl, i, t), sc2 = scale_of(pl, i + 1, t);
    sc1.e_out = nullptr;                                   // tanh intermediate: static exponent
    sc1.amax_out = pl->d_amax + 2 * a.dst + (t & 1);       // measured for the underflow check the last CTA makes
    sc1.amax_zero = pl->d_amax + 2 * a.dst + ((t + 1) & 1);
    sc2.amax_chk = nullptr;
    if (pd < 0) { sc2.e_out = nullptr; sc2.amax_out = nullptr; sc2.amax_zero = nullptr; }
    DlwpConvDesc d1 = conv_desc_of(pl, a, N), d2 = conv_desc_of(pl, b, N);
    float* y32 = (out.d.kind == DLWP_BUF_OUTPUT) ? out.ptr + (long long)b.dst_c0 * out.d.H * out.d.W : nullptr;
    __half* yp = nullptr;
    int wpad_out = 0, planes_out = 0, out_plane0 = 0;
    if (pd >= 0) {
        const Buffer& pb = pl->buffers[pd];
        yp = pb.P; wpad_out = pb.wpad; planes_out = pb.planes;
        out_plane0 = pd == b.dst ? PPC(pl) * (b.dst_c0 / 8) : 0;
    }
    return tc_pair_launch(d1, pl->tc_layers[i], w1.kst, w1.bimg, w1.has_bias ? w1.b : nullptr, d2, pl->tc_layers[i + 1],
                          w2.kst, w2.bimg, w2.has_bias ? w2.b : nullptr, s.P, s.planes, PPC(pl) * (a.src_c0 / 8), y32, yp,
                          wpad_out, planes_out, out_plane0, sc1, sc2, pl->d_counter, stream, pl->tc_opt);
}

static int run_one_tc(DlwpPlan* pl, int i, int N, cudaStream_t stream, int t, bool feedback_write) {
    if (pl->pair_first >= 0) {
        if (i == pl->pair_first) return 0;                 // runs with its partner
        if (i == pl->pair_first + 1) return run_pair_tc(pl, N, stream, t, feedback_write);
    }
    const DlwpOpDesc& op = pl->ops[i];
    const Buffer& s = pl->buffers[op.src];
    const Buffer& b = pl->buffers[op.dst];
    int pd = pl->tc_pdst[i];
    if (pd == pl->input_buf && i == pl->tc_feedback_op && !feedback_write) pd = -1;  // plain forward: nobody reads it
    TcScale sc = scale_of(pl, i, t);
    if (pd < 0) { sc.e_out = nullptr; sc.amax_out = nullptr; sc.amax_zero = nullptr; }
    if (op.kind != DLWP_OP_CONV) {  // data mover on P images
        if (pd < 0) return 0;  // nobody reads the result in P layout
        const Buffer& pb = pl->buffers[pd];
        return tc_ew_launch(op.kind, s.P, pb.P, N, PPC(pl) * (op.src_c / 8), s.d.H, s.d.W, s.wpad, PPC(pl) * (op.src_c0 / 8), s.planes,
                            pb.wpad, PPC(pl) * (op.dst_c0 / 8), pb.planes, stream, op.row_begin, op.row_end, sc,
                            pl->tc_opt.bf16);
    }
    const Weight& w = pl->weights[op.weight_id];
    DLWP_REQUIRE(w.set && w.bimg, DLWP_ESTATE, "weights %d were never set", op.weight_id);
    DlwpConvDesc d = conv_desc_of(pl, op, N);
    float* y32 = (b.d.kind == DLWP_BUF_OUTPUT) ? b.ptr + (long long)op.dst_c0 * b.d.H * b.d.W : nullptr;
    __half* yp = nullptr;
    int wpad_out = 0, planes_out = 0;
    TcWindow win;
    win.in_plane0 = PPC(pl) * (op.src_c0 / 8);
    win.in_planes_total = s.planes;
    TcPeers peers;
    if (pd >= 0) {
        const Buffer& pb = pl->buffers[pd];
        yp = pb.P; wpad_out = pb.wpad; planes_out = pb.planes;
        win.out_plane0 = pd == op.dst ? PPC(pl) * (op.dst_c0 / 8) : 0;
        if (pl->p2p_active && pd == pl->input_buf) {   // double-buffered input image + the neighbours' halo rows
            const int nxt = (pl->cur_parity + 1) & 1;
            yp = pl->P_in[nxt];
            if (pl->p2p_send) {
                peers.up = pl->peer_P[0][nxt];
                peers.down = pl->peer_P[1][nxt];
                peers.up_end = pl->p2p_up_end;
                peers.down_begin = pl->p2p_down_begin;
            }
        }
    }
    return tc_launch(d, pl->tc_layers[i], w.kst, s.P, w.bimg, w.has_bias ? w.b : nullptr, y32, yp, wpad_out,
                     planes_out, stream, win, sc, pl->tc_opt, peers);
}

static int amax_allreduce(void* comm, float* amax, cudaStream_t stream);   // (NCCL: defined with the lat-band exchange)

// One application of the chain at iteration t.  input_is_packed: the feedback conv of iteration t - 1 already wrote the
// input image (rollout).  feedback_write: let the feedback conv write the next input image (off for a plain forward).
static int run_ops_tc(DlwpPlan* pl, int N, cudaStream_t stream, int t, bool input_is_packed, bool feedback_write,
                      bool input_rows_all_valid) {
    Buffer& in = pl->buffers[pl->input_buf];
    pl->tc_last_t = t;
    TcPackScale ps;
    ps.e = pl->d_exp + 2 * pl->input_buf + (t & 1);
    ps.amax = pl->d_amax + 2 * pl->input_buf + (t & 1);
    ps.bf16 = pl->tc_opt.bf16;
    if (!input_is_packed) {
        ps.fresh = 1;
        ps.amax_zero = pl->d_amax + 2 * pl->input_buf + ((t + 1) & 1);
        // The exponent comes from max|x| over the whole state when every row is known to be valid (x0 of a rollout: the
        // same value on every rank of a latitude-band run), else over the rows this plan reads (a band's series slot
        // holds garbage outside band + halo).
        if (!input_rows_all_valid || pl->x0_amax_comm) { ps.amax_row0 = pl->tc_in_row0; ps.amax_row1 = pl->tc_in_row1; }
        if (pl->x0_amax_comm && t == 0) {   // x0 holds this band's rows only: the ranks agree on max|x0| before packing
            ps.after_amax = amax_allreduce;
            ps.after_amax_ctx = pl->x0_amax_comm;
        }
        int rc = tc_pack_state(in.ptr, in.P, N, in.d.C, in.d.H, in.d.W, in.wpad, in.sample_elems(),
                               (long long)in.d.H * in.d.W, in.d.W, stream, pl->tc_in_row0, pl->tc_in_row1, ps);
        if (rc) return rc;
    } else if (pl->halo_in_p) {
        pl->halo_in_p = false;          // dlwp_rollout_latband's exchange stored the received halo rows in the image already
    } else if (pl->tc_band_row1 > 0) {  // latitude band: pack the halo rows received from the neighbours
        ps.fresh = 0;
        const int lo[2] = {pl->tc_in_row0, pl->tc_band_row1}, hi[2] = {pl->tc_band_row0, pl->tc_in_row1};
        for (int k = 0; k < 2; ++k)
            if (hi[k] > lo[k]) {
                int rc = tc_pack_state(in.ptr, in.P, N, in.d.C, in.d.H, in.d.W, in.wpad, in.sample_elems(),
                                       (long long)in.d.H * in.d.W, in.d.W, stream, lo[k], hi[k], ps);
                if (rc) return rc;
            }
    }
    for (size_t i = 0; i < pl->ops.size(); ++i) {
        int rc = run_one_tc(pl, (int)i, N, stream, t, feedback_write);
        if (rc) return rc;
    }
    return 0;
}

static int run_one(DlwpPlan* pl, size_t i, int N, cudaStream_t stream) {
    {
        const DlwpOpDesc& op = pl->ops[i];
        const Buffer& s = pl->buffers[op.src];
        const Buffer& t = pl->buffers[op.dst];
        const long long s_hw = (long long)s.d.H * s.d.W, t_hw = (long long)t.d.H * t.d.W;
        const float* x = s.ptr + op.src_c0 * s_hw;
        float* y = t.ptr + op.dst_c0 * t_hw;
        const long long xs_n = s.sample_elems(), xs_c = s_hw, xs_h = s.d.W;
        const long long ys_n = t.sample_elems(), ys_c = t_hw, ys_h = t.d.W;
        int rc = 0;
        switch (op.kind) {
            case DLWP_OP_CONV: {
                const Weight& w = pl->weights[op.weight_id];
                DLWP_REQUIRE(w.set, DLWP_ESTATE, "weights %d were never set", op.weight_id);
                DlwpConvDesc d;
                memset(&d, 0, sizeof(d));
                d.N = N; d.Cin = op.src_c; d.H = s.d.H; d.W = s.d.W;
                d.Cout = op.Cout; d.kh = op.kh; d.kw = op.kw; d.dil_h = op.dil_h; d.dil_w = op.dil_w;
                d.pad_t = op.pad_t; d.pad_b = op.pad_b; d.pad_l = op.pad_l; d.pad_r = op.pad_r;
                d.pad_mode_h = op.pad_mode_h; d.pad_mode_w = op.pad_mode_w;
                d.act = op.act; d.pre_op = op.pre_op; d.rowwise = op.rowwise; d.impl = op.impl;
                d.row_begin = op.row_begin; d.row_end = op.row_end;
                d.x_stride_n = xs_n; d.x_stride_c = xs_c; d.x_stride_h = xs_h;
                d.y_stride_n = ys_n; d.y_stride_c = ys_c; d.y_stride_h = ys_h;
                rc = conv2d_fwd(d, x, w.k, w.has_bias ? w.b : nullptr, y, stream);
                break;
            }
            case DLWP_OP_PAD:
            case DLWP_OP_MAXPOOL:
            case DLWP_OP_UPSAMPLE:
            case DLWP_OP_COPY:
                rc = dlwp_rows_op(op.kind, x, y, N, op.src_c, s.d.H, s.d.W, op.pad_t, op.pad_b, op.pad_l, op.pad_r,
                                  op.pad_mode_h, op.pad_mode_w, xs_n, xs_c, xs_h, ys_n, ys_c, ys_h, op.row_begin,
                                  op.row_end, stream);
                break;
            case DLWP_OP_TO_NCHW:
            case DLWP_OP_TO_NHWC:
                DLWP_REQUIRE(op.src_c0 == 0 && op.dst_c0 == 0 && op.src_c == s.d.C && s.d.C == t.d.C, DLWP_ESHAPE,
                             "layout ops move whole buffers");
                rc = dlwp_layout2d(x, y, N, s.d.C, s.d.H, s.d.W, op.kind == DLWP_OP_TO_NCHW ? 1 : 0, stream);
                break;
            case DLWP_OP_LSTM: {
                const Buffer& cbuf = pl->buffers[op.aux];
                float* c = cbuf.ptr + op.aux_c0 * t_hw;
                const bool first = op.src_c == 4 * op.Cout;   // h_-1 = 0, c_-1 = 0: no recurrent half
                rc = dlwp_convlstm_gates(x, first ? nullptr : x + 4LL * op.Cout * s_hw, first ? nullptr : c, c, y, N,
                                         op.Cout, t.d.H, t.d.W, xs_n, xs_n, cbuf.sample_elems(), ys_n, op.act, op.act2,
                                         op.row_begin, op.row_end, stream);
                break;
            }
            default: DLWP_REQUIRE(false, DLWP_EINVAL, "op %zu: unknown kind %d", i, op.kind);
        }
        return rc;
    }
}

static int run_ops(DlwpPlan* pl, int N, cudaStream_t stream) {
    for (size_t i = 0; i < pl->ops.size(); ++i) {
        int rc = run_one(pl, i, N, stream);
        if (rc) return rc;
    }
    return 0;
}

static int check_rollout_shapes(const DlwpPlan* pl) {
    const Buffer& in = pl->buffers[pl->input_buf];
    for (int o : pl->outputs) {
        const Buffer& b = pl->buffers[o];
        DLWP_REQUIRE(b.d.C == in.d.C && b.d.H == in.d.H && b.d.W == in.d.W, DLWP_ESHAPE,
                     "rollout needs output shape == input shape, got (%d,%d,%d) vs (%d,%d,%d)", b.d.C, b.d.H, b.d.W,
                     in.d.C, in.d.H, in.d.W);
    }
    return 0;
}

// iterations [t0, t1) of the feedback loop, launched on `stream`
static int rollout_range(DlwpPlan* pl, int N, const float* x0, float* series, int t0, int t1, cudaStream_t stream) {
    const long long slot = (long long)N * pl->buffers[pl->input_buf].sample_elems();
    const int n_out = (int)pl->outputs.size();
    if (pl->tc && t0 == 0)  // a new rollout: every image's amax slots start from zero
        DLWP_CUDA_TRY(cudaMemsetAsync(pl->d_amax, 0, sizeof(float) * 2 * pl->buffers.size(), stream));
    for (int t = t0; t < t1; ++t) {
        pl->buffers[pl->input_buf].ptr =
            const_cast<float*>(t == 0 ? x0 : series + ((long long)t * n_out - 1) * slot);
        for (int k = 0; k < n_out; ++k) pl->buffers[pl->outputs[k]].ptr = series + ((long long)t * n_out + k) * slot;
        int rc = pl->tc ? run_ops_tc(pl, N, stream, t, t > 0 && pl->tc_feedback_op >= 0, true, t == 0)
                        : run_ops(pl, N, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace dlwp

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" const char* dlwp_last_error_string(void) { return t_error.c_str(); }
extern "C" int dlwp_abi_version(void) { return DLWP_B200_ABI_VERSION; }
extern "C" int64_t dlwp_kernel_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" const char* dlwp_conv2d_impl_name(const DlwpConvDesc* desc) {
    if (!desc) return "invalid";
    return conv2d_impl_name(*desc);
}

extern "C" int dlwp_conv2d_fwd(const DlwpConvDesc* desc, const float* x, const float* w, const float* bias, float* y,
                               dlwp_stream_t stream) {
    DLWP_REQUIRE(desc != nullptr, DLWP_EINVAL, "null descriptor");
    return conv2d_fwd(*desc, x, w, bias, y, (cudaStream_t)stream);
}

extern "C" int dlwp_conv2d_bwd_input(const DlwpConvDesc* desc, const float* dy, const float* w, float* dx,
                                     dlwp_stream_t stream) {
    DLWP_REQUIRE(desc && dy && w && dx, DLWP_EINVAL, "null argument");
    int rc = check_device();
    if (rc) return rc;
    return conv2d_bwd_input(*desc, dy, w, dx, (cudaStream_t)stream);
}

extern "C" int dlwp_conv2d_bwd_weight(const DlwpConvDesc* desc, const float* x, const float* dy, float* dw, float* db,
                                      dlwp_stream_t stream) {
    DLWP_REQUIRE(desc && x && dy && dw, DLWP_EINVAL, "null argument");
    int rc = check_device();
    if (rc) return rc;
    return conv2d_bwd_weight(*desc, x, dy, dw, db, (cudaStream_t)stream);
}

extern "C" int dlwp_plan_create(const DlwpNetDesc* net, DlwpPlan** out) {
    return dlwp_plan_create_opts(net, nullptr, out);
}

extern "C" int dlwp_plan_create_opts(const DlwpNetDesc* net, const DlwpPlanOptions* opts, DlwpPlan** out) {
    DLWP_REQUIRE(net && out, DLWP_EINVAL, "null argument");
    *out = nullptr;
    DLWP_REQUIRE(net->n_buffers > 0 && net->n_ops > 0 && net->max_batch > 0 && net->buffers && net->ops,
                 DLWP_EINVAL, "empty network description");
    int rc = check_device();
    if (rc) return rc;
    DlwpPlan* pl = new DlwpPlan();
    memset(&pl->opt, 0, sizeof(pl->opt));
    pl->opt.tc_taps_in_k = -1;
    if (opts) pl->opt = *opts;
    pl->tc_opt.generic = pl->opt.tc_generic;
    pl->tc_opt.bands = pl->opt.tc_bands;
    pl->tc_opt.no_tma = pl->opt.tc_no_tma;
    pl->tc_opt.taps_in_k = pl->opt.tc_taps_in_k;
    pl->tc_opt.debug = pl->opt.tc_debug;
    pl->tc_opt.bf16 = pl->opt.precision == 1 ? 1 : 0;
    pl->max_batch = net->max_batch;
    pl->buffers.resize(net->n_buffers);
    int n_outputs = 0;
    for (int i = 0; i < net->n_buffers; ++i) {
        pl->buffers[i].d = net->buffers[i];
        const DlwpBufferDesc& b = net->buffers[i];
        if (b.C <= 0 || b.H <= 0 || b.W <= 0) {
            delete pl;
            DLWP_REQUIRE(false, DLWP_ESHAPE, "buffer %d has non-positive dims", i);
        }
        if (b.kind == DLWP_BUF_INPUT) {
            if (pl->input_buf >= 0) {
                delete pl;
                DLWP_REQUIRE(false, DLWP_EINVAL, "more than one input buffer");
            }
            pl->input_buf = i;
        }
        if (b.kind == DLWP_BUF_OUTPUT) n_outputs = std::max(n_outputs, b.output_index + 1);
    }
    if (pl->input_buf < 0 || n_outputs == 0) {
        delete pl;
        DLWP_REQUIRE(false, DLWP_EINVAL, "network needs one input and at least one output buffer");
    }
    pl->outputs.assign(n_outputs, -1);
    for (int i = 0; i < net->n_buffers; ++i)
        if (net->buffers[i].kind == DLWP_BUF_OUTPUT) pl->outputs[net->buffers[i].output_index] = i;
    for (int o : pl->outputs)
        if (o < 0) {
            delete pl;
            DLWP_REQUIRE(false, DLWP_EINVAL, "output indices are not contiguous");
        }
    pl->weights.resize(std::max(0, net->n_weights));
    pl->ops.assign(net->ops, net->ops + net->n_ops);
    // validate ops and size the weight table
    for (int i = 0; i < net->n_ops; ++i) {
        const DlwpOpDesc& op = pl->ops[i];
        bool ok = op.src >= 0 && op.src < net->n_buffers && op.dst >= 0 && op.dst < net->n_buffers && op.src_c > 0 &&
                  op.src_c0 >= 0 && op.src_c0 + op.src_c <= pl->buffers[op.src].d.C && op.dst_c0 >= 0;
        int C = 0, H = 0, W = 0;
        if (ok) {
            out_dims(pl, op, C, H, W);
            const Buffer& t = pl->buffers[op.dst];
            ok = H == t.d.H && W == t.d.W && op.dst_c0 + C <= t.d.C;
        }
        if (ok && op.kind == DLWP_OP_CONV) ok = op.weight_id >= 0 && op.weight_id < net->n_weights;
        if (ok && op.kind == DLWP_OP_LSTM)
            ok = op.Cout > 0 && (op.src_c == 4 * op.Cout || op.src_c == 8 * op.Cout) && op.aux >= 0 &&
                 op.aux < net->n_buffers && op.aux_c0 >= 0 && op.aux_c0 + op.Cout <= pl->buffers[op.aux].d.C &&
                 pl->buffers[op.aux].d.H == H && pl->buffers[op.aux].d.W == W &&
                 pl->buffers[op.aux].d.kind == DLWP_BUF_INTERNAL;
        if (!ok) {
            delete pl;
            DLWP_REQUIRE(false, DLWP_ESHAPE, "op %d is inconsistent with its buffers (computed out dims %d,%d,%d)", i,
                         C, H, W);
        }
        if (op.kind == DLWP_OP_CONV) {
            Weight& w = pl->weights[op.weight_id];
            long long ke = (long long)op.kh * op.kw * op.src_c * op.Cout, be = op.Cout;
            if (op.rowwise) { ke *= H; be *= H; }
            if (w.k_elems == 0) { w.k_elems = ke; w.b_elems = be; }
            if (w.k_elems != ke) {
                delete pl;
                DLWP_REQUIRE(false, DLWP_ESHAPE, "weight %d is shared by layers of different shape", op.weight_id);
            }
        }
    }
    for (Buffer& b : pl->buffers)
        if (b.d.kind == DLWP_BUF_INTERNAL) {
            cudaError_t e = cudaMalloc(&b.ptr, sizeof(float) * b.sample_elems() * pl->max_batch);
            if (e != cudaSuccess) {
                dlwp_plan_destroy(pl);
                DLWP_REQUIRE(false, DLWP_ENOMEM, "cudaMalloc of an activation buffer failed: %s",
                             cudaGetErrorString(e));
            }
        }
    for (Weight& w : pl->weights) {
        if (w.k_elems == 0) continue;
        if (cudaMalloc(&w.k, sizeof(float) * w.k_elems) != cudaSuccess ||
            cudaMalloc(&w.b, sizeof(float) * w.b_elems) != cudaSuccess) {
            dlwp_plan_destroy(pl);
            DLWP_REQUIRE(false, DLWP_ENOMEM, "cudaMalloc of a weight tensor failed");
        }
    }
    rc = tc_setup(pl);
    if (rc) {
        dlwp_plan_destroy(pl);
        DLWP_REQUIRE(false, rc, "tensor-core chain setup failed");
    }
    *out = pl;
    return 0;
}

extern "C" void dlwp_plan_destroy(DlwpPlan* pl) {
    if (!pl) return;
    for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& kv : pl->band_graphs) cudaGraphExecDestroy(kv.second);
    for (float* h : pl->halo_stage)
        if (h) cudaFree(h);
    for (float* g : pl->gbuf)
        if (g) cudaFree(g);
    for (float* g : pl->out_store)
        if (g) cudaFree(g);
    if (pl->acc_stats) cudaFree(pl->acc_stats);
    for (float* g : {pl->flat_g, pl->flat_m, pl->flat_v, pl->stats, pl->loss_wmap, pl->bwd_scratch, pl->wt_scratch, pl->acc_mean})
        if (g) cudaFree(g);
    for (Buffer& b : pl->buffers)
        if (b.d.kind == DLWP_BUF_INTERNAL && b.ptr) cudaFree(b.ptr);
    for (Buffer& b : pl->buffers)
        if (b.P) cudaFree(b.P);
    for (float* p : pl->seq_state)
        if (p) cudaFree(p);
    for (void* p : pl->ipc_opened) cudaIpcCloseMemHandle(p);
    if (pl->p2p) {   // P_in[0] is the input buffer's own image (freed with the buffers)
        if (pl->P_in[1]) cudaFree(pl->P_in[1]);
        pl->buffers[pl->input_buf].P = pl->P_in[0];
    }
    if (pl->flags) cudaFree(pl->flags);
    if (pl->d_counter) cudaFree(pl->d_counter);
    if (pl->d_exp) cudaFree(pl->d_exp);
    if (pl->d_amax) cudaFree(pl->d_amax);
    for (Weight& w : pl->weights) {
        if (w.bimg) cudaFree(w.bimg);
        if (w.k) cudaFree(w.k);
        if (w.b) cudaFree(w.b);
    }
    if (pl->d_x0) cudaFree(pl->d_x0);
    if (pl->d_series) cudaFree(pl->d_series);
    for (cudaEvent_t e : pl->events) cudaEventDestroy(e);
    if (pl->s_compute) cudaStreamDestroy(pl->s_compute);
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
    if (pl->s_d2h) cudaStreamDestroy(pl->s_d2h);
    for (cudaEvent_t e : pl->d2h_events) cudaEventDestroy(e);
    delete pl;
}

extern "C" int dlwp_plan_set_weights(DlwpPlan* pl, int32_t id, const float* kernel, int64_t kernel_elems,
                                     const float* bias, int64_t bias_elems) {
    DLWP_REQUIRE(pl && kernel, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(id >= 0 && id < (int)pl->weights.size(), DLWP_EINVAL, "weight id %d out of range", id);
    Weight& w = pl->weights[id];
    DLWP_REQUIRE(kernel_elems == w.k_elems, DLWP_ESHAPE, "kernel %d has %lld elements, expected %lld", id,
                 (long long)kernel_elems, w.k_elems);
    DLWP_REQUIRE(!bias || bias_elems == w.b_elems, DLWP_ESHAPE, "bias %d has %lld elements, expected %lld", id,
                 (long long)bias_elems, w.b_elems);
    // weights may be in use by queued work: drain first (set_weights is not on the hot path)
    DLWP_CUDA_TRY(cudaDeviceSynchronize());
    DLWP_CUDA_TRY(cudaMemcpy(w.k, kernel, sizeof(float) * w.k_elems, cudaMemcpyHostToDevice));
    if (bias) DLWP_CUDA_TRY(cudaMemcpy(w.b, bias, sizeof(float) * w.b_elems, cudaMemcpyHostToDevice));
    w.has_bias = bias != nullptr;
    w.set = true;
    if (pl->tc) return tc_pack_plan_weights(pl, id, kernel, bias);
    return 0;
}

extern "C" int dlwp_plan_get_weights(DlwpPlan* pl, int32_t id, float* kernel, int64_t kernel_elems, float* bias,
                                     int64_t bias_elems) {
    DLWP_REQUIRE(pl && kernel, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(id >= 0 && id < (int)pl->weights.size(), DLWP_EINVAL, "weight id %d out of range", id);
    Weight& w = pl->weights[id];
    DLWP_REQUIRE(w.set, DLWP_ESTATE, "weights %d were never set", id);
    DLWP_REQUIRE(kernel_elems == w.k_elems && (!bias || bias_elems == w.b_elems), DLWP_ESHAPE, "size mismatch");
    DLWP_CUDA_TRY(cudaDeviceSynchronize());
    DLWP_CUDA_TRY(cudaMemcpy(kernel, w.k, sizeof(float) * w.k_elems, cudaMemcpyDeviceToHost));
    if (bias && w.has_bias) DLWP_CUDA_TRY(cudaMemcpy(bias, w.b, sizeof(float) * w.b_elems, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int dlwp_plan_forward(DlwpPlan* pl, int32_t N, const float* x, float* const* outputs,
                                 dlwp_stream_t stream) {
    DLWP_REQUIRE(pl && x && outputs, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch, DLWP_ESHAPE, "batch %d outside (0, %d]", N, pl->max_batch);
    pl->buffers[pl->input_buf].ptr = const_cast<float*>(x);
    for (size_t k = 0; k < pl->outputs.size(); ++k) {
        DLWP_REQUIRE(outputs[k] != nullptr, DLWP_EINVAL, "output %zu is null", k);
        pl->buffers[pl->outputs[k]].ptr = outputs[k];
    }
    if (!pl->tc) return run_ops(pl, N, (cudaStream_t)stream);
    DLWP_CUDA_TRY(cudaMemsetAsync(pl->d_amax, 0, sizeof(float) * 2 * pl->buffers.size(), (cudaStream_t)stream));
    return run_ops_tc(pl, N, (cudaStream_t)stream, 0, false, false, false);
}

// predict_timeseries(step_sequence=True), DLWP/model/models.py:280-290: the model advances ONE time slice per application.
// The state holds time_dim slices of C / time_dim channels; after application t the next input is the old input without
// its first slice, followed by the first slice of the prediction; series slot t keeps the whole prediction.
extern "C" int dlwp_rollout_step_sequence(DlwpPlan* pl, int32_t N, const float* x0, float* series, int32_t iterations,
                                          int32_t time_dim, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && x0 && series, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch && iterations > 0, DLWP_ESHAPE, "bad batch / iterations");
    DLWP_REQUIRE(pl->outputs.size() == 1, DLWP_ESHAPE, "step_sequence needs a single-output model");
    int rc = check_rollout_shapes(pl);
    if (rc) return rc;
    Buffer& in = pl->buffers[pl->input_buf];
    const int C = in.d.C, H = in.d.H, W = in.d.W;
    DLWP_REQUIRE(time_dim >= 1 && C % time_dim == 0, DLWP_ESHAPE, "channels %d are not time_dim %d slices", C, time_dim);
    if (time_dim == 1) return dlwp_rollout(pl, N, x0, series, iterations, 0, stream_);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long slot = (long long)N * in.sample_elems(), hw = (long long)H * W, sn = in.sample_elems();
    if (pl->seq_cap < slot) {
        for (float*& p : pl->seq_state) {
            if (p) cudaFree(p);
            p = nullptr;
            DLWP_CUDA_TRY(cudaMalloc(&p, sizeof(float) * slot));
        }
        pl->seq_cap = slot;
    }
    const int Cs = C / time_dim;
    if (pl->tc) DLWP_CUDA_TRY(cudaMemsetAsync(pl->d_amax, 0, sizeof(float) * 2 * pl->buffers.size(), stream));
    for (int t = 0; t < iterations; ++t) {
        const float* cur = t == 0 ? x0 : pl->seq_state[t & 1];
        float* pred = series + (long long)t * slot;
        in.ptr = const_cast<float*>(cur);
        pl->buffers[pl->outputs[0]].ptr = pred;
        rc = pl->tc ? run_ops_tc(pl, N, stream, t, false, false, true) : run_ops(pl, N, stream);
        if (rc) return rc;
        if (t + 1 == iterations) break;
        float* nxt = pl->seq_state[(t + 1) & 1];
        rc = dlwp_copy4d(cur + (long long)Cs * hw, nxt, N, C - Cs, H, W, sn, hw, W, sn, hw, W, stream_);
        if (!rc) rc = dlwp_copy4d(pred, nxt + (long long)(C - Cs) * hw, N, Cs, H, W, sn, hw, W, sn, hw, W, stream_);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int dlwp_rollout(DlwpPlan* pl, int32_t N, const float* x0, float* series, int32_t iterations,
                            int32_t use_graph, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && x0 && series, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch, DLWP_ESHAPE, "batch %d outside (0, %d]", N, pl->max_batch);
    DLWP_REQUIRE(iterations > 0, DLWP_EINVAL, "iterations must be > 0");
    int rc = check_rollout_shapes(pl);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!use_graph) return rollout_range(pl, N, x0, series, 0, iterations, stream);

    GraphKey key{N, iterations, x0, series};
    auto it = pl->graphs.find(key);
    if (it == pl->graphs.end()) {
        if (!pl->s_compute) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_compute, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        DLWP_CUDA_TRY(cudaStreamBeginCapture(pl->s_compute, cudaStreamCaptureModeThreadLocal));
        const long long before = g_launches.load();
        rc = rollout_range(pl, N, x0, series, 0, iterations, pl->s_compute);
        cudaError_t e = cudaStreamEndCapture(pl->s_compute, &graph);
        g_launches.store(before);  // capture does not execute anything
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        DLWP_CUDA_TRY(e);
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        DLWP_CUDA_TRY(e);
        if (pl->graphs.size() >= 8) {  // bounded cache
            cudaGraphExecDestroy(pl->graphs.begin()->second);
            pl->graphs.erase(pl->graphs.begin());
        }
        it = pl->graphs.emplace(key, exec).first;
    }
    DLWP_CUDA_TRY(cudaGraphLaunch(it->second, stream));
    int convs = 0;
    for (const DlwpOpDesc& op : pl->ops) (void)op, ++convs;
    g_launches.fetch_add((long long)convs * iterations);
    return 0;
}

namespace dlwp {
// device copies of the host-side x0 and series of dlwp_rollout_host / dlwp_rollout_latband_host (grown on demand)
static int ensure_host_staging(DlwpPlan* pl, long long slot, long long need) {
    if (pl->d_x0_cap < slot) {
        if (pl->d_x0) cudaFree(pl->d_x0);
        pl->d_x0 = nullptr; pl->d_x0_cap = 0;
        DLWP_CUDA_TRY(cudaMalloc(&pl->d_x0, sizeof(float) * slot));
        DLWP_CUDA_TRY(cudaMemset(pl->d_x0, 0, sizeof(float) * slot));   // (a band uploads only its rows; the rest stays finite)
        pl->d_x0_cap = slot;
    }
    if (pl->d_series_cap < need) {
        if (pl->d_series) cudaFree(pl->d_series);
        pl->d_series = nullptr; pl->d_series_cap = 0;
        cudaError_t e = cudaMalloc(&pl->d_series, sizeof(float) * need);
        DLWP_REQUIRE(e == cudaSuccess, DLWP_ENOMEM, "cudaMalloc of the %lld-byte device series failed: %s",
                     (long long)(sizeof(float) * need), cudaGetErrorString(e));
        pl->d_series_cap = need;
    }
    return 0;
}
}  // namespace dlwp

extern "C" int dlwp_rollout_host(DlwpPlan* pl, int32_t N, const float* x0_host, float* series_host,
                                 int32_t iterations, int32_t d2h_group) {
    DLWP_REQUIRE(pl && x0_host && series_host, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch, DLWP_ESHAPE, "batch %d outside (0, %d]", N, pl->max_batch);
    DLWP_REQUIRE(iterations > 0, DLWP_EINVAL, "iterations must be > 0");
    int rc = check_rollout_shapes(pl);
    if (rc) return rc;
    const long long slot = (long long)N * pl->buffers[pl->input_buf].sample_elems();
    const int n_out = (int)pl->outputs.size();
    const long long need = slot * n_out * iterations;
    if (!pl->s_compute) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_compute, cudaStreamNonBlocking));
    if (!pl->s_copy) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_copy, cudaStreamNonBlocking));
    rc = ensure_host_staging(pl, slot, need);
    if (rc) return rc;
    if (d2h_group <= 0) d2h_group = std::max(1, iterations / 8);
    const int groups = (iterations + d2h_group - 1) / d2h_group;
    while ((int)pl->events.size() < groups) {
        cudaEvent_t ev;
        DLWP_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        pl->events.push_back(ev);
    }
    DLWP_CUDA_TRY(cudaMemcpyAsync(pl->d_x0, x0_host, sizeof(float) * slot, cudaMemcpyHostToDevice, pl->s_compute));
    for (int g = 0; g < groups; ++g) {
        const int t0 = g * d2h_group, t1 = std::min(iterations, t0 + d2h_group);
        rc = rollout_range(pl, N, pl->d_x0, pl->d_series, t0, t1, pl->s_compute);
        if (rc) return rc;
        DLWP_CUDA_TRY(cudaEventRecord(pl->events[g], pl->s_compute));
        DLWP_CUDA_TRY(cudaStreamWaitEvent(pl->s_copy, pl->events[g], 0));
        const long long off = (long long)t0 * n_out * slot, cnt = (long long)(t1 - t0) * n_out * slot;
        DLWP_CUDA_TRY(cudaMemcpyAsync(series_host + off, pl->d_series + off, sizeof(float) * cnt,
                                      cudaMemcpyDeviceToHost, pl->s_copy));
    }
    DLWP_CUDA_TRY(cudaStreamSynchronize(pl->s_copy));
    DLWP_CUDA_TRY(cudaStreamSynchronize(pl->s_compute));
    return 0;
}

extern "C" int dlwp_plan_profile_op(DlwpPlan* pl, int32_t N, int32_t op_index, int32_t iters, float* ms_per_launch,
                                    dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && ms_per_launch, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(op_index >= 0 && op_index < (int)pl->ops.size() && iters > 0, DLWP_EINVAL, "bad op index / iters");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch, DLWP_ESHAPE, "batch %d outside (0, %d]", N, pl->max_batch);
    DLWP_REQUIRE(pl->buffers[pl->input_buf].ptr != nullptr, DLWP_ESTATE, "run a forward / rollout first");
    cudaStream_t stream = (cudaStream_t)stream_;
    cudaEvent_t e0, e1;
    DLWP_CUDA_TRY(cudaEventCreate(&e0));
    DLWP_CUDA_TRY(cudaEventCreate(&e1));
    int rc = 0;
    const int pt = pl->tc_last_t;
    for (int w = 0; w < 2 && !rc; ++w) rc = pl->tc ? run_one_tc(pl, op_index, N, stream, pt, true) : run_one(pl, op_index, N, stream);
    cudaEventRecord(e0, stream);
    for (int k = 0; k < iters && !rc; ++k)
        rc = pl->tc ? run_one_tc(pl, op_index, N, stream, pt, true) : run_one(pl, op_index, N, stream);
    cudaEventRecord(e1, stream);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return rc;
    DLWP_CUDA_TRY(e);
    *ms_per_launch = ms / iters;
    return 0;
}

extern "C" int dlwp_plan_uses_tensor_cores(DlwpPlan* pl) { return pl && pl->tc ? 1 : 0; }
extern "C" int dlwp_plan_fused_pair(DlwpPlan* pl) { return pl && pl->tc ? pl->pair_first : -1; }

// =================================================================================================================
// Latitude-band rollout with the halo exchange inside the library (NCCL resolved at run time, no link dependency)
// =================================================================================================================
namespace dlwp {
typedef struct { char internal[128]; } NcclId;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(const char* path) {
    if (g_nccl.handle) return 0;
    void* h = dlopen(path && *path ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    DLWP_REQUIRE(h != nullptr, DLWP_EARCH, "cannot load NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
    g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    DLWP_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.Send &&
                     g_nccl.Recv,
                 DLWP_EARCH, "NCCL library lacks the point-to-point API");
    g_nccl.handle = h;
    return 0;
}

#define DLWP_NCCL_TRY(expr)                                                                            \
    do {                                                                                               \
        int _r = (expr);                                                                               \
        if (_r != 0) {                                                                                 \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return 1000 + _r;                                                                          \
        }                                                                                              \
    } while (0)

static int amax_allreduce(void* comm, float* amax, cudaStream_t stream) {
    DLWP_REQUIRE(g_nccl.AllReduce != nullptr, DLWP_EARCH, "the loaded NCCL library has no ncclAllReduce");
    DLWP_NCCL_TRY(g_nccl.AllReduce(amax, amax, 1, 7 /* ncclFloat32 */, 2 /* ncclMax */, comm, stream));
    return 0;
}

// One grouped SendRecv of the halo rows of `slot` (N, C, H, W): my top/bottom band rows out, the neighbours' rows in.
// Three launches per iteration: ONE kernel packs the rows for both neighbours, one NCCL group moves them, ONE kernel
// stores what arrived -- on the tensor-core chain straight into the halo rows of the next input's P image (iteration
// t + 1's exponent word, which the feedback conv of iteration t has already published), so the fp32 halo rows of the
// series slot are never written (they belong to the neighbours' output); otherwise into the fp32 slot.
static int halo_exchange(DlwpPlan* pl, void* comm, int N, float* slot, const DlwpBandInfo& b, cudaStream_t stream, int t) {
    const Buffer& in = pl->buffers[pl->input_buf];
    const int C = in.d.C, H = in.d.H, W = in.d.W;
    const long long sn = (long long)C * H * W, sc = (long long)H * W;
    const bool up = b.rank > 0, down = b.rank + 1 < b.world;
    const int s_rows[2] = {up ? b.send_up : 0, down ? b.send_down : 0};
    const int r_rows[2] = {up ? b.recv_top : 0, down ? b.recv_bot : 0};
    int rc = 0;
    {
        const float* src[2] = {s_rows[0] ? slot + (long long)b.band_lo * W : nullptr,
                               s_rows[1] ? slot + (long long)(b.band_hi - b.send_down) * W : nullptr};
        float* dst[2] = {pl->halo_stage[0], pl->halo_stage[1]};
        const long long ss_n[2] = {sn, sn}, ss_c[2] = {sc, sc};
        const long long ds_n[2] = {(long long)C * s_rows[0] * W, (long long)C * s_rows[1] * W};
        const long long ds_c[2] = {(long long)s_rows[0] * W, (long long)s_rows[1] * W};
        if ((rc = halo_copy(src, dst, s_rows, ss_n, ss_c, ds_n, ds_c, N, C, W, stream))) return rc;
    }
    const size_t per_row = (size_t)N * C * W;
    DLWP_NCCL_TRY(g_nccl.GroupStart());
    if (s_rows[0]) DLWP_NCCL_TRY(g_nccl.Send(pl->halo_stage[0], per_row * s_rows[0], 7, b.rank - 1, comm, stream));
    if (r_rows[0]) DLWP_NCCL_TRY(g_nccl.Recv(pl->halo_stage[2], per_row * r_rows[0], 7, b.rank - 1, comm, stream));
    if (s_rows[1]) DLWP_NCCL_TRY(g_nccl.Send(pl->halo_stage[1], per_row * s_rows[1], 7, b.rank + 1, comm, stream));
    if (r_rows[1]) DLWP_NCCL_TRY(g_nccl.Recv(pl->halo_stage[3], per_row * r_rows[1], 7, b.rank + 1, comm, stream));
    DLWP_NCCL_TRY(g_nccl.GroupEnd());
    const int r_row0[2] = {b.band_lo - b.recv_top, b.band_hi};
    if (pl->tc && pl->tc_feedback_op >= 0) {
        TcPackScale ps;
        ps.e = pl->d_exp + 2 * pl->input_buf + ((t + 1) & 1);
        ps.amax = pl->d_amax + 2 * pl->input_buf + ((t + 1) & 1);
        ps.fresh = 0;
        ps.bf16 = pl->tc_opt.bf16;
        const float* src[2] = {r_rows[0] ? pl->halo_stage[2] : nullptr, r_rows[1] ? pl->halo_stage[3] : nullptr};
        pl->halo_in_p = true;
        return tc_pack_halo(src, r_row0, r_rows, in.P, N, C, H, W, in.wpad, stream, ps);
    }
    const float* src[2] = {r_rows[0] ? pl->halo_stage[2] : nullptr, r_rows[1] ? pl->halo_stage[3] : nullptr};
    float* dst[2] = {slot + (long long)r_row0[0] * W, slot + (long long)r_row0[1] * W};
    const long long ss_n[2] = {(long long)C * r_rows[0] * W, (long long)C * r_rows[1] * W};
    const long long ss_c[2] = {(long long)r_rows[0] * W, (long long)r_rows[1] * W};
    const long long ds_n[2] = {sn, sn}, ds_c[2] = {sc, sc};
    return halo_copy(src, dst, r_rows, ss_n, ss_c, ds_n, ds_c, N, C, W, stream);
}

// Can the exchange of iteration t run BEHIND the start of iteration t + 1?  Yes when the plan is a tensor-core chain with a
// feedback conv (the received rows go straight into the next input's P image), exactly one op -- the first -- reads the
// input, it is a conv, and its destination image has a static exponent (so the rows it computes before the halo arrives
// are scaled like the rows it computes after).  Then op 0 of iteration t + 1 is split into the rows that only read this
// rank's own band (launched right after iteration t) and the edge rows that read halo rows (launched once the exchange,
// which runs on a second stream, has finished).
// ---- halo over peer memory: arrival counters ------------------------------------------------------------------------------
// After its feedback conv (whose epilogue stored the neighbours' halo rows through NVLink) a rank bumps the arrival counter
// in each neighbour's flag block; before the first layer of the next iteration it waits until its own counters have
// reached the number of arrivals it has consumed so far plus one.  All counters only grow, so a CUDA-graph replay of the
// whole rollout needs no reset.
__global__ void halo_signal_kernel(int* up_flags, int* down_flags) {
    __threadfence_system();
    if (up_flags) atomicAdd_system(up_flags + 1, 1);      // I am the upper neighbour's LOWER neighbour
    if (down_flags) atomicAdd_system(down_flags + 0, 1);
}
__global__ void halo_wait_kernel(int* flags, int has_up, int has_down) {
    for (int k = 0; k < 2; ++k) {
        if (!(k == 0 ? has_up : has_down)) continue;
        const int expected = ++flags[2 + k];
        volatile int* arr = flags + k;
        unsigned spins = 0;
        while (*arr < expected) {
            __nanosleep(200);
            if (++spins > (1u << 24)) {            // ~ seconds: a neighbour died; raise the timeout flag instead of hanging
                atomicOr(&g_device_flags_plan, 1);
                break;
            }
        }
    }
    __threadfence_system();
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static bool latband_can_overlap(const DlwpPlan* pl, const DlwpBandInfo& b) {
    if (b.world <= 1 || !pl->tc || pl->tc_feedback_op < 0 || pl->pair_first >= 0 || pl->opt.latband_spare_sms <= 0) return false;
    if (pl->ops.empty() || pl->ops[0].kind != DLWP_OP_CONV || pl->ops[0].src != pl->input_buf) return false;
    for (size_t i = 1; i < pl->ops.size(); ++i)
        if (pl->ops[i].src == pl->input_buf) return false;
    const int pd = pl->tc_pdst[0];
    if (pd < 0 || (!pl->tc_opt.bf16 && !pl->buffers[pd].e_static)) return false;
    if (pl->ops[0].row_begin == 0 && pl->ops[0].row_end == 0) return false;
    return true;
}

// Iterations [t0, t1) of a rollout of `iterations` steps (the whole loop for a device-resident rollout; groups of steps when
// the D2H of finished states is pipelined behind the compute, dlwp_rollout_latband_host).
static int latband_range(DlwpPlan* pl, void* comm, int N, const float* x0, float* series, int iterations,
                         const DlwpBandInfo& b, cudaStream_t stream, int t0 = 0, int t1 = -1) {
    if (t1 < 0) t1 = iterations;
    const long long slot = (long long)N * pl->buffers[pl->input_buf].sample_elems();
    const int n_out = (int)pl->outputs.size();
    if (pl->p2p && b.world > 1) {
        // Halo over peer memory: no packing, no NCCL -- the feedback conv's epilogue writes the neighbours' halo rows, two
        // one-thread kernels count arrivals.  A neighbour barrier opens every rollout (nobody may still read the image
        // parity this rollout's first feedback conv overwrites).
        Buffer& in = pl->buffers[pl->input_buf];
        const bool up = b.rank > 0, down = b.rank + 1 < b.world;
        pl->p2p_up_end = up ? b.band_lo + b.send_up : 0;
        pl->p2p_down_begin = down ? b.band_hi - b.send_down : 1 << 30;
        int rc = 0;
        if (t0 == 0) {
            halo_signal_kernel<<<1, 1, 0, stream>>>(up ? pl->peer_flags[0] : nullptr, down ? pl->peer_flags[1] : nullptr);
            halo_wait_kernel<<<1, 1, 0, stream>>>(pl->flags, up ? 1 : 0, down ? 1 : 0);
            if ((rc = after_launch("halo barrier"))) return rc;
        }
        pl->p2p_active = true;
        for (int t = t0; t < t1 && !rc; ++t) {
            pl->cur_parity = t & 1;
            in.P = pl->P_in[t & 1];
            pl->p2p_send = t + 1 < iterations;
            if (t > 0) {
                halo_wait_kernel<<<1, 1, 0, stream>>>(pl->flags, up ? 1 : 0, down ? 1 : 0);
                pl->halo_in_p = true;      // the halo rows are in the image already: nothing to pack from the fp32 slot
            }
            rc = rollout_range(pl, N, x0, series, t, t + 1, stream);
            if (!rc && t + 1 < iterations) {
                halo_signal_kernel<<<1, 1, 0, stream>>>(up ? pl->peer_flags[0] : nullptr, down ? pl->peer_flags[1] : nullptr);
                rc = after_launch("halo_signal_kernel");
            }
        }
        pl->p2p_active = false;
        pl->p2p_send = false;
        pl->cur_parity = 0;
        in.P = pl->P_in[0];
        return rc;
    }
    const bool overlap = latband_can_overlap(pl, b);
    if (overlap) {
        if (!pl->s_copy) {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            DLWP_CUDA_TRY(cudaStreamCreateWithPriority(&pl->s_copy, cudaStreamNonBlocking, hi_pri));
        }
        while (pl->events.size() < 2) {
            cudaEvent_t ev;
            DLWP_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            pl->events.push_back(ev);
        }
    }
    // rows of op 0 that read only this rank's band rows of the input
    DlwpOpDesc& op0 = pl->ops[0];
    const int w0 = op0.row_begin, w1 = op0.row_end, span = op0.dil_h * (op0.kh - 1);
    const int int_lo = b.rank > 0 ? std::max(w0, b.band_lo + op0.pad_t) : w0;
    const int int_hi = b.rank + 1 < b.world ? std::min(w1, b.band_hi - span + op0.pad_t) : w1;
    bool halo_pending = false;
    for (int t = t0; t < t1; ++t) {
        int rc = 0;
        if (overlap && halo_pending && int_lo < int_hi) {
            // iteration t with the exchange of iteration t - 1 still in flight on s_copy
            Buffer& in = pl->buffers[pl->input_buf];
            in.ptr = series + ((long long)t * n_out - 1) * slot;
            for (int k = 0; k < n_out; ++k) pl->buffers[pl->outputs[k]].ptr = series + ((long long)t * n_out + k) * slot;
            pl->tc_last_t = t;
            op0.row_begin = int_lo; op0.row_end = int_hi;
            // the exchange's kernels (packing, NCCL, unpacking) need SMs of their own: the conv CTAs are persistent and
            // fill an SM's shared memory, so nothing else starts on an SM until its CTA exits
            pl->tc_opt.max_ctas = std::max(1, sm_count() - pl->opt.latband_spare_sms);
            rc = run_one_tc(pl, 0, N, stream, t, true);
            pl->tc_opt.max_ctas = 0;
            if (!rc) rc = cudaStreamWaitEvent(stream, pl->events[1], 0) == cudaSuccess ? 0 : (int)cudaGetLastError();
            pl->halo_in_p = false;
            if (!rc && w0 < int_lo) { op0.row_begin = w0; op0.row_end = int_lo; rc = run_one_tc(pl, 0, N, stream, t, true); }
            if (!rc && int_hi < w1) { op0.row_begin = int_hi; op0.row_end = w1; rc = run_one_tc(pl, 0, N, stream, t, true); }
            op0.row_begin = w0; op0.row_end = w1;
            for (size_t i = 1; i < pl->ops.size() && !rc; ++i) rc = run_one_tc(pl, (int)i, N, stream, t, true);
            halo_pending = false;
        } else {
            if (halo_pending) {   // (no interior rows: plain order)
                DLWP_CUDA_TRY(cudaStreamWaitEvent(stream, pl->events[1], 0));
                halo_pending = false;
            }
            rc = rollout_range(pl, N, x0, series, t, t + 1, stream);
        }
        if (rc) return rc;
        if (t + 1 < iterations && b.world > 1) {
            float* last = series + ((long long)t * n_out + n_out - 1) * slot;
            if (overlap) {
                DLWP_CUDA_TRY(cudaEventRecord(pl->events[0], stream));
                DLWP_CUDA_TRY(cudaStreamWaitEvent(pl->s_copy, pl->events[0], 0));
                rc = halo_exchange(pl, comm, N, last, b, pl->s_copy, t);
                if (rc) return rc;
                DLWP_CUDA_TRY(cudaEventRecord(pl->events[1], pl->s_copy));
                halo_pending = true;
            } else {
                rc = halo_exchange(pl, comm, N, last, b, stream, t);
                if (rc) return rc;
            }
        }
    }
    if (halo_pending) DLWP_CUDA_TRY(cudaStreamWaitEvent(stream, pl->events[1], 0));
    return 0;
}
}  // namespace dlwp

extern "C" int dlwp_comm_unique_id(const char* libnccl_path, void* id128) {
    DLWP_REQUIRE(id128 != nullptr, DLWP_EINVAL, "null argument");
    int rc = nccl_load(libnccl_path);
    if (rc) return rc;
    DLWP_NCCL_TRY(g_nccl.GetUniqueId(reinterpret_cast<NcclId*>(id128)));
    return 0;
}

extern "C" int dlwp_comm_create(const char* libnccl_path, int32_t rank, int32_t world, const void* id128, void** comm) {
    DLWP_REQUIRE(id128 && comm && world > 0 && rank >= 0 && rank < world, DLWP_EINVAL, "bad argument");
    int rc = nccl_load(libnccl_path);
    if (rc) return rc;
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    DLWP_NCCL_TRY(g_nccl.CommInitRank(comm, world, id, rank));
    return 0;
}

extern "C" void dlwp_comm_destroy(void* comm) {
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
}

// ---- halo over peer memory: set-up ---------------------------------------------------------------------------------------
extern "C" int dlwp_plan_halo_enable(DlwpPlan* pl) {
    DLWP_REQUIRE(pl != nullptr, DLWP_EINVAL, "null plan");
    if (pl->p2p) return 0;
    // same preconditions as the overlapped NCCL exchange: a tensor-core chain whose last conv re-packs the next input, one
    // reader of the input (the first op) whose destination image has a data-independent exponent
    DLWP_REQUIRE(pl->tc && pl->tc_feedback_op >= 0 && pl->pair_first < 0, DLWP_ESTATE,
                 "the peer-memory halo needs a tensor-core chain with a feedback conv");
    DLWP_REQUIRE(!pl->ops.empty() && pl->ops[0].kind == DLWP_OP_CONV && pl->ops[0].src == pl->input_buf, DLWP_ESTATE,
                 "the peer-memory halo needs a conv as the only reader of the input");
    for (size_t i = 1; i < pl->ops.size(); ++i)
        DLWP_REQUIRE(pl->ops[i].src != pl->input_buf, DLWP_ESTATE, "the peer-memory halo needs ONE reader of the input");
    const int pd = pl->tc_pdst[0];
    DLWP_REQUIRE(pd >= 0 && (pl->tc_opt.bf16 || pl->buffers[pd].e_static), DLWP_ESTATE,
                 "the first layer's output exponent depends on the data (rank-dependent without an amax exchange)");
    Buffer& in = pl->buffers[pl->input_buf];
    const size_t bytes = tc_p_bytes(pl->max_batch, in.planes, in.d.H, in.d.W + 2 * in.wpad);
    pl->P_in[0] = in.P;
    DLWP_CUDA_TRY(cudaMalloc(&pl->P_in[1], bytes));
    DLWP_CUDA_TRY(cudaMemset(pl->P_in[1], 0, bytes));
    DLWP_CUDA_TRY(cudaMalloc(&pl->flags, 64));
    DLWP_CUDA_TRY(cudaMemset(pl->flags, 0, 64));
    DLWP_CUDA_TRY(cudaDeviceSynchronize());
    // everything the peer-memory rollout launches is loaded now, not at its first launch (sw_preload_one)
    {
        const int fb = pl->tc_feedback_op;
        const DlwpOpDesc& op = pl->ops[fb];
        DlwpConvDesc d = conv_desc_of(pl, op, pl->max_batch);
        tc_preload(d, pl->tc_layers[fb], 3, 1, pl->tc_opt);
        for (size_t i = 0; i < pl->ops.size(); ++i)
            if (pl->ops[i].kind == DLWP_OP_CONV)
                for (int mode = 1; mode <= 3; ++mode)
                    tc_preload(conv_desc_of(pl, pl->ops[i], pl->max_batch), pl->tc_layers[i], mode, 0, pl->tc_opt);
        tc_preload_aux();
        cudaFuncAttributes attr;
        cudaFuncGetAttributes(&attr, halo_signal_kernel);
        cudaFuncGetAttributes(&attr, halo_wait_kernel);
    }
    pl->p2p = true;
    return 0;
}

// handles: 3 x 64 bytes (cudaIpcMemHandle_t of the two input images and of the flag block), to be sent to the neighbours
extern "C" int dlwp_plan_halo_export(DlwpPlan* pl, void* handles192) {
    DLWP_REQUIRE(pl && handles192 && pl->p2p, DLWP_EINVAL, "halo not enabled");
    cudaIpcMemHandle_t h[3];
    DLWP_CUDA_TRY(cudaIpcGetMemHandle(&h[0], pl->P_in[0]));
    DLWP_CUDA_TRY(cudaIpcGetMemHandle(&h[1], pl->P_in[1]));
    DLWP_CUDA_TRY(cudaIpcGetMemHandle(&h[2], pl->flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles192, h, sizeof(h));
    return 0;
}

// which: 0 = upper neighbour (rank - 1), 1 = lower neighbour (rank + 1); handles192 as exported by THAT rank
extern "C" int dlwp_plan_halo_import(DlwpPlan* pl, int32_t which, const void* handles192) {
    DLWP_REQUIRE(pl && handles192 && pl->p2p && (which == 0 || which == 1), DLWP_EINVAL, "bad argument");
    cudaIpcMemHandle_t h[3];
    memcpy(h, handles192, sizeof(h));
    void* ptr[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; ++k) {
        DLWP_CUDA_TRY(cudaIpcOpenMemHandle(&ptr[k], h[k], cudaIpcMemLazyEnablePeerAccess));
        pl->ipc_opened.push_back(ptr[k]);
    }
    pl->peer_P[which][0] = (__half*)ptr[0];
    pl->peer_P[which][1] = (__half*)ptr[1];
    pl->peer_flags[which] = (int*)ptr[2];
    return 0;
}

// Same-process neighbours (tests: several bands on one GPU, one plan each): raw device pointers of the other plan.
extern "C" int dlwp_plan_halo_connect(DlwpPlan* pl, int32_t which, DlwpPlan* neighbour) {
    DLWP_REQUIRE(pl && neighbour && pl->p2p && neighbour->p2p && (which == 0 || which == 1), DLWP_EINVAL, "bad argument");
    pl->peer_P[which][0] = neighbour->P_in[0];
    pl->peer_P[which][1] = neighbour->P_in[1];
    pl->peer_flags[which] = neighbour->flags;
    return 0;
}

namespace dlwp {
static int ensure_halo_stage(DlwpPlan* pl, const DlwpBandInfo& b) {
    const Buffer& in = pl->buffers[pl->input_buf];
    const int max_rows = std::max(std::max(b.send_up, b.send_down), std::max(b.recv_top, b.recv_bot));
    const long long need = (long long)pl->max_batch * in.d.C * std::max(1, max_rows) * in.d.W;
    if (pl->halo_cap < need) {
        for (float*& h : pl->halo_stage) {
            if (h) cudaFree(h);
            h = nullptr;
            DLWP_CUDA_TRY(cudaMalloc(&h, sizeof(float) * need));
        }
        pl->halo_cap = need;
    }
    return 0;
}

static int check_latband_args(DlwpPlan* pl, void* comm, int N, int iterations, const DlwpBandInfo* band) {
    DLWP_REQUIRE(band->world == 1 || comm != nullptr || pl->p2p, DLWP_EINVAL,
                 "a communicator (or dlwp_plan_halo_enable + peers) is required for world > 1");
    if (pl->p2p && band->world > 1) {
        DLWP_REQUIRE(band->rank == 0 || pl->peer_flags[0], DLWP_ESTATE, "upper neighbour not connected");
        DLWP_REQUIRE(band->rank + 1 == band->world || pl->peer_flags[1], DLWP_ESTATE, "lower neighbour not connected");
    }
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch && iterations > 0, DLWP_ESHAPE, "bad batch / iterations");
    return check_rollout_shapes(pl);
}
}  // namespace dlwp

extern "C" int dlwp_rollout_latband(DlwpPlan* pl, void* comm, int32_t N, const float* x0, float* series,
                                    int32_t iterations, const DlwpBandInfo* band, int32_t use_graph,
                                    dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && x0 && series && band, DLWP_EINVAL, "null argument");
    int rc = check_latband_args(pl, comm, N, iterations, band);
    if (rc) return rc;
    rc = ensure_halo_stage(pl, *band);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!use_graph) return latband_range(pl, comm, N, x0, series, iterations, *band, stream);
    GraphKey key{N, iterations, x0, series};
    auto it = pl->band_graphs.find(key);
    if (it == pl->band_graphs.end()) {
        if (!pl->s_compute) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_compute, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        DLWP_CUDA_TRY(cudaStreamBeginCapture(pl->s_compute, cudaStreamCaptureModeThreadLocal));
        const long long before = g_launches.load();
        rc = latband_range(pl, comm, N, x0, series, iterations, *band, pl->s_compute);
        cudaError_t e = cudaStreamEndCapture(pl->s_compute, &graph);
        g_launches.store(before);
        if (rc) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        DLWP_CUDA_TRY(e);
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        DLWP_CUDA_TRY(e);
        if (pl->band_graphs.size() >= 8) {  // bounded cache, like the single-domain graphs
            cudaGraphExecDestroy(pl->band_graphs.begin()->second);
            pl->band_graphs.erase(pl->band_graphs.begin());
        }
        it = pl->band_graphs.emplace(key, exec).first;
    }
    DLWP_CUDA_TRY(cudaGraphLaunch(it->second, stream));
    g_launches.fetch_add((long long)pl->ops.size() * iterations);
    return 0;
}

extern "C" int dlwp_rollout_latband_host(DlwpPlan* pl, void* comm, int32_t N, const float* x0_host, float* band_host,
                                         int32_t iterations, const DlwpBandInfo* band, int32_t d2h_group) {
    DLWP_REQUIRE(pl && x0_host && band_host && band, DLWP_EINVAL, "null argument");
    int rc = check_latband_args(pl, comm, N, iterations, band);
    if (rc) return rc;
    rc = ensure_halo_stage(pl, *band);
    if (rc) return rc;
    const Buffer& in = pl->buffers[pl->input_buf];
    const int H = in.d.H, W = in.d.W, rows = band->band_hi - band->band_lo;
    DLWP_REQUIRE(band->band_lo >= 0 && band->band_hi <= H && rows > 0, DLWP_ESHAPE, "band [%d, %d) outside [0, %d)",
                 band->band_lo, band->band_hi, H);
    const long long slot = (long long)N * in.sample_elems();
    const int n_out = (int)pl->outputs.size();
    rc = ensure_host_staging(pl, slot, slot * n_out * iterations);
    if (rc) return rc;
    if (!pl->s_compute) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_compute, cudaStreamNonBlocking));
    if (!pl->s_d2h) DLWP_CUDA_TRY(cudaStreamCreateWithFlags(&pl->s_d2h, cudaStreamNonBlocking));
    if (d2h_group <= 0) d2h_group = std::max(1, iterations / 8);
    const int groups = (iterations + d2h_group - 1) / d2h_group;
    while ((int)pl->d2h_events.size() < groups) {
        cudaEvent_t ev;
        DLWP_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        pl->d2h_events.push_back(ev);
    }
    const long long planes = (long long)N * in.d.C;        // (sample, channel) planes per state
    // Only the rows this rank reads (band + halo) go up.  The exponent of the x0 image comes from max|x0| over the FULL state
    // (every rank must scale alike, and like a single-domain rollout, for the bands to stay bit-identical to it): each rank
    // measures its rows and one 4-byte ncclAllReduce(max) on `comm` makes it global.  Without a communicator (peer-memory
    // halo set up by hand) the whole x0 is uploaded instead; bf16 images and the fp32 kernels have no exponent.
    int up_lo = std::max(0, band->band_lo - band->recv_top), up_hi = std::min(H, band->band_hi + band->recv_bot);
    if (pl->tc && pl->tc_in_row1 > 0) { up_lo = std::min(up_lo, pl->tc_in_row0); up_hi = std::max(up_hi, pl->tc_in_row1); }
    const bool needs_amax = pl->tc && !pl->tc_opt.bf16;
    const bool partial = band->world > 1 && (up_lo > 0 || up_hi < H) && (!needs_amax || comm != nullptr);
    if (partial) {
        const size_t up_width = sizeof(float) * (size_t)(up_hi - up_lo) * W, pitch = sizeof(float) * (size_t)H * W;
        DLWP_CUDA_TRY(cudaMemcpy2DAsync(pl->d_x0 + (long long)up_lo * W, pitch, x0_host + (long long)up_lo * W, pitch,
                                        up_width, (size_t)planes, cudaMemcpyHostToDevice, pl->s_compute));
    } else {
        DLWP_CUDA_TRY(cudaMemcpyAsync(pl->d_x0, x0_host, sizeof(float) * slot, cudaMemcpyHostToDevice, pl->s_compute));
    }
    pl->x0_amax_comm = (partial && needs_amax) ? comm : nullptr;
    const size_t width = sizeof(float) * (size_t)rows * W;  // a band's rows of one plane are contiguous
    for (int g = 0; g < groups; ++g) {
        const int t0 = g * d2h_group, t1 = std::min(iterations, t0 + d2h_group);
        rc = latband_range(pl, comm, N, pl->d_x0, pl->d_series, iterations, *band, pl->s_compute, t0, t1);
        pl->x0_amax_comm = nullptr;     // (only iteration 0 packs x0)
        if (rc) return rc;
        DLWP_CUDA_TRY(cudaEventRecord(pl->d2h_events[g], pl->s_compute));
        DLWP_CUDA_TRY(cudaStreamWaitEvent(pl->s_d2h, pl->d2h_events[g], 0));
        const long long states = (long long)(t1 - t0) * n_out;
        DLWP_CUDA_TRY(cudaMemcpy2DAsync(band_host + (long long)t0 * n_out * planes * rows * W, width,
                                        pl->d_series + (long long)t0 * n_out * slot + (long long)band->band_lo * W,
                                        sizeof(float) * (size_t)H * W, width, (size_t)(states * planes),
                                        cudaMemcpyDeviceToHost, pl->s_d2h));
    }
    DLWP_CUDA_TRY(cudaStreamSynchronize(pl->s_d2h));
    DLWP_CUDA_TRY(cudaStreamSynchronize(pl->s_compute));
    return 0;
}

// =================================================================================================================
// Training: forward + sum_k w_k MSE_k + backward, gradient buffer for the data-parallel all-reduce, Adam
// (keras.Model.fit_generator / train_on_batch as driven by DLWP/model/models.py:216-228, :394-402)
// =================================================================================================================
namespace dlwp {
static int train_setup(DlwpPlan* pl) {
    if (pl->train_ready) return 0;
    DLWP_REQUIRE(!pl->tc, DLWP_ESTATE, "training needs the fp32 plan (create it with DlwpPlanOptions.math = 1)");
    for (const DlwpOpDesc& op : pl->ops) {
        DLWP_REQUIRE(op.kind != DLWP_OP_PAD, DLWP_ESHAPE, "stand-alone padding layers are not differentiable here yet");
        DLWP_REQUIRE(op.kind != DLWP_OP_LSTM, DLWP_ESHAPE, "ConvLSTM2D is inference-only here (no backward kernels yet)");
        DLWP_REQUIRE(op.kind != DLWP_OP_TO_NCHW && op.kind != DLWP_OP_TO_NHWC, DLWP_ESHAPE,
                     "channels_last models are inference-only here");
        DLWP_REQUIRE(!(op.kind == DLWP_OP_CONV && (op.rowwise || op.pre_op)), DLWP_ESHAPE,
                     "RowConnected2D / fused pre-ops are not differentiable here yet");
        DLWP_REQUIRE(!op.row_begin && !op.row_end, DLWP_ESHAPE, "row-windowed plans cannot be trained");
    }
    pl->gbuf.assign(pl->buffers.size(), nullptr);
    for (size_t i = 0; i < pl->buffers.size(); ++i)
        DLWP_CUDA_TRY(cudaMalloc(&pl->gbuf[i], sizeof(float) * pl->buffers[i].sample_elems() * pl->max_batch));
    pl->out_store.assign(pl->outputs.size(), nullptr);
    for (size_t k = 0; k < pl->outputs.size(); ++k)
        DLWP_CUDA_TRY(cudaMalloc(&pl->out_store[k],
                                 sizeof(float) * pl->buffers[pl->outputs[k]].sample_elems() * pl->max_batch));
    long long off = 0;
    pl->gk_off.assign(pl->weights.size(), 0);
    pl->gb_off.assign(pl->weights.size(), 0);
    for (size_t w = 0; w < pl->weights.size(); ++w) {
        pl->gk_off[w] = off; off += pl->weights[w].k_elems;
        pl->gb_off[w] = off; off += pl->weights[w].b_elems;
    }
    pl->flat_elems = off;
    DLWP_CUDA_TRY(cudaMalloc(&pl->flat_g, sizeof(float) * off));
    DLWP_CUDA_TRY(cudaMalloc(&pl->flat_m, sizeof(float) * off));
    DLWP_CUDA_TRY(cudaMalloc(&pl->flat_v, sizeof(float) * off));
    DLWP_CUDA_TRY(cudaMemset(pl->flat_m, 0, sizeof(float) * off));
    DLWP_CUDA_TRY(cudaMemset(pl->flat_v, 0, sizeof(float) * off));
    DLWP_CUDA_TRY(cudaMalloc(&pl->stats, sizeof(float) * 2 * pl->outputs.size()));
    long long max_buf = 0, max_w = 0;
    for (const Buffer& b : pl->buffers) max_buf = std::max(max_buf, b.sample_elems() * pl->max_batch);
    for (const Weight& w : pl->weights) max_w = std::max(max_w, w.k_elems);
    DLWP_CUDA_TRY(cudaMalloc(&pl->bwd_scratch, sizeof(float) * max_buf));
    DLWP_CUDA_TRY(cudaMalloc(&pl->wt_scratch, sizeof(float) * std::max<long long>(max_w, 1)));
    pl->train_ready = true;
    return 0;
}
}  // namespace dlwp

extern "C" int dlwp_train_step(DlwpPlan* pl, int32_t N, const float* x, const float* const* targets,
                               const float* loss_weights, int32_t backward, int32_t input_grad, float* losses,
                               float* maes, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && x && targets && losses, DLWP_EINVAL, "null argument");
    DLWP_REQUIRE(N > 0 && N <= pl->max_batch, DLWP_ESHAPE, "batch %d outside (0, %d]", N, pl->max_batch);
    int rc = train_setup(pl);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_out = (int)pl->outputs.size();
    // ---- forward (every intermediate stays in its INTERNAL buffer) ----
    pl->buffers[pl->input_buf].ptr = const_cast<float*>(x);
    for (int k = 0; k < n_out; ++k) pl->buffers[pl->outputs[k]].ptr = pl->out_store[k];
    rc = run_ops(pl, N, stream);
    if (rc) return rc;
    // ---- loss and its gradient ----
    DLWP_CUDA_TRY(cudaMemsetAsync(pl->stats, 0, sizeof(float) * 2 * n_out, stream));
    if (backward) {
        for (size_t i = 0; i < pl->buffers.size(); ++i)
            DLWP_CUDA_TRY(cudaMemsetAsync(pl->gbuf[i], 0, sizeof(float) * pl->buffers[i].sample_elems() * N, stream));
        DLWP_CUDA_TRY(cudaMemsetAsync(pl->flat_g, 0, sizeof(float) * pl->flat_elems, stream));
    }
    for (int k = 0; k < n_out; ++k) {
        DLWP_REQUIRE(targets[k] != nullptr, DLWP_EINVAL, "target %d is null", k);
        const long long nk = (long long)N * pl->buffers[pl->outputs[k]].sample_elems();
        const float lw = loss_weights ? loss_weights[k] : 1.f;
        const Buffer& ob = pl->buffers[pl->outputs[k]];
        if (pl->loss_kind == 1) {
            if (!pl->acc_stats) DLWP_CUDA_TRY(cudaMalloc(&pl->acc_stats, sizeof(double) * 5 * n_out));
            if (k == 0) DLWP_CUDA_TRY(cudaMemsetAsync(pl->acc_stats, 0, sizeof(double) * 5 * n_out, stream));
            DLWP_REQUIRE(!pl->acc_mean || pl->acc_mean_elems == ob.sample_elems(), DLWP_ESHAPE,
                         "the ACC loss mean has %lld elements, output %d has %lld per sample", pl->acc_mean_elems, k,
                         ob.sample_elems());
            rc = acc_loss_grad(pl->out_store[k], targets[k], pl->acc_mean, ob.sample_elems(),
                               backward ? pl->gbuf[pl->outputs[k]] : nullptr, nk, pl->acc_stats + 5 * k,
                               pl->acc_reverse ? lw : -lw, pl->acc_regularize, stream);
        } else
            rc = mse_grad(pl->out_store[k], targets[k], backward ? pl->gbuf[pl->outputs[k]] : nullptr, nk, lw * 2.f / (float)nk,
                          pl->stats + 2 * k, stream, pl->loss_wmap, (long long)ob.d.H * ob.d.W);
        if (rc) return rc;
    }
    // ---- backward: ops in reverse order; gradients accumulate into channel windows of the per-buffer gradient ----
    if (backward) {
        pl->g_touched.assign(pl->buffers.size(), 0);
        for (int o : pl->outputs) pl->g_touched[o] = 1;     // the loss gradient is already there
        for (int i = (int)pl->ops.size() - 1; i >= 0; --i) {
            const DlwpOpDesc& op = pl->ops[i];
            const Buffer& s = pl->buffers[op.src];
            const Buffer& t = pl->buffers[op.dst];
            const long long s_hw = (long long)s.d.H * s.d.W, t_hw = (long long)t.d.H * t.d.W;
            const long long ss[3] = {s.sample_elems(), s_hw, s.d.W}, ts[3] = {t.sample_elems(), t_hw, t.d.W};
            const float* xs = s.ptr + op.src_c0 * s_hw;
            float* gs = pl->gbuf[op.src] + op.src_c0 * s_hw;
            float* gt = pl->gbuf[op.dst] + op.dst_c0 * t_hw;
            const float* yt = t.ptr + op.dst_c0 * t_hw;
            switch (op.kind) {
                case DLWP_OP_CONV: {
                    Weight& w = pl->weights[op.weight_id];
                    DlwpConvDesc d = conv_desc_of(pl, op, N);
                    rc = act_bwd(yt, gt, op.act, N, op.Cout, t.d.H, t.d.W, ts, ts, stream);
                    if (!rc) rc = conv2d_bwd_weight(d, xs, gt, pl->flat_g + pl->gk_off[op.weight_id],
                                                    w.has_bias ? pl->flat_g + pl->gb_off[op.weight_id] : nullptr, stream);
                    if (!rc && (op.src != pl->input_buf || input_grad)) {
                        // dX through the forward kernels (conv2d_bwd_input_fwd): straight into the gradient window when
                        // nothing has accumulated there yet (it was zeroed above), else via scratch + add
                        const bool whole = op.src_c0 == 0 && op.src_c == s.d.C;
                        float* dst = (!pl->g_touched[op.src] && whole) ? gs : pl->bwd_scratch;
                        DlwpConvDesc dd = d;
                        if (dst != gs) { dd.x_stride_n = (long long)op.src_c * s_hw; }
                        int r2 = conv2d_bwd_input_fwd(dd, gt, w.k, dst, pl->wt_scratch, stream);
                        if (r2 == -1) rc = conv2d_bwd_input(d, gt, w.k, gs, stream);
                        else if (r2) rc = r2;
                        else if (dst != gs) {
                            const long long sc[3] = {(long long)op.src_c * s_hw, s_hw, s.d.W};
                            rc = add_bwd(pl->bwd_scratch, gs, N, op.src_c, s.d.H, s.d.W, sc, ss, stream);
                        }
                    }
                    pl->g_touched[op.src] = 1;
                    break;
                }
                case DLWP_OP_MAXPOOL: rc = maxpool_bwd(xs, gt, gs, N, op.src_c, t.d.H, t.d.W, ss, ts, ss, stream); pl->g_touched[op.src] = 1; break;
                case DLWP_OP_UPSAMPLE: rc = upsample_bwd(gt, gs, N, op.src_c, s.d.H, s.d.W, ts, ss, stream); pl->g_touched[op.src] = 1; break;
                case DLWP_OP_COPY: rc = add_bwd(gt, gs, N, op.src_c, s.d.H, s.d.W, ts, ss, stream); pl->g_touched[op.src] = 1; break;
                default: DLWP_REQUIRE(false, DLWP_ESHAPE, "op %d has no backward", i);
            }
            if (rc) return rc;
        }
    }
    if (pl->loss_kind == 1) {
        std::vector<double> hs(5 * n_out);
        DLWP_CUDA_TRY(cudaMemcpyAsync(hs.data(), pl->acc_stats, sizeof(double) * 5 * n_out, cudaMemcpyDeviceToHost, stream));
        DLWP_CUDA_TRY(cudaStreamSynchronize(stream));
        for (int k = 0; k < n_out; ++k) {
            const double nk = (double)N * pl->buffers[pl->outputs[k]].sample_elems();
            const double* v = hs.data() + 5 * k;
            const double a = v[0] / sqrt(v[1] * v[2]);
            const double m = pl->acc_regularize == 1 ? v[3] / nk : (pl->acc_regularize == 2 ? v[4] / nk : 0.0);
            losses[k] = (float)(pl->acc_reverse ? m - a : a - m);
            if (maes) maes[k] = (float)(v[4] / nk);
        }
        return 0;
    }
    std::vector<float> h(2 * n_out);
    DLWP_CUDA_TRY(cudaMemcpyAsync(h.data(), pl->stats, sizeof(float) * 2 * n_out, cudaMemcpyDeviceToHost, stream));
    DLWP_CUDA_TRY(cudaStreamSynchronize(stream));
    for (int k = 0; k < n_out; ++k) {
        const double nk = (double)N * pl->buffers[pl->outputs[k]].sample_elems();
        losses[k] = (float)(h[2 * k] / nk);
        if (maes) maes[k] = (float)(h[2 * k + 1] / nk);
    }
    return 0;
}

extern "C" int dlwp_train_loss_weights(DlwpPlan* pl, const float* wmap_host, int64_t elems) {
    DLWP_REQUIRE(pl != nullptr, DLWP_EINVAL, "null plan");
    if (pl->loss_wmap) cudaFree(pl->loss_wmap);
    pl->loss_wmap = nullptr;
    if (!wmap_host) return 0;
    const Buffer& ob = pl->buffers[pl->outputs[0]];
    DLWP_REQUIRE(elems == (int64_t)ob.d.H * ob.d.W, DLWP_ESHAPE, "loss weight map must have H*W = %d elements",
                 ob.d.H * ob.d.W);
    DLWP_CUDA_TRY(cudaMalloc(&pl->loss_wmap, sizeof(float) * elems));
    DLWP_CUDA_TRY(cudaMemcpy(pl->loss_wmap, wmap_host, sizeof(float) * elems, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int dlwp_train_loss_kind(DlwpPlan* pl, int32_t kind, int32_t regularize, int32_t reverse, const float* mean_host,
                                    int64_t mean_elems) {
    DLWP_REQUIRE(pl != nullptr, DLWP_EINVAL, "null plan");
    DLWP_REQUIRE(kind == 0 || kind == 1, DLWP_EINVAL, "loss kind %d (0 = MSE, 1 = anomaly correlation)", kind);
    DLWP_REQUIRE(regularize >= 0 && regularize <= 2, DLWP_EINVAL, "regularize %d (0 none, 1 mse, 2 mae)", regularize);
    DLWP_REQUIRE(regularize == 0 || reverse, DLWP_EINVAL, "a regularized ACC loss is always reversed (custom.py:1053-1055)");
    if (pl->acc_mean) cudaFree(pl->acc_mean);
    pl->acc_mean = nullptr;
    pl->acc_mean_elems = 0;
    pl->loss_kind = kind;
    pl->acc_regularize = regularize;
    pl->acc_reverse = reverse ? 1 : 0;
    if (kind == 1 && mean_host) {
        const Buffer& ob = pl->buffers[pl->outputs[0]];
        DLWP_REQUIRE(mean_elems == (int64_t)ob.sample_elems(), DLWP_ESHAPE, "the mean must have C*H*W = %lld elements",
                     ob.sample_elems());
        DLWP_CUDA_TRY(cudaMalloc(&pl->acc_mean, sizeof(float) * mean_elems));
        DLWP_CUDA_TRY(cudaMemcpy(pl->acc_mean, mean_host, sizeof(float) * mean_elems, cudaMemcpyHostToDevice));
        pl->acc_mean_elems = mean_elems;
    }
    return 0;
}

extern "C" int dlwp_train_allreduce(DlwpPlan* pl, void* comm, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && pl->train_ready, DLWP_ESTATE, "run dlwp_train_step first");
    DLWP_REQUIRE(comm != nullptr, DLWP_EINVAL, "null communicator");
    DLWP_REQUIRE(g_nccl.AllReduce != nullptr, DLWP_EARCH, "the loaded NCCL library has no ncclAllReduce");
    // ncclFloat32 = 7, ncclAvg = 4 (NCCL >= 2.10): the mean over the ranks, in place, ordered on `stream` after the backward
    DLWP_NCCL_TRY(g_nccl.AllReduce(pl->flat_g, pl->flat_g, (size_t)pl->flat_elems, 7, 4, comm, (cudaStream_t)stream_));
    return 0;
}

extern "C" int dlwp_train_buffers(DlwpPlan* pl, float** flat_grad, int64_t* elems, float** input_grad) {
    DLWP_REQUIRE(pl && pl->train_ready, DLWP_ESTATE, "run dlwp_train_step first");
    if (flat_grad) *flat_grad = pl->flat_g;
    if (elems) *elems = pl->flat_elems;
    if (input_grad) *input_grad = pl->gbuf[pl->input_buf];
    return 0;
}

extern "C" int dlwp_train_weight_offsets(DlwpPlan* pl, int32_t weight_id, int64_t* kernel_off, int64_t* bias_off) {
    DLWP_REQUIRE(pl && pl->train_ready && weight_id >= 0 && weight_id < (int)pl->weights.size(), DLWP_ESTATE, "bad state");
    *kernel_off = pl->gk_off[weight_id];
    *bias_off = pl->gb_off[weight_id];
    return 0;
}

extern "C" int dlwp_train_regularize(DlwpPlan* pl, int32_t weight_id, float kernel_l1, float kernel_l2, float bias_l1,
                                     float bias_l2, float* penalty, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && pl->train_ready && penalty, DLWP_ESTATE, "run dlwp_train_step first");
    DLWP_REQUIRE(weight_id >= 0 && weight_id < (int)pl->weights.size(), DLWP_EINVAL, "weight id %d out of range", weight_id);
    cudaStream_t stream = (cudaStream_t)stream_;
    Weight& W = pl->weights[weight_id];
    DLWP_CUDA_TRY(cudaMemsetAsync(pl->stats, 0, sizeof(float), stream));
    int rc = 0;
    if (kernel_l1 != 0.f || kernel_l2 != 0.f)
        rc = regularize_grad(W.k, pl->flat_g + pl->gk_off[weight_id], W.k_elems, kernel_l1, kernel_l2, pl->stats, stream);
    if (!rc && W.has_bias && (bias_l1 != 0.f || bias_l2 != 0.f))
        rc = regularize_grad(W.b, pl->flat_g + pl->gb_off[weight_id], W.b_elems, bias_l1, bias_l2, pl->stats, stream);
    if (rc) return rc;
    DLWP_CUDA_TRY(cudaMemcpyAsync(penalty, pl->stats, sizeof(float), cudaMemcpyDeviceToHost, stream));
    DLWP_CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int dlwp_train_adam_state(DlwpPlan* pl, float* m_host, float* v_host, int64_t elems, int64_t* step,
                                     int32_t set) {
    DLWP_REQUIRE(pl && m_host && v_host && step, DLWP_EINVAL, "null argument");
    if (set) {
        int rc = train_setup(pl);
        if (rc) return rc;
    }
    DLWP_REQUIRE(pl->train_ready, DLWP_ESTATE, "the plan has no optimizer state yet");
    DLWP_REQUIRE(elems == pl->flat_elems, DLWP_ESHAPE, "optimizer state has %lld elements, expected %lld",
                 (long long)elems, pl->flat_elems);
    DLWP_CUDA_TRY(cudaDeviceSynchronize());
    const cudaMemcpyKind kind = set ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    DLWP_CUDA_TRY(cudaMemcpy(set ? (void*)pl->flat_m : (void*)m_host, set ? (void*)m_host : (void*)pl->flat_m,
                             sizeof(float) * elems, kind));
    DLWP_CUDA_TRY(cudaMemcpy(set ? (void*)pl->flat_v : (void*)v_host, set ? (void*)v_host : (void*)pl->flat_v,
                             sizeof(float) * elems, kind));
    if (set) pl->adam_t = *step;
    else *step = pl->adam_t;
    return 0;
}

extern "C" int dlwp_train_adam(DlwpPlan* pl, float lr, float beta1, float beta2, float eps, dlwp_stream_t stream_) {
    DLWP_REQUIRE(pl && pl->train_ready, DLWP_ESTATE, "run dlwp_train_step first");
    cudaStream_t stream = (cudaStream_t)stream_;
    pl->adam_t += 1;
    const double t = (double)pl->adam_t;
    const float lr_t = (float)(lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
    for (size_t w = 0; w < pl->weights.size(); ++w) {
        Weight& W = pl->weights[w];
        if (!W.k_elems) continue;
        int rc = adam_step(W.k, pl->flat_g + pl->gk_off[w], pl->flat_m + pl->gk_off[w], pl->flat_v + pl->gk_off[w],
                           W.k_elems, lr_t, beta1, beta2, eps, stream);
        if (!rc && W.has_bias)
            rc = adam_step(W.b, pl->flat_g + pl->gb_off[w], pl->flat_m + pl->gb_off[w], pl->flat_v + pl->gb_off[w], W.b_elems,
                           lr_t, beta1, beta2, eps, stream);
        if (rc) return rc;
    }
    return 0;
}
