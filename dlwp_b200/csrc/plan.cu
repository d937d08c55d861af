// The network plan and the device-resident rollout engine.
//
// A DlwpPlan is the executable form of the Keras graph that DLWPNeuralNet.build_model / DLWPFunctional.build_model
// compile (DLWP/model/models.py:96-112, :349-373): a list of ops over (N,C,H,W) buffers with channel windows, so that
// slice_layer (DLWP/custom.py:675-692) and concatenate(axis=1) are views.  dlwp_rollout replaces the Python feedback
// loops of DLWP/model/models.py:277-293 and :439-447: the forecast state never leaves HBM -- iteration t reads slot
// t*n_out-1 of the output series directly and writes slots t*n_out .. t*n_out+n_out-1 -- and the whole loop can be
// replayed as one CUDA graph.
#include "internal.h"

#include <stdarg.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace dlwp {

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
};

struct Weight {
    float* k = nullptr;
    float* b = nullptr;
    long long k_elems = 0, b_elems = 0;
    bool has_bias = false, set = false;
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
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    std::vector<cudaEvent_t> events;
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
        default: break;
    }
}

static int run_ops(DlwpPlan* pl, int N, cudaStream_t stream) {
    for (size_t i = 0; i < pl->ops.size(); ++i) {
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
            default: DLWP_REQUIRE(false, DLWP_EINVAL, "op %zu: unknown kind %d", i, op.kind);
        }
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
    for (int t = t0; t < t1; ++t) {
        pl->buffers[pl->input_buf].ptr =
            const_cast<float*>(t == 0 ? x0 : series + ((long long)t * n_out - 1) * slot);
        for (int k = 0; k < n_out; ++k) pl->buffers[pl->outputs[k]].ptr = series + ((long long)t * n_out + k) * slot;
        int rc = run_ops(pl, N, stream);
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

extern "C" int dlwp_plan_create(const DlwpNetDesc* net, DlwpPlan** out) {
    DLWP_REQUIRE(net && out, DLWP_EINVAL, "null argument");
    *out = nullptr;
    DLWP_REQUIRE(net->n_buffers > 0 && net->n_ops > 0 && net->max_batch > 0 && net->buffers && net->ops,
                 DLWP_EINVAL, "empty network description");
    int rc = check_device();
    if (rc) return rc;
    DlwpPlan* pl = new DlwpPlan();
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
    *out = pl;
    return 0;
}

extern "C" void dlwp_plan_destroy(DlwpPlan* pl) {
    if (!pl) return;
    for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second);
    for (Buffer& b : pl->buffers)
        if (b.d.kind == DLWP_BUF_INTERNAL && b.ptr) cudaFree(b.ptr);
    for (Weight& w : pl->weights) {
        if (w.k) cudaFree(w.k);
        if (w.b) cudaFree(w.b);
    }
    if (pl->d_x0) cudaFree(pl->d_x0);
    if (pl->d_series) cudaFree(pl->d_series);
    for (cudaEvent_t e : pl->events) cudaEventDestroy(e);
    if (pl->s_compute) cudaStreamDestroy(pl->s_compute);
    if (pl->s_copy) cudaStreamDestroy(pl->s_copy);
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
    return run_ops(pl, N, (cudaStream_t)stream);
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
    if (pl->d_x0_cap < slot) {
        if (pl->d_x0) cudaFree(pl->d_x0);
        pl->d_x0 = nullptr; pl->d_x0_cap = 0;
        DLWP_CUDA_TRY(cudaMalloc(&pl->d_x0, sizeof(float) * slot));
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
