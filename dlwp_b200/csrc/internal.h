// Internal declarations shared by the translation units of libdlwp_b200.so.  Not part of the ABI.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/dlwp_b200.h"

namespace dlwp {

// ---- error plumbing ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define DLWP_CUDA_TRY(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::dlwp::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                 \
                              cudaGetErrorString(_e));                                             \
            return (int)_e;                                                                        \
        }                                                                                          \
    } while (0)

#define DLWP_REQUIRE(cond, code, ...)                                                              \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            ::dlwp::set_error(__VA_ARGS__);                                                        \
            return (code);                                                                         \
        }                                                                                          \
    } while (0)

inline int after_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

// ---- kernel launchers (conv.cu / elementwise.cu) ------------------------------------------------------------------
int conv2d_fwd(const DlwpConvDesc& d, const float* x, const float* w, const float* bias, float* y,
               cudaStream_t stream);
const char* conv2d_impl_name(const DlwpConvDesc& d);
int check_device();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute: a process that drives several GPUs
// (keras multi_gpu_model: one plan per device) has to opt in on each of them.  Returns true the first time it is called for
// (this mask, current device).
#include <atomic>
inline bool first_use_on_device(std::atomic<unsigned long long>& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}
int plan_flags_read_clear();   // plan.cu: bit 0 = a latitude-band halo wait timed out
int halo_copy(const float* const src[2], float* const dst[2], const int rows[2], const long long ss_n[2],
              const long long ss_c[2], const long long ds_n[2], const long long ds_c[2], int N, int C, int W,
              cudaStream_t stream);

// ---- training kernels (train.cu); stride arrays are {n, c, h} element strides -------------------------------------------
int conv2d_bwd_input(const DlwpConvDesc& d, const float* dy, const float* w, float* dx, cudaStream_t stream);
int conv2d_bwd_weight(const DlwpConvDesc& d, const float* x, const float* dy, float* dw, float* db, cudaStream_t stream);
int conv2d_bwd_input_fwd(const DlwpConvDesc& d, const float* dy, const float* w, float* dx, float* wt, cudaStream_t stream);
int act_bwd(const float* y, float* g, int act, int N, int C, int H, int W, const long long* ys, const long long* gs,
            cudaStream_t stream);
int maxpool_bwd(const float* x, const float* g, float* dx, int N, int C, int Hp, int Wp, const long long* xs,
                const long long* gs, const long long* dxs, cudaStream_t stream);
int upsample_bwd(const float* g, float* dx, int N, int C, int Hl, int Wl, const long long* gs, const long long* dxs,
                 cudaStream_t stream);
int add_bwd(const float* g, float* dx, int N, int C, int H, int W, const long long* gs, const long long* dxs,
            cudaStream_t stream);
int acc_loss_grad(const float* yhat, const float* y, const float* mean, long long per, float* g, long long n, double* stats,
                  float scale, int regularize, cudaStream_t stream);
int mse_grad(const float* yhat, const float* y, float* g, long long n, float scale, float* stats, cudaStream_t stream,
             const float* wmap = nullptr, long long hw = 1);
int regularize_grad(const float* w, float* g, long long n, float l1, float l2, float* stat, cudaStream_t stream);
int adam_step(float* w, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
              cudaStream_t stream);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
int encode_tensor_map_4d(CUtensorMap* map, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4]);
int encode_tensor_map_u64(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                          const uint32_t* box);
int encode_tensor_map_any(CUtensorMap* map, const void* base, int is_f32, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace dlwp

// ---- device helpers -------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace dlwp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// tanh(x) as the degree-13 / degree-6 odd rational of x on [-9, 9] (the float coefficients of Eigen's fast tanh) with an
// approximate reciprocal: relative error <= 4e-7 over the whole range INCLUDING tiny |x| (scripts/tanh_check.py sweeps
// 2.4e6 points against float64).  Round 1 used 1 - 2 / (1 + 2^(2x log2 e)): 5 instructions, absolute error 2e-7 -- fine for
// O(1) pre-activations but with no relative accuracy left once they are small (x = 1e-4: 1.7e-4 relative), which broke the
// per-layer parity bar for down-scaled inputs / weights.  tanh.approx.f32 (2^-11) would not pass the 1e-4 / 50-step gate.
__device__ __forceinline__ float tanh_accurate(float x) {
    const float x_in = x;
    x = fminf(fmaxf(x, -9.0f), 9.0f);   // (drops a NaN; restored below)
    const float x2 = x * x;
    float p = fmaf(x2, -2.76076847742355e-16f, 2.00018790482477e-13f);
    p = fmaf(x2, p, -8.60467152213735e-11f);
    p = fmaf(x2, p, 5.12229709037114e-08f);
    p = fmaf(x2, p, 1.48572235717979e-05f);
    p = fmaf(x2, p, 6.37261928875436e-04f);
    p = fmaf(x2, p, 4.89352455891786e-03f);
    float q = fmaf(x2, 1.19825839466702e-06f, 1.18534705686654e-04f);
    q = fmaf(x2, q, 2.26843463243900e-03f);
    q = fmaf(x2, q, 4.89352518554385e-03f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
    return x_in != x_in ? x_in : x * p * r;   // tanh(NaN) = NaN, as in the reference's fp32 arithmetic
}

// Two tanh at once on packed fp32 pairs (sm_100 fma.rn.f32x2, SASS FFMA2): same arithmetic, half the FMA instructions.
// The epilogues of the tanh layers are issue / store bound, so the polynomial has to be cheap.
__device__ __forceinline__ void tanh_accurate2(float& a, float& b) {
    typedef unsigned long long f2;
    auto pk = [](float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; };
    auto fma2 = [](f2 x, f2 y, f2 z) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z)); return d; };
    const float a_in = a, b_in = b;
    a = fminf(fmaxf(a, -9.0f), 9.0f);
    b = fminf(fmaxf(b, -9.0f), 9.0f);
    const f2 x = pk(a, b), zero = pk(0.f, 0.f);
    const f2 x2 = fma2(x, x, zero);
    f2 p = fma2(x2, pk(-2.76076847742355e-16f, -2.76076847742355e-16f), pk(2.00018790482477e-13f, 2.00018790482477e-13f));
    p = fma2(x2, p, pk(-8.60467152213735e-11f, -8.60467152213735e-11f));
    p = fma2(x2, p, pk(5.12229709037114e-08f, 5.12229709037114e-08f));
    p = fma2(x2, p, pk(1.48572235717979e-05f, 1.48572235717979e-05f));
    p = fma2(x2, p, pk(6.37261928875436e-04f, 6.37261928875436e-04f));
    p = fma2(x2, p, pk(4.89352455891786e-03f, 4.89352455891786e-03f));
    f2 q = fma2(x2, pk(1.19825839466702e-06f, 1.19825839466702e-06f), pk(1.18534705686654e-04f, 1.18534705686654e-04f));
    q = fma2(x2, q, pk(2.26843463243900e-03f, 2.26843463243900e-03f));
    q = fma2(x2, q, pk(4.89352518554385e-03f, 4.89352518554385e-03f));
    p = fma2(x, p, zero);
    float q0, q1, p0, p1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(p));
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
    a = a_in != a_in ? a_in : p0 * r0;
    b = b_in != b_in ? b_in : p1 * r1;
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == DLWP_ACT_TANH) return tanh_accurate(v);
    if (act == DLWP_ACT_RELU) return fmaxf(v, 0.0f);
    return v;
}

// floor-mod for possibly negative a, b > 0
__device__ __forceinline__ int wrap_index(int a, int b) {
    int m = a % b;
    return m < 0 ? m + b : m;
}

// ---- cp.async (LDGSTS) ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_4(uint32_t dst, const float* src, bool valid) {
    const int sz = valid ? 4 : 0;  // src-size 0 => destination is zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const float* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- mbarrier + TMA (conv.cu only: g_device_flags lives there) ----------------------------------------------------
#ifdef DLWP_CONV_TU
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe (mbarrier.test_wait): issued a row ahead by the MMA issuer, consumed after that row's MMAs have been
// issued, so the barrier's round trip overlaps the issue stream instead of heading every row.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never completes (bad descriptor) raises a flag the host can read (dlwp_debug_flags) and
// lets the kernel finish with garbage instead of hanging the GPU.
// try_wait with a suspend-time hint: the warp may sleep up to ~hint ns per attempt instead of re-polling (idle epilogue /
// producer warps otherwise burn issue slots and power in the poll loop while the MMA issuer is the critical path).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if (++spins > (1u << 17)) {  // bounded (<~ 3 s): a broken pipeline raises the flag instead of hanging the GPU
            atomicOr(&g_device_flags, 1);
            break;
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) {
            atomicOr(&g_device_flags, 1);
            break;
        }
    }
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
#endif  // DLWP_CONV_TU

}  // namespace dlwp
#endif
