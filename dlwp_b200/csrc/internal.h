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

// ---- training kernels (train.cu); stride arrays are {n, c, h} element strides -------------------------------------------
int conv2d_bwd_input(const DlwpConvDesc& d, const float* dy, const float* w, float* dx, cudaStream_t stream);
int conv2d_bwd_weight(const DlwpConvDesc& d, const float* x, const float* dy, float* dw, float* db, cudaStream_t stream);
int act_bwd(const float* y, float* g, int act, int N, int C, int H, int W, const long long* ys, const long long* gs,
            cudaStream_t stream);
int maxpool_bwd(const float* x, const float* g, float* dx, int N, int C, int Hp, int Wp, const long long* xs,
                const long long* gs, const long long* dxs, cudaStream_t stream);
int upsample_bwd(const float* g, float* dx, int N, int C, int Hl, int Wl, const long long* gs, const long long* dxs,
                 cudaStream_t stream);
int add_bwd(const float* g, float* dx, int N, int C, int H, int W, const long long* gs, const long long* dxs,
            cudaStream_t stream);
int mse_grad(const float* yhat, const float* y, float* g, long long n, float scale, float* stats, cudaStream_t stream,
             const float* wmap = nullptr, long long hw = 1);
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

// tanh(x) = 1 - 2 / (1 + exp(2x)) with ex2.approx / rcp.approx: 5 instructions, absolute error ~2e-7 everywhere
// (ex2.approx is good to 2 ulp, rcp.approx to 1 ulp; saturates correctly to +-1).  tanh.approx.f32 (2^-11 relative)
// would not pass the 1e-4 / 50-step parity gate.
__device__ __forceinline__ float tanh_accurate(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));  // exp(2x) = 2^(2x log2 e)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
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
