// Stand-alone probe of the TMA path used by conv_ffma_kernel<STAGE_TMA>: one 4-D box load with out-of-bounds
// coordinates into shared memory, copied back to global for inspection.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__device__ int g_flag;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int box_floats, int cx, int cy, int cc, int cn,
                      int use_prefetch) {
    extern __shared__ __align__(128) float smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ((box_floats + 31) / 32) * 32);
    if (threadIdx.x == 0) {
        if (use_prefetch) asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar)), "r"(box_floats * 4) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
            ::"r"(s32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(cx), "r"(cy), "r"(cc), "r"(cn), "r"(s32(bar)) : "memory");
    }
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
        if (++spins > (1u << 20)) { if (threadIdx.x == 0) g_flag = 1; break; }
    }
    for (int i = threadIdx.x; i < box_floats; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int W = 36, H = 23, C = 6, N = 3;
    int bw = 40, bh = 8, bc = 6;
    int cx = -2, cy = -2, cc = 0, cn = 1;
    int use_prefetch = argc > 1 ? atoi(argv[1]) : 1;
    if (argc > 2) { cx = atoi(argv[2]); cy = atoi(argv[3]); }
    std::vector<float> h((size_t)W * H * C * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    int box = bw * bh * bc;
    CK(cudaMalloc(&o, box * 4));
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    printf("entry point %p query %d\n", fp, (int)q);
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    cuuint32_t bx[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = ((encode_t)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 2;
    size_t smem = ((box + 31) / 32) * 32 * 4 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe<<<1, 128, smem>>>(map, o, box, cx, cy, cc, cn, use_prefetch);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int flag = 0;
    CK(cudaMemcpyFromSymbol(&flag, g_flag, sizeof(int)));
    std::vector<float> ho(box);
    CK(cudaMemcpy(ho.data(), o, box * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < bc; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
        int gx = cx + x, gy = cy + y, gc = cc + c;
        float ref = (gx >= 0 && gx < W && gy >= 0 && gy < H && gc < C) ? h[((size_t)(cn * C + gc) * H + gy) * W + gx] : 0.f;
        if (ho[(c * bh + y) * bw + x] != ref) { if (bad < 5) printf("mismatch c%d y%d x%d got %f want %f\n", c, y, x, ho[(c * bh + y) * bw + x], ref); ++bad; }
    }
    printf("timeout flag %d, mismatches %d of %d -> %s\n", flag, bad, box, (!flag && !bad) ? "TMA PROBE OK" : "TMA PROBE FAILED");
    return 0;
}
