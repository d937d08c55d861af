// FP32 pipe probe: what does a clean register-tiled outer product (the inner loop shape of conv_ffma_kernel) reach
// with scalar FFMA vs packed fma.rn.f32x2 on sm_100a?  Prints TFLOP/s for a few tile shapes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int P, int C>
__global__ void __launch_bounds__(256) k_scalar(const float* __restrict__ in, float* out, int iters) {
    float acc[P][C];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < C; ++c) acc[p][c] = 0.f;
    __shared__ float sv[256 * 4], sw[64];
    for (int i = threadIdx.x; i < 256 * 4; i += 256) sv[i] = in[i];
    if (threadIdx.x < 64) sw[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        float v[P], w[C];
#pragma unroll
        for (int p = 0; p < P; ++p) v[p] = sv[(threadIdx.x * 4 + p + it) & 1023];
#pragma unroll
        for (int c = 0; c < C; ++c) w[c] = sw[(c + it) & 63];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[p][c] = fmaf(v[p], w[c], acc[p][c]);
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < C; ++c) s += acc[p][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int P, int C>  // C even: acc[p][c/2] holds filters (c, c+1)
__global__ void __launch_bounds__(256) k_packed(const float* __restrict__ in, float* out, int iters) {
    unsigned long long acc[P][C / 2];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < C / 2; ++c) acc[p][c] = 0ull;
    __shared__ float sv[256 * 4], sw[64];
    for (int i = threadIdx.x; i < 256 * 4; i += 256) sv[i] = in[i];
    if (threadIdx.x < 64) sw[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        unsigned long long v2[P], w2[C / 2];
#pragma unroll
        for (int p = 0; p < P; ++p) { float t = sv[(threadIdx.x * 4 + p + it) & 1023]; v2[p] = pack2(t, t); }
#pragma unroll
        for (int c = 0; c < C / 2; ++c) w2[c] = pack2(sw[(2 * c + it) & 63], sw[(2 * c + 1 + it) & 63]);
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int c = 0; c < C / 2; ++c) acc[p][c] = fma2(v2[p], w2[c], acc[p][c]);
    }
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int c = 0; c < C / 2; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[p][c])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
void run(const char* name, K kern, int P, int C, const float* in, float* out) {
    int iters = 4096, blocks = 148 * 4;
    kern<<<blocks, 256>>>(in, out, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(in, out, iters);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * P * C * (double)iters * blocks * 256;
    printf("%-28s P=%d C=%2d  %.3f ms  %.1f TFLOP/s  (%s)\n", name, P, C, ms, flops / ms / 1e9, cudaGetErrorString(e));
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 4 * 256 * 4);
    cudaMemset(in, 0, 4096 * 4);
    run("scalar FFMA", k_scalar<8, 8>, 8, 8, in, out);
    run("scalar FFMA", k_scalar<8, 6>, 8, 6, in, out);
    run("scalar FFMA", k_scalar<8, 16>, 8, 16, in, out);
    run("scalar FFMA", k_scalar<4, 8>, 4, 8, in, out);
    run("packed fma.rn.f32x2", k_packed<8, 8>, 8, 8, in, out);
    run("packed fma.rn.f32x2", k_packed<8, 6>, 8, 6, in, out);
    run("packed fma.rn.f32x2", k_packed<8, 16>, 8, 16, in, out);
    run("packed fma.rn.f32x2", k_packed<4, 8>, 4, 8, in, out);
    return 0;
}
