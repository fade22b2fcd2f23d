// micro-benchmark: cost of the 4-lane sum-of-squares chain variants inside a 288-thread CTA (other warps parked at a barrier)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fast-llama_b200/csrc/kernels.cuh"
using namespace fl;

__device__ __forceinline__ float chain_v1(const float* xf, int n, int lane) {   // 8-batch, no double buffering
    float acc = 0.0f;
    if (lane < 4) {
        const float* p = xf + lane; int i = 0; const int steps = n / 4;
        for (; i + 8 <= steps; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[(i + u) * 4];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fmaf_rn(v[u], v[u], acc);
        }
    }
    return acc;
}
__device__ __forceinline__ float chain_v3(const float* xf, int n, int lane) {   // one lane, four interleaved chains, LDS.128
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (lane == 0) {
        const float4* p = reinterpret_cast<const float4*>(xf); const int steps = n / 4;
        float4 v[8];
        for (int i = 0; i < steps; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[i + u];
#pragma unroll
            for (int u = 0; u < 8; ++u) { a0 = __fmaf_rn(v[u].x, v[u].x, a0); a1 = __fmaf_rn(v[u].y, v[u].y, a1); a2 = __fmaf_rn(v[u].z, v[u].z, a2); a3 = __fmaf_rn(v[u].w, v[u].w, a3); }
        }
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(0.f, a0), a1), a2), a3);
}

__global__ void __launch_bounds__(288, 1) k(const float* x, int n, float* out, long long* cyc, int variant, int spin_producer) {
    extern __shared__ float xf[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) {
        if (spin_producer && lane == 0) { volatile int* f = (volatile int*)(xf + n); while (*f == 0) { } }
        return;
    }
    for (int i = tid; i < n; i += 256) xf[i] = x[i];
    if (tid == 0) *(volatile int*)(xf + n) = 0;
    asm volatile("bar.sync 1, 256;");
    long long t0 = clock64();
    float r = 0.f;
    if (warp == 0) {
        for (int rep = 0; rep < 8; ++rep) {
            if (variant == 1) r += chain_v1(xf, n, lane);
            else if (variant == 2) r += sumsq_chain_warp0(xf, n, lane);
            else r += chain_v3(xf, n, lane);
        }
    }
    long long t1 = clock64();
    asm volatile("bar.sync 1, 256;");
    if (tid == 0) { *(volatile int*)(xf + n) = 1; out[blockIdx.x] = r; cyc[blockIdx.x] = (t1 - t0) / 8; }
}

int main() {
    const int n = 4096;
    float* x; float* out; long long* cyc;
    cudaMalloc(&x, n * 4); cudaMalloc(&out, 148 * 4); cudaMalloc(&cyc, 148 * 8);
    float h[n]; for (int i = 0; i < n; ++i) h[i] = (float)(i % 17) * 0.01f;
    cudaMemcpy(x, h, n * 4, cudaMemcpyHostToDevice);
    for (int spin = 0; spin < 2; ++spin)
        for (int v = 1; v <= 3; ++v) {
            k<<<148, 288, (n + 32) * 4>>>(x, n, out, cyc, v, spin);
            cudaDeviceSynchronize();
            long long c[148]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
            long long mn = 1 << 30, mx = 0; for (int i = 0; i < 148; ++i) { if (c[i] < mn) mn = c[i]; if (c[i] > mx) mx = c[i]; }
            printf("variant %d spin_producer %d: cycles per chain(n=%d) min %lld max %lld  (%s)\n", v, spin, n, mn, mx, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
