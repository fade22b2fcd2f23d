// pv_bench.cu — cycles per row of the single-warp PV chain (pv_rows<32>) with the rest of the CTA (a) parked at a barrier,
// (b) one warp spinning on an mbarrier try_wait + nanosleep (what the TMA producer does while its ring is full)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o pv_bench pv_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../fast-llama_b200/csrc/megakernel.cuh"
using namespace fl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(320, 1) k(int variant, int which_warp, int n_rows, long long* cyc, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* att = reinterpret_cast<float*>(smem);              // 4 KB
    float* v = att + 1024;                                    // 9 chunks x 32 rows x 32 dims
    uint64_t* bar = reinterpret_cast<uint64_t*>(v + 9 * 1024);
    volatile int* stop = reinterpret_cast<volatile int*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 1024 + 9 * 1024; i += 320) reinterpret_cast<float*>(smem)[i] = 0.001f * (float)((i * 37) % 101) + 0.01f;
    if (tid == 0) { mbar_init(bar, 1); *stop = 0; }
    __syncthreads();
    if (warp == 8) {
        if (variant == 1 && lane == 0) { while (!*stop) { if (!mbar_try(bar, 0)) __nanosleep(32); } }
        return;
    }
    if (warp == 9) return;
    float o = 0.f;
    long long t0 = 0, t1 = 0;
    if (warp == which_warp) {
        t0 = clock64();
        for (int rep = 0; rep < 8; ++rep) {
            const float* wp = att;
            for (int c = 0; c < n_rows / 32; ++c) { o = pv_rows(v + c * 1024 + lane * 4, wp, 32, c == 0 ? 1 : 0, 32, 0u, o); wp += 32; }
        }
        t1 = clock64();
        if (lane == 0) { cyc[blockIdx.x] = (t1 - t0) / 8; *stop = 1; }
    }
    consumer_sync();
    if (o == 1.2345f) sink[0] = o;
}

int main() {
    int G = 148; long long* cyc; float* sink;
    CK(cudaMalloc(&cyc, G * 8)); CK(cudaMalloc(&sink, 64));
    const size_t smem = 4096 + 9 * 4096 + 64;
    for (int variant = 0; variant < 2; ++variant) for (int w : {0, 7}) {
        k<<<G, 320, smem>>>(variant, w, 288, cyc, sink);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(G); CK(cudaMemcpy(h.data(), cyc, G * 8, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        printf("variant %d (%s) on warp %d: %lld cycles per 288 rows = %.1f cycles/row\n", variant, variant ? "producer warp spinning on try_wait" : "others parked", w, h[G / 2], h[G / 2] / 288.0);
    }
    return 0;
}
