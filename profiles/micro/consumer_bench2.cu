// consumer_bench2.cu — cost of one "row per lane" stage (stage_pairs) per warp, 8 warps per CTA, data already in shared memory.
// variants: 0 full; 1 no x loads (x constant); 2 no weight loads (w constant); 3 loads only (no dp4a); 4 full, but the loop
// body repeated 4x straight-line (instruction-cache pressure like the unrolled superblock loop of the kernel)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o consumer_bench2 consumer_bench2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../fast-llama_b200/csrc/megakernel.cuh"
using namespace fl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int QT, int GS, int V>
__device__ __forceinline__ void stage_var(const uint8_t* sp, int R, const uint4* xp, const float* xsp, int lane,
                                          float (&prod)[Rk<QT, GS>::GPS], float (&fdot)[Rk<QT, GS>::GPS]) {
    using RK = Rk<QT, GS>;
    const uint4* wp = reinterpret_cast<const uint4*>(sp) + lane;
    const float* ssp = reinterpret_cast<const float*>(sp + R * kStageRowBytes) + lane;
#pragma unroll
    for (int g = 0; g < RK::GPS; ++g) {
        int d = 0;
#pragma unroll
        for (int j = 0; j < RK::PPG; ++j) {
            uint4 w = (V == 2) ? make_uint4(lane, g, j, 7) : wp[(g * RK::PPG + j) * R];
            uint4 x = (V == 1) ? make_uint4(lane + 1, g + 3, j + 5, 9) : xp[g * RK::PPG + j];
            if (V == 3) d += (int)(w.x ^ x.x ^ w.y ^ x.y ^ w.z ^ x.z ^ w.w ^ x.w);
            else d += dot16<QT>(w, x, 0);
        }
        prod[g] = __fmul_rn(ssp[g * R], xsp[g]);
        fdot[g] = __int2float_rn(d);
    }
}

template <int QT, int GS>
__global__ void __launch_bounds__(288, 1) k(int variant, int iters, int n_slots, int R, long long* cyc, float* sink) {
    using RK = Rk<QT, GS>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* xq = smem;                                   // 11264 B
    float* xs = reinterpret_cast<float*>(smem + 11264);   // 1024 B
    uint8_t* ring = smem + 12288;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) return;
    for (int i = tid; i < (12288 + n_slots * RK::SLOT_BYTES) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = (uint32_t)(i * 2654435761u) ^ 0x5bd1e995u;
    consumer_sync();
    for (int i = tid; i < 256; i += 256) xs[i] = 0.02f;
    consumer_sync();
    const uint4* xq4 = reinterpret_cast<const uint4*>(xq);
    float acc = 0.f;
    const bool live = lane < R;
    long long t0 = clock64();
#define BODY(VV, IT) { const int slot = ((IT) * 8 + warp) % n_slots; const int kc = (IT) & 31; float pr[RK::GPS], fd[RK::GPS]; \
        if (live) { stage_var<QT, GS, VV>(ring + (size_t)slot * RK::SLOT_BYTES, R, xq4 + kc * RK::PIECES, xs + kc * RK::GPS, lane, pr, fd); \
        _Pragma("unroll") for (int g = 0; g < RK::GPS; ++g) acc = __fmaf_rn(pr[g], fd[g], acc); } __syncwarp(); }
    if (variant == 0) { for (int it = 0; it < iters; ++it) BODY(0, it) }
    else if (variant == 1) { for (int it = 0; it < iters; ++it) BODY(1, it) }
    else if (variant == 2) { for (int it = 0; it < iters; ++it) BODY(2, it) }
    else if (variant == 3) { for (int it = 0; it < iters; ++it) BODY(3, it) }
    else { for (int it = 0; it < iters; it += 4) { BODY(0, it) BODY(0, it + 1) BODY(0, it + 2) BODY(0, it + 3) } }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 1.2345f) sink[0] = acc;
}

int main() {
    int G = 148; long long* cyc; float* sink;
    CK(cudaMalloc(&cyc, G * 8 * 8)); CK(cudaMalloc(&sink, 64));
    const int n_slots = 20, iters = 2000;
    auto run = [&](auto kern, const char* name, int slot_bytes) {
        const size_t smem = 12288 + (size_t)n_slots * slot_bytes;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int R : {28, 32}) for (int v = 0; v < 5; ++v) {
            kern<<<G, 288, smem>>>(v, iters, n_slots, R, cyc, sink);
            CK(cudaDeviceSynchronize());
            std::vector<long long> h(G * 8); CK(cudaMemcpy(h.data(), cyc, G * 64, cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.end());
            const double c = (double)h[h.size() / 2] / iters;
            printf("%s R=%d variant %d: %.0f cycles per stage per warp (8 warps) -> %.1f B/cycle/SM = %.0f GB/s/SM\n", name, R, v, c, 8.0 * R * 272 / c, 8.0 * R * 272 / c * 1.965);
        }
    };
    run(k<Q_INT8, 64>, "int8/g64", Rk<Q_INT8, 64>::SLOT_BYTES);
    return 0;
}
