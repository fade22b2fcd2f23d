// consumer_bench.cu — how fast can the 8 consumer warps of one CTA drain ring stages that are already in shared memory?
//   variant 0: stage_chain exactly as the persistent kernel runs it (integer dots with dp4a + FP32 group chain)
//   variant 1: the same without the FP32 chain pass (pass 1 only)   variant 2: dp4a issue rate alone (independent IDP4A)
//   variant 3: IMAD issue rate alone (reference)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o consumer_bench consumer_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../fast-llama_b200/csrc/megakernel.cuh"
using namespace fl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int QT, int GS>
__global__ void __launch_bounds__(288, 1) k(int variant, int iters, int n_slots, long long* cyc, float* sink) {
    using T = Traits<QT, GS>; using R = Ring<QT, GS>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* xq = smem;                                   // 11264 B
    float* xs = reinterpret_cast<float*>(smem + 11264);   // 1024 B
    float* csb = reinterpret_cast<float*>(smem + 12288);  // 8 KB
    uint8_t* ring = smem + 12288 + 8192;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) return;
    for (int i = tid; i < (12288 + 8192 + n_slots * R::SLOT_BYTES) / 4; i += 256) {
        uint32_t v = (uint32_t)(i * 2654435761u) ^ 0x5bd1e995u;
        if (i * 4 >= 12288 + 8192 && ((i * 4 - 12288 - 8192) % T::UNIT_BYTES) >= T::W_BYTES) v = __float_as_uint(0.01f + (v & 255) * 1e-4f);   // scale area
        reinterpret_cast<uint32_t*>(smem)[i] = v;
    }
    for (int i = tid; i < 256; i += 256) xs[i] = 0.02f;
    consumer_sync();
    float* cs = csb + (size_t)warp * (R::U * 32 * 2 * T::GPL);
    const uint4* xq4 = reinterpret_cast<const uint4*>(xq);
    float acc = 0.f; int iacc = lane;
    long long t0 = clock64();
    if (variant <= 1) {
        for (int it = 0; it < iters; ++it) {
            const int slot = (it * 8 + warp) % n_slots;
            const int kb = (it & 1) * R::U;
            acc = stage_chain<QT, GS>(ring + (size_t)slot * R::SLOT_BYTES, variant == 0 ? R::U : 0, xq4 + (size_t)kb * (T::KB_BYTES / 16), xs + kb * 8 * T::GPL, cs, lane, acc);
            __syncwarp();
        }
    } else if (variant == 2) {
        int a0 = lane, a1 = lane + 1, a2 = lane + 2, a3 = lane + 3, a4 = lane + 4, a5 = lane + 5, a6 = lane + 6, a7 = lane + 7;
        const int w = 0x01020304 * (lane + 1), x = 0x04030201 + lane;
        for (int it = 0; it < iters * 8; ++it) {
            a0 = __dp4a(w, x, a0); a1 = __dp4a(w, x, a1); a2 = __dp4a(w, x, a2); a3 = __dp4a(w, x, a3);
            a4 = __dp4a(w, x, a4); a5 = __dp4a(w, x, a5); a6 = __dp4a(w, x, a6); a7 = __dp4a(w, x, a7);
        }
        iacc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    } else {
        int a0 = lane, a1 = lane + 1, a2 = lane + 2, a3 = lane + 3, a4 = lane + 4, a5 = lane + 5, a6 = lane + 6, a7 = lane + 7;
        const int w = 0x01020304 * (lane + 1);
        for (int it = 0; it < iters * 8; ++it) {
            a0 = a0 * w + 1; a1 = a1 * w + 2; a2 = a2 * w + 3; a3 = a3 * w + 4; a4 = a4 * w + 5; a5 = a5 * w + 6; a6 = a6 * w + 7; a7 = a7 * w + 8;
        }
        iacc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 1.2345f || iacc == 123456789) sink[0] = acc + iacc;
}

int main() {
    int G = 148; long long* cyc; float* sink;
    CK(cudaMalloc(&cyc, G * 8 * 8)); CK(cudaMalloc(&sink, 64));
    const int n_slots = 20, iters = 2000;
    auto run = [&](auto kern, const char* name, int slot_bytes, int units) {
        const size_t smem = 12288 + 8192 + (size_t)n_slots * slot_bytes;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int v = 0; v < 4; ++v) {
            kern<<<G, 288, smem>>>(v, iters, n_slots, cyc, sink);
            CK(cudaDeviceSynchronize());
            std::vector<long long> h(G * 8); CK(cudaMemcpy(h.data(), cyc, G * 64, cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.end());
            const double c = (double)h[h.size() / 2] / iters;
            if (v <= 1) printf("%s variant %d: %.0f cycles per stage per warp (8 warps concurrently) -> %.1f B/cycle/SM = %.0f GB/s/SM at 1.965 GHz\n", name, v, c, 8.0 * slot_bytes / c, 8.0 * slot_bytes / c * 1.965);
            else printf("%s variant %d (%s): %.2f cycles per warp-instruction per warp, 2 warps per scheduler -> rt_SMSP = %.2f\n", name, v, v == 2 ? "IDP4A" : "IMAD", c / 64.0, c / 64.0 / 2.0);
        }
    };
    run(k<Q_INT8, 64>, "int8/g64", Ring<Q_INT8, 64>::SLOT_BYTES, 4);
    run(k<Q_INT16, 64>, "int16/g64", Ring<Q_INT16, 64>::SLOT_BYTES, 2);
    return 0;
}
