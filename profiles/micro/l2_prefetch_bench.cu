// l2_prefetch_bench.cu - does prefetching the weight stream into L2 during a serial section pay?
// Emulates the persistent kernel's phase structure: every CTA (one per SM) streams B bytes per phase from HBM through a
// shared-memory ring of bulk copies; between two phases there is a "serial section" of S microseconds during which the ring
// producer is gated (only the ring's capacity is prefetched after the gate opens at the START of the section, as in the
// kernel: the gate opens when the input has been polled in, the rebuild then takes S).  Mode 1 additionally lets the producer
// issue cp.async.bulk.prefetch.L2 for up to D bytes ahead of its ring cursor at all times (across phase boundaries and while
// gated), so HBM keeps working during the serial section and the drain is then served from L2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_prefetch_bench l2_prefetch_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* b, uint32_t par) {
    uint32_t ok = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t lds_v(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p))); return v; }
__device__ __forceinline__ void sts_v(uint32_t* p, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory"); }

constexpr int NS = 23;
constexpr uint32_t SB = 7680;
constexpr uint32_t PFC = 7680;          // bytes per L2 prefetch instruction

// phase i of CTA c reads [base + (i * gridDim.x + c) * B, + B)
__global__ void __launch_bounds__(64, 1) k(const uint8_t* base, uint32_t B, int n_phases, int serial_ns, uint32_t D, int consume_ns_per_stage,
                                          unsigned long long* out_ns, unsigned long long* drain_ns) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + NS;
    uint32_t* gate = reinterpret_cast<uint32_t*>(smem + 512);
    uint8_t* ring = smem + 1024;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        *gate = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t nst = B / SB;
    const unsigned long long t0 = gtimer();
    if (tid == 0) {
        // producer
        uint32_t slot = 0, par = 1;
        // L2 cursor: (phase, offset)
        int pph = 0; uint32_t poff = 0; long long ahead = 0;       // bytes prefetched into L2 and not yet requested by the ring
        auto pump = [&]() {
            while (D && ahead + PFC <= (long long)D && pph < n_phases) {
                l2_prefetch(base + ((size_t)pph * gridDim.x + blockIdx.x) * B + poff, PFC);
                poff += PFC; ahead += PFC;
                if (poff >= nst * SB) { poff = 0; ++pph; }
            }
        };
        for (int i = 0; i < n_phases; ++i) {
            while ((int)(lds_v(gate) - (uint32_t)(i + 1)) < 0) { pump(); __nanosleep(64); }
            const uint8_t* src = base + ((size_t)i * gridDim.x + blockIdx.x) * B;
            for (uint32_t s = 0; s < nst; ++s) {
                pump();
                while (!mbar_test(&empty[slot], par)) { pump(); __nanosleep(32); }
                mbar_expect(&full[slot], SB);
                bulk_g2s(ring + (size_t)slot * SB, src + (size_t)s * SB, SB, &full[slot]);
                ahead -= SB; if (ahead < 0) { ahead = 0; pph = i; poff = (s + 1) * SB; if (poff >= nst * SB) { poff = 0; ++pph; } }
                if (++slot == NS) { slot = 0; par ^= 1u; }
            }
        }
    } else if (tid == 32) {
        // consumer
        uint32_t slot = 0, par = 0;
        unsigned long long dsum = 0;
        for (int i = 0; i < n_phases; ++i) {
            // "input polled in": gate opens, then the serial section (rebuild) runs
            sts_v(gate, (uint32_t)(i + 1));
            const unsigned long long ts = gtimer();
            while (gtimer() - ts < (unsigned long long)serial_ns) __nanosleep(100);
            const unsigned long long td = gtimer();
            for (uint32_t s = 0; s < nst; ++s) {
                mbar_wait(&full[slot], par);
                if (consume_ns_per_stage) { const unsigned long long tc = gtimer(); while (gtimer() - tc < (unsigned long long)consume_ns_per_stage) { } }
                mbar_arrive(&empty[slot]);
                if (++slot == NS) { slot = 0; par ^= 1u; }
            }
            dsum += gtimer() - td;
        }
        out_ns[blockIdx.x] = gtimer() - t0;
        drain_ns[blockIdx.x] = dsum;
    }
}

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    const int nsm = pr.multiProcessorCount;
    const size_t smem = 1024 + (size_t)NS * SB;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_phases = 48;
    struct PhaseCfg { const char* name; uint32_t stages; } cfgs[] = { {"W13 (79 stages, 607 KB/SM)", 79}, {"QKV (44 stages, 338 KB/SM)", 44}, {"W2 (40 stages, 307 KB/SM)", 40} };
    const size_t maxB = 79 * (size_t)SB;
    const size_t total = maxB * nsm * n_phases;
    uint8_t* buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 1, total));
    unsigned long long *d_ns, *d_dr; CK(cudaMalloc(&d_ns, nsm * 8)); CK(cudaMalloc(&d_dr, nsm * 8));
    std::vector<unsigned long long> h(nsm), hd(nsm);
    for (auto& cf : cfgs) {
        const uint32_t B = cf.stages * SB;
        printf("== phase %s; ideal stream at 6.9 TB/s: %.2f us per phase\n", cf.name, (double)B * nsm / 6.9e6);
        for (int serial_us : {0, 4, 8, 12}) {
            for (uint32_t Dk : {0u, 120u, 240u, 360u, 480u}) {
                double best = 1e30, bestd = 0;
                for (int rep = 0; rep < 3; ++rep) {
                    k<<<nsm, 64, smem>>>(buf, B, n_phases, serial_us * 1000, Dk * 1024u, 40, d_ns, d_dr);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(h.data(), d_ns, nsm * 8, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(hd.data(), d_dr, nsm * 8, cudaMemcpyDeviceToHost));
                    const double t = (double)*std::max_element(h.begin(), h.end()) / n_phases / 1000.0;
                    std::sort(hd.begin(), hd.end());
                    if (t < best) { best = t; bestd = (double)hd[nsm / 2] / n_phases / 1000.0; }
                }
                printf("serial %2d us, L2 prefetch depth %3u KB/SM: %.2f us per phase (drain %.2f us, %.0f GB/s per SM in the drain)\n", serial_us, Dk, best, bestd,
                       (double)B / bestd / 1000.0);
            }
        }
    }
    return 0;
}
