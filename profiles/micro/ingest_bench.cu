// ingest_bench.cu — how fast can ONE SM pull bytes in?  (a) LDG.128 into registers, (b) cp.async.bulk into shared memory;
// source resident in L2 (small private region re-read) or streamed from HBM (large region); few or all SMs active.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ingest_bench ingest_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ldg_nc(const void* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode 0: LDG.128, 8 warps, 8 loads in flight per thread;  mode 1: cp.async.bulk ring (16 x 8 KB), one producer thread
__global__ void __launch_bounds__(288, 1) k(const uint8_t* base, size_t region, int passes, int mode, unsigned long long* out_ns, unsigned int* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint8_t* src = base + (size_t)blockIdx.x * region;
    const int tid = threadIdx.x;
    unsigned int acc = 0;
    __syncthreads();
    const unsigned long long t0 = gtimer();
    if (mode == 0) {
        if (tid < 256) {
            for (int p = 0; p < passes; ++p)
                for (size_t off = (size_t)tid * 16; off + 8 * 4096 <= region + (size_t)tid * 16 && off + 7 * 4096 + 16 <= region; off += 8 * 4096) {
                    uint4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = ldg_nc(src + off + (size_t)u * 4096);
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
                }
        }
    } else {
        uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
        uint8_t* ring = smem + 256;
        const int NS = 16; const uint32_t SB = 8192;
        if (tid == 0) { for (int i = 0; i < NS; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bars[i])), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncthreads();
        if (tid == 0) {
            const size_t total = (region / SB) * (size_t)passes;
            size_t issued = 0, done = 0; uint32_t par[16] = {0};
            while (done < total) {
                while (issued < total && issued - done < (size_t)NS) {
                    const int s = (int)(issued % NS);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[s])), "r"(SB) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(ring + s * SB)), "l"(src + (issued % (region / SB)) * SB), "r"(SB), "r"(smem_u32(&bars[s])) : "memory");
                    ++issued;
                }
                const int s = (int)(done % NS);
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(par[s]) : "memory");
                par[s] ^= 1u; ++done;
            }
        }
    }
    __syncthreads();
    if (tid == 0) out_ns[blockIdx.x] = gtimer() - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const size_t WB = (size_t)6 << 30;
    uint8_t* buf; CK(cudaMalloc(&buf, WB)); CK(cudaMemset(buf, 1, WB));
    unsigned long long* out; unsigned int* sink; CK(cudaMalloc(&out, 8 * 148)); CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 + 16 * 8192));
    struct Case { int grid; size_t region; int passes; const char* what; };
    std::vector<Case> cases = {
        {148, 256 << 10, 200, "L2-resident (148 x 256 KB)"}, {16, 256 << 10, 200, "L2-resident (16 x 256 KB)"}, {1, 256 << 10, 200, "L2-resident (1 x 256 KB)"},
        {148, 32 << 20, 2, "HBM stream (148 x 32 MB)"}, {16, 256 << 20, 1, "HBM stream (16 x 256 MB)"}, {1, 512 << 20, 1, "HBM stream (1 x 512 MB)"},
    };
    for (int mode = 0; mode < 2; ++mode) for (auto& c : cases) {
        k<<<c.grid, 288, 256 + 16 * 8192>>>(buf, c.region, c.passes, mode, out, sink);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(c.grid); CK(cudaMemcpy(h.data(), out, 8 * c.grid, cudaMemcpyDeviceToHost));
        std::sort(h.begin(), h.end());
        const double bytes = (double)(c.region / 32768 * 32768) * c.passes;
        printf("%-9s %-28s: per SM %.1f GB/s (median), total %.0f GB/s\n", mode ? "bulk/TMA" : "LDG.128", c.what, bytes / h[c.grid / 2], bytes * c.grid / h[c.grid - 1]);
    }
    return 0;
}
