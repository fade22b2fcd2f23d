// sync_bench.cu — design micro-benchmarks for the persistent decode kernel (148 CTAs x 288 threads, 1 CTA/SM).
//   A  grid barrier latency (red.release + ld.acquire poll), quiet and under a saturating TMA weight stream
//   B  tagged-word ("LL") all-to-all vector exchange: every CTA writes its slice of a K-word vector as (value, tag)
//      8-byte words, every CTA poll-loads the whole vector; quiet and under load
//   C  sum-of-squares chain variants (cycles per dependent step)
//   D  stream with stalls: ring of bulk copies, consumers stall S us every P bytes; with/without an L2 prefetch
//      running `ahead` bytes in front of the ring
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sync_bench sync_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kCons = 256, kThreads = 288;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kCons) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed_add_u64(unsigned long long* p, unsigned long long v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* p, uint32_t a, uint32_t b) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(a), "r"(b) : "memory"); }

// background weight stream of warp 8 (lane 0): 4 slots x 16 KB in flight, walks this CTA's private region
struct LoadGen {
    const uint8_t* base; size_t region; int on;
    int ns; uint32_t sb;                       // slots in flight, bytes per slot (ns * sb <= 64 KB)
    unsigned long long* bytes;                 // optional: total bytes moved (atomicAdd at the end)
};
__device__ __forceinline__ void loadgen_run(const LoadGen& lg, uint8_t* ring, uint64_t* bars, volatile int* stop) {
    if (!lg.on) return;
    const uint32_t SB = lg.sb; const int NS = lg.ns;
    const uint8_t* src = lg.base + (size_t)blockIdx.x * lg.region;
    size_t off = 0; int s = 0; uint32_t par = 0; unsigned long long moved = 0;
    for (int k = 0; k < NS; ++k) { mbar_arrive_expect_tx(&bars[k], SB); bulk_g2s(ring + k * SB, src + off, SB, &bars[k]); off += SB; }
    while (!*stop) {
        while (!mbar_try(&bars[s], par)) { }
        moved += SB;
        if (off + SB > lg.region) off = 0;
        mbar_arrive_expect_tx(&bars[s], SB); bulk_g2s(ring + s * SB, src + off, SB, &bars[s]); off += SB;
        if (++s == NS) { s = 0; par ^= 1u; }
    }
    for (int k = 0; k < NS; ++k) { while (!mbar_try(&bars[s], par)) { } if (++s == NS) { s = 0; par ^= 1u; } }
    if (lg.bytes) atomicAdd(lg.bytes, moved);
}

// ---------------------------------------------------------------------------------------------- A: grid barrier
__global__ void __launch_bounds__(kThreads, 1) bar_kernel(unsigned long long* ctr, int iters, int variant, LoadGen lg, unsigned long long* out_ns) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 64);
    uint8_t* ring = smem + 128;
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid >= kCons) { if (tid == kCons) loadgen_run(lg, ring, bars, stop); return; }
    unsigned long long target = gridDim.x;
    unsigned long long t0 = 0;
    for (int it = 0; it < iters + 20; ++it) {
        if (it == 20) t0 = gtimer();
        consumer_sync();
        if (tid == 0) {
            if (variant == 0) { red_release_add_u64(ctr, 1ull); while (ld_acquire_u64(ctr) < target) { } }
            else { __threadfence(); red_relaxed_add_u64(ctr, 1ull); while (ld_relaxed_u64(ctr) < target) { } __threadfence(); }
        }
        consumer_sync();
        target += gridDim.x;
    }
    if (tid == 0) { out_ns[blockIdx.x] = gtimer() - t0; *stop = 1; }
}

// ---------------------------------------------------------------------------------------------- B: LL exchange
__global__ void __launch_bounds__(kThreads, 1) ll_kernel(uint2* vec, int K, int iters, LoadGen lg, unsigned long long* out_ns, unsigned long long* out_polls) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    volatile int* stop = reinterpret_cast<volatile int*>(smem + 64);
    uint8_t* ring = smem + 128;
    float* xf = reinterpret_cast<float*>(smem + 128 + 65536);
    const int tid = threadIdx.x;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid >= kCons) { if (tid == kCons) loadgen_run(lg, ring, bars, stop); return; }
    const int w0 = (int)((long long)K * blockIdx.x / gridDim.x), w1 = (int)((long long)K * (blockIdx.x + 1) / gridDim.x);
    unsigned long long t0 = 0, polls = 0;
    float accum = 0.f;
    for (int it = 0; it < iters + 20; ++it) {
        if (it == 20) { t0 = gtimer(); polls = 0; }
        const uint32_t tag = (uint32_t)it + 1u;
        uint2* buf = vec + (size_t)(it & 1) * K;
        for (int i = w0 + tid; i < w1; i += kCons) st_relaxed_v2(buf + i, __float_as_uint((float)i + accum * 1e-30f), tag);
        // poll-load the whole vector, two words per 16-byte load; all loads of a batch in flight before any check
        constexpr int NB = 12;
        for (int base = 0; base < K; base += kCons * 2 * NB) {
            uint4 v[NB];
            bool ok[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) ok[u] = base + (u * kCons + tid) * 2 >= K;
            bool all_ok;
            do {
#pragma unroll
                for (int u = 0; u < NB; ++u) { const int i = base + (u * kCons + tid) * 2; if (!ok[u]) { v[u] = ld_relaxed_v4(buf + i); ++polls; } }
                all_ok = true;
#pragma unroll
                for (int u = 0; u < NB; ++u) { if (!ok[u]) { ok[u] = (v[u].y == tag && v[u].w == tag); all_ok = all_ok && ok[u]; } }
            } while (!all_ok);
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const int i = base + (u * kCons + tid) * 2;
                if (i < K) { xf[i] = __uint_as_float(v[u].x); xf[i + 1] = __uint_as_float(v[u].z); }
            }
        }
        consumer_sync();
        accum += xf[(it * 37) % K];
        consumer_sync();
    }
    if (tid == 0) { out_ns[blockIdx.x] = gtimer() - t0; *stop = 1; }
    atomicAdd(out_polls, polls);
    if (accum == 123.456f) out_ns[0] = 0;
}

// ---------------------------------------------------------------------------------------------- C: chain variants
__device__ __forceinline__ float chain_a(const float* xf, int n, int lane) {          // current: 8 loads then 8 fmas
    float acc = 0.0f;
    if (lane < 4) {
        const float* p = xf + lane; const int steps = n / 4;
#pragma unroll 1
        for (int i = 0; i + 8 <= steps; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[(i + u) * 4];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fmaf_rn(v[u], v[u], acc);
        }
    }
    return acc;
}
template <int B>
__device__ __forceinline__ float chain_b(const float* xt, int n, int lane) {          // transposed image, LDS.128, double-buffered B float4
    float acc = 0.0f;
    if (lane < 4) {
        const int steps = n / 4;
        const float4* p = reinterpret_cast<const float4*>(xt + lane * steps);
        float4 cur[B], nxt[B];
#pragma unroll
        for (int u = 0; u < B; ++u) cur[u] = p[u];
#pragma unroll 1
        for (int i = 0; i < steps / 4; i += B) {
#pragma unroll
            for (int u = 0; u < B; ++u) nxt[u] = p[min(i + B + u, steps / 4 - 1)];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                acc = __fmaf_rn(cur[u].x, cur[u].x, acc); acc = __fmaf_rn(cur[u].y, cur[u].y, acc);
                acc = __fmaf_rn(cur[u].z, cur[u].z, acc); acc = __fmaf_rn(cur[u].w, cur[u].w, acc);
            }
#pragma unroll
            for (int u = 0; u < B; ++u) cur[u] = nxt[u];
        }
    }
    return acc;
}
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const float* x, int n, float* out, long long* cyc, int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* xf = reinterpret_cast<float*>(smem);
    float* xt = xf + n;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) return;
    for (int i = tid; i < n; i += kCons) { xf[i] = x[i]; xt[(i & 3) * (n / 4) + (i >> 2)] = x[i]; }
    consumer_sync();
    long long t0 = clock64();
    float r = 0.f;
    if (warp == 0) {
        for (int rep = 0; rep < 8; ++rep) {
            float a;
            if (variant == 0) a = chain_a(xf, n, lane);
            else if (variant == 1) a = chain_b<2>(xt, n, lane);
            else if (variant == 2) a = chain_b<4>(xt, n, lane);
            else a = chain_b<8>(xt, n, lane);
            r += a;
        }
    }
    long long t1 = clock64();
    consumer_sync();
    if (tid < 4) { out[blockIdx.x * 4 + tid] = r; }
    if (tid == 0) cyc[blockIdx.x] = (t1 - t0) / 8;
}

// ---------------------------------------------------------------------------------------------- D: stream + stalls
struct StreamArgs {
    const uint8_t* base; size_t region;     // per-CTA private region (bytes)
    int n_slots; uint32_t slot_bytes;
    int n_phases; uint32_t phase_bytes;     // per CTA
    uint32_t stall_ns;
    uint32_t ahead;                         // L2 prefetch distance in bytes (0 = off)
    uint32_t pf_chunk;
    int touch;                              // consumers read the slot through shared memory
    int window;                             // max stages in flight (issued, not yet landed); >= n_slots: unlimited
    int fence;                              // 1: __threadfence_block + volatile issue counter after every stage (as the decode kernel did), 2: counter only
};
__global__ void __launch_bounds__(kThreads, 1) stream_kernel(StreamArgs a, unsigned long long* ctr, unsigned long long* out_ns, float* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 64;
    uint8_t* ring = smem + 1024;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < a.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint8_t* src = a.base + (size_t)blockIdx.x * a.region;
    const uint32_t spp = a.phase_bytes / a.slot_bytes;          // stages per phase
    const unsigned long long total = (unsigned long long)a.n_phases * spp;
    if (warp == 8) {
        if (lane == 0) {
            // lean loop: incrementing slot/parity/offsets, no divisions
            const uint32_t total32 = (uint32_t)total;
            uint32_t slot = 0, par = 1;                  // waiting parity for empty[] starts at 1 (passes immediately)
            uint32_t off = 0;                            // ring-issue offset inside the region
            uint32_t pf_off = 0;                         // prefetch offset inside the region
            int lead = 0;                                // prefetched stages in front of the ring issue point
            const int max_lead = (int)(a.ahead / a.slot_bytes);
            const uint32_t region = (uint32_t)a.region;
            uint32_t cslot = 0, cpar = 0; int inflight = 0;
            for (uint32_t idx = 0; idx < total32; ++idx) {
                if (a.window < a.n_slots) {
                    while (inflight > 0 && (inflight >= a.window || true)) {
                        if (mbar_try(&full[cslot], cpar)) { --inflight; if (++cslot == (uint32_t)a.n_slots) { cslot = 0; cpar ^= 1u; } }
                        else if (inflight < a.window) break;
                    }
                    ++inflight;
                }
                while (lead < max_lead) { bulk_prefetch_l2(src + pf_off, a.slot_bytes); pf_off += a.slot_bytes; if (pf_off >= region) pf_off = 0; ++lead; }
                while (!mbar_try(&empty[slot], par)) { __nanosleep(20); }
                mbar_arrive_expect_tx(&full[slot], a.slot_bytes);
                bulk_g2s(ring + (size_t)slot * a.slot_bytes, src + off, a.slot_bytes, &full[slot]);
                if (a.fence == 1) __threadfence_block();
                if (a.fence) *reinterpret_cast<volatile uint32_t*>(smem + 1016) = idx + 1;
                off += a.slot_bytes; if (off >= region) off = 0;
                if (lead > 0) --lead; else { pf_off = off; }
                if (++slot == (uint32_t)a.n_slots) { slot = 0; par ^= 1u; }
            }
        }
        return;
    }
    unsigned long long t0 = gtimer();
    float acc = 0.f;
    unsigned long long idx0 = 0;
    // stage idx is consumed by warp idx % 8 and lives in slot idx % n_slots; n_slots % 8 == 0, so every slot has ONE consumer
    // warp and a warp can never wait on a slot whose previous revolution it has not itself finished (parity aliasing).
    uint32_t my = warp;                                   // next stage of this warp
    uint32_t slot = warp, par = 0;
    for (int ph = 0; ph < a.n_phases; ++ph) {
        if (a.stall_ns) { const unsigned long long ts = gtimer(); while (gtimer() - ts < a.stall_ns) { } }
        const uint32_t end = (uint32_t)(ph + 1) * spp;
        for (; my < end; my += 8) {
            while (!mbar_try(&full[slot], par)) { }
            if (a.touch) {
                const uint4* p = reinterpret_cast<const uint4*>(ring + (size_t)slot * a.slot_bytes);
                for (uint32_t i = lane; i < a.slot_bytes / 16; i += 32) { const uint4 v = p[i]; acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            slot += 8; if (slot >= (uint32_t)a.n_slots) { slot -= a.n_slots; par ^= 1u; }
        }
        consumer_sync();
    }
    if (tid == 0) out_ns[blockIdx.x] = gtimer() - t0;
    if (acc == 1.2345f) sink[0] = acc;
}

static void stats(const char* name, std::vector<unsigned long long>& v, double per) {
    std::sort(v.begin(), v.end());
    printf("%-60s min %9.3f  med %9.3f  max %9.3f\n", name, v.front() / per, v[v.size() / 2] / per, v.back() / per);
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    int G = prop.multiProcessorCount;
    if (getenv("SB_GRID")) G = atoi(getenv("SB_GRID"));     // fewer CTAs: is one SM's ingest rate the cap, or HBM?
    printf("device %s, %d SMs\n", prop.name, G);
    const size_t WB = (size_t)4 << 30;
    uint8_t* wbuf; CK(cudaMalloc(&wbuf, WB)); CK(cudaMemset(wbuf, 1, WB));
    const size_t region = (WB / G) & ~(size_t)0xffff;
    unsigned long long *ctr, *out_ns, *polls;
    CK(cudaMalloc(&ctr, 64)); CK(cudaMalloc(&out_ns, 8 * G)); CK(cudaMalloc(&polls, 8));
    std::vector<unsigned long long> h(G);
    auto coop = [&](const void* fn, void** args, size_t smem) {
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(kThreads), args, smem, 0));
        CK(cudaDeviceSynchronize());
    };
    const char* sel = argc > 1 ? argv[1] : "ABCD";
    auto has = [&](char c) { for (const char* q = sel; *q; ++q) if (*q == c) return true; return false; };
    unsigned long long* lbytes; CK(cudaMalloc(&lbytes, 8));
    struct Load { int on, ns; uint32_t sb; };
    const Load loads[] = {{0, 0, 0}, {1, 1, 8704}, {1, 2, 8704}, {1, 3, 8704}, {1, 4, 8704}, {1, 7, 8704}, {1, 4, 16384}};
    // ---- A
    if (has('A')) for (const Load& ld : loads) for (int variant = 0; variant < 2; ++variant) {
        if (variant == 1 && ld.on && ld.ns != 4) continue;
        CK(cudaMemset(ctr, 0, 64)); CK(cudaMemset(lbytes, 0, 8));
        int iters = 2000; LoadGen lg{wbuf, region, ld.on, ld.ns, ld.sb, lbytes};
        void* args[] = {&ctr, &iters, &variant, &lg, &out_ns};
        coop((const void*)bar_kernel, args, 128 + 65536);
        CK(cudaMemcpy(h.data(), out_ns, 8 * G, cudaMemcpyDeviceToHost));
        unsigned long long nb; CK(cudaMemcpy(&nb, lbytes, 8, cudaMemcpyDeviceToHost));
        char nm[160]; snprintf(nm, sizeof nm, "A barrier v%d (%s) load %dx%u B (stream %.0f GB/s): us/barrier", variant, variant ? "fence+relaxed" : "rel/acq", ld.ns, ld.sb, nb / (double)h[0]);
        stats(nm, h, 1e3 * iters);
    }
    // ---- B
    uint2* vec; CK(cudaMalloc(&vec, 2 * 16384 * sizeof(uint2)));
    if (has('B')) for (const Load& ld : loads) for (int K : {4096, 11008}) {
        CK(cudaMemset(vec, 0, 2 * 16384 * sizeof(uint2))); CK(cudaMemset(polls, 0, 8)); CK(cudaMemset(lbytes, 0, 8));
        int iters = 2000; LoadGen lg{wbuf, region, ld.on, ld.ns, ld.sb, lbytes};
        void* args[] = {&vec, &K, &iters, &lg, &out_ns, &polls};
        coop((const void*)ll_kernel, args, 128 + 65536 + 16384 * 4);
        CK(cudaMemcpy(h.data(), out_ns, 8 * G, cudaMemcpyDeviceToHost));
        unsigned long long np; CK(cudaMemcpy(&np, polls, 8, cudaMemcpyDeviceToHost));
        unsigned long long nb; CK(cudaMemcpy(&nb, lbytes, 8, cudaMemcpyDeviceToHost));
        char nm[160]; snprintf(nm, sizeof nm, "B LL K=%d load %dx%u B (stream %.0f GB/s, polls/load %.2f): us/round", K, ld.ns, ld.sb, nb / (double)h[0], (double)np / ((double)iters * G * (K / 2)));
        stats(nm, h, 1e3 * iters);
    }
    // ---- C
    if (has('C')) {
        const int n = 4096; float* x; float* out; long long* cyc;
        CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&out, G * 16)); CK(cudaMalloc(&cyc, G * 8));
        std::vector<float> hx(n); for (int i = 0; i < n; ++i) hx[i] = (float)(i % 17) * 0.01f + 0.001f * (i % 5);
        CK(cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * n * 4));
        std::vector<float> ref(4);
        for (int v = 0; v < 4; ++v) {
            chain_kernel<<<G, kThreads, 2 * n * 4>>>(x, n, out, cyc, v);
            CK(cudaDeviceSynchronize());
            std::vector<long long> c(G); CK(cudaMemcpy(c.data(), cyc, 8 * G, cudaMemcpyDeviceToHost));
            std::vector<float> o(4); CK(cudaMemcpy(o.data(), out, 16, cudaMemcpyDeviceToHost));
            if (v == 0) ref = o;
            std::sort(c.begin(), c.end());
            printf("C chain variant %d: cycles per 1024-step chain min %lld med %lld (%.2f cyc/step) bits %s\n", v, c.front(), c[G / 2], c[G / 2] / 1024.0,
                   (o[0] == ref[0] && o[1] == ref[1] && o[2] == ref[2] && o[3] == ref[3]) ? "same" : "DIFFERENT");
        }
    }
    // ---- D
    if (has('D')) {
        float* sink; CK(cudaMalloc(&sink, 64));
        const uint32_t slot = 8704;
        struct Case { int n_slots; uint32_t phase_kb; uint32_t stall_ns; uint32_t ahead_kb; int touch; int window; int fence; };
        std::vector<Case> cases = {
            {24, 306, 0, 0, 0, 99, 0}, {24, 306, 0, 0, 1, 99, 0}, {24, 306, 8000, 0, 0, 99, 0}, {24, 306, 8000, 256, 0, 99, 0}, {24, 306, 8000, 512, 0, 99, 0},
        };
        int case_limit = argc > 3 ? atoi(argv[3]) : 1000;
        for (auto& c : cases) {
            if (case_limit-- <= 0) break;
            StreamArgs a{};
            a.base = wbuf; a.region = region; a.n_slots = c.n_slots; a.slot_bytes = slot;
            a.phase_bytes = (c.phase_kb * 1024u / slot) * slot; a.n_phases = argc > 2 ? atoi(argv[2]) : 300; a.stall_ns = c.stall_ns; a.ahead = c.ahead_kb * 1024u; a.pf_chunk = slot; a.touch = c.touch; a.window = c.window; a.fence = c.fence;
            a.region = (region / slot) * slot;
            const size_t smem = 1024 + (size_t)c.n_slots * slot;
            void* args[] = {&a, &ctr, &out_ns, &sink};
            coop((const void*)stream_kernel, args, smem);
            CK(cudaMemcpy(h.data(), out_ns, 8 * G, cudaMemcpyDeviceToHost));
            std::sort(h.begin(), h.end());
            const double t = h.back() * 1e-9;
            const double bytes = (double)G * a.n_phases * a.phase_bytes;
            const double stall_total = a.n_phases * (double)c.stall_ns * 1e-9;
            printf("D fence %d slots %2d win %2d phase %3u KB stall %5u ns ahead %4u KB touch %d: %.3f ms, %.0f GB/s overall, per phase %.2f us (stall %.1f + stream %.2f us = %.0f GB/s while streaming)\n",
                   c.fence, c.n_slots, c.window, a.phase_bytes / 1024, c.stall_ns, c.ahead_kb, c.touch, t * 1e3, bytes / t / 1e9, t / a.n_phases * 1e6, c.stall_ns * 1e-3,
                   (t - stall_total) / a.n_phases * 1e6, bytes / (t - stall_total) / 1e9);
        }
    }
    return 0;
}
