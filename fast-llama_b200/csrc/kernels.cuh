// kernels.cuh — sm_100a kernels of the decode hot path (see DESIGN.md for the layout and the rooflines).
//
//   pack_weights_kernel   row-major payload + scale table  ->  streaming "unit" layout
//   gemv_kernel           group-scaled integer GEMV (quant::matmul, quant_operators.cpp:252-284) with the
//                         producer ops fused in front (rmsnorm + quantize / quantize) and the consumer ops
//                         fused behind (store / residual add / swiglu)
//   attn_decode_kernel    execute_attn (transformer.cpp:397-455) for one new token
//   embed_kernel, argmax_kernel, small op kernels for the per-operator C-ABI
//
// Bit-exactness rules the code is written around (DESIGN.md "Exactness"):
//   * integer dots are exact, so they may be split across lanes in any way (dp4a);
//   * every FP32 accumulation the reference does in a fixed order is done in that order by ONE thread
//     (group chain of the matmul, 4-lane sum of squares, 8-lane QK dot, softmax sum, PV chain).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "exact_math.cuh"

namespace fl {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kKBlockElems = 512;      // one K-block = 8 lanes x 64 elements
constexpr unsigned kFull = 0xffffffffu;

enum { Q_INT16 = 1, Q_INT8 = 2 };

// ---------------------------------------------------------------------------------------------
// Streaming weight layout.  A matrix [M][K] (M padded to 4, K padded to 512) is cut into units of
// 4 rows x 512 columns.  Lane = r*8 + l owns row r, elements [l*64, l*64+64) of the K-block, i.e. whole
// quantisation groups.  Inside a unit the j-th 16-byte chunk of every lane is stored lane-contiguously:
//     byte  (j*32 + lane)*16 + b   <-  row 4*rt + r, element-byte  l*LB + j*16 + b      (LB = 64*sizeof(T))
// so each LDG.128 / bulk copy of a warp is one contiguous 512 B span, followed by the lane's GPL scales
//     float (32*LB)/4 + lane*GPL + gg  <-  scale[row][(kb*8 + l)*GPL + gg]
// Units are ordered [row tile][K-block]; a warp streams whole row tiles, so its accumulation order over
// groups is the reference's.
// ---------------------------------------------------------------------------------------------
template <int QT, int GS>
struct Traits {
    static constexpr int ES = (QT == Q_INT8) ? 1 : 2;          // element bytes
    static constexpr int LB = 64 * ES;                         // bytes per lane per unit
    static constexpr int NJ = LB / 16;                         // 16-byte chunks per lane per unit
    static constexpr int GPL = 64 / GS;                        // groups per lane per unit
    static constexpr int KB_BYTES = 8 * LB;                    // bytes of one row inside a K-block
    static constexpr int W_BYTES = 32 * LB;                    // payload bytes per unit
    static constexpr int UNIT_BYTES = W_BYTES + 32 * GPL * 4;  // + scales
    static_assert(GS == 64 || (GS == 32 && QT == Q_INT8), "group size");
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// element index e of an activation vector -> byte offset inside the permuted shared-memory image
template <int QT, int GS>
__device__ __forceinline__ int x_perm_offset(int e) {
    using T = Traits<QT, GS>;
    const int o = e * T::ES;
    const int kb = o / T::KB_BYTES, w = o % T::KB_BYTES;
    const int l = w / T::LB, j = (w % T::LB) / 16, b = w % 16;
    return kb * T::KB_BYTES + (j * 8 + l) * 16 + b;
}

template <int QT, int GS>
__global__ void pack_weights_kernel(const uint8_t* __restrict__ raw, const float* __restrict__ scales,
                                    uint8_t* __restrict__ packed, int M, int K, int n_tiles, int nkb,
                                    int tile_stride, int tile_offset) {
    // output tile index = tile_offset + rt * tile_stride (lets W1/W3 interleave into one stream)
    using T = Traits<QT, GS>;
    const int G = K / GS;
    const size_t n_units = (size_t)n_tiles * nkb;
    const int chunks_per_unit = T::UNIT_BYTES / 16;
    const size_t total = n_units * chunks_per_unit;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t u = idx / chunks_per_unit;
        const int c = (int)(idx % chunks_per_unit);
        const int rt = (int)(u / nkb), kb = (int)(u % nkb);
        uint8_t* dst = packed + ((size_t)(tile_offset + rt * tile_stride) * nkb + kb) * T::UNIT_BYTES + (size_t)c * 16;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < T::W_BYTES / 16) {
            const int j = c / 32, lane = c % 32, r = lane / 8, l = lane % 8;
            const int row = rt * 4 + r;
            const int eb = kb * T::KB_BYTES + l * T::LB + j * 16;     // byte offset inside the row
            if (row < M && eb < K * T::ES) v = *reinterpret_cast<const uint4*>(raw + (size_t)row * K * T::ES + eb);
        } else {
            float f[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int si = (c - T::W_BYTES / 16) * 4 + q;         // float index in the scale area
                const int lane = si / T::GPL, gg = si % T::GPL, r = lane / 8, l = lane % 8;
                const int row = rt * 4 + r;
                const int g = (kb * 8 + l) * T::GPL + gg;
                f[q] = (row < M && g < G) ? scales[(size_t)row * G + g] : 0.0f;
            }
            v = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
        }
        *reinterpret_cast<uint4*>(dst) = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Producer side: rmsnorm / quantize of the activation vector into shared memory (every CTA does it
// redundantly; the vector is 16-44 KB and comes from L2).
// ---------------------------------------------------------------------------------------------

// simd::rmsnorm's sum of squares (x86_simd.cpp:941-962 via the __AVX2 typo at :1093): four FMA chains over
// x[4i+j], then 0 + l0 + l1 + l2 + l3.  xf is the fp32 vector in shared memory, n % 4 == 0.
// Called by warp 0; returns the value in all its lanes.
__device__ __forceinline__ float sumsq_chain_warp0(const float* xf, int n, int lane) {
    float acc = 0.0f;
    if (lane < 4) {
        // FFMA-latency bound (profiles/micro/chain_bench.cu: 6.2 cycles per step on B200; deeper software pipelining is slower)
        const float* p = xf + lane;
        int i = 0;
        const int steps = n / 4;
#pragma unroll 1
        for (; i + 8 <= steps; i += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = p[(i + u) * 4];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = __fmaf_rn(v[u], v[u], acc);
        }
        for (; i < steps; ++i) { const float v = p[i * 4]; acc = __fmaf_rn(v, v, acc); }
    }
    const float l0 = __shfl_sync(kFull, acc, 0), l1 = __shfl_sync(kFull, acc, 1);
    const float l2 = __shfl_sync(kFull, acc, 2), l3 = __shfl_sync(kFull, acc, 3);
    float res = __fadd_rn(0.0f, l0);
    res = __fadd_rn(res, l1);
    res = __fadd_rn(res, l2);
    res = __fadd_rn(res, l3);
    return res;
}

// r = 1/sqrtf(ss/n + 1e-5f)   (x86_simd.cpp:1755 as compiled: vdivss, vaddss, vsqrtss, vdivss)
// K cache row layout.  A row is read by 8 lanes, lane j walking the reference's AVX lane j (elements 8 i + j, i ascending: the
// order of its FMA chain).  Lane j's q-th float4 (i = 4q .. 4q + 3) lives at float4 index 8 q + j, so the 8 lanes of a row read
// 128 contiguous bytes per load: coalesced from global memory, conflict-free from the shared-memory ring of the long-context
// attention (round 2; before, lane j's 16 values were contiguous: 4-way bank conflicts once K rows were staged unpadded).
__host__ __device__ __forceinline__ int k_cache_index(int e) { const int j = e & 7, i = e >> 3; return (((i >> 2) << 3) + j) * 4 + (i & 3); }

__device__ __forceinline__ float rms_scale(float ss, int n) {
    return __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)n), 1e-5f)));
}

// quant::quantize (quant_operators.cpp:26-47) of n values produced by val(e) into the permuted shared image
// xq (+ scales xs[n/GS]); optionally also into natural-order global buffers (per-op entry point).
// 8 threads cooperate on one group.  Block-wide; caller syncs afterwards.
template <int QT, int GS, typename ValFn>
__device__ __forceinline__ void quantize_block(ValFn val, int n, uint8_t* xq, float* xs,
                                               void* q_nat, float* s_nat) {
    constexpr int PER = GS / 8;                      // elements per thread
    const float QF = (QT == Q_INT8) ? 127.0f : 5792.0f;
    const int G = n / GS;
    const int sub = threadIdx.x & 7;
    for (int g = threadIdx.x >> 3; g < ceil_div(G, kThreads / 8) * (kThreads / 8); g += kThreads / 8) {
        const bool live = g < G;
        float v[PER];
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            v[i] = live ? val(g * GS + sub * PER + i) : 0.0f;
            m = fmaxf(m, fabsf(v[i]));
        }
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 1));
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 4));
        if (!live) continue;
        const float r = __fdiv_rn(m, QF);
        if (sub == 0) { xs[g] = r; if (s_nat) s_nat[g] = r; }
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = g * GS + sub * PER + i;
            const int q = cvtt_x86(__fdiv_rn(v[i], r));
            if (QT == Q_INT8) {
                xq[x_perm_offset<QT, GS>(e)] = (uint8_t)(q & 0xff);
                if (q_nat) reinterpret_cast<uint8_t*>(q_nat)[e] = (uint8_t)(q & 0xff);
            } else {
                *reinterpret_cast<uint16_t*>(xq + x_perm_offset<QT, GS>(e)) = (uint16_t)(q & 0xffff);
                if (q_nat) reinterpret_cast<uint16_t*>(q_nat)[e] = (uint16_t)(q & 0xffff);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GEMV
// ---------------------------------------------------------------------------------------------
enum { PRO_RMS_QUANT = 0, PRO_QUANT = 1, PRO_LOADQ = 2 };
enum { EPI_STORE = 0, EPI_RESADD = 1, EPI_SWIGLU = 2 };

struct GemvArgs {
    const uint8_t* w;        // packed units
    int M;                   // logical output rows (EPI_SWIGLU: hidden; the stream holds 2*M rows, W1/W3 tiles interleaved)
    int K;                   // input columns
    int n_tasks;             // ceil(M/4)
    int nkb;                 // K-blocks per row tile
    const float* in;         // PRO_RMS_QUANT / PRO_QUANT: fp32 activations [K]
    const float* gain;       // PRO_RMS_QUANT: rmsnorm gain [K]
    const void* in_q;        // PRO_LOADQ: quantised activations, natural order
    const float* in_s;       // PRO_LOADQ: their scales
    float* out;              // [M]
    float* tap;              // optional copy of the fp32 producer output (debug), may be NULL
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f32(const void* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int QT>
__device__ __forceinline__ int dot16(uint4 w, uint4 x, int d) {
    if (QT == Q_INT8) {
        d = __dp4a((int)w.x, (int)x.x, d);
        d = __dp4a((int)w.y, (int)x.y, d);
        d = __dp4a((int)w.z, (int)x.z, d);
        d = __dp4a((int)w.w, (int)x.w, d);
    } else {
        // int16 x int16 products accumulated in wrapping int32, like _mm256_mullo_epi32/_mm256_add_epi32
        // (x86_simd.cpp:1524-1552)
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int wl = (int)(short)(ww[i] & 0xffffu), wh = ((int)ww[i]) >> 16;
            const int xl = (int)(short)(xx[i] & 0xffffu), xh = ((int)xx[i]) >> 16;
            d += wl * xl;
            d += wh * xh;
        }
    }
    return d;
}

template <int QT, int GS>
struct UnitRegs {
    uint4 w[Traits<QT, GS>::NJ];
    float s[Traits<QT, GS>::GPL];
};

// One unit = 4 rows x 512 columns: integer dots of the lane's group(s), then the reference's FP32 chain
//     o[j] += s * dot  ==  fma(scales1*scales2, (float)dot, o[j]),  groups ascending   (quant_operators.cpp:274-275)
// The (s, f) pairs of a row's 8 lanes are transposed through a per-warp shared-memory slot so that every lane of the
// row walks them in group order (redundantly: same instruction stream, no shuffles, no divergence).
// Padded groups (K % 512 != 0) carry zero weights/scales/activations: fma(0, 0, acc) == acc exactly (acc is never -0).
template <int QT, int GS>
__device__ __forceinline__ float unit_chain(const uint4 (&w)[Traits<QT, GS>::NJ], const float (&ws)[Traits<QT, GS>::GPL],
                                            const uint4* __restrict__ xk, const float* __restrict__ xsk,
                                            float* __restrict__ cs, int lane, float acc) {
    using T = Traits<QT, GS>;
    const int l = lane & 7, r = lane >> 3;
    // independent partial dots per 16-byte chunk (integer adds are exact in any order), then one add tree per group
    int dj[T::NJ];
#pragma unroll
    for (int j = 0; j < T::NJ; ++j) dj[j] = dot16<QT>(w[j], xk[j * 8 + l], 0);
    int d[T::GPL];
#pragma unroll
    for (int gg = 0; gg < T::GPL; ++gg) d[gg] = 0;
#pragma unroll
    for (int j = 0; j < T::NJ; ++j) d[(j * 16) / (GS * T::ES)] += dj[j];
    if (T::GPL == 1) {
        reinterpret_cast<float2*>(cs)[lane] = make_float2(__fmul_rn(ws[0], xsk[l]), __int2float_rn(d[0]));
        __syncwarp();
        const float4* row = reinterpret_cast<const float4*>(cs) + r * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 p = row[i];
            acc = __fmaf_rn(p.x, p.y, acc);
            acc = __fmaf_rn(p.z, p.w, acc);
        }
    } else {
        reinterpret_cast<float4*>(cs)[lane] = make_float4(__fmul_rn(ws[0], xsk[l * 2]), __int2float_rn(d[0]),
                                                          __fmul_rn(ws[T::GPL - 1], xsk[l * 2 + 1]), __int2float_rn(d[T::GPL - 1]));
        __syncwarp();
        const float4* row = reinterpret_cast<const float4*>(cs) + r * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 p = row[i];
            acc = __fmaf_rn(p.x, p.y, acc);
            acc = __fmaf_rn(p.z, p.w, acc);
        }
    }
    return acc;
}

template <int QT, int GS, int PRO, int EPI>
__global__ void __launch_bounds__(kThreads) gemv_kernel(const GemvArgs a) {
    using T = Traits<QT, GS>;
    constexpr int TT = (EPI == EPI_SWIGLU) ? 2 : 1;        // row tiles per task
    constexpr int DEPTH = (QT == Q_INT8) ? 4 : 2;          // units in flight per warp
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(16) float chain_slots[kWarps][2][32 * 2 * T::GPL];
    const int K = a.K, G = K / GS;
    const int kpad = a.nkb * kKBlockElems;
    uint8_t* xq = smem;                                                  // kpad * ES bytes, permuted
    float* xs = reinterpret_cast<float*>(smem + (size_t)kpad * T::ES);   // G scales (padded to nkb*8*GPL)
    float* xf = xs + a.nkb * 8 * T::GPL;                                 // K floats (PRO_RMS_QUANT only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 3, l = lane & 7;
    const int gw = blockIdx.x * kWarps + warp, GW = gridDim.x * kWarps;

    // ---- start the weight stream before touching activations: the first DEPTH units of this warp
    const int units_per_task = TT * a.nkb;
    int n_units = 0;
    if (gw < a.n_tasks) n_units = ((a.n_tasks - 1 - gw) / GW + 1) * units_per_task;
    const size_t task_jump = ((size_t)GW - 1) * units_per_task * T::UNIT_BYTES;      // from the end of a task to the warp's next
    const uint8_t* ld_w = a.w + (size_t)gw * units_per_task * T::UNIT_BYTES + (size_t)lane * 16;
    const uint8_t* ld_s = a.w + (size_t)gw * units_per_task * T::UNIT_BYTES + T::W_BYTES + (size_t)lane * T::GPL * 4;
    UnitRegs<QT, GS> buf[DEPTH];
    int ld_u = 0, ld_in_task = 0;
    auto issue_load = [&](UnitRegs<QT, GS>& b) {
        if (ld_u < n_units) {
#pragma unroll
            for (int j = 0; j < T::NJ; ++j) b.w[j] = ldg_stream(ld_w + j * 512);
#pragma unroll
            for (int gg = 0; gg < T::GPL; ++gg) b.s[gg] = ldg_stream_f32(ld_s + gg * 4);
            ++ld_u;
            size_t adv = T::UNIT_BYTES;
            if (++ld_in_task == units_per_task) { ld_in_task = 0; adv += task_jump; }
            ld_w += adv; ld_s += adv;
        }
    };
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) issue_load(buf[i]);

    // ---- producer: bring the activation vector into shared memory, quantised and permuted
    for (int i = threadIdx.x; i < (kpad * T::ES) / 16; i += kThreads) reinterpret_cast<uint4*>(xq)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < a.nkb * 8 * T::GPL; i += kThreads) xs[i] = 0.0f;
    __syncthreads();
    if (PRO == PRO_RMS_QUANT) {
        for (int i = threadIdx.x; i < K / 4; i += kThreads)
            reinterpret_cast<float4*>(xf)[i] = __ldcg(reinterpret_cast<const float4*>(a.in) + i);
        __shared__ float s_r;
        __syncthreads();
        if (warp == 0) {
            const float ss = sumsq_chain_warp0(xf, K, lane);
            if (lane == 0) s_r = rms_scale(ss, K);
        }
        __syncthreads();
        const float rr = s_r;
        const float* gain = a.gain;
        float* tap = (blockIdx.x == 0) ? a.tap : nullptr;
        quantize_block<QT, GS>([&](int e) {
            const float y = __fmul_rn(__fmul_rn(xf[e], gain[e]), rr);   // (x*w)*r, multiply_avx256 :1359
            if (tap) tap[e] = y;
            return y;
        }, K, xq, xs, nullptr, nullptr);
    } else if (PRO == PRO_QUANT) {
        const float* in = a.in;
        quantize_block<QT, GS>([&](int e) { return __ldcg(in + e); }, K, xq, xs, nullptr, nullptr);
    } else {
        if (QT == Q_INT8) {
            const uint8_t* q = reinterpret_cast<const uint8_t*>(a.in_q);
            for (int e = threadIdx.x; e < K; e += kThreads) xq[x_perm_offset<QT, GS>(e)] = q[e];
        } else {
            const uint16_t* q = reinterpret_cast<const uint16_t*>(a.in_q);
            for (int e = threadIdx.x; e < K; e += kThreads) *reinterpret_cast<uint16_t*>(xq + x_perm_offset<QT, GS>(e)) = q[e];
        }
        for (int g = threadIdx.x; g < G; g += kThreads) xs[g] = a.in_s[g];
    }
    __syncthreads();

    // ---- stream: each warp walks its row tiles, K-block by K-block
    float acc = 0.0f, acc_first = 0.0f;
    int cu = 0, kb = 0, tt = 0, task = gw;
    const uint4* xq4 = reinterpret_cast<const uint4*>(xq);
    while (cu < n_units) {
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {
            if (cu < n_units) {
                UnitRegs<QT, GS> cur = buf[i];
                issue_load(buf[i]);          // refill the slot; the loads fly while the chain below runs
                acc = unit_chain<QT, GS>(cur.w, cur.s, xq4 + (size_t)kb * (T::KB_BYTES / 16), xs + kb * 8 * T::GPL,
                                         chain_slots[warp][cu & 1], lane, acc);
                ++cu;
                if (++kb == a.nkb) {
                    kb = 0;
                    if (TT == 2 && tt == 0) {
                        acc_first = acc; acc = 0.0f; tt = 1;
                    } else {
                        const int row = task * 4 + r;
                        if (l == 0 && row < a.M) {
                            if (EPI == EPI_STORE) a.out[row] = acc;
                            else if (EPI == EPI_RESADD) a.out[row] = __fadd_rn(a.out[row], acc);   // x1 += tmp (tensor.cpp:723)
                            else a.out[row] = swiglu_exact(acc_first, acc);
                        }
                        acc = 0.0f; tt = 0; task += GW;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// embedding row fetch (transformer.cpp:115-122); the table is fp32 on the device
// ---------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ table, const int* __restrict__ token, float* __restrict__ x1, int dim) {
    const int t = *token;
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)t * dim);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dim / 4; i += gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(x1)[i] = src[i];
}

// ---------------------------------------------------------------------------------------------
// Attention for one new token — execute_attn (transformer.cpp:397-455).  One CTA per query head.
//   K cache row layout: lane-permuted for the reference's 8-lane AVX2 dot (x86_simd.cpp:1447-1468), see k_cache_index:
//     float ((i / 4) * 8 + j) * 4 + i % 4  <-  k[8*i + j]      (lane j's q-th float4 at float4 index 8 q + j)
//   V cache row layout: natural.
// ---------------------------------------------------------------------------------------------
struct AttnArgs {
    const float* qkv;      // [dim + 2*kv_dim] fp32 (q | k | v), pre-RoPE
    float* k_cache;        // this layer, this sequence: [n_kv_heads][max_seq][HS] (permuted rows)
    float* v_cache;        // [n_kv_heads][HS / v_dw][max_seq][v_dw]: head dims in column blocks of v_dw (v_dw == HS: natural rows)
    const float* rope;     // [max_pos][HS/2][2] = (cos, sin) built on the host with glibc sincosf
    const int* pos_ptr;    // device: position of the new token
    const int* bs_ptr;     // device: tokens in the enclosing forward() call
    float* out;            // [dim]
    float* tap_qkv;        // optional: [dim + 2 kv_dim] post-RoPE copy
    int n_heads, n_kv_heads, max_seq;
    float attn_scale;      // 1/sqrtf(HS) computed on the host (transformer.cpp:416)
    int v_dw;              // column-block width of the V cache (the persistent kernel streams one block per CTA)
};

constexpr int kVChunk = 64;    // V rows per cp.async stage

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

template <int HS>
__global__ void __launch_bounds__(kThreads) attn_decode_kernel(const AttnArgs a) {
    constexpr int EPL = HS / 8;            // elements per AVX lane
    extern __shared__ __align__(16) uint8_t smem[];
    float* q_s = reinterpret_cast<float*>(smem);            // [HS] roped q, natural
    float* k_s = q_s + HS;                                  // [HS] roped new k, natural
    float* v_s = k_s + HS;                                  // [HS] new v
    float* red = v_s + HS;                                  // [32] reduction scratch
    float* v_stage = red + 32;                              // [2][kVChunk][HS]
    float* att = v_stage + 2 * kVChunk * HS;                // [max_seq]

    const int qh = blockIdx.x;
    const int hgs = a.n_heads / a.n_kv_heads;
    const int kvh = qh / hgs, g = qh % hgs;
    const int dim = a.n_heads * HS, kv_dim = a.n_kv_heads * HS;
    const int pos = *a.pos_ptr;
    const int n = pos + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* kc = a.k_cache + (size_t)kvh * a.max_seq * HS;
    float* vc = a.v_cache + (size_t)kvh * a.max_seq * HS;

    // ---- start fetching V (does not depend on anything computed here)
    auto issue_v_chunk = [&](int c) {
        const int t0 = c * kVChunk;
        float* dst = v_stage + (size_t)(c & 1) * kVChunk * HS;
        const int rows = min(kVChunk, pos - t0);           // cached rows only (row `pos` comes from v_s)
        for (int i = tid; i < rows * (HS / 4); i += kThreads) {
            const int row = i / (HS / 4), d = (i - row * (HS / 4)) * 4;
            cp_async16(dst + (size_t)i * 4, vc + ((size_t)(d / a.v_dw) * a.max_seq + t0 + row) * a.v_dw + d % a.v_dw);
        }
        cp_async_commit();
    };
    const int n_chunks = ceil_div(n, kVChunk);
    issue_v_chunk(0);
    if (n_chunks > 1) issue_v_chunk(1); else cp_async_commit();

    // ---- RoPE (rope_v2, tf_operators.cpp:355-402) + KV append (transformer.cpp:431-439)
    if (tid < HS / 2) {
        // sequence_rope_v2 (tensor.h:262-270) walks all bs*hgs rows of the q tensor with position pos0 + row, so
        // query head g of a GQA group, token i of a bs-token forward, is rotated at pos0 + g*bs + i = pos + g*bs
        // (== pos when n_heads == n_kv_heads).  Reproduced, not fixed: parity with the reference is the contract.
        const float2 cs = reinterpret_cast<const float2*>(a.rope)[(size_t)(pos + g * (*a.bs_ptr)) * (HS / 2) + tid];
        const float2 x = reinterpret_cast<const float2*>(a.qkv + (size_t)qh * HS)[tid];
        float o0, o1;
        rope_pair(cs.x, cs.y, x.x, x.y, o0, o1);
        q_s[2 * tid] = o0; q_s[2 * tid + 1] = o1;
        if (a.tap_qkv) { a.tap_qkv[(size_t)qh * HS + 2 * tid] = o0; a.tap_qkv[(size_t)qh * HS + 2 * tid + 1] = o1; }
    } else if (tid < HS) {
        const int i = tid - HS / 2;
        const float2 cs = reinterpret_cast<const float2*>(a.rope)[(size_t)pos * (HS / 2) + i];
        const float2 x = reinterpret_cast<const float2*>(a.qkv + dim + (size_t)kvh * HS)[i];
        float o0, o1;
        rope_pair(cs.x, cs.y, x.x, x.y, o0, o1);
        k_s[2 * i] = o0; k_s[2 * i + 1] = o1;
        if (g == 0) {
            float* krow = kc + (size_t)pos * HS;
            krow[k_cache_index(2 * i)] = o0;
            krow[k_cache_index(2 * i + 1)] = o1;
            if (a.tap_qkv) { a.tap_qkv[dim + (size_t)kvh * HS + 2 * i] = o0; a.tap_qkv[dim + (size_t)kvh * HS + 2 * i + 1] = o1; }
        }
    } else if (tid < HS + HS / 4) {
        const int i = tid - HS;
        const float4 v = reinterpret_cast<const float4*>(a.qkv + dim + kv_dim + (size_t)kvh * HS)[i];
        reinterpret_cast<float4*>(v_s)[i] = v;
        if (g == 0) {
            *reinterpret_cast<float4*>(vc + ((size_t)((4 * i) / a.v_dw) * a.max_seq + pos) * a.v_dw + (4 * i) % a.v_dw) = v;
            if (a.tap_qkv) reinterpret_cast<float4*>(a.tap_qkv + dim + kv_dim + (size_t)kvh * HS)[i] = v;
        }
    }
    __syncthreads();

    // ---- att[t] = (K[t] . q) * scale   (float dot_product_avx256: 8 FMA chains, then 0 + l0 + ... + l7)
    {
        const int rr = lane >> 3, j = lane & 7;
        float qr[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qr[i] = q_s[8 * i + j];
        constexpr int U = 4;                       // row groups in flight
        for (int base = 0; base < n; base += kWarps * 4 * U) {
            float4 kv[U][EPL / 4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + (u * kWarps + warp) * 4 + rr;
                if (t < pos) {
                    const float4* p = reinterpret_cast<const float4*>(kc + (size_t)t * HS) + j;
#pragma unroll
                    for (int c = 0; c < EPL / 4; ++c) kv[u][c] = __ldcg(p + 8 * c);
                } else if (t == pos) {
#pragma unroll
                    for (int c = 0; c < EPL / 4; ++c)
                        kv[u][c] = make_float4(k_s[8 * (4 * c) + j], k_s[8 * (4 * c + 1) + j], k_s[8 * (4 * c + 2) + j], k_s[8 * (4 * c + 3) + j]);
                } else {
#pragma unroll
                    for (int c = 0; c < EPL / 4; ++c) kv[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + (u * kWarps + warp) * 4 + rr;
                float acc = 0.0f;
#pragma unroll
                for (int c = 0; c < EPL / 4; ++c) {
                    acc = __fmaf_rn(kv[u][c].x, qr[4 * c], acc);
                    acc = __fmaf_rn(kv[u][c].y, qr[4 * c + 1], acc);
                    acc = __fmaf_rn(kv[u][c].z, qr[4 * c + 2], acc);
                    acc = __fmaf_rn(kv[u][c].w, qr[4 * c + 3], acc);
                }
                float tot = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) tot = __fadd_rn(tot, __shfl_sync(kFull, acc, (rr << 3) + k));
                if (j == 0 && t < n) att[t] = __fmul_rn(tot, a.attn_scale);      // att.multiply(attn_scale) :443
            }
        }
    }
    __syncthreads();

    // ---- softmax_sisd (tf_operators.cpp:176-186): max, expf(x - max), serial sum, divide
    float m = -INFINITY;
    for (int t = tid; t < n; t += kThreads) m = fmaxf(m, att[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) m = fmaxf(m, red[w]);
    for (int t = tid; t < n; t += kThreads) att[t] = expf_exact(__fsub_rn(att[t], m));
    __syncthreads();
    if (tid == 0) {
        float sum = 0.0f;
        int t = 0;
        for (; t + 8 <= n; t += 8) {
            float e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) e[u] = att[t + u];
#pragma unroll
            for (int u = 0; u < 8; ++u) sum = __fadd_rn(sum, e[u]);
        }
        for (; t < n; ++t) sum = __fadd_rn(sum, att[t]);
        red[16] = sum;
    }
    __syncthreads();
    const float sum = red[16];
    for (int t = tid; t < n; t += kThreads) att[t] = __fdiv_rn(att[t], sum);
    __syncthreads();

    // ---- weighted_sum (tf_operators.cpp:325-350): o = V[0]*w0; t>=1: if |w_t| > 1e-15: o = fma(V[t], w_t, o)
    float o = 0.0f;
    for (int c = 0; c < n_chunks; ++c) {
        cp_async_wait<1>();
        __syncthreads();
        if (tid < HS) {
            const float* vb = v_stage + (size_t)(c & 1) * kVChunk * HS;
            const int t0 = c * kVChunk, t1 = min(n, t0 + kVChunk);
            for (int t = t0; t < t1; ++t) {
                const float v = (t == pos) ? v_s[tid] : vb[(size_t)(t - t0) * HS + tid];
                const float w = att[t];
                if (t == 0) o = __fmul_rn(v, w);
                else if (fabsf(w) > 1e-15f) o = __fmaf_rn(v, w, o);
            }
        }
        __syncthreads();
        if (c + 2 < n_chunks) issue_v_chunk(c + 2); else cp_async_commit();
    }
    if (tid < HS) a.out[(size_t)qh * HS + tid] = o;
}

// ---------------------------------------------------------------------------------------------
// Sampler::sample_argmax (sampler.cpp:36-46): first index of the strict maximum; then advance the
// sequence state on the device (token fed back, pos + 1) so decode needs no host round trip.
// ---------------------------------------------------------------------------------------------
struct SeqState {      // lives in device memory, one per KV slot
    int token;         // input token of the next step
    int pos;           // its position
    int n_out;         // tokens appended to out_tokens so far
    int bs;            // tokens in the forward() call this step belongs to (1 while decoding); see attn kernel RoPE note
};

__global__ void argmax_kernel(const float* __restrict__ logits, int n, SeqState* st, int* out_tokens, int out_cap,
                              int* argmax_out, int advance) {
    __shared__ float sv[32];
    __shared__ int si[32];
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = logits[i];
        if (v > bv) { bv = v; bi = i; }
    }
    // the reference starts from logits[0]: an index is best if its value is larger, or equal with a lower index
    if (bi == 0x7fffffff) { bi = threadIdx.x < n ? threadIdx.x : 0; bv = logits[bi]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(kFull, bv, o);
        const int oi = __shfl_xor_sync(kFull, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        if (argmax_out) *argmax_out = bi;
        if (advance) {
            if (st->n_out < out_cap) out_tokens[st->n_out] = bi;
            st->n_out += 1;
            st->token = bi;
            st->pos += 1;
            st->bs = 1;
        }
    }
}

__global__ void set_state_kernel(SeqState* st, const int* tokens, int idx, int pos, int bs, int reset_out) {
    st->token = tokens[idx];
    st->pos = pos;
    st->bs = bs;
    if (reset_out) st->n_out = 0;
}

// ---------------------------------------------------------------------------------------------
// small kernels behind the per-operator C-ABI (same device functions as the fused kernels)
// ---------------------------------------------------------------------------------------------
template <int QT, int GS>
__global__ void __launch_bounds__(kThreads) op_quantize_kernel(const float* x, int n, void* q_out, float* s_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int nkb = ceil_div(n, kKBlockElems);
    uint8_t* xq = smem;
    float* xs = reinterpret_cast<float*>(smem + (size_t)nkb * kKBlockElems * Traits<QT, GS>::ES);
    quantize_block<QT, GS>([&](int e) { return x[e]; }, n, xq, xs, q_out, s_out);
}

__global__ void __launch_bounds__(kThreads) op_rmsnorm_kernel(const float* x, const float* w, int n, float* out) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* xf = reinterpret_cast<float*>(smem);
    __shared__ float s_r;
    for (int i = threadIdx.x; i < n; i += kThreads) xf[i] = x[i];
    __syncthreads();
    if (threadIdx.x < 32) {
        const float ss = sumsq_chain_warp0(xf, n, threadIdx.x);
        if (threadIdx.x == 0) s_r = rms_scale(ss, n);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kThreads) out[i] = __fmul_rn(__fmul_rn(xf[i], w[i]), s_r);
}

__global__ void op_swiglu_kernel(const float* a, const float* b, int n, float* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = swiglu_exact(a[i], b[i]);
}

__global__ void op_expf_kernel(const float* x, int n, float* out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = expf_exact(x[i]);
}

__global__ void op_rope_kernel(const float* x, const float* cs, int n_dims, float* out) {
    const int i = threadIdx.x;
    if (i < n_dims / 2) {
        float o0, o1;
        rope_pair(cs[2 * i], cs[2 * i + 1], x[2 * i], x[2 * i + 1], o0, o1);
        out[2 * i] = o0; out[2 * i + 1] = o1;
    }
}

__global__ void __launch_bounds__(kThreads) op_softmax_kernel(const float* x, int n, float* out) {
    __shared__ float red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float m = -INFINITY;
    for (int t = tid; t < n; t += kThreads) m = fmaxf(m, x[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < kWarps; ++w) m = fmaxf(m, red[w]);
    for (int t = tid; t < n; t += kThreads) out[t] = expf_exact(__fsub_rn(x[t], m));
    __syncthreads();
    if (tid == 0) {
        float sum = 0.0f;
        for (int t = 0; t < n; ++t) sum = __fadd_rn(sum, out[t]);
        red[16] = sum;
    }
    __syncthreads();
    const float sum = red[16];
    for (int t = tid; t < n; t += kThreads) out[t] = __fdiv_rn(out[t], sum);
}

// natural-order K rows -> lane-permuted cache rows (per-op attention entry point / tests)
__global__ void permute_k_rows_kernel(const float* src, float* dst, int rows, int hs) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * hs; i += gridDim.x * blockDim.x) {
        const int row = i / hs, e = i % hs;
        dst[(size_t)row * hs + k_cache_index(e)] = src[i];
    }
}

}  // namespace fl
