// megakernel.cuh — the decode step as ONE persistent sm_100a kernel (one CTA per SM).
//
// Why: a decode token is 161 dependent matrix-vector phases of 18-140 MB each.  Launched as separate kernels
// every phase pays launch latency, a cold pipeline and its activation prologue (for rmsnorm a 1024-step serial FP32
// chain) with HBM idle.  Here the weight stream never stops:
//
//   * warp 8 (one elected lane) is a TMA producer: it walks the token's static weight schedule and issues
//     cp.async.bulk global->shared copies into a ring of stages guarded by full/empty mbarriers.  Weights do not
//     depend on activations, so the producer runs ahead across phase, layer and token boundaries; only ring
//     capacity (~180 KB per SM = ~4 us of this SM's HBM share) limits it.
//   * warps 0-7 are consumers: per phase they (1) pass a grid-wide barrier, (2) rebuild the quantised activation
//     vector in shared memory (rmsnorm chain, quantise), (3) drain their stages with the exact per-unit arithmetic
//     of kernels.cuh and (4) write their rows.  While they do (1)+(2) the ring fills, so HBM stays busy.
//   * attention runs between the QKV and Wo phases on n_heads * CPH CTAs (CPH CTAs share one head: keys are split
//     for QK^T, head dims are split for the PV chains), exchanging the score vector through L2.
//
// Each CTA owns a contiguous range of 4-row tiles of every matrix (rows/148 +- 4), so its share of a phase is ONE
// contiguous byte range of the streaming layout and every stage is a single bulk copy.
//
// Two hardware facts shape the code (both measured, see DESIGN.md "What the profiler taught us"):
//   1. With ~227 KB of shared memory carved out there is practically no L1 left: every local-memory (stack) access and
//      every re-read of a global word is an L2 round trip (~0.3 us).  Nothing here may spill or take the address of a
//      local; everything is force-inlined, parameters stay in the constant bank, loops keep state in registers.
//   2. The kernel body must stay small: code that runs once per phase is fetched from L2 when it does not fit the
//      instruction cache.  Hence ONE instance of each phase routine inside a flat, rolled loop over phases.
#pragma once
#include "kernels.cuh"

namespace fl {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kMegaThreads = kConsumerThreads + 32;

struct MegaLayer {
    const uint8_t* qkv;
    const uint8_t* wo;
    const uint8_t* w13;
    const uint8_t* w2;
    const float* att_norm;
    const float* ffn_norm;
};

struct MegaParams {
    const MegaLayer* layers;
    const uint8_t* cls;
    const float* out_norm;
    const float* emb;
    float* x1; float* qkv; float* attn; float* hd; float* logits;
    float* att_scratch;               // [n_heads][max_seq] raw scores exchanged between the CTAs of a head
    float* k_cache; float* v_cache;   // this sequence: [n_layers][n_kv_heads][max_seq][HS]
    const float* rope;
    SeqState* st;
    int* out_tokens; int out_cap; int* argmax_out;
    unsigned long long* bar_ctr;      // [0] grid barrier counter, [1] its value at the end of the previous launch
    unsigned long long* head_ctr;     // [n_heads] per-head arrival counters, [n_heads .. 2 n_heads) their launch bases
    float* am_val; int* am_idx;       // [gridDim] per-CTA argmax partials
    float* tap_norm;
    unsigned long long* prof;         // optional [gridDim][32] ns per category, see fl_profile_read
    int dim, hidden, n_layers, n_heads, n_kv_heads, vocab, max_seq;
    int qkv_rows;
    float attn_scale;
    int n_steps;
    int cph;                          // CTAs per head (1, 2 or 4)
    int n_slots;                      // ring stages
    // dynamic shared memory carve-up (byte offsets)
    int off_ring, off_xq, off_xs, off_xf, off_chain, off_att, off_misc, off_bars, off_vstage;
    int v_chunk_rows;
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) { }
}
// Producer-side wait: back off instead of spinning, the ring holds microseconds of data.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// per-CTA phase timing, only when MegaParams::prof is set; lives in registers.  The LAST consumer thread keeps the
// clock: it takes part in none of the serial single-warp sections (thread 0 timing them made lane 0 diverge from the
// other chain lanes and doubled their cost), and every stop sits right after a consumer_sync, so it still sees
// every interval end.
struct Prof {
    unsigned long long* p; unsigned long long t0;
    __device__ __forceinline__ void stop(int tid, int cat) {
        if (p && tid == kConsumerThreads - 1) { const unsigned long long t = gtimer(); atomicAdd(p + cat, t - t0); t0 = t; }
    }
};

// grid-wide barrier for the consumer warps of all CTAs (all CTAs are co-resident: cooperative launch).
// bar.sync makes every consumer thread's stores happen-before thread 0's release; the acquire poll + bar.sync make the
// other CTAs' stores visible to every consumer thread (which then read them with L2 loads, never through L1).
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long& target, int tid) {
    consumer_sync();
    if (tid == 0) {
        red_release_add_u64(ctr, 1ull);
        while (ld_acquire_u64(ctr) < target) { }
    }
    consumer_sync();
    target += gridDim.x;
}

// ---------------------------------------------------------------------------------------------- schedule
template <int QT, int GS>
struct Ring {
    using T = Traits<QT, GS>;
    static constexpr int U = (QT == Q_INT8) ? 4 : 2;             // units per stage
    static constexpr int SLOT_BYTES = U * T::UNIT_BYTES;
};

// Phase p of a token (p = 4*layer + {0 QKV, 1 Wo, 2 W1/W3, 3 W2}, p = 4*n_layers: classifier): where its weights are
// and how they are cut into tasks (row tiles / W1-W3 tile pairs).  Pure function of (params, p): the producer and
// every consumer warp evaluate it independently and must agree.
struct PhaseShape {
    const uint8_t* w;
    int n_tasks;          // row tiles (or W1/W3 tile pairs)
    int tt;               // row tiles per task (2 for the interleaved W1/W3 stream)
    int nkb;              // K-blocks (units) per row tile; a stage never straddles two tiles
    int K;                // input length
    int M;                // output rows
};

__device__ __forceinline__ PhaseShape phase_shape(const MegaParams& p, int pi) {
    const int nkb_d = ceil_div(p.dim, kKBlockElems), nkb_h = ceil_div(p.hidden, kKBlockElems);
    PhaseShape s;
    if (pi == 4 * p.n_layers) {
        s.w = p.cls; s.n_tasks = ceil_div(p.vocab, 4); s.tt = 1; s.nkb = nkb_d; s.K = p.dim; s.M = p.vocab;
        return s;
    }
    const MegaLayer* L = p.layers + (pi >> 2);
    const int ph = pi & 3;
    if (ph == 0)      { s.w = L->qkv; s.n_tasks = ceil_div(p.qkv_rows, 4); s.tt = 1; s.nkb = nkb_d; s.K = p.dim;    s.M = p.qkv_rows; }
    else if (ph == 1) { s.w = L->wo;  s.n_tasks = ceil_div(p.dim, 4);      s.tt = 1; s.nkb = nkb_d; s.K = p.dim;    s.M = p.dim; }
    else if (ph == 2) { s.w = L->w13; s.n_tasks = ceil_div(p.hidden, 4);   s.tt = 2; s.nkb = nkb_d; s.K = p.dim;    s.M = p.hidden; }
    else              { s.w = L->w2;  s.n_tasks = ceil_div(p.dim, 4);      s.tt = 1; s.nkb = nkb_h; s.K = p.hidden; s.M = p.dim; }
    return s;
}

// ---------------------------------------------------------------------------------------------- activation rebuild
__device__ __forceinline__ float group_max8(float m) {
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 1));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 4));
    return m;
}

// quantise one lane's share of a group (PER consecutive values, 8 lanes per group) given the group's max |value|:
// quant::quantize (quant_operators.cpp:26-47).  The PER values are PER consecutive elements of one 16-byte chunk of
// the permuted image, so they are packed and stored as words.
template <int QT, int GS>
__device__ __forceinline__ void quant_store(uint8_t* xq, float* xs, const float (&y)[GS / 8], float m, int g, int sub, float* tap) {
    constexpr int PER = GS / 8;
    const float QF = (QT == Q_INT8) ? 127.0f : 5792.0f;
    const float sc = __fdiv_rn(m, QF);
    if (sub == 0) xs[g] = sc;
    const int e0 = g * GS + sub * PER;
    constexpr int EPW = (QT == Q_INT8) ? 4 : 2;            // elements per 32-bit word
    uint32_t pk[PER / EPW];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const uint32_t q = (uint32_t)cvtt_x86(__fdiv_rn(y[i], sc)) & ((QT == Q_INT8) ? 0xffu : 0xffffu);
        pk[i / EPW] = (i % EPW == 0) ? q : (pk[i / EPW] | (q << ((32 / EPW) * (i % EPW))));
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(xq + x_perm_offset<QT, GS>(e0));
#pragma unroll
    for (int i = 0; i < PER / EPW; ++i) dst[i] = pk[i];
    if (tap) {
#pragma unroll
        for (int i = 0; i < PER; ++i) tap[e0 + i] = y[i];
    }
}

// Rebuild the quantised activation vector of a phase in shared memory (every CTA, redundantly).
//   gain != NULL: y = (x*w)*r, r = 1/sqrt(mean(x^2)+eps) (simd::rmsnorm, x86_simd.cpp:1754);  gain == NULL: y = x.
template <int QT, int GS>
__device__ __forceinline__ void build_activation(uint8_t* xq, float* xs, float* xf, float* misc, const float* in, const float* gain,
                                                 int K, int nkb, float* tap, int tid, Prof& pf) {
    using T = Traits<QT, GS>;
    constexpr int PER = GS / 8;                 // values per thread per group
    constexpr int GPP = kConsumerThreads / 8;   // groups per pass
    const int warp = tid >> 5, lane = tid & 31;
    const int kpad = nkb * kKBlockElems;
    const int G = K / GS;
    const int sub = tid & 7, g0 = tid >> 3;
    // zero the padded tail so padded groups contribute fma(0, 0, acc) == acc
    for (int i = K * T::ES + tid * 4; i < kpad * T::ES; i += kConsumerThreads * 4) *reinterpret_cast<uint32_t*>(xq + i) = 0u;
    for (int i = G + tid; i < nkb * 8 * T::GPL; i += kConsumerThreads) xs[i] = 0.0f;
    const int n_pass = ceil_div(G, GPP);
    float rr = 1.0f;
    if (gain) {
#pragma unroll 1
        for (int i = tid; i < K / 4; i += kConsumerThreads)
            reinterpret_cast<float4*>(xf)[i] = __ldcg(reinterpret_cast<const float4*>(in) + i);
        consumer_sync();
        pf.stop(tid, 8);
        if (warp == 0) {
            const float ss = sumsq_chain_warp0(xf, K, lane);
            if (lane == 0) misc[0] = rms_scale(ss, K);
        }
        consumer_sync();
        pf.stop(tid, 9);
        rr = misc[0];
    }
    // software-pipelined over passes: the next pass's operands are in flight while this pass is divided and stored
    float4 nx[PER / 4], nw[PER / 4];
#pragma unroll
    for (int q = 0; q < PER / 4; ++q) { nx[q] = make_float4(0.f, 0.f, 0.f, 0.f); nw[q] = nx[q]; }
    auto fetch = [&](int ps) {
        const int g = g0 + ps * GPP;
        if (ps < n_pass && g < G) {
#pragma unroll
            for (int q = 0; q < PER / 4; ++q) {
                if (gain) {
                    nx[q] = reinterpret_cast<const float4*>(xf + g * GS + sub * PER)[q];
                    nw[q] = __ldg(reinterpret_cast<const float4*>(gain + g * GS + sub * PER) + q);
                } else {
                    nx[q] = __ldcg(reinterpret_cast<const float4*>(in + g * GS + sub * PER) + q);
                }
            }
        }
    };
    fetch(0);
#pragma unroll 1
    for (int ps = 0; ps < n_pass; ++ps) {
        const int g = g0 + ps * GPP;
        float y[PER];
#pragma unroll
        for (int q = 0; q < PER / 4; ++q) {
            if (gain) {     // (x*w)*r, multiply_avx256 x86_simd.cpp:1359
                y[4 * q] = __fmul_rn(__fmul_rn(nx[q].x, nw[q].x), rr); y[4 * q + 1] = __fmul_rn(__fmul_rn(nx[q].y, nw[q].y), rr);
                y[4 * q + 2] = __fmul_rn(__fmul_rn(nx[q].z, nw[q].z), rr); y[4 * q + 3] = __fmul_rn(__fmul_rn(nx[q].w, nw[q].w), rr);
            } else {
                y[4 * q] = nx[q].x; y[4 * q + 1] = nx[q].y; y[4 * q + 2] = nx[q].z; y[4 * q + 3] = nx[q].w;
            }
        }
        fetch(ps + 1);
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < PER; ++i) m = fmaxf(m, fabsf(y[i]));
        m = group_max8(m);
        if (g < G) quant_store<QT, GS>(xq, xs, y, m, g, sub, tap);
    }
    consumer_sync();
}

// ---------------------------------------------------------------------------------------------- consumer GEMV phase
// One stage = up to U consecutive units of one row tile.  Pass 1 (independent work, lots of ILP): integer dots of all
// units of the stage, (scale product, float(dot)) pairs into the warp's shared-memory slot.  Pass 2: every lane of a
// row walks its row's pairs in group order — the reference's FP32 chain fma(s, f, acc) (quant_operators.cpp:274-275).
template <int QT, int GS>
__device__ __forceinline__ float stage_chain(const uint8_t* sp, int nu, const uint4* xk, const float* xsk, float* cs, int lane, float acc) {
    using T = Traits<QT, GS>;
    using R = Ring<QT, GS>;
    const int l = lane & 7, r = lane >> 3;
    constexpr int PW = 2 * T::GPL;                 // floats per lane per unit in the slot
#pragma unroll
    for (int u = 0; u < R::U; ++u) {
        if (u < nu) {
            const uint8_t* up = sp + (size_t)u * T::UNIT_BYTES;
            int dj[T::NJ];
#pragma unroll
            for (int j = 0; j < T::NJ; ++j)
                dj[j] = dot16<QT>(reinterpret_cast<const uint4*>(up)[j * 32 + lane], xk[(size_t)u * (T::KB_BYTES / 16) + j * 8 + l], 0);
            int d[T::GPL];
#pragma unroll
            for (int gg = 0; gg < T::GPL; ++gg) d[gg] = 0;
#pragma unroll
            for (int j = 0; j < T::NJ; ++j) d[(j * 16) / (GS * T::ES)] += dj[j];
            const float* wsp = reinterpret_cast<const float*>(up + T::W_BYTES) + lane * T::GPL;
            float* dst = cs + (size_t)(u * 32 + lane) * PW;
            if (T::GPL == 1) {
                *reinterpret_cast<float2*>(dst) = make_float2(__fmul_rn(wsp[0], xsk[u * 8 + l]), __int2float_rn(d[0]));
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(__fmul_rn(wsp[0], xsk[(u * 8 + l) * 2]), __int2float_rn(d[0]),
                                                              __fmul_rn(wsp[T::GPL - 1], xsk[(u * 8 + l) * 2 + 1]), __int2float_rn(d[T::GPL - 1]));
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < R::U; ++u) {
        if (u < nu) {
            const float4* row = reinterpret_cast<const float4*>(cs + (size_t)(u * 32 + r * 8) * PW);
#pragma unroll
            for (int i = 0; i < 4 * T::GPL; ++i) {
                const float4 q = row[i];
                acc = __fmaf_rn(q.x, q.y, acc);
                acc = __fmaf_rn(q.z, q.w, acc);
            }
        }
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------- attention part
// execute_attn (transformer.cpp:397-455) for query head qh, CTA `part` of `cph`: scores for a contiguous share of
// the keys, score exchange through L2, full softmax (redundantly per part), PV chains for HS/cph head dims.
template <int HS>
__device__ __forceinline__ void attention_part(const MegaParams& p, uint8_t* smem, int layer, int qh, int part,
                                               unsigned long long head_target, int tid, Prof& pf) {
    constexpr int EPL = HS / 8;
    const int cph = p.cph;
    const int DW = HS / cph;                                // head dims owned by this part
    float* att = reinterpret_cast<float*>(smem + p.off_att);
    float* q_s = reinterpret_cast<float*>(smem + p.off_misc) + 32;
    float* k_s = q_s + HS;
    float* v_s = k_s + HS;
    float* red = reinterpret_cast<float*>(smem + p.off_misc);
    float* v_stage = reinterpret_cast<float*>(smem + p.off_vstage);
    const int VR = p.v_chunk_rows;

    const int hgs = p.n_heads / p.n_kv_heads;
    const int kvh = qh / hgs, g = qh % hgs;
    const int dim = p.n_heads * HS, kv_dim = p.n_kv_heads * HS;
    const int pos = __ldcg(&p.st->pos), bs = __ldcg(&p.st->bs);     // state changes between steps of one launch: bypass L1
    const int n = pos + 1;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t cache_off = ((size_t)layer * p.n_kv_heads + kvh) * p.max_seq * HS;
    float* kc = p.k_cache + cache_off;
    float* vc = p.v_cache + cache_off;
    const int d0 = part * DW;

    // V stream for this part's dims: rows [0, pos) from the cache in chunks of VR rows, 3 chunks in flight
    const int n_chunks = ceil_div(n, VR);
    const int ppr = DW / 4;                                 // 16-byte pieces per row
    auto issue_v_chunk = [&](int ch) {
        if (ch < n_chunks) {
            const int t0 = ch * VR;
            float* dst = v_stage + (size_t)(ch % 3) * VR * DW;
            const int rows = min(VR, pos - t0);
            for (int i = tid; i < rows * ppr; i += kConsumerThreads) {
                const int row = i / ppr, pc = i - row * ppr;
                cp_async16(dst + (size_t)row * DW + pc * 4, vc + (size_t)(t0 + row) * HS + d0 + pc * 4);
            }
        }
        cp_async_commit();
    };
    issue_v_chunk(0); issue_v_chunk(1); issue_v_chunk(2);

    // RoPE + KV append (rope_v2 tf_operators.cpp:355-402; transformer.cpp:431-439)
    const float* qkv = p.qkv;
    if (tid < HS / 2) {
        // sequence_rope_v2 (tensor.h:262-270) walks all bs*hgs rows of the q tensor with position pos0 + row: query head g
        // of a GQA group is rotated at pos + g*bs (== pos when n_heads == n_kv_heads).  Reproduced, not fixed.
        const float2 cs2 = __ldg(reinterpret_cast<const float2*>(p.rope) + (size_t)(pos + g * bs) * (HS / 2) + tid);
        const float2 x = __ldcg(reinterpret_cast<const float2*>(qkv + (size_t)qh * HS) + tid);
        float o0, o1;
        rope_pair(cs2.x, cs2.y, x.x, x.y, o0, o1);
        q_s[2 * tid] = o0; q_s[2 * tid + 1] = o1;
    } else if (tid < HS) {
        const int i = tid - HS / 2;
        const float2 cs2 = __ldg(reinterpret_cast<const float2*>(p.rope) + (size_t)pos * (HS / 2) + i);
        const float2 x = __ldcg(reinterpret_cast<const float2*>(qkv + dim + (size_t)kvh * HS) + i);
        float o0, o1;
        rope_pair(cs2.x, cs2.y, x.x, x.y, o0, o1);
        k_s[2 * i] = o0; k_s[2 * i + 1] = o1;
        if (g == 0 && part == 0) {
            float* krow = kc + (size_t)pos * HS;
            krow[((2 * i) & 7) * EPL + ((2 * i) >> 3)] = o0;
            krow[((2 * i + 1) & 7) * EPL + ((2 * i + 1) >> 3)] = o1;
        }
    } else if (tid < HS + HS / 4) {
        const int i = tid - HS;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(qkv + dim + kv_dim + (size_t)kvh * HS) + i);
        reinterpret_cast<float4*>(v_s)[i] = v;
        if (g == 0 && part == 0) reinterpret_cast<float4*>(vc + (size_t)pos * HS)[i] = v;
    }
    consumer_sync();
    pf.stop(tid, 10);

    // scores for this part's keys: float dot_product_avx256 (x86_simd.cpp:1447-1468): 8 FMA chains, then 0 + l0 + ... + l7
    const int per = ceil_div(ceil_div(n, cph), 4) * 4;
    const int tb = part * per, te = min(n, tb + per);
    float* att_g = p.att_scratch + (size_t)qh * p.max_seq;
    {
        const int rr = lane >> 3, j = lane & 7;
        float qr[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qr[i] = q_s[8 * i + j];
        constexpr int UU = 2;
#pragma unroll 1
        for (int base = tb; base < te; base += kConsumerWarps * 4 * UU) {
            float4 kv[UU][EPL / 4];
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                const int t = base + (u * kConsumerWarps + warp) * 4 + rr;
                if (t < pos && t < te) {
                    const float4* kp = reinterpret_cast<const float4*>(kc + (size_t)t * HS + j * EPL);
#pragma unroll
                    for (int q = 0; q < EPL / 4; ++q) kv[u][q] = __ldcg(kp + q);
                } else if (t == pos && t < te) {
#pragma unroll
                    for (int q = 0; q < EPL / 4; ++q)
                        kv[u][q] = make_float4(k_s[8 * (4 * q) + j], k_s[8 * (4 * q + 1) + j], k_s[8 * (4 * q + 2) + j], k_s[8 * (4 * q + 3) + j]);
                } else {
#pragma unroll
                    for (int q = 0; q < EPL / 4; ++q) kv[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                const int t = base + (u * kConsumerWarps + warp) * 4 + rr;
                float acc = 0.0f;
#pragma unroll
                for (int q = 0; q < EPL / 4; ++q) {
                    acc = __fmaf_rn(kv[u][q].x, qr[4 * q], acc);
                    acc = __fmaf_rn(kv[u][q].y, qr[4 * q + 1], acc);
                    acc = __fmaf_rn(kv[u][q].z, qr[4 * q + 2], acc);
                    acc = __fmaf_rn(kv[u][q].w, qr[4 * q + 3], acc);
                }
                float tot = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) tot = __fadd_rn(tot, __shfl_sync(kFull, acc, (rr << 3) + k));
                if (j == 0 && t < te) {
                    const float sc = __fmul_rn(tot, p.attn_scale);      // att.multiply(attn_scale), transformer.cpp:443
                    att[t] = sc;
                    if (cph > 1) att_g[t] = sc;
                }
            }
        }
    }
    // exchange: wait until all parts of this head have published their scores, then fetch the others'
    consumer_sync();
    pf.stop(tid, 11);
    if (cph > 1) {
        if (tid == 0) {
            red_release_add_u64(p.head_ctr + qh, 1ull);
            while (ld_acquire_u64(p.head_ctr + qh) < head_target) { }
        }
        consumer_sync();
        for (int t = tid; t < n; t += kConsumerThreads)
            if (t < tb || t >= te) att[t] = __ldcg(att_g + t);
        consumer_sync();
    }
    pf.stop(tid, 12);

    // softmax_sisd (tf_operators.cpp:176-186): max, expf(x - max), serial sum, divide
    float m = -INFINITY;
    for (int t = tid; t < n; t += kConsumerThreads) m = fmaxf(m, att[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) red[warp] = m;
    consumer_sync();
    m = red[0];
#pragma unroll
    for (int w = 1; w < kConsumerWarps; ++w) m = fmaxf(m, red[w]);
    for (int t = tid; t < n; t += kConsumerThreads) att[t] = expf_exact(__fsub_rn(att[t], m));
    consumer_sync();
    if (tid == 0) {
        float sum = 0.0f;
        int t = 0;
#pragma unroll 1
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
    consumer_sync();
    const float sum = red[16];
    for (int t = tid; t < n; t += kConsumerThreads) att[t] = __fdiv_rn(att[t], sum);
    consumer_sync();
    pf.stop(tid, 13);

    // weighted_sum (tf_operators.cpp:325-350): o = V[0]*w0; t >= 1: if |w_t| > 1e-15: o = fma(V[t], w_t, o) — one chain
    // per head dim, branch-free (a skipped row keeps o through a select)
    float o = 0.0f;
#pragma unroll 1
    for (int ch = 0; ch < n_chunks; ++ch) {
        cp_async_wait<2>();
        consumer_sync();
        pf.stop(tid, 14);
        if (tid < DW) {
            float* vb = v_stage + (size_t)(ch % 3) * VR * DW;
            const int t0 = ch * VR, t1 = min(n, t0 + VR);
            if (pos >= t0 && pos < t1) vb[(size_t)(pos - t0) * DW + tid] = v_s[d0 + tid];   // the new row joins its chunk
            int t = t0;
            if (t == 0) { o = __fmul_rn(vb[tid], att[0]); t = 1; }
            constexpr int PB = 8;
#pragma unroll 1
            for (; t + PB <= t1; t += PB) {
                float vv[PB], ww[PB];
#pragma unroll
                for (int u = 0; u < PB; ++u) { ww[u] = att[t + u]; vv[u] = vb[(size_t)(t + u - t0) * DW + tid]; }
#pragma unroll
                for (int u = 0; u < PB; ++u) {
                    const float nf = __fmaf_rn(vv[u], ww[u], o);
                    o = (fabsf(ww[u]) > 1e-15f) ? nf : o;
                }
            }
            for (; t < t1; ++t) {
                const float w = att[t];
                const float nf = __fmaf_rn(vb[(size_t)(t - t0) * DW + tid], w, o);
                o = (fabsf(w) > 1e-15f) ? nf : o;
            }
        }
        consumer_sync();
        pf.stop(tid, 15);
        issue_v_chunk(ch + 3);
        pf.stop(tid, 16);
    }
    cp_async_wait<0>();
    if (tid < DW) p.attn[(size_t)qh * HS + d0 + tid] = o;
    pf.stop(tid, 17);
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int QT, int GS, int HS>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel(const __grid_constant__ MegaParams p) {
    using T = Traits<QT, GS>;
    using R = Ring<QT, GS>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem + p.off_ring;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_bars);
    uint64_t* empty = full + p.n_slots;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_slots = p.n_slots;

    if (tid == 0) {
        for (int i = 0; i < n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_phases = 4 * p.n_layers + 1;

    if (warp == kConsumerWarps) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t sc = 0;
#pragma unroll 1
            for (int step = 0; step < p.n_steps; ++step) {
#pragma unroll 1
                for (int pi = 0; pi < n_phases; ++pi) {
                    const PhaseShape ph = phase_shape(p, pi);
                    const int t0 = (int)((long long)ph.n_tasks * blockIdx.x / gridDim.x);
                    const int t1 = (int)((long long)ph.n_tasks * (blockIdx.x + 1) / gridDim.x);
                    const int spt_tile = ceil_div(ph.nkb, R::U), spt = ph.tt * spt_tile;
#pragma unroll 1
                    for (int r0 = t0; r0 < t1; r0 += kConsumerWarps) {
                        const int nw = min(kConsumerWarps, t1 - r0);
#pragma unroll 1
                        for (int s = 0; s < spt; ++s) {
                            const int tile = s / spt_tile, ks = s - tile * spt_tile;
                            const uint32_t bytes = (uint32_t)min(R::U, ph.nkb - ks * R::U) * T::UNIT_BYTES;
#pragma unroll 1
                            for (int w = 0; w < nw; ++w) {
                                const uint32_t idx = sc + (uint32_t)(s * nw + w);
                                const uint32_t slot = idx % (uint32_t)n_slots, k = idx / (uint32_t)n_slots;
                                mbar_wait_sleep(&empty[slot], (k & 1u) ^ 1u);
                                mbar_arrive_expect_tx(&full[slot], bytes);
                                const uint8_t* src = ph.w + (((size_t)(r0 + w) * ph.tt + tile) * ph.nkb + (size_t)ks * R::U) * T::UNIT_BYTES;
                                bulk_g2s(ring + (size_t)slot * R::SLOT_BYTES, src, bytes, &full[slot]);
                            }
                        }
                        sc += (uint32_t)(nw * spt);
                    }
                }
            }
        }
        return;
    }

    // ================= consumers =================
    uint8_t* xq = smem + p.off_xq;
    float* xs = reinterpret_cast<float*>(smem + p.off_xs);
    float* xf = reinterpret_cast<float*>(smem + p.off_xf);
    float* misc = reinterpret_cast<float*>(smem + p.off_misc);
    float* cs = reinterpret_cast<float*>(smem + p.off_chain) + (size_t)warp * (R::U * 32 * 2 * T::GPL);
    const uint4* xq4 = reinterpret_cast<const uint4*>(xq);
    const int r = lane >> 3, l = lane & 7;

    unsigned long long bar_target = __ldcg(p.bar_ctr + 1) + gridDim.x;
    const int n_attn_ctas = p.n_heads * p.cph;
    const bool attn_cta = (int)blockIdx.x < n_attn_ctas;
    const int my_head = blockIdx.x / p.cph, my_part = blockIdx.x % p.cph;
    unsigned long long head_base = attn_cta ? __ldcg(p.head_ctr + p.n_heads + my_head) : 0ull;
    unsigned long long attn_rounds = 0;
    uint32_t sc = 0;
    Prof pf;
    pf.p = p.prof ? p.prof + (size_t)blockIdx.x * 32 : nullptr;
    pf.t0 = pf.p ? gtimer() : 0ull;

#pragma unroll 1
    for (int step = 0; step < p.n_steps; ++step) {
        const int token = __ldcg(&p.st->token);
        const float* emb_row = p.emb + (size_t)token * p.dim;
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
#pragma unroll 1
        for (int pi = 0; pi < n_phases; ++pi) {
            const PhaseShape ph = phase_shape(p, pi);
            const int layer = pi >> 2, pk = (pi == n_phases - 1) ? 4 : (pi & 3);
            const MegaLayer* L = p.layers + (pk == 4 ? 0 : layer);
            // layer 0 reads the embedding row instead of x1 (transformer.cpp:115-122)
            const float* x1_in = (layer == 0 && pk < 2) ? emb_row : p.x1;
            // ---- producer side of the math: the activation vector of this phase
            const float* in = (pk == 1) ? p.attn : (pk == 3) ? p.hd : x1_in;
            const float* gain = (pk == 0) ? L->att_norm : (pk == 2) ? L->ffn_norm : (pk == 4) ? p.out_norm : nullptr;
            build_activation<QT, GS>(xq, xs, xf, misc, in, gain, ph.K, ph.nkb, (pk == 4 && blockIdx.x == 0) ? p.tap_norm : nullptr, tid, pf);
            pf.stop(tid, 1);
            // ---- drain this CTA's stages of the phase
            float* out = (pk == 0) ? p.qkv : (pk == 2) ? p.hd : (pk == 4) ? p.logits : p.x1;
            {
                const int t0 = (int)((long long)ph.n_tasks * blockIdx.x / gridDim.x);
                const int t1 = (int)((long long)ph.n_tasks * (blockIdx.x + 1) / gridDim.x);
                const int spt_tile = ceil_div(ph.nkb, R::U), spt = ph.tt * spt_tile;
#pragma unroll 1
                for (int r0 = t0; r0 < t1; r0 += kConsumerWarps) {
                    const int nw = min(kConsumerWarps, t1 - r0);
                    if (warp < nw) {
                        float acc = 0.0f, acc_first = 0.0f;
#pragma unroll 1
                        for (int s = 0; s < spt; ++s) {
                            const uint32_t idx = sc + (uint32_t)(s * nw + warp);
                            const uint32_t slot = idx % (uint32_t)n_slots, k = idx / (uint32_t)n_slots;
                            const int tile = s / spt_tile, ks = s - tile * spt_tile, kb = ks * R::U;
                            const int nu = min(R::U, ph.nkb - kb);
                            if (tile == 1 && ks == 0) { acc_first = acc; acc = 0.0f; }       // W1 tile done, W3 tile starts
                            mbar_wait(&full[slot], k & 1u);
                            acc = stage_chain<QT, GS>(ring + (size_t)slot * R::SLOT_BYTES, nu, xq4 + (size_t)kb * (T::KB_BYTES / 16),
                                                      xs + kb * 8 * T::GPL, cs, lane, acc);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&empty[slot]);
                        }
                        const int row = (r0 + warp) * 4 + r;
                        if (l == 0 && row < ph.M) {
                            float v;
                            if (pk == 0 || pk == 4) v = acc;
                            else if (pk == 2) v = swiglu_exact(acc_first, acc);
                            else v = __fadd_rn(__ldcg(((pk == 1) ? x1_in : p.x1) + row), acc);     // x1 += tmp (tensor.cpp:723)
                            out[row] = v;
                            if (pk == 4 && (v > best_v || (v == best_v && row < best_i))) { best_v = v; best_i = row; }
                        }
                    }
                    sc += (uint32_t)(nw * spt);
                }
            }
            pf.stop(tid, 2 + pk);
            if (pk == 4) break;
            grid_barrier(p.bar_ctr, bar_target, tid);
            pf.stop(tid, 0);
            if (pk == 0) {
                // ---- attention (transformer.cpp:136, :397-455)
                if (attn_cta) {
                    ++attn_rounds;
                    attention_part<HS>(p, smem, layer, my_head, my_part, head_base + attn_rounds * (unsigned long long)p.cph, tid, pf);
                }
                pf.stop(tid, 7);
                grid_barrier(p.bar_ctr, bar_target, tid);
                pf.stop(tid, 0);
            }
        }
        // ---- argmax (sampler.cpp:36-46: first index of the strict maximum) and state advance
        {
            float bv = best_v; int bi = best_i;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, bv, o);
                const int oi = __shfl_xor_sync(kFull, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            float* sv = misc + 8; int* si = reinterpret_cast<int*>(misc + 16);
            if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
            consumer_sync();
            if (tid == 0) {
                for (int w = 1; w < kConsumerWarps; ++w)
                    if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
                p.am_val[blockIdx.x] = bv; p.am_idx[blockIdx.x] = bi;
            }
        }
        grid_barrier(p.bar_ctr, bar_target, tid);
        if (blockIdx.x == 0 && warp == 0) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int i = lane; i < (int)gridDim.x; i += 32) {
                const float v = __ldcg(p.am_val + i); const int ix = __ldcg(p.am_idx + i);
                if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, bv, o);
                const int oi = __shfl_xor_sync(kFull, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) {
                if (bi == 0x7fffffff) bi = 0;
                SeqState* st = p.st;
                *p.argmax_out = bi;
                const int no = __ldcg(&st->n_out);
                if (no < p.out_cap) p.out_tokens[no] = bi;
                st->n_out = no + 1; st->token = bi; st->pos = __ldcg(&st->pos) + 1; st->bs = 1;
            }
        }
        grid_barrier(p.bar_ctr, bar_target, tid);      // the new state is visible to every CTA before the next token
        pf.stop(tid, 0);
    }
    // publish the counter bases for the next launch (every CTA has passed the last barrier)
    if (tid == 0) {
        if (blockIdx.x == 0) p.bar_ctr[1] = bar_target - gridDim.x;
        if (attn_cta && my_part == 0) p.head_ctr[p.n_heads + my_head] = head_base + attn_rounds * (unsigned long long)p.cph;
    }
}

}  // namespace fl
