// megakernel.cuh — the decode step as ONE persistent sm_100a kernel (one CTA per SM).
//
// Why: a decode token is 161 dependent matrix-vector phases of 18-140 MB each.  Launched as separate kernels
// every phase pays launch latency, a cold pipeline and its activation prologue (for rmsnorm a 1024-step serial FP32
// chain, 2 us) with HBM idle.  Here the weight stream never stops:
//
//   * warp 8 (one elected lane) is a TMA producer: it walks the token's static weight schedule and issues
//     cp.async.bulk global->shared copies into a ring of stages guarded by full/empty mbarriers.  Weights do not
//     depend on activations, so the producer runs ahead across phase, layer and token boundaries; only ring
//     capacity (~150-170 KB per SM = ~4 us of this SM's HBM share) limits it.
//   * warps 0-7 are consumers: per phase they (1) pass a grid-wide barrier, (2) rebuild the quantised activation
//     vector in shared memory (rmsnorm chain, quantise), (3) drain their stages with the exact per-unit code of
//     kernels.cuh (unit_chain) and (4) write their rows.  While they do (1)+(2) the ring fills, so HBM stays busy.
//   * attention runs between the QKV and Wo phases on n_heads * CPH CTAs (CPH CTAs share one head: keys are split
//     for QK^T, head dims are split for the PV chains), exchanging the score vector through L2.
//
// Each CTA owns a contiguous range of 4-row tiles of every matrix (rows/148 +- 4), so its share of a phase is ONE
// contiguous byte range of the streaming layout and every stage is a single bulk copy.
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(a), "r"(parity) : "memory");
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

// grid-wide barrier for the consumer warps of all CTAs (all CTAs are co-resident: cooperative launch)
__device__ __forceinline__ void grid_barrier(unsigned long long* ctr, unsigned long long& target, int tid) {
    consumer_sync();
    if (tid == 0) {
        __threadfence();
        atomicAdd(ctr, 1ull);
        while (ld_acquire_u64(ctr) < target) { }
        __threadfence();
    }
    consumer_sync();
    target += gridDim.x;
}

// ---------------------------------------------------------------------------------------------- schedule
// Both the producer and every consumer warp walk this; it must be a pure function of (phase shape, CTA id).
struct PhaseShape {
    const uint8_t* w;
    int n_tasks;          // row tiles (or W1/W3 tile pairs)
    int upt;              // units per task
};

template <int QT, int GS>
struct Ring {
    using T = Traits<QT, GS>;
    static constexpr int U = (QT == Q_INT8) ? 4 : 2;             // units per stage
    static constexpr int SLOT_BYTES = U * T::UNIT_BYTES;
};

template <int QT, int GS>
__device__ __forceinline__ void produce_phase(const PhaseShape& ph, uint8_t* ring, uint64_t* full, uint64_t* empty,
                                              int n_slots, uint32_t& stage_count) {
    using R = Ring<QT, GS>;
    using T = Traits<QT, GS>;
    const int t0 = (int)((long long)ph.n_tasks * blockIdx.x / gridDim.x);
    const int t1 = (int)((long long)ph.n_tasks * (blockIdx.x + 1) / gridDim.x);
    const int spt = ceil_div(ph.upt, R::U);
    for (int r0 = t0; r0 < t1; r0 += kConsumerWarps) {
        const int nw = min(kConsumerWarps, t1 - r0);
        for (int s = 0; s < spt; ++s) {
            const uint32_t bytes = (uint32_t)min(R::U, ph.upt - s * R::U) * T::UNIT_BYTES;
            for (int w = 0; w < nw; ++w) {
                const uint32_t idx = stage_count + (uint32_t)(s * nw + w);
                const uint32_t slot = idx % (uint32_t)n_slots, k = idx / (uint32_t)n_slots;
                mbar_wait(&empty[slot], (k & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[slot], bytes);
                const uint8_t* src = ph.w + ((size_t)(r0 + w) * ph.upt + (size_t)s * R::U) * T::UNIT_BYTES;
                bulk_g2s(ring + (size_t)slot * R::SLOT_BYTES, src, bytes, &full[slot]);
            }
        }
        stage_count += (uint32_t)(nw * spt);
    }
}

// ---------------------------------------------------------------------------------------------- consumer GEMV phase
struct ConsumerCtx {
    uint8_t* ring; uint64_t* full; uint64_t* empty; int n_slots;
    uint8_t* xq; float* xs; float* xf; float* chain; float* misc;
    int tid, warp, lane;
};

template <int QT, int GS, int PRO>
__device__ __forceinline__ void build_activation(const ConsumerCtx& c, const float* in, const float* gain, int K, int nkb, float* tap) {
    using T = Traits<QT, GS>;
    const int kpad = nkb * kKBlockElems;
    // zero the padded tail so padded groups contribute fma(0, 0, acc) == acc
    for (int i = K * T::ES + c.tid * 4; i < kpad * T::ES; i += kConsumerThreads * 4) *reinterpret_cast<uint32_t*>(c.xq + i) = 0u;
    for (int i = K / GS + c.tid; i < nkb * 8 * T::GPL; i += kConsumerThreads) c.xs[i] = 0.0f;
    if (PRO == PRO_RMS_QUANT) {
        for (int i = c.tid; i < K / 4; i += kConsumerThreads)
            reinterpret_cast<float4*>(c.xf)[i] = __ldcg(reinterpret_cast<const float4*>(in) + i);
        consumer_sync();
        if (c.warp == 0) {
            const float ss = sumsq_chain_warp0(c.xf, K, c.lane);
            if (c.lane == 0) c.misc[0] = rms_scale(ss, K);
        }
        consumer_sync();
        const float rr = c.misc[0];
        const float* xf = c.xf;
        quantize_block<QT, GS>([&](int e) {
            const float y = __fmul_rn(__fmul_rn(xf[e], __ldg(gain + e)), rr);
            if (tap) tap[e] = y;
            return y;
        }, K, c.xq, c.xs, nullptr, nullptr);
    } else {
        quantize_block<QT, GS>([&](int e) { return __ldcg(in + e); }, K, c.xq, c.xs, nullptr, nullptr);
    }
    consumer_sync();
}

// argmax partial of this warp's finished rows (cls phase): strict maximum, lowest index on ties
struct ArgBest { float v; int i; };

template <int QT, int GS, int EPI, bool TRACK_ARGMAX>
__device__ __forceinline__ void consume_phase(const ConsumerCtx& c, const PhaseShape& ph, int nkb, int M, float* out,
                                              const float* resid, uint32_t& stage_count, ArgBest& best) {
    using R = Ring<QT, GS>;
    using T = Traits<QT, GS>;
    const int t0 = (int)((long long)ph.n_tasks * blockIdx.x / gridDim.x);
    const int t1 = (int)((long long)ph.n_tasks * (blockIdx.x + 1) / gridDim.x);
    const int spt = ceil_div(ph.upt, R::U);
    const int r = c.lane >> 3, l = c.lane & 7;
    const uint4* xq4 = reinterpret_cast<const uint4*>(c.xq);
    float* cs = c.chain + (size_t)c.warp * (2 * 32 * 2 * T::GPL);
    for (int r0 = t0; r0 < t1; r0 += kConsumerWarps) {
        const int nw = min(kConsumerWarps, t1 - r0);
        if (c.warp < nw) {
            const int task = r0 + c.warp;
            float acc = 0.0f, acc_first = 0.0f;
            int kb = 0, unit = 0;
            for (int s = 0; s < spt; ++s) {
                const uint32_t idx = stage_count + (uint32_t)(s * nw + c.warp);
                const uint32_t slot = idx % (uint32_t)c.n_slots, k = idx / (uint32_t)c.n_slots;
                mbar_wait(&c.full[slot], k & 1u);
                const uint8_t* sp = c.ring + (size_t)slot * R::SLOT_BYTES;
                const int nu = min(R::U, ph.upt - s * R::U);
                for (int u = 0; u < nu; ++u) {
                    const uint8_t* up = sp + (size_t)u * T::UNIT_BYTES;
                    uint4 wv[T::NJ];
                    float wsv[T::GPL];
#pragma unroll
                    for (int j = 0; j < T::NJ; ++j) wv[j] = reinterpret_cast<const uint4*>(up)[j * 32 + c.lane];
#pragma unroll
                    for (int gg = 0; gg < T::GPL; ++gg) wsv[gg] = reinterpret_cast<const float*>(up + T::W_BYTES)[c.lane * T::GPL + gg];
                    acc = unit_chain<QT, GS>(wv, wsv, xq4 + (size_t)kb * (T::KB_BYTES / 16), c.xs + kb * 8 * T::GPL,
                                             cs + (unit & 1) * (32 * 2 * T::GPL), c.lane, acc);
                    ++unit;
                    if (++kb == nkb) {
                        kb = 0;
                        if (EPI == EPI_SWIGLU && unit == nkb) { acc_first = acc; acc = 0.0f; }
                    }
                }
                __syncwarp();
                if (c.lane == 0) mbar_arrive(&c.empty[slot]);
            }
            const int row = task * 4 + r;
            if (l == 0 && row < M) {
                float v;
                if (EPI == EPI_STORE) v = acc;
                else if (EPI == EPI_RESADD) v = __fadd_rn(__ldcg(resid + row), acc);      // x1 += tmp (tensor.cpp:723)
                else v = swiglu_exact(acc_first, acc);
                out[row] = v;
                if (TRACK_ARGMAX && (v > best.v || (v == best.v && row < best.i))) { best.v = v; best.i = row; }
            }
        }
        stage_count += (uint32_t)(nw * spt);
    }
}

// ---------------------------------------------------------------------------------------------- attention part
// execute_attn (transformer.cpp:397-455) for query head qh, CTA `part` of `cph`: scores for a contiguous share of
// the keys, score exchange through L2, full softmax (redundantly per part), PV chains for HS/cph head dims.
template <int HS>
__device__ __forceinline__ void attention_part(const MegaParams& p, const ConsumerCtx& c, int layer, int qh, int part,
                                               unsigned long long head_target, uint8_t* smem) {
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
    const int tid = c.tid, warp = c.warp, lane = c.lane;
    const size_t cache_off = ((size_t)layer * p.n_kv_heads + kvh) * p.max_seq * HS;
    float* kc = p.k_cache + cache_off;
    float* vc = p.v_cache + cache_off;
    const int d0 = part * DW;

    // V stream for this part's dims: rows [0, pos) from the cache in chunks of VR rows, 3 chunks in flight
    const int n_chunks = ceil_div(n, VR);
    auto issue_v_chunk = [&](int ch) {
        if (ch < n_chunks) {
            const int t0 = ch * VR;
            float* dst = v_stage + (size_t)(ch % 3) * VR * DW;
            const int rows = min(VR, pos - t0);
            const int ppr = DW / 4;                         // 16-byte pieces per row
            for (int i = tid; i < rows * ppr; i += kConsumerThreads) {
                const int row = i / ppr, pc = i % ppr;
                cp_async16(dst + (size_t)row * DW + pc * 4, vc + (size_t)(t0 + row) * HS + d0 + pc * 4);
            }
        }
        cp_async_commit();
    };
    issue_v_chunk(0); issue_v_chunk(1); issue_v_chunk(2);

    // RoPE + KV append (rope_v2 tf_operators.cpp:355-402; transformer.cpp:431-439)
    const float* qkv = p.qkv;
    if (tid < HS / 2) {
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

    // scores for this part's keys
    const int per = ceil_div(ceil_div(n, cph), 4) * 4;
    const int tb = part * per, te = min(n, tb + per);
    float* att_g = p.att_scratch + (size_t)qh * p.max_seq;
    {
        const int rr = lane >> 3, j = lane & 7;
        float qr[EPL];
#pragma unroll
        for (int i = 0; i < EPL; ++i) qr[i] = q_s[8 * i + j];
        constexpr int UU = 4;
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
                    const float sc = __fmul_rn(tot, p.attn_scale);
                    att[t] = sc;
                    if (cph > 1) att_g[t] = sc;
                }
            }
        }
    }
    // exchange: wait until all parts of this head have published their scores, then fetch the others'
    consumer_sync();
    if (cph > 1) {
        if (tid == 0) {
            __threadfence();
            atomicAdd(p.head_ctr + qh, 1ull);
            while (ld_acquire_u64(p.head_ctr + qh) < head_target) { }
            __threadfence();
        }
        consumer_sync();
        for (int t = tid; t < n; t += kConsumerThreads)
            if (t < tb || t >= te) att[t] = __ldcg(att_g + t);
        consumer_sync();
    }

    // softmax_sisd (tf_operators.cpp:176-186)
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

    // weighted_sum (tf_operators.cpp:325-350): one chain per head dim
    float o = 0.0f;
    for (int ch = 0; ch < n_chunks; ++ch) {
        cp_async_wait<2>();
        consumer_sync();
        if (tid < DW) {
            const float* vb = v_stage + (size_t)(ch % 3) * VR * DW;
            const int t0 = ch * VR, t1 = min(n, t0 + VR);
            for (int t = t0; t < t1; ++t) {
                const float v = (t == pos) ? v_s[d0 + tid] : vb[(size_t)(t - t0) * DW + tid];
                const float w = att[t];
                if (t == 0) o = __fmul_rn(v, w);
                else if (fabsf(w) > 1e-15f) o = __fmaf_rn(v, w, o);
            }
        }
        consumer_sync();
        issue_v_chunk(ch + 3);
    }
    cp_async_wait<0>();
    if (tid < DW) p.attn[(size_t)qh * HS + d0 + tid] = o;
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int QT, int GS, int HS>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel(const MegaParams p) {
    using T = Traits<QT, GS>;
    using R = Ring<QT, GS>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem + p.off_ring;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_bars);
    uint64_t* empty = full + p.n_slots;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < p.n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int kv_dim = p.n_kv_heads * HS;
    const int nkb_d = ceil_div(p.dim, kKBlockElems), nkb_h = ceil_div(p.hidden, kKBlockElems);
    const int tiles_d = ceil_div(p.dim, 4);
    const int n_layers = p.n_layers;
    (void)kv_dim;

    if (warp == kConsumerWarps) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t sc = 0;
            for (int step = 0; step < p.n_steps; ++step) {
                for (int l = 0; l < n_layers; ++l) {
                    const MegaLayer L = p.layers[l];
                    produce_phase<QT, GS>(PhaseShape{L.qkv, ceil_div(p.qkv_rows, 4), nkb_d}, ring, full, empty, p.n_slots, sc);
                    produce_phase<QT, GS>(PhaseShape{L.wo, tiles_d, nkb_d}, ring, full, empty, p.n_slots, sc);
                    produce_phase<QT, GS>(PhaseShape{L.w13, ceil_div(p.hidden, 4), 2 * nkb_d}, ring, full, empty, p.n_slots, sc);
                    produce_phase<QT, GS>(PhaseShape{L.w2, tiles_d, nkb_h}, ring, full, empty, p.n_slots, sc);
                }
                produce_phase<QT, GS>(PhaseShape{p.cls, ceil_div(p.vocab, 4), nkb_d}, ring, full, empty, p.n_slots, sc);
            }
        }
        return;
    }

    // ================= consumers =================
    ConsumerCtx c;
    c.ring = ring; c.full = full; c.empty = empty; c.n_slots = p.n_slots;
    c.xq = smem + p.off_xq; c.xs = reinterpret_cast<float*>(smem + p.off_xs); c.xf = reinterpret_cast<float*>(smem + p.off_xf);
    c.chain = reinterpret_cast<float*>(smem + p.off_chain); c.misc = reinterpret_cast<float*>(smem + p.off_misc);
    c.tid = tid; c.warp = warp; c.lane = lane;

    unsigned long long bar_target = p.bar_ctr[1] + gridDim.x;
    unsigned long long head_base = 0;
    const int n_attn_ctas = p.n_heads * p.cph;
    const bool attn_cta = (int)blockIdx.x < n_attn_ctas;
    const int my_head = blockIdx.x / p.cph, my_part = blockIdx.x % p.cph;
    if (attn_cta) head_base = p.head_ctr[p.n_heads + my_head];
    unsigned long long attn_rounds = 0;
    uint32_t sc = 0;
    ArgBest best;

    for (int step = 0; step < p.n_steps; ++step) {
        const int token = __ldcg(&p.st->token);
        const float* emb_row = p.emb + (size_t)token * p.dim;
        for (int l = 0; l < n_layers; ++l) {
            const MegaLayer L = p.layers[l];
            const float* resid_in = (l == 0) ? emb_row : p.x1;        // layer 0 reads the embedding row (transformer.cpp:115-122)
            // ---- QKV: qkv = Wqkv * quantize(rmsnorm(x1))                         (:132-135)
            build_activation<QT, GS, PRO_RMS_QUANT>(c, resid_in, L.att_norm, p.dim, nkb_d, nullptr);
            consume_phase<QT, GS, EPI_STORE, false>(c, PhaseShape{L.qkv, ceil_div(p.qkv_rows, 4), nkb_d}, nkb_d, p.qkv_rows, p.qkv, nullptr, sc, best);
            grid_barrier(p.bar_ctr, bar_target, tid);
            // ---- attention                                                     (:136, :397-455)
            if (attn_cta) {
                ++attn_rounds;
                attention_part<HS>(p, c, l, my_head, my_part, head_base + attn_rounds * (unsigned long long)p.cph, smem);
            }
            grid_barrier(p.bar_ctr, bar_target, tid);
            // ---- x1 += Wo * quantize(attn)                                       (:138-139)
            build_activation<QT, GS, PRO_QUANT>(c, p.attn, nullptr, p.dim, nkb_d, nullptr);
            consume_phase<QT, GS, EPI_RESADD, false>(c, PhaseShape{L.wo, tiles_d, nkb_d}, nkb_d, p.dim, p.x1, resid_in, sc, best);
            grid_barrier(p.bar_ctr, bar_target, tid);
            // ---- hd = swiglu(W1 q, W3 q), q = quantize(rmsnorm(x1))              (:144-147)
            build_activation<QT, GS, PRO_RMS_QUANT>(c, p.x1, L.ffn_norm, p.dim, nkb_d, nullptr);
            consume_phase<QT, GS, EPI_SWIGLU, false>(c, PhaseShape{L.w13, ceil_div(p.hidden, 4), 2 * nkb_d}, nkb_d, p.hidden, p.hd, nullptr, sc, best);
            grid_barrier(p.bar_ctr, bar_target, tid);
            // ---- x1 += W2 * quantize(hd)                                         (:149-150)
            build_activation<QT, GS, PRO_QUANT>(c, p.hd, nullptr, p.hidden, nkb_h, nullptr);
            consume_phase<QT, GS, EPI_RESADD, false>(c, PhaseShape{L.w2, tiles_d, nkb_h}, nkb_h, p.dim, p.x1, p.x1, sc, best);
            grid_barrier(p.bar_ctr, bar_target, tid);
        }
        // ---- logits = Wcls * quantize(rmsnorm(x1)); argmax                        (:154-160, sampler.cpp:36-46)
        build_activation<QT, GS, PRO_RMS_QUANT>(c, p.x1, p.out_norm, p.dim, nkb_d, blockIdx.x == 0 ? p.tap_norm : nullptr);
        best.v = -INFINITY; best.i = 0x7fffffff;
        consume_phase<QT, GS, EPI_STORE, true>(c, PhaseShape{p.cls, ceil_div(p.vocab, 4), nkb_d}, nkb_d, p.vocab, p.logits, nullptr, sc, best);
        {
            float bv = best.v; int bi = best.i;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, bv, o);
                const int oi = __shfl_xor_sync(kFull, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            float* sv = c.misc + 8; int* si = reinterpret_cast<int*>(c.misc + 16);
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
                if (st->n_out < p.out_cap) p.out_tokens[st->n_out] = bi;
                st->n_out += 1; st->token = bi; st->pos += 1; st->bs = 1;
                __threadfence();
            }
        }
        grid_barrier(p.bar_ctr, bar_target, tid);      // the new state is visible to every CTA before the next token
    }
    // publish the counter bases for the next launch (every CTA has passed the last barrier)
    if (tid == 0) {
        if (blockIdx.x == 0) p.bar_ctr[1] = bar_target - gridDim.x;
        if (attn_cta && my_part == 0) p.head_ctr[p.n_heads + my_head] = head_base + attn_rounds * (unsigned long long)p.cph;
    }
}

}  // namespace fl
