// batch_kernels.cuh — the per-row operators around the tensor-core GEMM (tc_gemm.cuh) when forward() carries several
// activation rows: a prompt chunk of one sequence (transformer.cpp:105-151 with bs > 1) or one new token of each of several
// sequences (fl_forward_batch).  Every row's arithmetic is the single-row arithmetic of kernels.cuh / megakernel.cuh,
// in the same order, so a row's bits do not depend on which rows travel with it.
//
//   embed_rows_kernel        embedding fetch                               transformer.cpp:115-122
//   rms_quant_rows_kernel    simd::rmsnorm + quant::quantize -> MMA image  x86_simd.cpp:1754, quant_operators.cpp:26-47
//   quant_rows_kernel        quant::quantize -> MMA image                  quant_operators.cpp:26-47
//   kv_append_rows_kernel    RoPE(k), KV append                            transformer.cpp:423-439
//   attn_rows_kernel         RoPE(q), QK^T, softmax, PV                    transformer.cpp:440-453
//   argmax_rows_kernel       Sampler::sample_argmax + state advance        sampler.cpp:36-46
//
// All of them are launched with programmatic stream serialisation: they release their dependents at once (the next GEMM
// then fills its weight ring while they run) and wait for their own producers before touching data.
#pragma once
#include "tc_gemm.cuh"

namespace fl {

struct RowMeta {       // one activation row of a batched forward (device memory)
    int token;         // input token id
    int slot;          // KV slot (sequence)
    int pos;           // position of the token in its sequence
    int bs;            // tokens in the reference forward() call this row belongs to (GQA RoPE quirk, see attn kernels)
};

// V cache addressing: the persistent decode kernel keeps, per (layer, kv head), `cph` column blocks of DW = HS / cph head
// dims with 4 consecutive positions adjacent ([block][max_seq / 4][DW][4]); the per-phase kernels keep natural rows.
struct VLayout {
    int dw;            // column block width (HS: natural rows)
    int blocked4;      // 1: positions interleaved by 4 inside a block
    int max_seq;
};
__device__ __forceinline__ size_t v_index(const VLayout& L, int t, int d) {
    if (L.blocked4) return (size_t)(d / L.dw) * L.max_seq * L.dw + ((size_t)(t >> 2) * L.dw + d % L.dw) * 4 + (t & 3);
    return (size_t)(d / L.dw) * L.max_seq * L.dw + (size_t)t * L.dw + d % L.dw;
}

// rows from the device-resident sequence states (decode of several sequences without a host round trip)
__global__ void rows_from_states_kernel(const SeqState* __restrict__ st, RowMeta* __restrict__ rows, int n) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = threadIdx.x;
    if (i < n) { rows[i].token = st[i].token; rows[i].slot = i; rows[i].pos = st[i].pos; rows[i].bs = 1; }
}

__global__ void embed_rows_kernel(const float* __restrict__ table, const RowMeta* __restrict__ rows, float* __restrict__ x1, int dim) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(table + (size_t)rows[i].token * dim);
    float4* dst = reinterpret_cast<float4*>(x1 + (size_t)i * dim);
    for (int k = threadIdx.x; k < dim / 4; k += blockDim.x) dst[k] = src[k];
}

// quantise 8-lane groups of one row into the activation image; val(e) produces element e
template <int GS, typename ValFn>
__device__ __forceinline__ void quant_row_to_image(ValFn val, int n, uint8_t* img, int i, int N, float* tap) {
    constexpr int PER = GS / 8;
    const int G = n / GS, sub = threadIdx.x & 7;
    for (int g = threadIdx.x >> 3; g < ceil_div(G, kThreads / 8) * (kThreads / 8); g += kThreads / 8) {
        const bool live = g < G;
        float v[PER];
        float m = 0.0f;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            v[u] = live ? val(g * GS + sub * PER + u) : 0.0f;
            m = fmaxf(m, fabsf(v[u]));
        }
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 1));
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
        m = fmaxf(m, __shfl_xor_sync(kFull, m, 4));
        if (!live) continue;
        const float sc = __fdiv_rn(m, 127.0f);                               // quant_operators.cpp:26-47
        if (sub == 0) *reinterpret_cast<float*>(img + tc_xs_offset(i, g, N, GS)) = sc;
        const int e0 = g * GS + sub * PER;
        uint32_t pk[PER / 4];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const uint32_t q = (uint32_t)cvtt_x86(__fdiv_rn(v[u], sc)) & 0xffu;
            pk[u / 4] = (u % 4 == 0) ? q : (pk[u / 4] | (q << (8 * (u % 4))));
            if (tap) tap[e0 + u] = v[u];
        }
        uint32_t* dst = reinterpret_cast<uint32_t*>(img + tc_xq_offset(i, e0, N, GS));     // PER consecutive bytes stay inside one 16-byte piece
#pragma unroll
        for (int u = 0; u < PER / 4; ++u) dst[u] = pk[u];
    }
}

// x[row0 + i * row_stride] -> rmsnorm -> quantise -> image row i.  grid = rows, kThreads threads, dim * 4 bytes of shared memory.
template <int GS>
__global__ void __launch_bounds__(kThreads) rms_quant_rows_kernel(const float* __restrict__ x, size_t row_stride, const float* __restrict__ gain,
                                                                  int dim, uint8_t* __restrict__ img, int N, float* tap) {
    extern __shared__ __align__(16) uint8_t bk_smem[];
    float* xt = reinterpret_cast<float*>(bk_smem);                 // transposed: xt[j * dim/4 + i] = x[4i + j]
    __shared__ float s_r;
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x;
    const float* xr = x + (size_t)i * row_stride;
    for (int k = threadIdx.x; k < dim / 4; k += kThreads) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(xr) + k);
        xt[k] = v.x; xt[(dim >> 2) + k] = v.y; xt[2 * (dim >> 2) + k] = v.z; xt[3 * (dim >> 2) + k] = v.w;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const float ss = sumsq_chain_t(xt, dim, threadIdx.x);
        if (threadIdx.x == 0) s_r = rms_scale(ss, dim);
    }
    __syncthreads();
    const float rr = s_r;
    quant_row_to_image<GS>([&](int e) { return __fmul_rn(__fmul_rn(xt[(e & 3) * (dim >> 2) + (e >> 2)], __ldg(gain + e)), rr); },   // (x*w)*r, x86_simd.cpp:1359
                           dim, img, i, N, (tap && i == (int)gridDim.x - 1) ? tap : nullptr);
}

// fp32 rows [T][K] -> image.  grid (T, ceil(K / GS / 32)): 32 groups per CTA
template <int GS>
__global__ void __launch_bounds__(kThreads) quant_rows_kernel(const float* __restrict__ x, int K, uint8_t* __restrict__ img, int N) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PER = GS / 8;
    const int i = blockIdx.x, G = K / GS;
    const int g = blockIdx.y * (kThreads / 8) + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    const bool live = g < G;
    const float* xr = x + (size_t)i * K;
    float v[PER];
    float m = 0.0f;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        v[u] = live ? __ldcg(xr + g * GS + sub * PER + u) : 0.0f;
        m = fmaxf(m, fabsf(v[u]));
    }
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 1));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 4));
    if (!live) return;
    const float sc = __fdiv_rn(m, 127.0f);
    if (sub == 0) *reinterpret_cast<float*>(img + tc_xs_offset(i, g, N, GS)) = sc;
    const int e0 = g * GS + sub * PER;
    uint32_t pk[PER / 4];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const uint32_t q = (uint32_t)cvtt_x86(__fdiv_rn(v[u], sc)) & 0xffu;
        pk[u / 4] = (u % 4 == 0) ? q : (pk[u / 4] | (q << (8 * (u % 4))));
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(img + tc_xq_offset(i, e0, N, GS));
#pragma unroll
    for (int u = 0; u < PER / 4; ++u) dst[u] = pk[u];
}

struct AttnRowsArgs {
    const float* qkv;          // [T][dim + 2 kv_dim] fp32 (q | k | v), pre-RoPE
    float* k_cache;            // this layer, slot 0: [n_kv_heads][max_seq][HS], rows lane-permuted (kernels.cuh)
    float* v_cache;            // this layer, slot 0: per kv head a block of max_seq * HS floats, see VLayout
    size_t slot_stride;        // floats between the caches of two KV slots
    const float* rope;         // [max_pos][HS/2][2] = (cos, sin), host table (glibc sincosf)
    const RowMeta* rows;
    float* out;                // [T][dim]
    int n_heads, n_kv_heads;
    float attn_scale;
    VLayout vl;
    float* tap_qkv;            // optional: post-RoPE q | k | v of row `tap_row` (debug tap), else NULL
    int tap_row;
};

// RoPE(k) + KV append of every row (transformer.cpp:431-439).  grid (n_kv_heads, T), HS threads.
template <int HS>
__global__ void __launch_bounds__(HS) kv_append_rows_kernel(const AttnRowsArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int kvh = blockIdx.x, i = blockIdx.y, tid = threadIdx.x;
    const RowMeta rm = a.rows[i];
    const int dim = a.n_heads * HS, kv_dim = a.n_kv_heads * HS;
    const float* row = a.qkv + (size_t)i * (dim + 2 * kv_dim);
    float* kc = a.k_cache + (size_t)rm.slot * a.slot_stride + (size_t)kvh * a.vl.max_seq * HS;
    float* vc = a.v_cache + (size_t)rm.slot * a.slot_stride + (size_t)kvh * a.vl.max_seq * HS;
    if (tid < HS / 2) {
        const float2 cs = reinterpret_cast<const float2*>(a.rope)[(size_t)rm.pos * (HS / 2) + tid];
        const float2 x = reinterpret_cast<const float2*>(row + dim + (size_t)kvh * HS)[tid];
        float o0, o1;
        rope_pair(cs.x, cs.y, x.x, x.y, o0, o1);
        float* krow = kc + (size_t)rm.pos * HS;
        krow[k_cache_index(2 * tid)] = o0;
        krow[k_cache_index(2 * tid + 1)] = o1;
        if (a.tap_qkv && i == a.tap_row) { a.tap_qkv[dim + (size_t)kvh * HS + 2 * tid] = o0; a.tap_qkv[dim + (size_t)kvh * HS + 2 * tid + 1] = o1; }
    } else {
        const int d = 2 * (tid - HS / 2);
        const float2 v = reinterpret_cast<const float2*>(row + dim + kv_dim + (size_t)kvh * HS)[tid - HS / 2];
        vc[v_index(a.vl, rm.pos, d)] = v.x;
        vc[v_index(a.vl, rm.pos, d + 1)] = v.y;
        if (a.tap_qkv && i == a.tap_row) { a.tap_qkv[dim + kv_dim + (size_t)kvh * HS + d] = v.x; a.tap_qkv[dim + kv_dim + (size_t)kvh * HS + d + 1] = v.y; }
    }
}

// One (query head, row): scores over the row's cache [0, pos], softmax, PV.  Everything comes from the cache (the row's own
// K/V were appended by kv_append_rows_kernel).  grid (n_heads, T), kThreads threads, (HS + 64 + max_seq + 8) floats of shared memory.
template <int HS>
__global__ void __launch_bounds__(kThreads) attn_rows_kernel(const AttnRowsArgs a) {
    constexpr int EPL = HS / 8;
    extern __shared__ __align__(16) uint8_t bk_smem[];
    float* q_s = reinterpret_cast<float*>(bk_smem);          // [HS] roped q
    float* red = q_s + HS;                                   // [64]
    float* att = red + 64;                                   // [max_seq + 8]
    pdl_launch_dependents();
    pdl_wait();
    const int qh = blockIdx.x, i = blockIdx.y;
    const RowMeta rm = a.rows[i];
    const int hgs = a.n_heads / a.n_kv_heads, kvh = qh / hgs, g = qh % hgs;
    const int dim = a.n_heads * HS, kv_dim = a.n_kv_heads * HS;
    const int pos = rm.pos, n = pos + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* kc = a.k_cache + (size_t)rm.slot * a.slot_stride + (size_t)kvh * a.vl.max_seq * HS;
    const float* vc = a.v_cache + (size_t)rm.slot * a.slot_stride + (size_t)kvh * a.vl.max_seq * HS;

    if (tid < HS / 2) {
        // sequence_rope_v2 (tensor.h:262-270) walks all bs*hgs rows of the q tensor with position pos0 + row: query head g of a
        // GQA group, token i of a bs-token forward, is rotated at pos + g*bs (== pos when n_heads == n_kv_heads).  Reproduced.
        const float2 cs = reinterpret_cast<const float2*>(a.rope)[(size_t)(pos + g * rm.bs) * (HS / 2) + tid];
        const float2 x = reinterpret_cast<const float2*>(a.qkv + (size_t)i * (dim + 2 * kv_dim) + (size_t)qh * HS)[tid];
        float o0, o1;
        rope_pair(cs.x, cs.y, x.x, x.y, o0, o1);
        q_s[2 * tid] = o0; q_s[2 * tid + 1] = o1;
        if (a.tap_qkv && i == a.tap_row) { a.tap_qkv[(size_t)qh * HS + 2 * tid] = o0; a.tap_qkv[(size_t)qh * HS + 2 * tid + 1] = o1; }
    }
    __syncthreads();

    // ---- att[t] = (K[t] . q) * scale: float dot_product_avx256 (x86_simd.cpp:1447-1468): 8 FMA chains, then 0 + l0 + ... + l7
    {
        const int rr = lane >> 3, j = lane & 7;
        float qr[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) qr[u] = q_s[8 * u + j];
        constexpr int U = 4;
        for (int base = 0; base < n; base += kWarps * 4 * U) {
            float4 kv[U][EPL / 4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = base + (u * kWarps + warp) * 4 + rr;
                if (t < n) {
                    const float4* p = reinterpret_cast<const float4*>(kc + (size_t)t * HS) + j;
#pragma unroll
                    for (int c = 0; c < EPL / 4; ++c) kv[u][c] = __ldcg(p + 8 * c);
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
    if (tid < 8) att[n + tid] = 0.0f;
    __syncthreads();
    if (tid == 0) {
        const float4* a4 = reinterpret_cast<const float4*>(att);
        const int nv = n >> 2;
        float sum = 0.0f;
        int k = 0;
        if (nv >= 2) {
            float4 c0 = a4[0], c1 = a4[1];
#pragma unroll 1
            for (; k + 2 <= nv; k += 2) {
                const float4 n0 = a4[k + 2], n1 = a4[k + 3];          // att is padded by 8 zeros: the read-ahead stays inside
                sum = __fadd_rn(sum, c0.x); sum = __fadd_rn(sum, c0.y); sum = __fadd_rn(sum, c0.z); sum = __fadd_rn(sum, c0.w);
                sum = __fadd_rn(sum, c1.x); sum = __fadd_rn(sum, c1.y); sum = __fadd_rn(sum, c1.z); sum = __fadd_rn(sum, c1.w);
                c0 = n0; c1 = n1;
            }
        }
        for (int t = 4 * k; t < n; ++t) sum = __fadd_rn(sum, att[t]);
        red[32] = sum;
    }
    __syncthreads();
    const float sum = red[32];
    for (int t = tid; t < n; t += kThreads) att[t] = __fdiv_rn(att[t], sum);
    __syncthreads();

    // ---- weighted_sum (tf_operators.cpp:325-350): o = V[0]*w0; t >= 1: if |w_t| > 1e-15: o = fma(V[t], w_t, o); one chain per head dim
    if (tid < HS) {
        const int d = tid;
        float o = 0.0f;
        if (a.vl.blocked4) {
            // 4 positions of a dim are one 16-byte load; the loads of the next 16 positions fly while this block's chain runs
            const float* vb = vc + (size_t)(d / a.vl.dw) * a.vl.max_seq * a.vl.dw + (size_t)(d % a.vl.dw) * 4;
            const size_t bstride = (size_t)a.vl.dw * 4;
            const int nb = (n + 3) >> 2;
            float4 cur[4], nxt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cur[u] = (u < nb) ? __ldcg(reinterpret_cast<const float4*>(vb + (size_t)u * bstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int b = 0; b < nb; b += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nxt[u] = (b + 4 + u < nb) ? __ldcg(reinterpret_cast<const float4*>(vb + (size_t)(b + 4 + u) * bstride)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t0 = (b + u) * 4;
                    const float vv[4] = {cur[u].x, cur[u].y, cur[u].z, cur[u].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = t0 + k;
                        if (t < n) {
                            const float w = att[t];
                            if (t == 0) o = __fmul_rn(vv[k], w);
                            else if (fabsf(w) > 1e-15f) o = __fmaf_rn(vv[k], w, o);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
            }
        } else {
            const float* vb = vc + (size_t)(d / a.vl.dw) * a.vl.max_seq * a.vl.dw + d % a.vl.dw;
            float cur[8], nxt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = (u < n) ? __ldcg(vb + (size_t)u * a.vl.dw) : 0.0f;
            for (int t0 = 0; t0 < n; t0 += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) nxt[u] = (t0 + 8 + u < n) ? __ldcg(vb + (size_t)(t0 + 8 + u) * a.vl.dw) : 0.0f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int t = t0 + u;
                    if (t < n) {
                        const float w = att[t];
                        if (t == 0) o = __fmul_rn(cur[u], w);
                        else if (fabsf(w) > 1e-15f) o = __fmaf_rn(cur[u], w, o);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
            }
        }
        a.out[(size_t)i * dim + (size_t)qh * HS + d] = o;
    }
}

// Sampler::sample_argmax (sampler.cpp:36-46) per row; then, if `advance`, the row's sequence state moves on (token fed back on
// the device).  logits [n_rows][ld]; row i belongs to rows[row0 + i].  reset_out: the row's token becomes out_tokens[0] (prefill).
__global__ void argmax_rows_kernel(const float* __restrict__ logits, int n, int ld, const RowMeta* __restrict__ rows, int row0,
                                   SeqState* st, int* out_tokens, int out_cap, int* argmax_out, int advance, int reset_out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    pdl_launch_dependents();
    pdl_wait();
    const float* lg = logits + (size_t)blockIdx.x * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = lg[i];
        if (v > bv) { bv = v; bi = i; }
    }
    if (bi == 0x7fffffff) { bi = threadIdx.x < n ? threadIdx.x : 0; bv = lg[bi]; }
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
        const RowMeta rm = rows[row0 + blockIdx.x];
        argmax_out[rm.slot] = bi;
        if (advance) {
            SeqState* s = st + rm.slot;
            int n_out = reset_out ? 0 : s->n_out;
            if (n_out < out_cap) out_tokens[(size_t)rm.slot * out_cap + n_out] = bi;
            s->n_out = n_out + 1;
            s->token = bi;
            s->pos = rm.pos + 1;
            s->bs = 1;
        }
    }
}

}  // namespace fl
