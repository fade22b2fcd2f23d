// engine.cu — C-ABI (include/fastllama_b200.h) over the sm_100a kernels in kernels.cuh.
//
// One fl_engine per GPU: packed weights, fp32 KV cache, pre-allocated workspace, one CUDA graph per KV slot
// that replays the whole decode step (embedding -> n_layers x {QKV, attention, Wo, W1/W3+SwiGLU, W2} ->
// classifier -> argmax + state advance).  No allocation and no host<->device traffic per token unless the
// caller asks for logits.  There is no CPU path: every entry point needs a CUDA device.
#include "../../include/fastllama_b200.h"
#include "kernels.cuh"
#include "megakernel.cuh"
#include "tc_gemm.cuh"
#include "batch_kernels.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <type_traits>
#include <vector>

using namespace fl;

namespace {

thread_local std::string g_last_error;

int set_err(fl_engine* e, int code, const char* fmt, ...);

#define CK(e, call)                                                                                         \
    do {                                                                                                    \
        cudaError_t _st = (call);                                                                           \
        if (_st != cudaSuccess)                                                                             \
            return set_err(e, FL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), __FILE__, __LINE__); \
    } while (0)

struct PackedMat {
    uint8_t* d = nullptr;
    int M = 0;          // logical rows (hidden for the W1/W3 stream)
    int K = 0;
    int n_tiles = 0;    // row tiles in the stream
    int nkb = 0;
    size_t bytes = 0;
};

// a matrix in the persistent kernel's "row per lane" streaming layout (megakernel.cuh)
struct RkMat {
    uint8_t* d = nullptr;
    size_t bytes = 0;
};
enum { RK_QKV = 0, RK_WO, RK_W13, RK_W2, RK_CLS, RK__COUNT };
constexpr int kMaxRows = 64;          // activation rows per pass of the tensor-core path (MMA N)

}  // namespace

struct fl_engine {
    fl_config c{};
    int device = 0;
    int n_sms = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    // weights
    float* emb = nullptr;
    float* att_norm = nullptr;
    float* ffn_norm = nullptr;
    float* out_norm = nullptr;
    std::vector<PackedMat> qkv, wo, w13, w2;   // per-phase kernels' layout (FL_FLAG_NO_MEGAKERNEL only)
    PackedMat cls;
    int tile_cap = 32;                            // rows of the tallest weight tile (MegaParams::tile_cap)
    int longctx_rows = 1 << 30;                   // a launch that can reach a longer context uses the long-context kernel variant
    bool want_mega = false;                       // decided in fl_create: which of the two layouts the uploads fill
    std::vector<RkMat> rk_qkv, rk_wo, rk_w13, rk_w2;
    RkMat rk_cls;
    unsigned long long* rk_off[RK__COUNT] = {};   // device: per-CTA stream offsets (n_sms + 1 entries), same for every layer
    size_t rk_bytes[RK__COUNT] = {};
    std::vector<uint8_t> have;          // [kind][layer]
    uint8_t* staging = nullptr;
    size_t staging_bytes = 0;
    // runtime
    float *x1 = nullptr, *qkv_buf = nullptr, *attn = nullptr, *hd = nullptr, *logits = nullptr;
    float *tap_qkv = nullptr, *tap_norm = nullptr;
    float *k_cache = nullptr, *v_cache = nullptr;
    float* rope = nullptr;
    SeqState* states = nullptr;
    int* out_tokens = nullptr;
    int out_cap = 0;
    int* in_tokens = nullptr;
    int in_cap = 0;
    int* argmax_dev = nullptr;
    int* h_tokens = nullptr;            // pinned
    float* h_logits = nullptr;          // pinned
    int* h_argmax = nullptr;            // pinned
    std::vector<cudaGraphExec_t> graphs;
    // megakernel state
    bool use_mega = false;
    int cph = 1;                        // CTAs per head in the persistent kernel; also the column blocking of the V cache
    MegaLayer* mega_layers = nullptr;
    uint2 *x1t = nullptr, *qkvt = nullptr, *attnt = nullptr, *hdt = nullptr, *hdqt = nullptr, *score_t = nullptr;   // tagged exchange buffers (sequence 0)
    uint8_t* xchg = nullptr;          // one block per sequence slot: [x1t | qkvt | attnt | hdt | score_t | am], xchg_stride bytes apart
    size_t xchg_stride = 0;
    uint4* am = nullptr;
    uint32_t epoch = 0;                 // last tag handed out (see megakernel.cuh)
    unsigned long long* prof = nullptr;
    unsigned long long* evlog = nullptr;
    MegaParams mega{};
    size_t mega_smem = 0;
    bool finalized = false;
    int64_t launches = 0;
    int kernels_per_step = 0;
    // tensor-core path (tc_gemm.cuh + batch_kernels.cuh): several activation rows per weight pass (prompt chunks, several sequences)
    bool tc = false;
    std::vector<RkMat> tc_qkv, tc_wo, tc_w13, tc_w2;   // stage-stream layout of tc_gemm.cuh
    RkMat tc_cls;
    unsigned long long* tc_off[RK__COUNT] = {};
    size_t tc_bytes[RK__COUNT] = {};
    float *bx1 = nullptr, *bqkv = nullptr, *batt = nullptr, *bhd = nullptr, *blogits = nullptr;   // [kMaxRows][...]
    uint8_t *img_x = nullptr, *img_h = nullptr, *img_c = nullptr;       // activation images: dim-wide, hidden-wide, classifier input
    RowMeta* rows = nullptr;            // device
    RowMeta* h_rows = nullptr;          // pinned
    int tc_smem = 0;
    bool tap_rows = false;              // the last forward went through the rows path: taps come from its last row
    int tap_row = 0;
    std::map<int, cudaGraphExec_t> batch_graphs;     // device-resident decode step of n sequences (fl_decode_batch_async)
    std::vector<int> h_pos;             // host mirror of every slot's next position (bounds checks)
    // NCCL (dlopen'ed; only when a communicator is bound)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    int* ag_send = nullptr;
    int* ag_recv = nullptr;
};

namespace {

int set_err(fl_engine* e, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (e) e->err = buf;
    return code;
}

inline int es_of(int qt) { return qt == FL_Q_INT8 ? 1 : 2; }
inline int unit_bytes(int qt, int gs) {
    const int lb = 64 * es_of(qt), gpl = 64 / gs;
    return 32 * lb + 32 * gpl * 4;
}

int alloc_packed(fl_engine* e, PackedMat& m, int rows_in_stream, int logical_rows, int K) {
    m.M = logical_rows;
    m.K = K;
    m.n_tiles = ceil_div(rows_in_stream, 4);
    m.nkb = ceil_div(K, kKBlockElems);
    m.bytes = (size_t)m.n_tiles * m.nkb * unit_bytes(e->c.quant_type, e->c.group_size);
    CK(e, cudaMalloc(&m.d, m.bytes));
    CK(e, cudaMemsetAsync(m.d, 0, m.bytes, e->stream));
    return FL_OK;
}

int rk_stage_bytes(int qt, int gs, int R) {
    const int gps = (kStageRowBytes / es_of(qt)) / gs;
    return R * kStageRowBytes + ((R * gps * 4 + 15) & ~15);
}

// per-CTA byte offsets of a matrix in the row-per-lane layout; must mirror pack_rk_kernel / the producer loop
int build_rk_table(fl_engine* e, int kind, int M, int K, int tt) {
    const int G = e->n_sms, qt = e->c.quant_type, gs = e->c.group_size;
    const int nkc = ceil_div(K * es_of(qt), kStageRowBytes);
    std::vector<unsigned long long> off(G + 1, 0);
    for (int c = 0; c < G; ++c) {
        const RkPart pt = rk_part(M, c, G, e->tile_cap);
        unsigned long long bytes = 0;
        for (int t = 0; t < pt.nt; ++t) { int lr0, R; rk_tile(pt, t, lr0, R); bytes += (unsigned long long)tt * nkc * rk_stage_bytes(qt, gs, R); }
        off[c + 1] = off[c] + bytes;
    }
    e->rk_bytes[kind] = off[G];
    CK(e, cudaMalloc(&e->rk_off[kind], sizeof(unsigned long long) * (G + 1)));
    CK(e, cudaMemcpyAsync(e->rk_off[kind], off.data(), sizeof(unsigned long long) * (G + 1), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return FL_OK;
}

int alloc_rk(fl_engine* e, RkMat& m, int kind) {
    m.bytes = e->rk_bytes[kind];
    CK(e, cudaMalloc(&m.d, m.bytes + 256));
    CK(e, cudaMemsetAsync(m.d, 0, m.bytes + 256, e->stream));
    return FL_OK;
}

bool mega_supported(const fl_config& c, int n_sms) {
    const int max_rows = (c.vocab_size > c.hidden_dim ? c.vocab_size : c.hidden_dim) / n_sms + 1;     // rows of a CTA: at most kGeomMaxTiles tiles
    if (max_rows > kGeomMaxTiles * 16 || c.dim + 2 * c.head_size * c.n_kv_heads > n_sms * kGeomMaxTiles * kTileRows) return false;
    return !(c.flags & FL_FLAG_NO_MEGAKERNEL) && c.n_heads <= n_sms && c.dim <= 6144;
}

// dispatch helpers over (quant type, group size)
template <typename F>
int dispatch_q(int qt, int gs, F&& f) {
    if (qt == FL_Q_INT8 && gs == 64) return f(std::integral_constant<int, Q_INT8>{}, std::integral_constant<int, 64>{});
    if (qt == FL_Q_INT8 && gs == 32) return f(std::integral_constant<int, Q_INT8>{}, std::integral_constant<int, 32>{});
    if (qt == FL_Q_INT16 && gs == 64) return f(std::integral_constant<int, Q_INT16>{}, std::integral_constant<int, 64>{});
    return FL_ERR_UNSUPPORTED;
}

size_t gemv_smem_bytes(int qt, int gs, int K, bool with_xf) {
    const int nkb = ceil_div(K, kKBlockElems);
    return (size_t)nkb * kKBlockElems * es_of(qt) + (size_t)nkb * 8 * (64 / gs) * 4 + (with_xf ? (size_t)K * 4 : 0);
}

template <int PRO, int EPI>
int launch_gemv(fl_engine* e, int qt, int gs, const GemvArgs& a, int grid, cudaStream_t st) {
    const size_t smem = gemv_smem_bytes(qt, gs, a.K, PRO == PRO_RMS_QUANT);
    return dispatch_q(qt, gs, [&](auto QT, auto GS) -> int {
        auto kern = gemv_kernel<decltype(QT)::value, decltype(GS)::value, PRO, EPI>;
        if (smem > 48 * 1024) {
            cudaError_t st2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (st2 != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaFuncSetAttribute(gemv): %s", cudaGetErrorString(st2));
        }
        kern<<<grid, kThreads, smem, st>>>(a);
        cudaError_t st3 = cudaGetLastError();
        if (st3 != cudaSuccess) return set_err(e, FL_ERR_CUDA, "gemv launch: %s", cudaGetErrorString(st3));
        return FL_OK;
    });
}

size_t attn_smem_bytes(int hs, int max_seq) {
    return (size_t)(3 * hs + 32 + 2 * kVChunk * hs + max_seq) * sizeof(float);
}

int launch_attn(fl_engine* e, int hs, const AttnArgs& a, int max_seq, cudaStream_t st) {
    const size_t smem = attn_smem_bytes(hs, max_seq);
    cudaError_t s1 = cudaSuccess;
    if (hs == 128) {
        s1 = cudaFuncSetAttribute(attn_decode_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (s1 == cudaSuccess) attn_decode_kernel<128><<<a.n_heads, kThreads, smem, st>>>(a);
    } else if (hs == 64) {
        s1 = cudaFuncSetAttribute(attn_decode_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (s1 == cudaSuccess) attn_decode_kernel<64><<<a.n_heads, kThreads, smem, st>>>(a);
    } else {
        return set_err(e, FL_ERR_UNSUPPORTED, "head_size %d not supported (64 or 128)", hs);
    }
    if (s1 != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaFuncSetAttribute(attn): %s", cudaGetErrorString(s1));
    s1 = cudaGetLastError();
    if (s1 != cudaSuccess) return set_err(e, FL_ERR_CUDA, "attn launch: %s", cudaGetErrorString(s1));
    return FL_OK;
}

// RoPE table exactly as rope_v2 walks it (tf_operators.cpp:367-395): theta_scale = powf(10000, -2/n),
// theta_0 = pos, theta_{k+1} = theta_k * theta_scale in float, glibc sincosf.
void build_rope_table(std::vector<float>& tab, int n_pos, int hs) {
    tab.resize((size_t)n_pos * hs);
    const float theta_scale = powf(10000.0f, -2.0f / (float)hs);
    for (int p = 0; p < n_pos; ++p) {
        float theta = (float)p;
        for (int i = 0; i < hs; i += 2) {
            float s, c;
            sincosf(theta, &s, &c);
            tab[(size_t)p * hs + i] = c;
            tab[(size_t)p * hs + i + 1] = s;
            theta *= theta_scale;
        }
    }
}

__global__ void dequant_rows_kernel(const uint8_t* q, const float* s, float* out, size_t n, int gs, int qt) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (qt == FL_Q_INT8) ? (int)reinterpret_cast<const int8_t*>(q)[i] : (int)reinterpret_cast<const int16_t*>(q)[i];
        out[i] = __fmul_rn(__int2float_rn(v), s[i / gs]);      // quant_operators.cpp:61
    }
}

int gemv_grid(const fl_engine* e, int n_tasks) {
    const int full = e->n_sms * 2;
    const int need = ceil_div(n_tasks, kWarps);
    return need < full ? (need < 1 ? 1 : need) : full;
}

// enqueue one decode step for `slot` on `st` (captured into a graph, or launched directly)
int enqueue_step(fl_engine* e, int slot, cudaStream_t st, int* n_kernels) {
    const fl_config& c = e->c;
    const int qt = c.quant_type, gs = c.group_size;
    const int kv_dim = c.head_size * c.n_kv_heads;
    SeqState* state = e->states + slot;
    int nk = 0;
    embed_kernel<<<4, 256, 0, st>>>(e->emb, &state->token, e->x1, c.dim);
    ++nk;
    const size_t cache_per_layer = (size_t)c.n_kv_heads * c.max_seq_len * c.head_size;
    const size_t cache_per_slot = cache_per_layer * c.n_layers;
    for (int l = 0; l < c.n_layers; ++l) {
        GemvArgs g{};
        // x2 = rmsnorm(x1); qkv = Wqkv * quantize(x2)            (transformer.cpp:132-135, :386-395)
        g.w = e->qkv[l].d; g.M = e->qkv[l].M; g.K = c.dim; g.n_tasks = e->qkv[l].n_tiles; g.nkb = e->qkv[l].nkb;
        g.in = e->x1; g.gain = e->att_norm + (size_t)l * c.dim; g.out = e->qkv_buf; g.tap = nullptr;
        int rc = launch_gemv<PRO_RMS_QUANT, EPI_STORE>(e, qt, gs, g, gemv_grid(e, g.n_tasks), st);
        if (rc) return rc;
        ++nk;
        // RoPE, KV append, QK^T, softmax, .V                      (:136, :397-455)
        AttnArgs a{};
        a.qkv = e->qkv_buf;
        a.k_cache = e->k_cache + slot * cache_per_slot + l * cache_per_layer;
        a.v_cache = e->v_cache + slot * cache_per_slot + l * cache_per_layer;
        a.rope = e->rope; a.pos_ptr = &state->pos; a.bs_ptr = &state->bs; a.out = e->attn;
        a.tap_qkv = (l == c.n_layers - 1) ? e->tap_qkv : nullptr;
        a.n_heads = c.n_heads; a.n_kv_heads = c.n_kv_heads; a.max_seq = c.max_seq_len;
        a.attn_scale = 1.0f / sqrtf((float)c.head_size);
        a.v_dw = c.head_size;          // the per-phase kernels keep natural V rows (the persistent kernel has its own blocked layout)
        rc = launch_attn(e, c.head_size, a, c.max_seq_len, st);
        if (rc) return rc;
        ++nk;
        // x1 += Wo * quantize(attn)                               (:138-139, :457-466)
        g = GemvArgs{};
        g.w = e->wo[l].d; g.M = c.dim; g.K = c.dim; g.n_tasks = e->wo[l].n_tiles; g.nkb = e->wo[l].nkb;
        g.in = e->attn; g.out = e->x1;
        rc = launch_gemv<PRO_QUANT, EPI_RESADD>(e, qt, gs, g, gemv_grid(e, g.n_tasks), st);
        if (rc) return rc;
        ++nk;
        // hd = swiglu(W1 q, W3 q), q = quantize(rmsnorm(x1))      (:144-147, :468-483)
        g = GemvArgs{};
        g.w = e->w13[l].d; g.M = c.hidden_dim; g.K = c.dim; g.n_tasks = e->w13[l].n_tiles / 2; g.nkb = e->w13[l].nkb;
        g.in = e->x1; g.gain = e->ffn_norm + (size_t)l * c.dim; g.out = e->hd;
        rc = launch_gemv<PRO_RMS_QUANT, EPI_SWIGLU>(e, qt, gs, g, gemv_grid(e, g.n_tasks), st);
        if (rc) return rc;
        ++nk;
        // x1 += W2 * quantize(hd)                                 (:149-150, :485-494)
        g = GemvArgs{};
        g.w = e->w2[l].d; g.M = c.dim; g.K = c.hidden_dim; g.n_tasks = e->w2[l].n_tiles; g.nkb = e->w2[l].nkb;
        g.in = e->hd; g.out = e->x1;
        rc = launch_gemv<PRO_QUANT, EPI_RESADD>(e, qt, gs, g, gemv_grid(e, g.n_tasks), st);
        if (rc) return rc;
        ++nk;
    }
    // logits = Wcls * quantize(rmsnorm(x1))                      (:154-160, :496-505)
    GemvArgs g{};
    g.w = e->cls.d; g.M = c.vocab_size; g.K = c.dim; g.n_tasks = e->cls.n_tiles; g.nkb = e->cls.nkb;
    g.in = e->x1; g.gain = e->out_norm; g.out = e->logits; g.tap = e->tap_norm;
    int rc = launch_gemv<PRO_RMS_QUANT, EPI_STORE>(e, qt, gs, g, gemv_grid(e, g.n_tasks), st);
    if (rc) return rc;
    ++nk;
    argmax_kernel<<<1, 1024, 0, st>>>(e->logits, c.vocab_size, state, e->out_tokens + (size_t)slot * e->out_cap, e->out_cap,
                                      e->argmax_dev + slot, 1);
    ++nk;
    (void)kv_dim;
    cudaError_t s = cudaGetLastError();
    if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "step launch: %s", cudaGetErrorString(s));
    if (n_kernels) *n_kernels = nk;
    return FL_OK;
}

// multi = true: the variant that walks several sequences per phase (fl_forward_batch)
template <typename F>
int dispatch_mega(int qt, int gs, int hs, bool multi, F&& f, bool relaxed = false, bool longctx = false) {
    // FL_FLAG_RELAXED: only the benchmark shape is instantiated (INT8, group 64, head 128, one sequence per launch)
    if (relaxed && !multi && qt == FL_Q_INT8 && gs == 64 && hs == 128) return f(decode_megakernel<Q_INT8, 64, 128, false, true>);
    // longctx: the variant whose attention part streams K and V through the weight ring (single sequence per launch only)
#define FL_MEGA_CASE(QT_, QTC, GS_, HS_) \
    if (qt == QT_ && gs == GS_ && hs == HS_) return multi ? f(decode_megakernel<QTC, GS_, HS_, true>) : longctx ? f(decode_megakernel<QTC, GS_, HS_, false, false, true>) : f(decode_megakernel<QTC, GS_, HS_, false>);
    FL_MEGA_CASE(FL_Q_INT8, Q_INT8, 64, 128)
    FL_MEGA_CASE(FL_Q_INT8, Q_INT8, 64, 64)
    FL_MEGA_CASE(FL_Q_INT8, Q_INT8, 32, 128)
    FL_MEGA_CASE(FL_Q_INT8, Q_INT8, 32, 64)
    FL_MEGA_CASE(FL_Q_INT16, Q_INT16, 64, 128)
    FL_MEGA_CASE(FL_Q_INT16, Q_INT16, 64, 64)
#undef FL_MEGA_CASE
    return FL_ERR_UNSUPPORTED;
}

// one-time set-up of the persistent decode kernel: layer table, counters, shared-memory carve-up
int setup_mega(fl_engine* e) {
    const fl_config& c = e->c;
    const int L = c.n_layers, qt = c.quant_type, gs = c.group_size, es = es_of(qt);
    std::vector<MegaLayer> tab(L);
    for (int l = 0; l < L; ++l) {
        tab[l].qkv = e->rk_qkv[l].d; tab[l].wo = e->rk_wo[l].d; tab[l].w13 = e->rk_w13[l].d; tab[l].w2 = e->rk_w2[l].d;
        tab[l].att_norm = e->att_norm + (size_t)l * c.dim; tab[l].ffn_norm = e->ffn_norm + (size_t)l * c.dim;
    }
    CK(e, cudaMalloc(&e->mega_layers, sizeof(MegaLayer) * L));
    CK(e, cudaMemcpyAsync(e->mega_layers, tab.data(), sizeof(MegaLayer) * L, cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    const int qkv_rows = c.dim + 2 * c.head_size * c.n_kv_heads;
    const int score_stride = (c.max_seq_len + 3) & ~1;
    {
        // exchange buffers: one block per sequence slot, the same layout in each (MegaParams::xchg_stride)
        size_t off = 0;
        auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        const size_t o_x1 = take((c.dim + 2) * sizeof(uint2)), o_qkv = take((qkv_rows + 2) * sizeof(uint2));
        const size_t o_attn = take((c.dim + 2) * sizeof(uint2)), o_hd = take((c.hidden_dim + 2) * sizeof(uint2));
        const size_t o_sc = take(((size_t)c.n_heads * score_stride + 2) * sizeof(uint2)), o_am = take(sizeof(uint4) * e->n_sms);
        // quantised hd: hidden * es / 4 payload words, then one scale per group, every word tagged (build_hd)
        const size_t o_hdq = take(((size_t)c.hidden_dim * es / 4 + c.hidden_dim / gs + 2) * sizeof(uint2));
        e->xchg_stride = off;
        CK(e, cudaMalloc(&e->xchg, off * c.max_seqs));
        CK(e, cudaMemsetAsync(e->xchg, 0, off * c.max_seqs, e->stream));      // tag 0 is never used
        e->x1t = reinterpret_cast<uint2*>(e->xchg + o_x1); e->qkvt = reinterpret_cast<uint2*>(e->xchg + o_qkv);
        e->attnt = reinterpret_cast<uint2*>(e->xchg + o_attn); e->hdt = reinterpret_cast<uint2*>(e->xchg + o_hd);
        e->score_t = reinterpret_cast<uint2*>(e->xchg + o_sc); e->am = reinterpret_cast<uint4*>(e->xchg + o_am);
        e->hdqt = reinterpret_cast<uint2*>(e->xchg + o_hdq);
    }
    CK(e, cudaMalloc(&e->prof, sizeof(unsigned long long) * 32 * e->n_sms));
    CK(e, cudaMemsetAsync(e->prof, 0, sizeof(unsigned long long) * 32 * e->n_sms, e->stream));
    CK(e, cudaMalloc(&e->evlog, sizeof(unsigned long long) * 2 * 4096));
    CK(e, cudaMemsetAsync(e->evlog, 0, sizeof(unsigned long long) * 2 * 4096, e->stream));
    MegaParams& p = e->mega;
    p.layers = e->mega_layers; p.cls = e->rk_cls.d; p.out_norm = e->out_norm; p.emb = e->emb;
    p.att_norm = e->att_norm; p.ffn_norm = e->ffn_norm;
    p.off_qkv = e->rk_off[RK_QKV]; p.off_wo = e->rk_off[RK_WO]; p.off_w13 = e->rk_off[RK_W13]; p.off_w2 = e->rk_off[RK_W2]; p.off_cls = e->rk_off[RK_CLS];
    p.x1t = e->x1t; p.qkvt = e->qkvt; p.attnt = e->attnt; p.hdt = e->hdt; p.hdqt = e->hdqt; p.score_t = e->score_t; p.am = e->am; p.logits = e->logits;
    p.rope = e->rope; p.out_cap = e->out_cap; p.score_stride = score_stride;
    p.tap_norm = e->tap_norm; p.prof = (c.flags & FL_FLAG_PROFILE) ? e->prof : nullptr;
    p.evlog = (c.flags & FL_FLAG_PROFILE) ? e->evlog : nullptr;
    p.dim = c.dim; p.hidden = c.hidden_dim; p.n_layers = L; p.n_heads = c.n_heads; p.n_kv_heads = c.n_kv_heads;
    p.vocab = c.vocab_size; p.max_seq = c.max_seq_len; p.qkv_rows = qkv_rows;
    p.attn_scale = 1.0f / sqrtf((float)c.head_size);
    const int cph = e->cph;
    p.cph = cph;
    const int dw = c.head_size / cph;
    // V chunks of kVChunkRows rows x dims-per-part fp32 (4 KB at 32 dims): small chunks keep a whole short context in flight
    // at once (the staging area holds ~10 of them); the PV loop hides the per-chunk cost (profiles/r02)
    const int v_chunk_bytes = kVChunkRows * dw * 4;
    // shared memory carve-up
    const int kmax = c.dim > c.hidden_dim ? c.dim : c.hidden_dim;
    const int nkc_max = ceil_div(kmax * es, kStageRowBytes);
    const int gps = (kStageRowBytes / es) / gs;
    auto al = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    size_t off = 0;
    p.off_misc = (int)off; off += 2048;
    p.off_att = (int)off; off += al((size_t)c.max_seq_len * 4 + 192, 128);      // + one chunk of zero weights past the last position
    p.off_xs = (int)off; off += al((size_t)nkc_max * gps * 4, 128);
    p.off_vbars = (int)off; off += 256;      // 32 mbarriers: the V ring of the attention part
    p.off_psrc = (int)off; off += al((size_t)(4 * L + 1) * 8, 128);      // this CTA's weight-stream start of every phase
    p.off_geom = (int)off; off += 5 * kGeomStride * 4;                   // this CTA's tile / superblock geometry of the five phase kinds
    // [activation image | pair buffers]: contiguous, because attention (which uses neither) turns the whole range into its ring
    // of V chunks.  The transposed fp32 vector of the rmsnorm rebuild lives on top of the pair buffers: it is written only
    // after the rebuild's input is complete, i.e. after this CTA's own chain warp has published (and stopped reading pairs).
    const size_t xq_bytes = al((size_t)nkc_max * kStageRowBytes, 128), xt_bytes = al((size_t)c.dim * 4, 128);
    const size_t pair_bytes = 2 * (size_t)kPairGroups * 32 * 8;
    p.off_vstage = (int)off;
    p.off_chain = (int)off;
    p.off_xq = (int)off; off += xq_bytes;
    p.off_pairs = (int)off; p.off_xt = (int)off; off += (pair_bytes > xt_bytes ? pair_bytes : xt_bytes);
    size_t vbytes = off - p.off_vstage;
    if (vbytes < 2 * (size_t)v_chunk_bytes) { off += 2 * v_chunk_bytes - vbytes; vbytes = 2 * v_chunk_bytes; }
    p.n_vchunks = (int)(vbytes / v_chunk_bytes) > 32 ? 32 : (int)(vbytes / v_chunk_bytes);
    int max_smem = 0;
    CK(e, cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device));
    const size_t slot_bytes = (size_t)rk_stage_bytes(qt, gs, e->tile_cap);
    p.tile_cap = e->tile_cap; p.slot_bytes = (int)slot_bytes;
    int n_slots = (int)(((size_t)max_smem - off - 1024) / (slot_bytes + 16));
    if (n_slots > 32) n_slots = 32;
    if (n_slots < 4) return set_err(e, FL_ERR_UNSUPPORTED, "megakernel: not enough shared memory for the weight ring (%d slots)", n_slots);
    p.n_slots = n_slots;
    p.window = 99;
    p.debug_skip = 0;
#ifdef FL_PROFILE                         // the profiling build only: FL_DEBUG_SKIP makes results garbage by design
    if (const char* w = getenv("FL_DEBUG_SKIP")) p.debug_skip = atoi(w);      // timing experiments only
    if (const char* w = getenv("FL_WINDOW")) { const int v = atoi(w); if (v >= 1) p.window = v; }      // tuning knob (profiles/)
#endif
    p.off_bars = (int)off; off += al((size_t)n_slots * 16, 128);
    p.off_ring = (int)off; off += (size_t)n_slots * slot_bytes;
    e->mega_smem = off;
    p.n_seqs = 1;
    p.xchg_stride = e->xchg_stride;
    p.cache_stride = (unsigned long long)c.n_layers * c.n_kv_heads * c.max_seq_len * c.head_size;
    auto prepare = [&](auto kern) -> int {
        cudaError_t s = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->mega_smem);
        if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaFuncSetAttribute(megakernel, %zu): %s", e->mega_smem, cudaGetErrorString(s));
        int nb = 0;
        s = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kMegaThreads, e->mega_smem);
        if (s != cudaSuccess || nb < 1) return set_err(e, FL_ERR_CUDA, "megakernel does not fit an SM (smem %zu)", e->mega_smem);
        return FL_OK;
    };
    int rc = dispatch_mega(qt, gs, c.head_size, false, prepare);
    e->longctx_rows = 2 * p.n_vchunks * kVChunkRows;       // contexts beyond this stage K and V in the weight ring (attention_part, `big`)
    if (rc == FL_OK && c.max_seq_len > e->longctx_rows) rc = dispatch_mega(qt, gs, c.head_size, false, prepare, false, true);
    if (rc == FL_OK && (c.flags & FL_FLAG_RELAXED)) rc = dispatch_mega(qt, gs, c.head_size, false, prepare, true);
    if (rc == FL_OK && c.max_seqs > 1) rc = dispatch_mega(qt, gs, c.head_size, true, prepare);
    if (rc == FL_ERR_UNSUPPORTED) return set_err(e, rc, "megakernel: unsupported quant/group/head combination");
    return rc;
}

// slots [slot, slot + n_seqs): n_seqs == 1 is the plain kernel; more walk every phase once per sequence in one launch
int launch_mega(fl_engine* e, int slot, int n_steps, int n_seqs = 1) {
    const fl_config& c = e->c;
    MegaParams p = e->mega;
    const size_t cache_per_slot = (size_t)c.n_layers * c.n_kv_heads * c.max_seq_len * c.head_size;
    p.k_cache = e->k_cache + slot * cache_per_slot;
    p.v_cache = e->v_cache + slot * cache_per_slot;
    p.st = e->states + slot;
    p.out_tokens = e->out_tokens + (size_t)slot * e->out_cap;
    p.argmax_out = e->argmax_dev + slot;
    p.n_steps = n_steps;
    p.n_seqs = n_seqs;
    if (slot > 0) {     // the exchange buffers of slot 0 serve a single sequence in any slot; a batch starts at its own block
        const size_t o = n_seqs > 1 ? (size_t)slot * e->xchg_stride : 0;
        auto sh = [&](auto*& ptr) { ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(reinterpret_cast<uint8_t*>(ptr) + o); };
        sh(p.x1t); sh(p.qkvt); sh(p.attnt); sh(p.hdt); sh(p.hdqt); sh(p.score_t); sh(p.am);
    }
    const uint64_t tags = (uint64_t)n_steps * (uint64_t)(c.n_layers + 1) * kTagsPerLayer;
    if ((uint64_t)e->epoch + tags >= 0xfff00000ull) {
        // the 32-bit tag space is about to wrap (~16 M tokens of a 32-layer model): start over with cleared exchange buffers
        // (tag 0 is never used); ordered on the engine stream like every launch
        CK(e, cudaMemsetAsync(e->xchg, 0, e->xchg_stride * c.max_seqs, e->stream));
        e->epoch = 0;
    }
    p.epoch = e->epoch;
    e->epoch += (uint32_t)tags;
    return dispatch_mega(c.quant_type, c.group_size, c.head_size, n_seqs > 1, [&](auto kern) -> int {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(e->n_sms); cfg.blockDim = dim3(kMegaThreads); cfg.dynamicSmemBytes = e->mega_smem; cfg.stream = e->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;        // every CTA resident: the grid barriers cannot deadlock
        attr[0].val.cooperative = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t s = cudaLaunchKernelEx(&cfg, kern, p);
        if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "megakernel launch: %s", cudaGetErrorString(s));
        e->launches += 1;
        return FL_OK;
    }, (c.flags & FL_FLAG_RELAXED) != 0, 
#ifdef FL_NO_LONGCTX       // A/B build: never the long-context variant
       false);
#else
       n_seqs == 1 && !(c.flags & FL_FLAG_RELAXED) && e->h_pos[slot] > e->longctx_rows);
#endif
}

int run_step(fl_engine* e, int slot) {
    if (e->use_mega) return launch_mega(e, slot, 1);
    if (e->c.flags & FL_FLAG_NO_GRAPH) {
        int nk = 0;
        int rc = enqueue_step(e, slot, e->stream, &nk);
        e->launches += nk;
        return rc;
    }
    if (!e->graphs[slot]) {
        cudaGraph_t graph = nullptr;
        CK(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        int nk = 0;
        int rc = enqueue_step(e, slot, e->stream, &nk);
        cudaError_t s = cudaStreamEndCapture(e->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(s));
        e->kernels_per_step = nk;
        s = cudaGraphInstantiate(&e->graphs[slot], graph, 0);
        cudaGraphDestroy(graph);
        if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(s));
    }
    CK(e, cudaGraphLaunch(e->graphs[slot], e->stream));
    e->launches += e->kernels_per_step;
    return FL_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// tensor-core path: several activation rows per weight pass
// ------------------------------------------------------------------------------------------------------------------
int build_tc_table(fl_engine* e, int kind, int M, int K, int tt) {
    const int G = e->n_sms, gs = e->c.group_size;
    const int nkc = ceil_div(K, kTcKC);
    std::vector<unsigned long long> off(G + 1, 0);
    for (int c = 0; c < G; ++c) {
        const TcPart pt = tc_part(M, c, G);
        unsigned long long bytes = 0;
        for (int t = 0; t < pt.nt; ++t) { int lr0, R; tc_tile(pt, t, lr0, R); bytes += (unsigned long long)nkc * tc_stage_a_bytes(R, tt, gs); }
        off[c + 1] = off[c] + bytes;
    }
    e->tc_bytes[kind] = off[G];
    CK(e, cudaMalloc(&e->tc_off[kind], sizeof(unsigned long long) * (G + 1)));
    CK(e, cudaMemcpyAsync(e->tc_off[kind], off.data(), sizeof(unsigned long long) * (G + 1), cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return FL_OK;
}

int alloc_tc(fl_engine* e, RkMat& m, int kind) {
    m.bytes = e->tc_bytes[kind];
    CK(e, cudaMalloc(&m.d, m.bytes + 4096));       // tail: an M = 128 MMA of a short last tile never reads past the allocation's stage
    CK(e, cudaMemsetAsync(m.d, 0, m.bytes + 4096, e->stream));
    return FL_OK;
}

// every kernel of the rows path is launched with programmatic stream serialisation (unless FL_FLAG_NO_PDL): it may start
// while its predecessor drains; each kernel executes griddepcontrol.wait before it touches dependent data
template <typename... KArgs, typename... Args>
int launch_pdl(fl_engine* e, bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t s = cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
    if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(s));
    if (e) e->launches += 1;
    return FL_OK;
}

template <int GS, int N>
int launch_qgemm_n(fl_engine* e, bool pdl, int epi, const TcGemmArgs& a, int grid, cudaStream_t st) {
    auto go = [&](auto kern) -> int {
        cudaError_t s = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
        if (s != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaFuncSetAttribute(qgemm, %d): %s", a.smem_bytes, cudaGetErrorString(s));
        return launch_pdl(e, pdl, kern, dim3(grid), dim3(kTcThreads), (size_t)a.smem_bytes, st, a);
    };
    if (epi == TC_EPI_STORE) return go(qgemm_kernel<GS, N, false, TC_EPI_STORE>);
    if (epi == TC_EPI_RESADD) return go(qgemm_kernel<GS, N, false, TC_EPI_RESADD>);
    return go(qgemm_kernel<GS, N, true, TC_EPI_SWIGLU>);
}

// N = activation rows padded to 8 / 16 / 32 / 64
int launch_qgemm(fl_engine* e, bool pdl, int gs, int N, int epi, const TcGemmArgs& a, int grid, cudaStream_t st) {
#define FL_QG(GS_, N_) if (gs == GS_ && N == N_) return launch_qgemm_n<GS_, N_>(e, pdl, epi, a, grid, st);
    FL_QG(64, 8) FL_QG(64, 16) FL_QG(64, 32) FL_QG(64, 64)
    FL_QG(32, 8) FL_QG(32, 16) FL_QG(32, 32) FL_QG(32, 64)
#undef FL_QG
    return set_err(e, FL_ERR_UNSUPPORTED, "qgemm: group %d / N %d not supported", gs, N);
}

// One forward over T activation rows described by e->rows (device): every layer is one weight pass for all rows.
//   n_cls == 0: no classifier;  n_cls == 1: only the last row (prompt chunk: transformer.cpp:140-142);  n_cls == T: every row.
//   advance: the classified rows' sequence states move on (token fed back on the device); reset_out: their n_out restarts.
int enqueue_rows(fl_engine* e, int T, int n_cls, int advance, int reset_out, cudaStream_t st) {
    const fl_config& c = e->c;
    const int gs = c.group_size, dim = c.dim, hid = c.hidden_dim, hs = c.head_size;
    const int kv_dim = hs * c.n_kv_heads, qkv_rows = dim + 2 * kv_dim;
    const int N = tc_pad_n(T);
    const bool pdl = !(c.flags & FL_FLAG_NO_PDL);
    const int G = e->n_sms;
    int rc = launch_pdl(e, pdl, embed_rows_kernel, dim3(T), dim3(256), 0, st, (const float*)e->emb, (const RowMeta*)e->rows, e->bx1, dim);
    if (rc) return rc;
    const size_t cache_per_layer = (size_t)c.n_kv_heads * c.max_seq_len * hs;
    const size_t cache_per_slot = cache_per_layer * c.n_layers;
    auto rms_quant = [&](const float* x, size_t stride, const float* gain, uint8_t* img, int rows, int Nimg, float* tap) -> int {
        if (gs == 64) return launch_pdl(e, pdl, rms_quant_rows_kernel<64>, dim3(rows), dim3(kThreads), (size_t)dim * 4, st, x, stride, gain, dim, img, Nimg, tap);
        return launch_pdl(e, pdl, rms_quant_rows_kernel<32>, dim3(rows), dim3(kThreads), (size_t)dim * 4, st, x, stride, gain, dim, img, Nimg, tap);
    };
    auto quant = [&](const float* x, int K, uint8_t* img) -> int {
        const dim3 grid(T, ceil_div(K / gs, kThreads / 8));
        if (gs == 64) return launch_pdl(e, pdl, quant_rows_kernel<64>, grid, dim3(kThreads), 0, st, x, K, img, N);
        return launch_pdl(e, pdl, quant_rows_kernel<32>, grid, dim3(kThreads), 0, st, x, K, img, N);
    };
    auto gemm = [&](const RkMat& w, int kind, const uint8_t* img, float* out, int M, int K, int rowsT, int Nimg, int ldo, int epi) -> int {
        TcGemmArgs a{};
        a.w = w.d; a.cta_off = e->tc_off[kind]; a.xq = img; a.out = out; a.M = M; a.K = K; a.T = rowsT; a.ldo = ldo;
        a.smem_bytes = e->tc_smem; a.variant = 0;
        return launch_qgemm(e, pdl, gs, Nimg, epi, a, G, st);
    };
    for (int l = 0; l < c.n_layers; ++l) {
        const bool last = l == c.n_layers - 1;
        // x2 = rmsnorm(x1); qkv = Wqkv * quantize(x2)                         (transformer.cpp:132-135)
        rc = rms_quant(e->bx1, (size_t)dim, e->att_norm + (size_t)l * dim, e->img_x, T, N, nullptr);
        if (!rc) rc = gemm(e->tc_qkv[l], RK_QKV, e->img_x, e->bqkv, qkv_rows, dim, T, N, qkv_rows, TC_EPI_STORE);
        if (rc) return rc;
        // RoPE, KV append, QK^T, softmax, PV                                  (:136, :397-455)
        AttnRowsArgs a{};
        a.qkv = e->bqkv;
        a.k_cache = e->k_cache + (size_t)l * cache_per_layer;
        a.v_cache = e->v_cache + (size_t)l * cache_per_layer;
        a.slot_stride = cache_per_slot;
        a.rope = e->rope; a.rows = e->rows; a.out = e->batt;
        a.n_heads = c.n_heads; a.n_kv_heads = c.n_kv_heads;
        a.attn_scale = 1.0f / sqrtf((float)hs);
        a.vl.max_seq = c.max_seq_len;
        a.vl.dw = e->want_mega ? hs / e->cph : hs;
        a.vl.blocked4 = e->want_mega ? 1 : 0;
        a.tap_qkv = last ? e->tap_qkv : nullptr;
        a.tap_row = T - 1;
        const size_t asmem = (size_t)(hs + 64 + c.max_seq_len + 8) * 4;
        if (hs == 128) {
            rc = launch_pdl(e, pdl, kv_append_rows_kernel<128>, dim3(c.n_kv_heads, T), dim3(128), 0, st, a);
            if (!rc) rc = launch_pdl(e, pdl, attn_rows_kernel<128>, dim3(c.n_heads, T), dim3(kThreads), asmem, st, a);
        } else {
            rc = launch_pdl(e, pdl, kv_append_rows_kernel<64>, dim3(c.n_kv_heads, T), dim3(64), 0, st, a);
            if (!rc) rc = launch_pdl(e, pdl, attn_rows_kernel<64>, dim3(c.n_heads, T), dim3(kThreads), asmem, st, a);
        }
        if (rc) return rc;
        // x1 += Wo * quantize(attn)                                           (:138-139)
        rc = quant(e->batt, dim, e->img_x);
        if (!rc) rc = gemm(e->tc_wo[l], RK_WO, e->img_x, e->bx1, dim, dim, T, N, dim, TC_EPI_RESADD);
        // hd = swiglu(W1 q, W3 q), q = quantize(rmsnorm(x1))                  (:144-147)
        if (!rc) rc = rms_quant(e->bx1, (size_t)dim, e->ffn_norm + (size_t)l * dim, e->img_x, T, N, nullptr);
        if (!rc) rc = gemm(e->tc_w13[l], RK_W13, e->img_x, e->bhd, hid, dim, T, N, hid, TC_EPI_SWIGLU);
        // x1 += W2 * quantize(hd)                                             (:149-150)
        if (!rc) rc = quant(e->bhd, hid, e->img_h);
        if (!rc) rc = gemm(e->tc_w2[l], RK_W2, e->img_h, e->bx1, dim, hid, T, N, dim, TC_EPI_RESADD);
        if (rc) return rc;
    }
    if (n_cls > 0) {
        // logits = Wcls * quantize(rmsnorm(x1))                               (:154-160)
        const int row0 = T - n_cls, Nc = tc_pad_n(n_cls);
        rc = rms_quant(e->bx1 + (size_t)row0 * dim, (size_t)dim, e->out_norm, e->img_c, n_cls, Nc, e->tap_norm);
        if (!rc) rc = gemm(e->tc_cls, RK_CLS, e->img_c, e->blogits, c.vocab_size, dim, n_cls, Nc, c.vocab_size, TC_EPI_STORE);
        if (!rc) rc = launch_pdl(e, pdl, argmax_rows_kernel, dim3(n_cls), dim3(1024), 0, st, (const float*)e->blogits, c.vocab_size, c.vocab_size,
                                 (const RowMeta*)e->rows, row0, e->states, e->out_tokens, e->out_cap, e->argmax_dev, advance, reset_out);
        if (rc) return rc;
    }
    e->tap_rows = true;
    e->tap_row = T - 1;
    return FL_OK;
}

}  // namespace

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 16) == cudaSuccess ? 0 : -1; }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};
int need_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return set_err(nullptr, FL_ERR_CUDA, "no CUDA device (this library has no CPU path)");
    return FL_OK;
}
#define CKO(call) CK(nullptr, call)
}  // namespace

// =================================================================================================
extern "C" {

const char* fl_last_error(const fl_engine* e) { return e ? e->err.c_str() : g_last_error.c_str(); }

int fl_create(const fl_config* cfg, int device, fl_engine** out) {
    if (!cfg || !out) return set_err(nullptr, FL_ERR_INVALID, "fl_create: null argument");
    *out = nullptr;
    const fl_config& c = *cfg;
    if (c.dim <= 0 || c.hidden_dim <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.n_kv_heads <= 0 || c.vocab_size <= 0 ||
        c.max_seq_len <= 0 || c.max_seqs <= 0)
        return set_err(nullptr, FL_ERR_INVALID, "fl_create: non-positive dimension");
    if (c.head_size * c.n_heads != c.dim || c.n_heads % c.n_kv_heads != 0)
        return set_err(nullptr, FL_ERR_INVALID, "fl_create: head_size*n_heads != dim or n_heads %% n_kv_heads != 0");
    if (c.head_size != 64 && c.head_size != 128)
        return set_err(nullptr, FL_ERR_UNSUPPORTED, "fl_create: head_size must be 64 or 128");
    if (!((c.quant_type == FL_Q_INT8 && (c.group_size == 64 || c.group_size == 32)) || (c.quant_type == FL_Q_INT16 && c.group_size == 64)))
        return set_err(nullptr, FL_ERR_UNSUPPORTED, "fl_create: quant_type/group_size combination not supported");
    if (c.dim % 64 || c.hidden_dim % 64)
        return set_err(nullptr, FL_ERR_INVALID, "fl_create: dim and hidden_dim must be multiples of 64");
    if (c.max_seq_len % 4)
        return set_err(nullptr, FL_ERR_INVALID, "fl_create: max_seq_len must be a multiple of 4 (the V cache keeps 4 positions per block)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return set_err(nullptr, FL_ERR_CUDA, "fl_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return set_err(nullptr, FL_ERR_INVALID, "fl_create: device %d out of range", device);

    fl_engine* e = new fl_engine();
    e->c = c;
    e->device = device;
    auto fail = [&](int rc) { std::string m = e->err; fl_destroy(e); g_last_error = m; return rc; };
#define CKF(call) do { int _rc = [&]() -> int { CK(e, call); return FL_OK; }(); if (_rc) return fail(_rc); } while (0)
    CKF(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKF(cudaGetDeviceProperties(&prop, device));
    e->n_sms = prop.multiProcessorCount;
    CKF(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->cph = 4;     // measured in round 2: 2 CTAs per head 412.6 against 431.0 tokens/s (7B, ctx 33-543)
    while (e->cph > 1 && (c.n_heads * e->cph > e->n_sms || c.head_size / e->cph < 16)) e->cph /= 2;

    const int L = c.n_layers, kv_dim = c.head_size * c.n_kv_heads;
    e->have.assign((size_t)FL_T__COUNT * L, 0);
    e->qkv.resize(L); e->wo.resize(L); e->w13.resize(L); e->w2.resize(L);
    e->graphs.assign(c.max_seqs, nullptr);
    CKF(cudaMalloc(&e->emb, (size_t)c.vocab_size * c.dim * 4));
    CKF(cudaMalloc(&e->att_norm, (size_t)L * c.dim * 4));
    CKF(cudaMalloc(&e->ffn_norm, (size_t)L * c.dim * 4));
    CKF(cudaMalloc(&e->out_norm, (size_t)c.dim * 4));
    e->want_mega = mega_supported(c, e->n_sms);
    if (e->want_mega) {
        e->rk_qkv.resize(L); e->rk_wo.resize(L); e->rk_w13.resize(L); e->rk_w2.resize(L);
        {   // tile cap = the tallest tile of the four layer matrices under the 32-lane limit (rk_part)
            int cap = 1;
            const int Ms[4] = {c.dim + 2 * kv_dim, c.dim, c.hidden_dim, c.dim};
            for (int M : Ms)
                for (int cta = 0; cta < e->n_sms; ++cta) {
                    const RkPart pt = rk_part(M, cta, e->n_sms, kTileRows);
                    for (int t = 0; t < pt.nt; ++t) { int lr0, R; rk_tile(pt, t, lr0, R); cap = R > cap ? R : cap; }
                }
            e->tile_cap = cap;
        }
        int rc = build_rk_table(e, RK_QKV, c.dim + 2 * kv_dim, c.dim, 1);
        if (!rc) rc = build_rk_table(e, RK_WO, c.dim, c.dim, 1);
        if (!rc) rc = build_rk_table(e, RK_W13, c.hidden_dim, c.dim, 2);
        if (!rc) rc = build_rk_table(e, RK_W2, c.dim, c.hidden_dim, 1);
        if (!rc) rc = build_rk_table(e, RK_CLS, c.vocab_size, c.dim, 1);
        for (int l = 0; l < L && !rc; ++l) {
            rc = alloc_rk(e, e->rk_qkv[l], RK_QKV);
            if (!rc) rc = alloc_rk(e, e->rk_wo[l], RK_WO);
            if (!rc) rc = alloc_rk(e, e->rk_w13[l], RK_W13);
            if (!rc) rc = alloc_rk(e, e->rk_w2[l], RK_W2);
        }
        if (!rc) rc = alloc_rk(e, e->rk_cls, RK_CLS);
        if (rc) return fail(rc);
    } else {
        for (int l = 0; l < L; ++l) {
            int rc = alloc_packed(e, e->qkv[l], c.dim + 2 * kv_dim, c.dim + 2 * kv_dim, c.dim);
            if (!rc) rc = alloc_packed(e, e->wo[l], c.dim, c.dim, c.dim);
            if (!rc) rc = alloc_packed(e, e->w13[l], 2 * ceil_div(c.hidden_dim, 4) * 4, c.hidden_dim, c.dim);
            if (!rc) rc = alloc_packed(e, e->w2[l], c.dim, c.dim, c.hidden_dim);
            if (rc) return fail(rc);
        }
        { int rc = alloc_packed(e, e->cls, c.vocab_size, c.vocab_size, c.dim); if (rc) return fail(rc); }
    }
    e->tc = c.quant_type == FL_Q_INT8 && !(c.flags & FL_FLAG_NO_TC) && prop.major >= 10;
    if (e->tc) {
        e->tc_qkv.resize(L); e->tc_wo.resize(L); e->tc_w13.resize(L); e->tc_w2.resize(L);
        int rc = build_tc_table(e, RK_QKV, c.dim + 2 * kv_dim, c.dim, 1);
        if (!rc) rc = build_tc_table(e, RK_WO, c.dim, c.dim, 1);
        if (!rc) rc = build_tc_table(e, RK_W13, c.hidden_dim, c.dim, 2);
        if (!rc) rc = build_tc_table(e, RK_W2, c.dim, c.hidden_dim, 1);
        if (!rc) rc = build_tc_table(e, RK_CLS, c.vocab_size, c.dim, 1);
        for (int l = 0; l < L && !rc; ++l) {
            rc = alloc_tc(e, e->tc_qkv[l], RK_QKV);
            if (!rc) rc = alloc_tc(e, e->tc_wo[l], RK_WO);
            if (!rc) rc = alloc_tc(e, e->tc_w13[l], RK_W13);
            if (!rc) rc = alloc_tc(e, e->tc_w2[l], RK_W2);
        }
        if (!rc) rc = alloc_tc(e, e->tc_cls, RK_CLS);
        if (rc) return fail(rc);
        const int qkv_rows = c.dim + 2 * kv_dim;
        CKF(cudaMalloc(&e->bx1, (size_t)kMaxRows * c.dim * 4));
        CKF(cudaMalloc(&e->bqkv, (size_t)kMaxRows * qkv_rows * 4));
        CKF(cudaMalloc(&e->batt, (size_t)kMaxRows * c.dim * 4));
        CKF(cudaMalloc(&e->bhd, (size_t)kMaxRows * c.hidden_dim * 4));
        CKF(cudaMalloc(&e->blogits, (size_t)kMaxRows * c.vocab_size * 4));
        const size_t ix = tc_image_bytes(c.dim, kMaxRows, c.group_size), ih = tc_image_bytes(c.hidden_dim, kMaxRows, c.group_size);
        CKF(cudaMalloc(&e->img_x, ix)); CKF(cudaMalloc(&e->img_h, ih)); CKF(cudaMalloc(&e->img_c, ix));
        CKF(cudaMemsetAsync(e->img_x, 0, ix, e->stream)); CKF(cudaMemsetAsync(e->img_h, 0, ih, e->stream)); CKF(cudaMemsetAsync(e->img_c, 0, ix, e->stream));
        CKF(cudaMalloc(&e->rows, sizeof(RowMeta) * kMaxRows));
        CKF(cudaMallocHost(&e->h_rows, sizeof(RowMeta) * kMaxRows));
        int max_smem = 0;
        CKF(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        e->tc_smem = max_smem - 28 * 1024;     // room for one small CTA of the neighbouring kernels next to a GEMM CTA
    }
    e->h_pos.assign(c.max_seqs, 0);
    const int es = es_of(c.quant_type);
    size_t max_elems = (size_t)c.vocab_size * c.dim;
    if ((size_t)c.hidden_dim * c.dim > max_elems) max_elems = (size_t)c.hidden_dim * c.dim;
    if ((size_t)c.dim * c.dim > max_elems) max_elems = (size_t)c.dim * c.dim;
    e->staging_bytes = ((max_elems * es + 255) & ~(size_t)255) + max_elems / c.group_size * 4 + 256;
    CKF(cudaMalloc(&e->staging, e->staging_bytes));

    CKF(cudaMalloc(&e->x1, (size_t)c.dim * 4));
    CKF(cudaMalloc(&e->qkv_buf, (size_t)(c.dim + 2 * kv_dim) * 4));
    CKF(cudaMalloc(&e->attn, (size_t)c.dim * 4));
    CKF(cudaMalloc(&e->hd, (size_t)c.hidden_dim * 4));
    CKF(cudaMalloc(&e->logits, (size_t)c.vocab_size * 4));
    CKF(cudaMalloc(&e->tap_qkv, (size_t)(c.dim + 2 * kv_dim) * 4));
    CKF(cudaMalloc(&e->tap_norm, (size_t)c.dim * 4));
    const size_t cache_elems = (size_t)c.max_seqs * L * c.n_kv_heads * c.max_seq_len * c.head_size;
    CKF(cudaMalloc(&e->k_cache, cache_elems * 4));
    CKF(cudaMalloc(&e->v_cache, cache_elems * 4));
    CKF(cudaMemsetAsync(e->k_cache, 0, cache_elems * 4, e->stream));
    CKF(cudaMemsetAsync(e->v_cache, 0, cache_elems * 4, e->stream));
    CKF(cudaMalloc(&e->states, sizeof(SeqState) * c.max_seqs));
    CKF(cudaMemsetAsync(e->states, 0, sizeof(SeqState) * c.max_seqs, e->stream));
    e->out_cap = c.max_seq_len + 8;
    CKF(cudaMalloc(&e->out_tokens, sizeof(int) * (size_t)e->out_cap * c.max_seqs));
    e->in_cap = c.max_seq_len > 64 ? c.max_seq_len : 64;
    if (c.max_seqs > e->in_cap) e->in_cap = c.max_seqs;
    CKF(cudaMalloc(&e->in_tokens, sizeof(int) * e->in_cap));
    CKF(cudaMalloc(&e->argmax_dev, sizeof(int) * c.max_seqs));
    CKF(cudaMalloc(&e->ag_send, sizeof(int) * 1024));
    CKF(cudaMalloc(&e->ag_recv, sizeof(int) * 1024 * 16));
    CKF(cudaMallocHost(&e->h_tokens, sizeof(int) * e->in_cap));
    CKF(cudaMallocHost(&e->h_logits, sizeof(float) * c.vocab_size));
    CKF(cudaMallocHost(&e->h_argmax, sizeof(int) * (c.max_seqs > 1024 ? c.max_seqs : 1024)));
    CKF(cudaStreamSynchronize(e->stream));
#undef CKF
    *out = e;
    return FL_OK;
}

void fl_destroy(fl_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto g : e->graphs) if (g) cudaGraphExecDestroy(g);
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(e->emb); fr(e->att_norm); fr(e->ffn_norm); fr(e->out_norm);
    for (auto& m : e->qkv) fr(m.d);
    for (auto& m : e->wo) fr(m.d);
    for (auto& m : e->w13) fr(m.d);
    for (auto& m : e->w2) fr(m.d);
    fr(e->cls.d); fr(e->staging);
    for (auto& m : e->rk_qkv) fr(m.d);
    for (auto& m : e->rk_wo) fr(m.d);
    for (auto& m : e->rk_w13) fr(m.d);
    for (auto& m : e->rk_w2) fr(m.d);
    fr(e->rk_cls.d);
    for (int k = 0; k < RK__COUNT; ++k) fr(e->rk_off[k]);
    fr(e->x1); fr(e->qkv_buf); fr(e->attn); fr(e->hd); fr(e->logits); fr(e->tap_qkv); fr(e->tap_norm);
    fr(e->k_cache); fr(e->v_cache); fr(e->rope); fr(e->states); fr(e->out_tokens); fr(e->in_tokens); fr(e->argmax_dev);
    fr(e->ag_send); fr(e->ag_recv);
    fr(e->mega_layers); fr(e->xchg); fr(e->prof); fr(e->evlog);
    for (auto& m : e->tc_qkv) fr(m.d);
    for (auto& m : e->tc_wo) fr(m.d);
    for (auto& m : e->tc_w13) fr(m.d);
    for (auto& m : e->tc_w2) fr(m.d);
    fr(e->tc_cls.d);
    for (int k = 0; k < RK__COUNT; ++k) fr(e->tc_off[k]);
    fr(e->bx1); fr(e->bqkv); fr(e->batt); fr(e->bhd); fr(e->blogits); fr(e->img_x); fr(e->img_h); fr(e->img_c); fr(e->rows);
    if (e->h_rows) cudaFreeHost(e->h_rows);
    for (auto& kv : e->batch_graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
    if (e->h_tokens) cudaFreeHost(e->h_tokens);
    if (e->h_logits) cudaFreeHost(e->h_logits);
    if (e->h_argmax) cudaFreeHost(e->h_argmax);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->nccl_lib) dlclose(e->nccl_lib);
    delete e;
}

int fl_upload(fl_engine* e, int kind, int layer, const void* q, const float* scales, int rows, int cols) {
    if (!e || !q) return set_err(e, FL_ERR_INVALID, "fl_upload: null argument");
    if (e->finalized) return set_err(e, FL_ERR_INVALID, "fl_upload: engine already finalized");
    const fl_config& c = e->c;
    if (kind < 0 || kind >= FL_T__COUNT) return set_err(e, FL_ERR_INVALID, "fl_upload: bad tensor kind %d", kind);
    const bool per_layer = !(kind == FL_T_TOK_EMB || kind == FL_T_OUT_NORM || kind == FL_T_CLS);
    if (layer < 0 || layer >= (per_layer ? c.n_layers : 1)) return set_err(e, FL_ERR_INVALID, "fl_upload: bad layer %d for kind %d", layer, kind);
    CK(e, cudaSetDevice(e->device));
    const int kv_dim = c.head_size * c.n_kv_heads;
    int exp_rows = 0, exp_cols = 0;
    switch (kind) {
        case FL_T_TOK_EMB: case FL_T_CLS: exp_rows = c.vocab_size; exp_cols = c.dim; break;
        case FL_T_ATT_NORM: case FL_T_FFN_NORM: case FL_T_OUT_NORM: exp_rows = 1; exp_cols = c.dim; break;
        case FL_T_WQ: case FL_T_WO: exp_rows = c.dim; exp_cols = c.dim; break;
        case FL_T_WK: case FL_T_WV: exp_rows = kv_dim; exp_cols = c.dim; break;
        case FL_T_W1: case FL_T_W3: exp_rows = c.hidden_dim; exp_cols = c.dim; break;
        case FL_T_W2: exp_rows = c.dim; exp_cols = c.hidden_dim; break;
    }
    if (rows != exp_rows || cols != exp_cols)
        return set_err(e, FL_ERR_INVALID, "fl_upload: kind %d expects %dx%d, got %dx%d", kind, exp_rows, exp_cols, rows, cols);
    const size_t n = (size_t)rows * cols;
    const int qt = c.quant_type, gs = c.group_size, es = es_of(qt);

    if (kind == FL_T_ATT_NORM || kind == FL_T_FFN_NORM || kind == FL_T_OUT_NORM) {
        float* dst = kind == FL_T_ATT_NORM ? e->att_norm + (size_t)layer * c.dim
                   : kind == FL_T_FFN_NORM ? e->ffn_norm + (size_t)layer * c.dim : e->out_norm;
        CK(e, cudaMemcpyAsync(dst, q, n * 4, cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    } else if (kind == FL_T_TOK_EMB && scales == nullptr) {
        CK(e, cudaMemcpyAsync(e->emb, q, n * 4, cudaMemcpyHostToDevice, e->stream));         // .flm keeps fp32 embedding rows
        CK(e, cudaStreamSynchronize(e->stream));
    } else {
        if (!scales) return set_err(e, FL_ERR_INVALID, "fl_upload: kind %d needs a scale table", kind);
        float* d_scales = reinterpret_cast<float*>(e->staging + ((n * es + 255) & ~(size_t)255));
        CK(e, cudaMemcpyAsync(e->staging, q, n * es, cudaMemcpyHostToDevice, e->stream));
        CK(e, cudaMemcpyAsync(d_scales, scales, n / gs * 4, cudaMemcpyHostToDevice, e->stream));
        if (kind == FL_T_TOK_EMB) {
            // dequantised once here; the reference dequantises the row per token (transformer.cpp:117-118), same bits
            dequant_rows_kernel<<<1024, 256, 0, e->stream>>>(e->staging, d_scales, e->emb, n, gs, qt);
        } else {
            if (e->tc) {
                RkMat* m = nullptr;
                int kindk = 0, m_total = rows, row_base = 0, tt = 1, sub = 0;
                switch (kind) {
                    case FL_T_WQ: m = &e->tc_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; break;
                    case FL_T_WK: m = &e->tc_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; row_base = c.dim; break;
                    case FL_T_WV: m = &e->tc_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; row_base = c.dim + kv_dim; break;
                    case FL_T_WO: m = &e->tc_wo[layer]; kindk = RK_WO; break;
                    case FL_T_W1: m = &e->tc_w13[layer]; kindk = RK_W13; tt = 2; break;
                    case FL_T_W3: m = &e->tc_w13[layer]; kindk = RK_W13; tt = 2; sub = 1; break;
                    case FL_T_W2: m = &e->tc_w2[layer]; kindk = RK_W2; break;
                    case FL_T_CLS: m = &e->tc_cls; kindk = RK_CLS; break;
                }
                if (gs == 64) pack_tc_kernel<64><<<dim3(e->n_sms, 16), 256, 0, e->stream>>>(e->staging, d_scales, m->d, e->tc_off[kindk], m_total, cols, row_base, rows, tt, sub);
                else pack_tc_kernel<32><<<dim3(e->n_sms, 16), 256, 0, e->stream>>>(e->staging, d_scales, m->d, e->tc_off[kindk], m_total, cols, row_base, rows, tt, sub);
            }
            if (e->want_mega) {
                RkMat* m = nullptr;
                int kindk = 0, m_total = rows, row_base = 0, tt = 1, sub = 0;
                switch (kind) {
                    case FL_T_WQ: m = &e->rk_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; break;
                    case FL_T_WK: m = &e->rk_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; row_base = c.dim; break;
                    case FL_T_WV: m = &e->rk_qkv[layer]; kindk = RK_QKV; m_total = c.dim + 2 * kv_dim; row_base = c.dim + kv_dim; break;
                    case FL_T_WO: m = &e->rk_wo[layer]; kindk = RK_WO; break;
                    case FL_T_W1: m = &e->rk_w13[layer]; kindk = RK_W13; tt = 2; break;
                    case FL_T_W3: m = &e->rk_w13[layer]; kindk = RK_W13; tt = 2; sub = 1; break;
                    case FL_T_W2: m = &e->rk_w2[layer]; kindk = RK_W2; break;
                    case FL_T_CLS: m = &e->rk_cls; kindk = RK_CLS; break;
                }
                int rc = dispatch_q(qt, gs, [&](auto QT, auto GS) -> int {
                    pack_rk_kernel<decltype(QT)::value, decltype(GS)::value><<<dim3(e->n_sms, 16), 256, 0, e->stream>>>(
                        e->staging, d_scales, m->d, e->rk_off[kindk], m_total, cols, row_base, rows, tt, sub, e->tile_cap);
                    return FL_OK;
                });
                if (rc) return set_err(e, rc, "fl_upload: unsupported quantisation");
            } else {
            PackedMat* m = nullptr;
            int tile_offset = 0, tile_stride = 1;
            switch (kind) {
                case FL_T_WQ: m = &e->qkv[layer]; break;
                case FL_T_WK: m = &e->qkv[layer]; tile_offset = c.dim / 4; break;
                case FL_T_WV: m = &e->qkv[layer]; tile_offset = (c.dim + kv_dim) / 4; break;
                case FL_T_WO: m = &e->wo[layer]; break;
                case FL_T_W1: m = &e->w13[layer]; tile_stride = 2; break;
                case FL_T_W3: m = &e->w13[layer]; tile_stride = 2; tile_offset = 1; break;
                case FL_T_W2: m = &e->w2[layer]; break;
                case FL_T_CLS: m = &e->cls; break;
            }
            const int n_tiles = ceil_div(rows, 4);
            int rc = dispatch_q(qt, gs, [&](auto QT, auto GS) -> int {
                pack_weights_kernel<decltype(QT)::value, decltype(GS)::value><<<2048, 256, 0, e->stream>>>(
                    e->staging, d_scales, m->d, rows, cols, n_tiles, m->nkb, tile_stride, tile_offset);
                return FL_OK;
            });
            if (rc) return set_err(e, rc, "fl_upload: unsupported quantisation");
            }
        }
        CK(e, cudaGetLastError());
        CK(e, cudaStreamSynchronize(e->stream));
    }
    e->have[(size_t)kind * c.n_layers + layer] = 1;
    return FL_OK;
}

int fl_finalize(fl_engine* e) {
    if (!e) return set_err(e, FL_ERR_INVALID, "fl_finalize: null engine");
    if (e->finalized) return FL_OK;
    const fl_config& c = e->c;
    for (int k = 0; k < FL_T__COUNT; ++k) {
        const bool per_layer = !(k == FL_T_TOK_EMB || k == FL_T_OUT_NORM || k == FL_T_CLS);
        for (int l = 0; l < (per_layer ? c.n_layers : 1); ++l)
            if (!e->have[(size_t)k * c.n_layers + l]) return set_err(e, FL_ERR_INVALID, "fl_finalize: tensor kind %d layer %d was never uploaded", k, l);
    }
    CK(e, cudaSetDevice(e->device));
    const int hgs = c.n_heads / c.n_kv_heads;
    std::vector<float> tab;
    build_rope_table(tab, c.max_seq_len * hgs + 1, c.head_size);   // q positions reach pos + g*bs (GQA quirk)
    CK(e, cudaMalloc(&e->rope, tab.size() * 4));
    CK(e, cudaMemcpyAsync(e->rope, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    if (e->staging) { cudaFree(e->staging); e->staging = nullptr; }
    if (e->want_mega) {
        int rc = setup_mega(e);
        if (rc) return rc;
        e->use_mega = true;
    }
    CK(e, cudaStreamSynchronize(e->stream));
    e->finalized = true;
    return FL_OK;
}

// copy `n` row descriptors to the device (pinned staging; the previous pass has been enqueued on the same stream)
static int push_rows(fl_engine* e, int n) {
    CK(e, cudaMemcpyAsync(e->rows, e->h_rows, sizeof(RowMeta) * n, cudaMemcpyHostToDevice, e->stream));
    return FL_OK;
}

int fl_forward(fl_engine* e, int seq_slot, const int32_t* tokens, int n_tokens, int pos, float* logits_out, int32_t* argmax_out) {
    if (!e || !tokens) return set_err(e, FL_ERR_INVALID, "fl_forward: null argument");
    if (!e->finalized) return set_err(e, FL_ERR_INVALID, "fl_forward: call fl_finalize first");
    const fl_config& c = e->c;
    if (seq_slot < 0 || seq_slot >= c.max_seqs) return set_err(e, FL_ERR_INVALID, "fl_forward: bad seq_slot %d", seq_slot);
    if (n_tokens < 1 || n_tokens > e->in_cap || pos < 0 || pos + n_tokens > c.max_seq_len)
        return set_err(e, FL_ERR_INVALID, "fl_forward: pos %d + n_tokens %d exceeds max_seq_len %d", pos, n_tokens, c.max_seq_len);
    for (int i = 0; i < n_tokens; ++i)
        if (tokens[i] < 0 || tokens[i] >= c.vocab_size) return set_err(e, FL_ERR_INVALID, "fl_forward: token id %d out of range", tokens[i]);
    CK(e, cudaSetDevice(e->device));
    const float* logits_src = e->logits;
    if (e->tc && n_tokens > 1) {
        // Prompt chunks on the tensor cores: up to kMaxRows tokens per weight pass (the reference forwards a prompt as batched
        // forward() calls too, transformer.cpp:105-151; a row's arithmetic does not depend on the rows travelling with it).
        for (int c0 = 0; c0 < n_tokens; c0 += kMaxRows) {
            const int T = n_tokens - c0 < kMaxRows ? n_tokens - c0 : kMaxRows;
            const bool last = c0 + T == n_tokens;
            if (c0 > 0) CK(e, cudaStreamSynchronize(e->stream));        // the pinned row table is being reused
            for (int i = 0; i < T; ++i) e->h_rows[i] = RowMeta{tokens[c0 + i], seq_slot, pos + c0 + i, n_tokens};
            int rc = push_rows(e, T);
            if (!rc) rc = enqueue_rows(e, T, last ? 1 : 0, 1, 1, e->stream);
            if (rc) return rc;
        }
        logits_src = e->blogits;
    } else {
        memcpy(e->h_tokens, tokens, sizeof(int) * n_tokens);
        CK(e, cudaMemcpyAsync(e->in_tokens, e->h_tokens, sizeof(int) * n_tokens, cudaMemcpyHostToDevice, e->stream));
        // token by token: bit-identical to the reference's bs > 1 forward, whose per-row arithmetic does not depend on the other rows
        for (int i = 0; i < n_tokens; ++i) {
            // n_out restarts at the last token of the call, so out_tokens[0] is the token sampled after the whole input
            set_state_kernel<<<1, 1, 0, e->stream>>>(e->states + seq_slot, e->in_tokens, i, pos + i, n_tokens, i == n_tokens - 1);
            e->launches += 1;
            e->h_pos[seq_slot] = pos + i + 1;       // launch_mega picks the long-context kernel variant by the position reached
            int rc = run_step(e, seq_slot);
            if (rc) return rc;
        }
        e->tap_rows = false;
    }
    e->h_pos[seq_slot] = pos + n_tokens;
    if (logits_out) CK(e, cudaMemcpyAsync(e->h_logits, logits_src, sizeof(float) * c.vocab_size, cudaMemcpyDeviceToHost, e->stream));
    if (logits_src != e->logits) CK(e, cudaMemcpyAsync(e->logits, logits_src, sizeof(float) * c.vocab_size, cudaMemcpyDeviceToDevice, e->stream));
    if (argmax_out) CK(e, cudaMemcpyAsync(e->h_argmax, e->argmax_dev + seq_slot, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    if (logits_out) memcpy(logits_out, e->h_logits, sizeof(float) * c.vocab_size);
    if (argmax_out) *argmax_out = e->h_argmax[0];
    return FL_OK;
}

int fl_forward_batch(fl_engine* e, int n_seqs, const int32_t* tokens, const int32_t* pos, int32_t* argmax_out) {
    if (!e || !tokens || !pos) return set_err(e, FL_ERR_INVALID, "fl_forward_batch: null argument");
    if (!e->finalized) return set_err(e, FL_ERR_INVALID, "fl_forward_batch: call fl_finalize first");
    if (n_seqs < 1 || n_seqs > e->c.max_seqs) return set_err(e, FL_ERR_INVALID, "fl_forward_batch: n_seqs %d > max_seqs %d", n_seqs, e->c.max_seqs);
    CK(e, cudaSetDevice(e->device));
    for (int i = 0; i < n_seqs; ++i) {
        if (tokens[i] < 0 || tokens[i] >= e->c.vocab_size || pos[i] < 0 || pos[i] >= e->c.max_seq_len)
            return set_err(e, FL_ERR_INVALID, "fl_forward_batch: sequence %d: token %d / pos %d out of range", i, tokens[i], pos[i]);
        e->h_tokens[i] = tokens[i];
    }
    if (e->tc && n_seqs > 1) {
        // ONE weight pass per step for all sequences (up to kMaxRows per pass): the projections are the M = n_seqs contraction
        // of SURVEY 8(d) on the tensor cores; per-sequence results are bit-identical to fl_forward on that sequence alone
        for (int s0 = 0; s0 < n_seqs; s0 += kMaxRows) {
            const int T = n_seqs - s0 < kMaxRows ? n_seqs - s0 : kMaxRows;
            if (s0 > 0) CK(e, cudaStreamSynchronize(e->stream));
            for (int i = 0; i < T; ++i) e->h_rows[i] = RowMeta{tokens[s0 + i], s0 + i, pos[s0 + i], 1};
            int rc = push_rows(e, T);
            if (!rc) rc = enqueue_rows(e, T, T, 1, 0, e->stream);
            if (rc) return rc;
        }
    } else {
    CK(e, cudaMemcpyAsync(e->in_tokens, e->h_tokens, sizeof(int) * n_seqs, cudaMemcpyHostToDevice, e->stream));
    if (e->use_mega && n_seqs > 1) {
        // one persistent launch per group of up to kMaxSeqsPerLaunch sequences: every phase is walked once per sequence, so the
        // exchanges and serial sections of one sequence hide behind the weight streaming of the others
        for (int i = 0; i < n_seqs; ++i) {
            set_state_kernel<<<1, 1, 0, e->stream>>>(e->states + i, e->in_tokens, i, pos[i], 1, 0);
            e->launches += 1;
        }
        for (int i = 0; i < n_seqs; i += kMaxSeqsPerLaunch) {
            const int n = n_seqs - i < kMaxSeqsPerLaunch ? n_seqs - i : kMaxSeqsPerLaunch;
            int rc = launch_mega(e, i, 1, n);
            if (rc) return rc;
        }
    } else {
        for (int i = 0; i < n_seqs; ++i) {
            set_state_kernel<<<1, 1, 0, e->stream>>>(e->states + i, e->in_tokens, i, pos[i], 1, 0);
            e->launches += 1;
            int rc = run_step(e, i);
            if (rc) return rc;
        }
    }
    e->tap_rows = false;
    }
    for (int i = 0; i < n_seqs; ++i) e->h_pos[i] = pos[i] + 1;
    if (argmax_out) CK(e, cudaMemcpyAsync(e->h_argmax, e->argmax_dev, sizeof(int) * n_seqs, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    if (argmax_out) memcpy(argmax_out, e->h_argmax, sizeof(int) * n_seqs);
    return FL_OK;
}

// n_steps greedy steps of sequences 0 .. n_seqs-1 from their device-resident states (set by fl_forward / fl_forward_batch),
// asynchronously on the engine stream: per step one captured graph = rows from the states, one weight pass, argmax + advance.
int fl_decode_batch_async(fl_engine* e, int n_seqs, int n_steps) {
    if (!e || !e->finalized) return set_err(e, FL_ERR_INVALID, "fl_decode_batch_async: engine not ready");
    if (!e->tc) return set_err(e, FL_ERR_UNSUPPORTED, "fl_decode_batch_async: needs the tensor-core path (INT8, FL_FLAG_NO_TC clear)");
    if (n_seqs < 1 || n_seqs > e->c.max_seqs || n_seqs > kMaxRows || n_steps < 0)
        return set_err(e, FL_ERR_INVALID, "fl_decode_batch_async: n_seqs %d (max %d) / n_steps %d", n_seqs, e->c.max_seqs < kMaxRows ? e->c.max_seqs : kMaxRows, n_steps);
    for (int i = 0; i < n_seqs; ++i)
        if (e->h_pos[i] + n_steps > e->c.max_seq_len)
            return set_err(e, FL_ERR_INVALID, "fl_decode_batch_async: sequence %d at position %d + %d steps exceeds max_seq_len %d", i, e->h_pos[i], n_steps, e->c.max_seq_len);
    CK(e, cudaSetDevice(e->device));
    const bool pdl = !(e->c.flags & FL_FLAG_NO_PDL);
    auto enqueue = [&]() -> int {
        int rc = launch_pdl(e, pdl, rows_from_states_kernel, dim3(1), dim3(kMaxRows), 0, e->stream, (const SeqState*)e->states, e->rows, n_seqs);
        if (!rc) rc = enqueue_rows(e, n_seqs, n_seqs, 1, 0, e->stream);
        return rc;
    };
    for (int s = 0; s < n_steps; ++s) {
        if (e->c.flags & FL_FLAG_NO_GRAPH) {
            if (int rc = enqueue()) return rc;
            continue;
        }
        cudaGraphExec_t& ge = e->batch_graphs[n_seqs];
        if (!ge) {
            cudaGraph_t graph = nullptr;
            const int64_t before = e->launches;
            CK(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue();
            cudaError_t st = cudaStreamEndCapture(e->stream, &graph);
            e->kernels_per_step = (int)(e->launches - before);
            e->launches = before;
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (st != cudaSuccess) return set_err(e, FL_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(st));
            st = cudaGraphInstantiate(&ge, graph, 0);
            cudaGraphDestroy(graph);
            if (st != cudaSuccess) { ge = nullptr; return set_err(e, FL_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(st)); }
        }
        CK(e, cudaGraphLaunch(ge, e->stream));
        e->launches += e->kernels_per_step;
    }
    for (int i = 0; i < n_seqs; ++i) e->h_pos[i] += n_steps;
    return FL_OK;
}

int fl_decode_async(fl_engine* e, int seq_slot, int n_steps) {
    if (!e || !e->finalized) return set_err(e, FL_ERR_INVALID, "fl_decode_async: engine not ready");
    if (seq_slot < 0 || seq_slot >= e->c.max_seqs || n_steps < 0) return set_err(e, FL_ERR_INVALID, "fl_decode_async: bad argument");
    if (e->h_pos[seq_slot] + n_steps > e->c.max_seq_len)
        return set_err(e, FL_ERR_INVALID, "fl_decode_async: position %d + %d steps exceeds max_seq_len %d", e->h_pos[seq_slot], n_steps, e->c.max_seq_len);
    CK(e, cudaSetDevice(e->device));
    e->h_pos[seq_slot] += n_steps;
    e->tap_rows = false;
    if (e->use_mega) return n_steps > 0 ? launch_mega(e, seq_slot, n_steps) : FL_OK;   // all steps inside one persistent launch
    for (int i = 0; i < n_steps; ++i) {
        int rc = run_step(e, seq_slot);
        if (rc) return rc;
    }
    return FL_OK;
}

int fl_generate_greedy(fl_engine* e, int seq_slot, const int32_t* prompt, int n_prompt, int max_new, int32_t* out_tokens, int* n_out) {
    if (!e || !prompt || !out_tokens || !n_out) return set_err(e, FL_ERR_INVALID, "fl_generate_greedy: null argument");
    const fl_config& c = e->c;
    if (n_prompt < 1 || n_prompt >= c.max_seq_len) return set_err(e, FL_ERR_INVALID, "fl_generate_greedy: prompt length %d not in [1, %d)", n_prompt, c.max_seq_len);
    if (max_new > c.max_seq_len - n_prompt) max_new = c.max_seq_len - n_prompt;      // transformer.cpp:85-87
    if (max_new < 0) max_new = 0;
    int rc = fl_forward(e, seq_slot, prompt, n_prompt, 0, nullptr, nullptr);
    if (rc) return rc;
    // transformer.cpp:93-101: one sampled token per forward; max_new further forwards at pos n_prompt ...
    rc = fl_decode_async(e, seq_slot, max_new);
    if (rc) return rc;
    const int total = 1 + max_new;
    std::vector<int> tmp(total);
    CK(e, cudaMemcpyAsync(tmp.data(), e->out_tokens + (size_t)seq_slot * e->out_cap, sizeof(int) * total, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    int n = 0;
    for (; n < total; ++n) {
        out_tokens[n] = tmp[n];
        if (tmp[n] == 0) { ++n; break; }           // generation ends on token id 0 (:93)
    }
    *n_out = n;
    return FL_OK;
}

/* ---- temperature / top-p sampling (host logic, as in the reference) ------------------------------------------------------
 * Sampler::sample (src/transformer/sampler.cpp:113-136) is host code in the reference and stays host code here: its softmax
 * (src/blas/tf_operators.cpp:188-209) is one 32 000-long FP32 running sum whose order fixes the bits, so a device version
 * would be a single-thread chain slower than the 128 KB logits copy it saves.  Greedy sampling (temperature 0, the path the
 * benchmark times) never comes here: it is fused into the classifier phase on the device. */
struct fl_sampler {
    int n = 0;
    uint64_t rng = 0;
    struct Cand { float p; int id; };
    std::vector<Cand> cand;
};

static int cand_order(const void* a, const void* b) {          // descending probability; ties compare equal (qsort decides)
    const float x = static_cast<const fl_sampler::Cand*>(a)->p, y = static_cast<const fl_sampler::Cand*>(b)->p;
    return (x < y) - (x > y);
}

static uint32_t xorshift_star(uint64_t& s) {                     // sampler.cpp:25-31
    s ^= s >> 12;
    s ^= s << 25;
    s ^= s >> 27;
    return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
}

// tf_operators.cpp:188-209.  libm expf is the reference's own; values more than 15 below the maximum are zeroed and left
// out of the sum; the normaliser is a double reciprocal rounded to float (simd::multiply takes a float).
static void sampler_softmax(float* x, int n) {
    float top = x[0];
    for (int i = 1; i < n; ++i) top = x[i] > top ? x[i] : top;
    float step[16];
    for (int i = 0; i < 16; ++i) step[i] = expf((float)(6 + i / 4));
    float total = 0.0f;
    for (int i = 0; i < n; ++i) {
        const float d = x[i] - top;
        float v = 0.0f;
        if (!(d < -15)) {
            v = d < 6.0 ? expf(d) : step[(int)((d - 6.0) * 4)];
            total += v;
        }
        x[i] = v;
    }
    const float scale = (float)(1.0 / total);
    for (int i = 0; i < n; ++i) x[i] *= scale;
}

int fl_sampler_create(int vocab_size, uint64_t seed, fl_sampler** out) {
    if (!out || vocab_size < 2) return set_err(nullptr, FL_ERR_INVALID, "fl_sampler_create: bad argument");
    fl_sampler* s = new fl_sampler();
    s->n = vocab_size;
    s->rng = seed;
    s->cand.resize((size_t)vocab_size);
    *out = s;
    return FL_OK;
}

void fl_sampler_destroy(fl_sampler* s) { delete s; }

uint64_t fl_sampler_state(const fl_sampler* s) { return s ? s->rng : 0; }

int fl_sampler_sample(fl_sampler* s, float* logits, float temperature, float topp, int32_t* token_out) {
    if (!s || !logits || !token_out) return set_err(nullptr, FL_ERR_INVALID, "fl_sampler_sample: null argument");
    const int n = s->n;
    if (temperature == 0.0f) {                                   // sampler.cpp:36-46; no random number is consumed
        int best = 0;
        for (int i = 1; i < n; ++i) if (logits[i] > logits[best]) best = i;
        *token_out = best;
        return FL_OK;
    }
    for (int i = 0; i < n; ++i) logits[i] /= temperature;
    sampler_softmax(logits, n);
    const float coin = (float)(xorshift_star(s->rng) >> 8) / 16777216.0f;
    if (topp <= 0 || topp >= 1) {                                // sample_mult, sampler.cpp:48-59
        float run = 0.0f;
        int tok = n - 1;
        for (int i = 0; i < n; ++i) {
            run += logits[i];
            if (coin < run) { tok = i; break; }
        }
        *token_out = tok;
        return FL_OK;
    }
    // sample_topp, sampler.cpp:70-111: candidates at or above (1-topp)/(n-1), sorted, cut where the mass passes topp
    const float least = (1.0f - topp) / (float)(n - 1);
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (logits[i] >= least) s->cand[(size_t)m++] = {logits[i], i};
    qsort(s->cand.data(), (size_t)m, sizeof(fl_sampler::Cand), cand_order);
    float mass = 0.0f;
    int last = m - 1;
    for (int i = 0; i < m; ++i) {
        mass += s->cand[(size_t)i].p;
        if (mass > topp) { last = i; break; }
    }
    const float r = coin * mass;
    float run = 0.0f;
    int tok = s->cand[(size_t)last].id;
    for (int i = 0; i <= last; ++i) {
        run += s->cand[(size_t)i].p;
        if (r < run) { tok = s->cand[(size_t)i].id; break; }
    }
    *token_out = tok;
    return FL_OK;
}

int fl_generate(fl_engine* e, int seq_slot, const int32_t* prompt, int n_prompt, int max_new, float temperature, float topp,
                uint64_t seed, int32_t* out_tokens, int* n_out) {
    if (!e || !prompt || !out_tokens || !n_out) return set_err(e, FL_ERR_INVALID, "fl_generate: null argument");
    if (temperature == 0.0f) return fl_generate_greedy(e, seq_slot, prompt, n_prompt, max_new, out_tokens, n_out);
    const fl_config& c = e->c;
    if (n_prompt < 1 || n_prompt >= c.max_seq_len) return set_err(e, FL_ERR_INVALID, "fl_generate: prompt length %d not in [1, %d)", n_prompt, c.max_seq_len);
    if (max_new > c.max_seq_len - n_prompt) max_new = c.max_seq_len - n_prompt;      // transformer.cpp:85-87
    if (max_new < 0) max_new = 0;
    fl_sampler* s = nullptr;
    if (int rc = fl_sampler_create(c.vocab_size, seed, &s)) return rc;
    std::vector<float> logits((size_t)c.vocab_size);
    int n = 0, pos = 0, rc = FL_OK;
    const int32_t* cur = prompt;
    int n_cur = n_prompt;
    int32_t tok = -1;
    while (tok != 0 && pos < n_prompt + max_new) {               // transformer.cpp:93-101
        rc = fl_forward(e, seq_slot, cur, n_cur, pos, logits.data(), nullptr);
        if (rc) break;
        fl_sampler_sample(s, logits.data(), temperature, topp, &tok);
        out_tokens[n++] = tok;
        pos += n_cur;
        cur = &tok;
        n_cur = 1;
    }
    fl_sampler_destroy(s);
    *n_out = n;
    return rc;
}

// the first n tokens sampled for `seq_slot` since its last prefill (out_tokens[0] is the token sampled after the prompt)
int fl_read_out_tokens(fl_engine* e, int seq_slot, int n, int32_t* out) {
    if (!e || !out || seq_slot < 0 || seq_slot >= e->c.max_seqs || n < 0 || n > e->out_cap) return set_err(e, FL_ERR_INVALID, "fl_read_out_tokens: bad argument");
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    CK(e, cudaMemcpy(out, e->out_tokens + (size_t)seq_slot * e->out_cap, sizeof(int) * n, cudaMemcpyDeviceToHost));
    return FL_OK;
}

void* fl_stream(fl_engine* e) { return e ? (void*)e->stream : nullptr; }

int fl_sync(fl_engine* e) {
    if (!e) return set_err(e, FL_ERR_INVALID, "fl_sync: null engine");
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    return FL_OK;
}

int64_t fl_launch_count(const fl_engine* e) { return e ? e->launches : 0; }

int fl_profile_read(fl_engine* e, uint64_t* out, int cap, int reset) {
    if (!e || !out || !e->prof) return set_err(e, FL_ERR_INVALID, "fl_profile_read: profiling not enabled (FL_FLAG_PROFILE) or null argument");
    const int n = 32 * e->n_sms;
    if (cap < n) return set_err(e, FL_ERR_INVALID, "fl_profile_read: buffer too small (%d < %d)", cap, n);
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    CK(e, cudaMemcpyAsync(out, e->prof, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    if (cap >= n + 2 * 4096 && e->evlog) {      // optional: the event log of the traced layer follows the counters
        CK(e, cudaMemcpyAsync(out + n, e->evlog, sizeof(uint64_t) * 2 * 4096, cudaMemcpyDeviceToHost, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    }
    if (reset) {
        CK(e, cudaMemsetAsync(e->prof, 0, sizeof(uint64_t) * n, e->stream));
        if (e->evlog) CK(e, cudaMemsetAsync(e->evlog, 0, sizeof(uint64_t) * 2 * 4096, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    }
    return n;
}

void* fl_device_ptr(fl_engine* e, const char* name, int seq_slot) {
    if (!e || !name || seq_slot < 0 || seq_slot >= e->c.max_seqs) return nullptr;
    if (!strcmp(name, "token")) return &e->states[seq_slot].token;          // int32: input token of the next step
    if (!strcmp(name, "pos")) return &e->states[seq_slot].pos;              // int32
    if (!strcmp(name, "argmax")) return e->argmax_dev + seq_slot;           // int32: last sampled token
    if (!strcmp(name, "out_tokens")) return e->out_tokens + (size_t)seq_slot * e->out_cap;
    if (!strcmp(name, "logits")) return e->logits;                          // fp32[vocab]
    if (!strcmp(name, "gathered")) return e->ag_recv;                       // int32[world * n_local]: result of the last fl_allgather_tokens
    return nullptr;
}

int64_t fl_step_bytes(const fl_engine* e, int ctx) {
    if (!e) return 0;
    const fl_config& c = e->c;
    const int64_t kv_dim = (int64_t)c.head_size * c.n_kv_heads;
    const int64_t P = 2 * (int64_t)c.dim * c.dim + 2 * kv_dim * c.dim + 3 * (int64_t)c.dim * c.hidden_dim;
    const int64_t params = (int64_t)c.n_layers * P + (int64_t)c.vocab_size * c.dim;
    const int64_t wbytes = params * es_of(c.quant_type) + params / c.group_size * 4;
    const int64_t kv = (int64_t)(ctx + 1) * 2 * c.n_layers * kv_dim * 4;
    return wbytes + kv;
}

int fl_tap(fl_engine* e, const char* name, float* out, int cap) {
    if (!e || !name || !out) return set_err(e, FL_ERR_INVALID, "fl_tap: null argument");
    const fl_config& c = e->c;
    const int kv_dim = c.head_size * c.n_kv_heads;
    const float* src = nullptr;
    int n = 0;
    if (!strcmp(name, "x1")) { src = e->x1; n = c.dim; }
    else if (!strcmp(name, "qkv")) { src = e->tap_qkv; n = c.dim + 2 * kv_dim; }
    else if (!strcmp(name, "attn")) { src = e->attn; n = c.dim; }
    else if (!strcmp(name, "hd")) { src = e->hd; n = c.hidden_dim; }
    else if (!strcmp(name, "final")) { src = e->tap_norm; n = c.dim; }
    else if (!strcmp(name, "logits")) { src = e->logits; n = c.vocab_size; }
    else return set_err(e, FL_ERR_INVALID, "fl_tap: unknown tap '%s'", name);
    if (cap < n) return set_err(e, FL_ERR_INVALID, "fl_tap: buffer too small (%d < %d)", cap, n);
    CK(e, cudaSetDevice(e->device));
    CK(e, cudaStreamSynchronize(e->stream));
    const uint2* tagged = nullptr;       // the persistent kernel keeps these vectors as (value, tag) words
    if (e->tap_rows) {                   // the last forward ran on the rows path: its last row
        if (!strcmp(name, "x1")) src = e->bx1 + (size_t)e->tap_row * c.dim;
        else if (!strcmp(name, "attn")) src = e->batt + (size_t)e->tap_row * c.dim;
        else if (!strcmp(name, "hd")) src = e->bhd + (size_t)e->tap_row * c.hidden_dim;
    } else if (e->use_mega) {
        if (!strcmp(name, "x1")) tagged = e->x1t;
        else if (!strcmp(name, "attn")) tagged = e->attnt;
        else if (!strcmp(name, "hd")) tagged = e->hdt;
        else if (!strcmp(name, "qkv")) return set_err(e, FL_ERR_UNSUPPORTED, "fl_tap: the persistent kernel keeps no post-RoPE qkv copy");
    }
    if (tagged) {
        std::vector<uint2> tmp(n);
        CK(e, cudaMemcpyAsync(tmp.data(), tagged, sizeof(uint2) * n, cudaMemcpyDeviceToHost, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
        for (int i = 0; i < n; ++i) memcpy(out + i, &tmp[i].x, 4);
        return n;
    }
    CK(e, cudaMemcpyAsync(out, src, sizeof(float) * n, cudaMemcpyDeviceToHost, e->stream));
    CK(e, cudaStreamSynchronize(e->stream));
    return n;
}

// ---- NCCL: weights replicated, request batch sharded; one all-gather of sampled tokens per step ------------
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);

int fl_set_comm(fl_engine* e, void* nccl_comm, int rank, int world) {
    if (!e || world < 1 || rank < 0 || rank >= world) return set_err(e, FL_ERR_INVALID, "fl_set_comm: bad argument");
    e->rank = rank; e->world = world; e->nccl_comm = nccl_comm;
    if (world > 1) {
        if (!nccl_comm) return set_err(e, FL_ERR_INVALID, "fl_set_comm: world > 1 needs a communicator");
        if (!e->nccl_lib) e->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!e->nccl_lib) return set_err(e, FL_ERR_NCCL, "fl_set_comm: cannot dlopen libnccl.so.2: %s", dlerror());
    }
    return FL_OK;
}

int fl_allgather_tokens(fl_engine* e, const int32_t* local, int n_local, int32_t* all) {
    if (!e || n_local < 1 || n_local > 1024 || e->world > 16 || (local && !all)) return set_err(e, FL_ERR_INVALID, "fl_allgather_tokens: bad argument");
    if (!local && n_local > e->c.max_seqs) return set_err(e, FL_ERR_INVALID, "fl_allgather_tokens: n_local %d > max_seqs %d", n_local, e->c.max_seqs);
    CK(e, cudaSetDevice(e->device));
    // local == NULL: the tokens just sampled on the device for slots 0 .. n_local-1 (no host round trip); the gathered tokens
    // stay in device memory (fl_device_ptr "gathered") unless `all` is given
    const int* src = local ? e->ag_send : e->argmax_dev;
    if (local) CK(e, cudaMemcpyAsync(e->ag_send, local, sizeof(int) * n_local, cudaMemcpyHostToDevice, e->stream));
    if (e->world == 1) {
        CK(e, cudaMemcpyAsync(e->ag_recv, src, sizeof(int) * n_local, cudaMemcpyDeviceToDevice, e->stream));
    } else {
        auto fn = (nccl_allgather_fn)dlsym(e->nccl_lib, "ncclAllGather");
        if (!fn) return set_err(e, FL_ERR_NCCL, "fl_allgather_tokens: ncclAllGather not found");
        const int ncclInt32 = 2;
        int rc = fn(src, e->ag_recv, (size_t)n_local, ncclInt32, e->nccl_comm, e->stream);
        if (rc != 0) return set_err(e, FL_ERR_NCCL, "ncclAllGather failed: %d", rc);
    }
    e->launches += 1;
    if (all) {
        CK(e, cudaMemcpyAsync(all, e->ag_recv, sizeof(int) * n_local * e->world, cudaMemcpyDeviceToHost, e->stream));
        CK(e, cudaStreamSynchronize(e->stream));
    }
    return FL_OK;
}

// =================================================================================================
// per-operator entry points
// =================================================================================================

int fl_op_quantize(int quant_type, int group_size, const float* x, int n, void* q_out, float* scales_out) {
    if (!x || !q_out || !scales_out || n < group_size || n % group_size) return set_err(nullptr, FL_ERR_INVALID, "fl_op_quantize: bad argument");
    if (int rc = need_device()) return rc;
    const int es = es_of(quant_type);
    DevBuf dx, dq, ds;
    if (dx.alloc((size_t)n * 4) || dq.alloc((size_t)n * es) || ds.alloc((size_t)n / group_size * 4)) return set_err(nullptr, FL_ERR_OOM, "fl_op_quantize: cudaMalloc");
    CKO(cudaMemcpy(dx.p, x, (size_t)n * 4, cudaMemcpyHostToDevice));
    const int nkb = ceil_div(n, kKBlockElems);
    const size_t smem = (size_t)nkb * kKBlockElems * es + (size_t)n / group_size * 4 + 64;
    int rc = dispatch_q(quant_type, group_size, [&](auto QT, auto GS) -> int {
        auto kern = op_quantize_kernel<decltype(QT)::value, decltype(GS)::value>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<1, kThreads, smem>>>(dx.as<float>(), n, dq.p, ds.as<float>());
        return FL_OK;
    });
    if (rc) return set_err(nullptr, rc, "fl_op_quantize: unsupported quantisation");
    CKO(cudaGetLastError());
    CKO(cudaMemcpy(q_out, dq.p, (size_t)n * es, cudaMemcpyDeviceToHost));
    CKO(cudaMemcpy(scales_out, ds.p, (size_t)n / group_size * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

int fl_op_matmul_q(int quant_type, int group_size, const void* w, const float* w_scales, int m, int n,
                   const void* x, const float* x_scales, int rows_x, float* out) {
    if (!w || !w_scales || !x || !x_scales || !out || m < 1 || n < group_size || n % group_size || rows_x < 1)
        return set_err(nullptr, FL_ERR_INVALID, "fl_op_matmul_q: bad argument");
    if (int rc = need_device()) return rc;
    const int es = es_of(quant_type);
    const int G = n / group_size;
    const int n_tiles = ceil_div(m, 4), nkb = ceil_div(n, kKBlockElems);
    const size_t pbytes = (size_t)n_tiles * nkb * unit_bytes(quant_type, group_size);
    DevBuf dw, dws, dp, dx, dxs, dout;
    if (dw.alloc((size_t)m * n * es) || dws.alloc((size_t)m * G * 4) || dp.alloc(pbytes) || dx.alloc((size_t)rows_x * n * es) ||
        dxs.alloc((size_t)rows_x * G * 4) || dout.alloc((size_t)rows_x * m * 4))
        return set_err(nullptr, FL_ERR_OOM, "fl_op_matmul_q: cudaMalloc");
    CKO(cudaMemcpy(dw.p, w, (size_t)m * n * es, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dws.p, w_scales, (size_t)m * G * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dx.p, x, (size_t)rows_x * n * es, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dxs.p, x_scales, (size_t)rows_x * G * 4, cudaMemcpyHostToDevice));
    int rc = dispatch_q(quant_type, group_size, [&](auto QT, auto GS) -> int {
        pack_weights_kernel<decltype(QT)::value, decltype(GS)::value><<<512, 256>>>(dw.as<uint8_t>(), dws.as<float>(), dp.as<uint8_t>(), m, n, n_tiles, nkb, 1, 0);
        return FL_OK;
    });
    if (rc) return set_err(nullptr, rc, "fl_op_matmul_q: unsupported quantisation");
    CKO(cudaGetLastError());
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    for (int i = 0; i < rows_x; ++i) {
        GemvArgs a{};
        a.w = dp.as<uint8_t>(); a.M = m; a.K = n; a.n_tasks = n_tiles; a.nkb = nkb;
        a.in_q = dx.as<uint8_t>() + (size_t)i * n * es; a.in_s = dxs.as<float>() + (size_t)i * G;
        a.out = dout.as<float>() + (size_t)i * m;
        int grid = ceil_div(n_tiles, kWarps);
        if (grid > sms * 2) grid = sms * 2;
        rc = launch_gemv<PRO_LOADQ, EPI_STORE>(nullptr, quant_type, group_size, a, grid, 0);
        if (rc) return rc;
    }
    CKO(cudaDeviceSynchronize());
    CKO(cudaMemcpy(out, dout.p, (size_t)rows_x * m * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

// quant::matmul on the tensor cores (tc_gemm.cuh), INT8 only: out[i*m + j] = W[j,:] . X[i,:] for rows_x <= 64 activation rows.
// w3 != NULL: the fused W1/W3 pass, out = swiglu(W X, W3 X) (x86_simd.cpp:1766-1770).  variant: debugging knobs of the kernel.
int fl_op_matmul_q_tc(int group_size, const void* w, const float* w_scales, const void* w3, const float* w3_scales, int m, int n,
                      const void* x, const float* x_scales, int rows_x, float* out, int variant) {
    if (!w || !w_scales || !x || !x_scales || !out || m < 1 || n < group_size || n % group_size || rows_x < 1 || rows_x > kMaxRows || (w3 && !w3_scales))
        return set_err(nullptr, FL_ERR_INVALID, "fl_op_matmul_q_tc: bad argument");
    if (group_size != 64 && group_size != 32) return set_err(nullptr, FL_ERR_UNSUPPORTED, "fl_op_matmul_q_tc: group size %d", group_size);
    if (int rc = need_device()) return rc;
    int dev = 0, sms = 148, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int G = n / group_size, tt = w3 ? 2 : 1, N = tc_pad_n(rows_x), nkc = ceil_div(n, kTcKC);
    std::vector<unsigned long long> off(sms + 1, 0);
    for (int c = 0; c < sms; ++c) {
        const TcPart pt = tc_part(m, c, sms);
        unsigned long long bytes = 0;
        for (int t = 0; t < pt.nt; ++t) { int lr0, R; tc_tile(pt, t, lr0, R); bytes += (unsigned long long)nkc * tc_stage_a_bytes(R, tt, group_size); }
        off[c + 1] = off[c] + bytes;
    }
    DevBuf dw, dws, dw3, dws3, dp, doff, dx, dxs, dimg, dout;
    const size_t img_bytes = tc_image_bytes(n, N, group_size);
    if (dw.alloc((size_t)m * n) || dws.alloc((size_t)m * G * 4) || dp.alloc(off[sms] + 4096) || doff.alloc(sizeof(unsigned long long) * (sms + 1)) ||
        dx.alloc((size_t)rows_x * n) || dxs.alloc((size_t)rows_x * G * 4) || dimg.alloc(img_bytes) || dout.alloc((size_t)rows_x * m * 4) ||
        (w3 && (dw3.alloc((size_t)m * n) || dws3.alloc((size_t)m * G * 4))))
        return set_err(nullptr, FL_ERR_OOM, "fl_op_matmul_q_tc: cudaMalloc");
    CKO(cudaMemcpy(dw.p, w, (size_t)m * n, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dws.p, w_scales, (size_t)m * G * 4, cudaMemcpyHostToDevice));
    if (w3) { CKO(cudaMemcpy(dw3.p, w3, (size_t)m * n, cudaMemcpyHostToDevice)); CKO(cudaMemcpy(dws3.p, w3_scales, (size_t)m * G * 4, cudaMemcpyHostToDevice)); }
    CKO(cudaMemcpy(doff.p, off.data(), sizeof(unsigned long long) * (sms + 1), cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dx.p, x, (size_t)rows_x * n, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dxs.p, x_scales, (size_t)rows_x * G * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemset(dp.p, 0, off[sms] + 4096));
    CKO(cudaMemset(dimg.p, 0, img_bytes));
    CKO(cudaMemset(dout.p, 0, (size_t)rows_x * m * 4));
    if (group_size == 64) {
        pack_tc_kernel<64><<<dim3(sms, 16), 256>>>(dw.as<uint8_t>(), dws.as<float>(), dp.as<uint8_t>(), doff.as<unsigned long long>(), m, n, 0, m, tt, 0);
        if (w3) pack_tc_kernel<64><<<dim3(sms, 16), 256>>>(dw3.as<uint8_t>(), dws3.as<float>(), dp.as<uint8_t>(), doff.as<unsigned long long>(), m, n, 0, m, tt, 1);
        tc_pack_x_kernel<64><<<256, 256>>>(dx.as<int8_t>(), dxs.as<float>(), dimg.as<uint8_t>(), rows_x, n, N);
    } else {
        pack_tc_kernel<32><<<dim3(sms, 16), 256>>>(dw.as<uint8_t>(), dws.as<float>(), dp.as<uint8_t>(), doff.as<unsigned long long>(), m, n, 0, m, tt, 0);
        if (w3) pack_tc_kernel<32><<<dim3(sms, 16), 256>>>(dw3.as<uint8_t>(), dws3.as<float>(), dp.as<uint8_t>(), doff.as<unsigned long long>(), m, n, 0, m, tt, 1);
        tc_pack_x_kernel<32><<<256, 256>>>(dx.as<int8_t>(), dxs.as<float>(), dimg.as<uint8_t>(), rows_x, n, N);
    }
    CKO(cudaGetLastError());
    TcGemmArgs a{};
    a.w = dp.as<uint8_t>(); a.cta_off = doff.as<unsigned long long>(); a.xq = dimg.as<uint8_t>(); a.out = dout.as<float>();
    a.M = m; a.K = n; a.T = rows_x; a.ldo = m; a.smem_bytes = max_smem - 28 * 1024; a.variant = variant;
    int rc = launch_qgemm(nullptr, false, group_size, N, w3 ? TC_EPI_SWIGLU : TC_EPI_STORE, a, sms, 0);
    if (rc) return rc;
    CKO(cudaDeviceSynchronize());
    CKO(cudaMemcpy(out, dout.p, (size_t)rows_x * m * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

int fl_op_rmsnorm(const float* x, const float* w, int n, float* out) {
    if (!x || !w || !out || n < 32 || n % 8) return set_err(nullptr, FL_ERR_INVALID, "fl_op_rmsnorm: n must be a multiple of 8, >= 32");
    if (int rc = need_device()) return rc;
    DevBuf dx, dw, dout;
    if (dx.alloc((size_t)n * 4) || dw.alloc((size_t)n * 4) || dout.alloc((size_t)n * 4)) return set_err(nullptr, FL_ERR_OOM, "fl_op_rmsnorm: cudaMalloc");
    CKO(cudaMemcpy(dx.p, x, (size_t)n * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dw.p, w, (size_t)n * 4, cudaMemcpyHostToDevice));
    cudaFuncSetAttribute(op_rmsnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 4);
    op_rmsnorm_kernel<<<1, kThreads, (size_t)n * 4>>>(dx.as<float>(), dw.as<float>(), n, dout.as<float>());
    CKO(cudaGetLastError());
    CKO(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

int fl_op_rope(const float* x, int n_dims, int pos, float* out) {
    if (!x || !out || n_dims < 2 || n_dims > 1024 || n_dims % 2 || pos < 0) return set_err(nullptr, FL_ERR_INVALID, "fl_op_rope: bad argument");
    if (int rc = need_device()) return rc;
    std::vector<float> tab((size_t)n_dims);
    {
        const float theta_scale = powf(10000.0f, -2.0f / (float)n_dims);
        float theta = (float)pos;
        for (int i = 0; i < n_dims; i += 2) { float s, c; sincosf(theta, &s, &c); tab[i] = c; tab[i + 1] = s; theta *= theta_scale; }
    }
    DevBuf dx, dt, dout;
    if (dx.alloc((size_t)n_dims * 4) || dt.alloc((size_t)n_dims * 4) || dout.alloc((size_t)n_dims * 4)) return set_err(nullptr, FL_ERR_OOM, "fl_op_rope: cudaMalloc");
    CKO(cudaMemcpy(dx.p, x, (size_t)n_dims * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dt.p, tab.data(), (size_t)n_dims * 4, cudaMemcpyHostToDevice));
    op_rope_kernel<<<1, 512>>>(dx.as<float>(), dt.as<float>(), n_dims, dout.as<float>());
    CKO(cudaGetLastError());
    CKO(cudaMemcpy(out, dout.p, (size_t)n_dims * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

static int unary_op(const float* x, int n, float* out, int which, const float* b) {
    if (!x || !out || n < 1) return set_err(nullptr, FL_ERR_INVALID, "fl_op: bad argument");
    if (int rc = need_device()) return rc;
    DevBuf dx, db, dout;
    if (dx.alloc((size_t)n * 4) || dout.alloc((size_t)n * 4) || (b && db.alloc((size_t)n * 4))) return set_err(nullptr, FL_ERR_OOM, "fl_op: cudaMalloc");
    CKO(cudaMemcpy(dx.p, x, (size_t)n * 4, cudaMemcpyHostToDevice));
    if (b) CKO(cudaMemcpy(db.p, b, (size_t)n * 4, cudaMemcpyHostToDevice));
    if (which == 0) op_softmax_kernel<<<1, kThreads>>>(dx.as<float>(), n, dout.as<float>());
    else if (which == 1) op_swiglu_kernel<<<296, 256>>>(dx.as<float>(), db.as<float>(), n, dout.as<float>());
    else op_expf_kernel<<<296, 256>>>(dx.as<float>(), n, dout.as<float>());
    CKO(cudaGetLastError());
    CKO(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

int fl_op_softmax(const float* x, int n, float* out) { return unary_op(x, n, out, 0, nullptr); }
int fl_op_swiglu(const float* a, const float* b, int n, float* out) {
    if (!b) return set_err(nullptr, FL_ERR_INVALID, "fl_op_swiglu: null argument");
    return unary_op(a, n, out, 1, b);
}
int fl_op_expf(const float* x, int n, float* out) { return unary_op(x, n, out, 2, nullptr); }

int fl_op_attn_decode(int n_heads, int n_kv_heads, int head_size, int pos, const float* qkv,
                      const float* k_cache, const float* v_cache, float* out, float* k_new, float* v_new) {
    if (!qkv || !out || n_heads < 1 || n_kv_heads < 1 || n_heads % n_kv_heads || pos < 0 || (pos > 0 && (!k_cache || !v_cache)))
        return set_err(nullptr, FL_ERR_INVALID, "fl_op_attn_decode: bad argument");
    if (head_size != 64 && head_size != 128) return set_err(nullptr, FL_ERR_UNSUPPORTED, "fl_op_attn_decode: head_size must be 64 or 128");
    if (int rc = need_device()) return rc;
    const int hs = head_size, dim = n_heads * hs, kv_dim = n_kv_heads * hs, hgs = n_heads / n_kv_heads;
    const int max_seq = pos + 1;
    std::vector<float> tab;
    build_rope_table(tab, max_seq + hgs, hs);
    DevBuf dqkv, dk, dv, dknat, dtab, dpos, dbs, dout;
    const size_t cache = (size_t)n_kv_heads * max_seq * hs;
    if (dqkv.alloc((size_t)(dim + 2 * kv_dim) * 4) || dk.alloc(cache * 4) || dv.alloc(cache * 4) || dknat.alloc(cache * 4) ||
        dtab.alloc(tab.size() * 4) || dpos.alloc(4) || dbs.alloc(4) || dout.alloc((size_t)dim * 4))
        return set_err(nullptr, FL_ERR_OOM, "fl_op_attn_decode: cudaMalloc");
    CKO(cudaMemset(dk.p, 0, cache * 4));
    CKO(cudaMemset(dv.p, 0, cache * 4));
    CKO(cudaMemcpy(dqkv.p, qkv, (size_t)(dim + 2 * kv_dim) * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dtab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    CKO(cudaMemcpy(dpos.p, &pos, 4, cudaMemcpyHostToDevice));
    const int one = 1;
    CKO(cudaMemcpy(dbs.p, &one, 4, cudaMemcpyHostToDevice));
    for (int h = 0; h < n_kv_heads && pos > 0; ++h) {   // host rows are [kv_head][pos][hs]; device rows [kv_head][max_seq][hs]
        CKO(cudaMemcpy(dknat.as<float>() + (size_t)h * max_seq * hs, k_cache + (size_t)h * pos * hs, (size_t)pos * hs * 4, cudaMemcpyHostToDevice));
        CKO(cudaMemcpy(dv.as<float>() + (size_t)h * max_seq * hs, v_cache + (size_t)h * pos * hs, (size_t)pos * hs * 4, cudaMemcpyHostToDevice));
    }
    permute_k_rows_kernel<<<256, 256>>>(dknat.as<float>(), dk.as<float>(), n_kv_heads * max_seq, hs);
    CKO(cudaGetLastError());
    AttnArgs a{};
    a.qkv = dqkv.as<float>(); a.k_cache = dk.as<float>(); a.v_cache = dv.as<float>(); a.rope = dtab.as<float>();
    a.pos_ptr = dpos.as<int>(); a.bs_ptr = dbs.as<int>(); a.out = dout.as<float>(); a.tap_qkv = nullptr;
    a.n_heads = n_heads; a.n_kv_heads = n_kv_heads; a.max_seq = max_seq; a.attn_scale = 1.0f / sqrtf((float)hs); a.v_dw = hs;
    int rc = launch_attn(nullptr, hs, a, max_seq, 0);
    if (rc) return rc;
    CKO(cudaDeviceSynchronize());
    CKO(cudaMemcpy(out, dout.p, (size_t)dim * 4, cudaMemcpyDeviceToHost));
    if (k_new || v_new) {
        std::vector<float> row(hs);
        for (int h = 0; h < n_kv_heads; ++h) {
            if (k_new) {
                CKO(cudaMemcpy(row.data(), dk.as<float>() + ((size_t)h * max_seq + pos) * hs, (size_t)hs * 4, cudaMemcpyDeviceToHost));
                for (int e2 = 0; e2 < hs; ++e2) k_new[(size_t)h * hs + e2] = row[k_cache_index(e2)];
            }
            if (v_new) CKO(cudaMemcpy(v_new + (size_t)h * hs, dv.as<float>() + ((size_t)h * max_seq + pos) * hs, (size_t)hs * 4, cudaMemcpyDeviceToHost));
        }
    }
    return FL_OK;
}

int fl_op_argmax(const float* logits, int n, int32_t* out) {
    if (!logits || !out || n < 1) return set_err(nullptr, FL_ERR_INVALID, "fl_op_argmax: bad argument");
    if (int rc = need_device()) return rc;
    DevBuf dl, dr;
    if (dl.alloc((size_t)n * 4) || dr.alloc(4)) return set_err(nullptr, FL_ERR_OOM, "fl_op_argmax: cudaMalloc");
    CKO(cudaMemcpy(dl.p, logits, (size_t)n * 4, cudaMemcpyHostToDevice));
    argmax_kernel<<<1, 1024>>>(dl.as<float>(), n, nullptr, nullptr, 0, dr.as<int>(), 0);
    CKO(cudaGetLastError());
    CKO(cudaMemcpy(out, dr.p, 4, cudaMemcpyDeviceToHost));
    return FL_OK;
}

}  // extern "C"
