/* oracle/ref_port.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the arithmetic on the reference's decode hot path
 * (CoderLSF/fast-llama @ c7817530; all file:line below are relative to /root/reference/).
 * It is the checker for the CUDA path: tests/ compares kernels against it bit for bit, and it
 * is itself pinned against the real reference (oracle/_ref/libref.so, built from the untouched
 * sources by oracle/build_ref.sh) in tests/test_oracle_port.py and through the committed golden
 * vectors in tests/golden/ (generated from libref.so by tests/golden/make_golden.py).
 * The product (libfastllama_b200.so, the host driver) never links or calls this file.
 *
 * Why a restatement and not just "the formulas": results must be BIT-identical, and those bits
 * depend on how g++ -O3 compiled the reference (FMA contraction, SIMD lane order).  Every such
 * choice below was read out of the disassembly of libref.so (-march=haswell and build.sh's
 * -march=native flags give the same code on this path; the AVX-512 kernels are dead code because
 * of the __AVX_512F__ macro typo, x86_simd.cpp:131).  Build with -ffp-contract=off so that only the
 * explicit fmaf() calls fuse.
 */
#define _GNU_SOURCE
#include "ref_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------- */
/* quantize  — src/blas/quant_operators.cpp:26-47, array_max_abs x86_simd.cpp:460-477            */
/*   r = maxabs(group) / QF ; q = (T)(x / r) with C truncation.  g++ vectorises the conversion as  */
/*   vcvttps2dq + vpand + vpackusdw, i.e. the LOW bits of the int32 truncation are kept and an     */
/*   unrepresentable quotient (NaN from 0/0, inf) becomes 0x80000000 -> low bits 0.                */
/* ------------------------------------------------------------------------------------------- */
#define QUANT16_FACTOR 5792.0f   /* quant_operators.h:32 */
#define QUANT8_FACTOR   127.0f   /* quant_operators.h:33 */

static inline int32_t cvtt_ss2si(float v) {
    if (v >= -2147483648.0f && v < 2147483648.0f) return (int32_t)v;
    return INT32_MIN;    /* x86 "integer indefinite" */
}

static float max_abs_f32(const float* x, size_t n) {
    float m = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float a = fabsf(x[i]);
        if (a > m) m = a;
    }
    return m;
}

void port_quantize(int qt, void* qx, float* qs, const float* x, size_t n, int gs) {
    const float QF = (qt == PORT_Q_INT8) ? QUANT8_FACTOR : QUANT16_FACTOR;
    size_t ng = (n + (size_t)gs - 1) / (size_t)gs;
    for (size_t g = 0; g < ng; ++g) {
        size_t gn = n - (size_t)gs * g;
        if (gn > (size_t)gs) gn = (size_t)gs;
        const float* xg = x + (size_t)gs * g;
        float r = max_abs_f32(xg, gn) / QF;
        qs[g] = r;
        for (size_t j = 0; j < gn; ++j) {
            int32_t iv = cvtt_ss2si(xg[j] / r);
            if (qt == PORT_Q_INT8) ((int8_t*)qx)[(size_t)gs * g + j] = (int8_t)(uint8_t)(iv & 0xff);
            else                   ((int16_t*)qx)[(size_t)gs * g + j] = (int16_t)(uint16_t)(iv & 0xffff);
        }
    }
}

/* dequantize — quant_operators.cpp:49-65: out[j] = q[j] * scale (int -> float, one multiply) */
void port_dequantize(int qt, float* out, const void* qx, const float* qs, size_t n, int gs) {
    for (size_t i = 0; i < n; ++i) {
        float q = (qt == PORT_Q_INT8) ? (float)((const int8_t*)qx)[i] : (float)((const int16_t*)qx)[i];
        out[i] = q * qs[i / (size_t)gs];
    }
}

/* ------------------------------------------------------------------------------------------- */
/* matmul — quant_operators.cpp:252-284 (+ dot_product x86_simd.cpp:1616 int8, :1524 int16)      */
/*   out[i*m + j] = sum over groups g ascending of fma(ws[j,g]*xs[i,g], (float)idot, acc), acc0=0  */
/*   (disassembly: vmulss, vcvtsi2ss, vfmadd213ss).  The BS=16 / gs blocking of the loops does not  */
/*   change the per-element order.  Integer dots wrap in int32 exactly like _mm256_add_epi32.      */
/* ------------------------------------------------------------------------------------------- */
static inline int32_t idot_i8(const int8_t* a, const int8_t* b, int n) {
    uint32_t s = 0;
    for (int k = 0; k < n; ++k) s += (uint32_t)((int32_t)a[k] * (int32_t)b[k]);
    return (int32_t)s;
}
static inline int32_t idot_i16(const int16_t* a, const int16_t* b, int n) {
    uint32_t s = 0;
    for (int k = 0; k < n; ++k) s += (uint32_t)((int32_t)a[k] * (int32_t)b[k]);
    return (int32_t)s;
}

void port_matmul(int qt, float* out, const void* w, const float* ws, const void* x, const float* xs,
                 int m, int n, int rows_x, int gs) {
    const int sn = (n + gs - 1) / gs;
    for (int i = 0; i < rows_x; ++i) {
        for (int j = 0; j < m; ++j) {
            float acc = 0.0f;
            for (int g = 0; g < sn; ++g) {
                int gn = n - g * gs < gs ? n - g * gs : gs;
                float s = ws[(size_t)j * sn + g] * xs[(size_t)i * sn + g];
                int32_t d;
                if (qt == PORT_Q_INT8)
                    d = idot_i8((const int8_t*)x + (size_t)i * n + g * gs, (const int8_t*)w + (size_t)j * n + g * gs, gn);
                else
                    d = idot_i16((const int16_t*)x + (size_t)i * n + g * gs, (const int16_t*)w + (size_t)j * n + g * gs, gn);
                acc = fmaf(s, (float)d, acc);
            }
            out[(size_t)i * m + j] = acc;
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* rmsnorm — x86_simd.cpp:1754-1764.  square_sum() dispatches to the SSE 4-lane kernel            */
/*   (x86_simd.cpp:941-962) because of the "#ifdef __AVX2" typo at :1093; each lane is an FMA chain */
/*   (vfmadd231ps), then res = 0 + l0 + l1 + l2 + l3.  r = 1/sqrtf(ss/n + 1e-5f) (g++ folds the     */
/*   double division to vdivss, which is the same value), o = (x*w)*r (multiply_avx256 :1359).      */
/*   Restricted to n % 8 == 0 and n >= 32 (every model dim is a multiple of the group size).       */
/* ------------------------------------------------------------------------------------------- */
float port_square_sum(const float* x, size_t n) {
    float l[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t i = 0; i + 3 < n; i += 4)
        for (int j = 0; j < 4; ++j) l[j] = fmaf(x[i + j], x[i + j], l[j]);
    float res = 0.0f;
    for (int j = 0; j < 4; ++j) res += l[j];
    return res;
}

void port_rmsnorm(float* o, const float* x, const float* w, size_t n) {
    const float ss = port_square_sum(x, n);
    const float r = 1.0f / sqrtf(ss / (float)n + 1e-5f);
    for (size_t i = 0; i < n; ++i) o[i] = (x[i] * w[i]) * r;
}

/* ------------------------------------------------------------------------------------------- */
/* rope_v2 — src/blas/tf_operators.cpp:355-402 via Tensor::sequence_rope_v2 (tensor.h:262-270).   */
/*   theta_scale = powf(10000, -2/n); theta_0 = pos; theta_{k+1} = theta_k * theta_scale (iterated  */
/*   float product); g++ calls glibc sincosf; o[i] = fma(cos,x0,-(sin*x1)), o[i+1] = fma(sin,x0,cos*x1) */
/*   (vmulss/vfmsub231ss and vmulss/vfmadd132ss).  ext_factor = 0, mscale = 1, zeta = 1 fold away.  */
/* ------------------------------------------------------------------------------------------- */
void port_rope_table(float* cos_sin, int n_dims, int pos) {
    const float theta_scale = powf(10000.0f, -2.0f / (float)n_dims);
    float theta = (float)pos;
    for (int i = 0; i < n_dims; i += 2) {
        float s, c;
        sincosf(theta, &s, &c);
        cos_sin[i] = c;
        cos_sin[i + 1] = s;
        theta *= theta_scale;
    }
}

void port_rope_v2(float* o, const float* x, int n_dims, int pos) {
    float tab[1024];
    port_rope_table(tab, n_dims, pos);
    for (int i = 0; i < n_dims; i += 2) {
        const float c = tab[i], s = tab[i + 1];
        const float x0 = x[i], x1 = x[i + 1];
        o[i]     = fmaf(c, x0, -(s * x1));
        o[i + 1] = fmaf(s, x0, c * x1);
    }
}

/* ------------------------------------------------------------------------------------------- */
/* float dot — x86_simd.cpp:1447-1468 (AVX2, n >= 32) / :1423-1445 (SSE, 16 <= n < 32):           */
/*   L lanes of FMA chains (vfmadd231ps), then total = 0 + l0 + l1 + ... in lane order.            */
/* ------------------------------------------------------------------------------------------- */
float port_dot_f32(const float* a, const float* b, size_t n) {
    const int L = n >= 32 ? 8 : 4;
    float l[8] = {0};
    size_t i = 0;
    for (; i + (size_t)L - 1 < n; i += (size_t)L)
        for (int j = 0; j < L; ++j) l[j] = fmaf(a[i + j], b[i + j], l[j]);
    float t = 0.0f;
    for (int j = 0; j < L; ++j) t += l[j];
    for (; i < n; ++i) t = fmaf(a[i], b[i], t);
    return t;
}

/* ------------------------------------------------------------------------------------------- */
/* expf as glibc 2.39 computes it (sysdeps/ieee754/flt-32/e_expf.c, Szabolcs Nagy's algorithm):   */
/*   exp(x) = 2^(k/32) * 2^(r/32), table of 32 doubles + cubic in double, one rounding to float.    */
/*   The reference calls libm expf (softmax tf_operators.cpp:180, swiglu x86_simd.cpp:1768); the     */
/*   CUDA kernels run this restatement in fp64.  tests check port_expf_emul(x) == expf(x) bitwise.  */
/* ------------------------------------------------------------------------------------------- */
static const uint64_t EXP2F_TAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};

float port_expf_emul(float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    uint32_t ix; memcpy(&ix, &x, 4);
    uint32_t abstop = (ix >> 20) & 0x7ff;
    if (abstop >= 0x42b) {                 /* |x| >= 88 or NaN */
        if (ix == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return x + x;
        if (x > 0x1.62e42ep6f) return INFINITY;             /* overflow */
        if (x < -0x1.9fe368p6f) return 0.0f;                /* underflow */
        if (x < -0x1.9d1d9ep6f) return 0x1p-149f;           /* __math_may_uflowf: 0x1.4p-75f^2 rounded */
    }
    double xd = (double)x;
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki; memcpy(&ki, &kd, 8);
    kd -= SHIFT;
    double r = fma(InvLn2N, xd, -kd);   /* contracted z - kd (see below) */
    uint64_t t = EXP2F_TAB[ki % 32];
    t += ki << (52 - 5);
    double s; memcpy(&s, &t, 8);
    /* glibc selects its FMA build of expf (__expf_fma, ifunc) on every AVX2+FMA host, in which gcc
     * contracted r = InvLn2N*xd - kd into one fma (the non-fused form differs from libm on 2 of the
     * 2^32 inputs; with it, all 2^32 inputs are bit-identical to glibc 2.39 expf — fused or unfused
     * polynomial makes no difference to any float result). */
    z = fma(C0, r, C1);
    double r2 = r * r;
    double y = fma(C2, r, 1.0);
    y = fma(z, r2, y);
    y = y * s;
    return (float)y;
}

/* ------------------------------------------------------------------------------------------- */
/* softmax_sisd — src/blas/tf_operators.cpp:176-186: max, e = expf(x-max), serial sum, e/sum      */
/* ------------------------------------------------------------------------------------------- */
void port_softmax_sisd(float* x, int n) {
    float mx = x[0];
    for (int i = 1; i < n; ++i) if (x[i] > mx) mx = x[i];
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        x[i] = expf(x[i] - mx);
        sum += x[i];
    }
    for (int i = 0; i < n; ++i) x[i] /= sum;
}

/* ------------------------------------------------------------------------------------------- */
/* weighted_sum (bs variant) — src/blas/tf_operators.cpp:325-350:                                 */
/*   o = V[0]*w0 ; for t>=1: if fabsf(w_t) > min_w: o[j] = fma(V[t][j], w_t, o[j])  (vfmadd213ps)    */
/*   weights laid out [bs][m], out [bs][n].                                                        */
/* ------------------------------------------------------------------------------------------- */
void port_weighted_sum(float* out, const float* matrix, const float* weights, int m, int n, int bs, float min_w) {
    for (int k = 0; k < bs; ++k) {
        float w = weights[(size_t)m * k];
        for (int j = 0; j < n; ++j) out[(size_t)k * n + j] = matrix[j] * w;
    }
    for (int i = 1; i < m; ++i) {
        const float* row = matrix + (size_t)n * i;
        for (int k = 0; k < bs; ++k) {
            float w = weights[(size_t)m * k + i];
            if (fabsf(w) <= min_w) continue;
            float* o = out + (size_t)k * n;
            for (int j = 0; j < n; ++j) o[j] = fmaf(row[j], w, o[j]);
        }
    }
}

/* ------------------------------------------------------------------------------------------- */
/* swiglu — x86_simd.cpp:1766-1770: xo = float( double(xo) / (1.0 + double(expf(-xo))) * double(xr) ) */
/* ------------------------------------------------------------------------------------------- */
void port_swiglu(float* xo, const float* xr, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        double e = (double)expf(-xo[i]);
        xo[i] = (float)((double)xo[i] / (e + 1.0) * (double)xr[i]);
    }
}

/* argmax — src/transformer/sampler.cpp:36-46: first index of the strict maximum */
int port_argmax(const float* logits, int n) {
    int best = 0;
    float bv = logits[0];
    for (int i = 1; i < n; ++i) if (logits[i] > bv) { bv = logits[i]; best = i; }
    return best;
}

/* ------------------------------------------------------------------------------------------- */
/* Sampler — src/transformer/sampler.cpp:25-136 with the softmax it calls (src/blas/tf_operators.cpp */
/* :188-209, NOT the attention's softmax_sisd): logits /= T; p = softmax; coin = xorshift* >> 8 /2^24; */
/* topp outside (0,1): first index whose running sum exceeds coin; else nucleus sampling over the     */
/* candidates p >= (1-topp)/(n-1), sorted by libc qsort (same comparator => same order on ties).     */
/* ------------------------------------------------------------------------------------------- */
void port_softmax(float* x, size_t n) {
    float mx = x[0];
    for (size_t i = 1; i < n; ++i) if (x[i] > mx) mx = x[i];          /* array_max: order-free */
    float big[16];
    for (int i = 0; i < 16; ++i) big[i] = expf((float)(6 + i / 4));   /* :192-194, integer i/4 */
    float sum = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float d = x[i] - mx;
        if (d < -15) { x[i] = 0.0f; continue; }                        /* dropped AND not summed */
        x[i] = d < 6.0 ? expf(d) : big[(int)((d - 6.0) * 4)];
        sum += x[i];
    }
    float inv = (float)(1.0 / sum);                                    /* double division, then float (simd.h:51) */
    for (size_t i = 0; i < n; ++i) x[i] *= inv;
}

uint32_t port_random_u32(uint64_t* st) {
    uint64_t s = *st;
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    *st = s;
    return (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
}

typedef struct { float p; int id; } port_cand;
static int cand_desc(const void* a, const void* b) {
    float pa = ((const port_cand*)a)->p, pb = ((const port_cand*)b)->p;
    return pa > pb ? -1 : (pa < pb ? 1 : 0);
}

int port_sample(float* logits, int n, float temperature, float topp, uint64_t* rng) {
    if (temperature == 0.0f) return port_argmax(logits, n);            /* no coin is drawn */
    for (int i = 0; i < n; ++i) logits[i] /= temperature;
    port_softmax(logits, (size_t)n);
    float coin = (float)(port_random_u32(rng) >> 8) / 16777216.0f;
    if (topp <= 0 || topp >= 1) {                                      /* sample_mult :48-59 */
        float run = 0.0f;
        for (int i = 0; i < n; ++i) { run += logits[i]; if (coin < run) return i; }
        return n - 1;
    }
    port_cand* c = (port_cand*)malloc(sizeof(port_cand) * (size_t)n);  /* sample_topp :70-111 */
    float floor_p = (1.0f - topp) / (float)(n - 1);
    int m = 0;
    for (int i = 0; i < n; ++i) if (logits[i] >= floor_p) { c[m].p = logits[i]; c[m].id = i; ++m; }
    qsort(c, (size_t)m, sizeof(port_cand), cand_desc);
    float mass = 0.0f;
    int last = m - 1;
    for (int i = 0; i < m; ++i) { mass += c[i].p; if (mass > topp) { last = i; break; } }
    float r = coin * mass, run = 0.0f;
    int tok = c[last].id;
    for (int i = 0; i <= last; ++i) { run += c[i].p; if (r < run) { tok = c[i].id; break; } }
    free(c);
    return tok;
}

/* ------------------------------------------------------------------------------------------- */
/* whole model: ParallelTransformer::forward — src/transformer/transformer.cpp:105-161 and the     */
/* six task bodies :386-505.  Row/head partitioning across worker threads has no cross-thread      */
/* reduction, so a single-threaded restatement is bit-identical for any -j.                        */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
    void*  q;        /* int8 / int16 payload, or fp32 when scales == NULL */
    float* scales;
    int rows, cols;
} port_tensor;

struct port_model {
    port_config c;
    port_tensor* t[PORT_T__COUNT];       /* [kind][layer] */
    float* k_cache;                      /* [layer][kv_head][max_seq][head_size] (transformer.cpp:366-374) */
    float* v_cache;
    /* taps of the last forward (last row) */
    float* tap_x1;      /* [n_layers][dim] residual stream after each layer */
    float* tap_attn;    /* [n_layers][dim] attention output (before Wo) */
    float* tap_qkv;     /* [n_layers][dim + 2 kv_dim] after rope */
    float* tap_hd;      /* [n_layers][hidden] */
    float* tap_final;   /* [dim] after final rmsnorm */
};

static size_t elem_size(int qt) { return qt == PORT_Q_INT8 ? 1 : (qt == PORT_Q_INT16 ? 2 : 4); }

port_model* port_model_create(const port_config* cfg) {
    port_model* m = (port_model*)calloc(1, sizeof(port_model));
    m->c = *cfg;
    for (int k = 0; k < PORT_T__COUNT; ++k) m->t[k] = (port_tensor*)calloc((size_t)cfg->n_layers, sizeof(port_tensor));
    size_t kv_dim = (size_t)cfg->head_size * cfg->n_kv_heads;
    size_t cache = (size_t)cfg->n_layers * cfg->max_seq_len * kv_dim;
    m->k_cache = (float*)calloc(cache, sizeof(float));
    m->v_cache = (float*)calloc(cache, sizeof(float));
    m->tap_x1 = (float*)calloc((size_t)cfg->n_layers * cfg->dim, sizeof(float));
    m->tap_attn = (float*)calloc((size_t)cfg->n_layers * cfg->dim, sizeof(float));
    m->tap_qkv = (float*)calloc((size_t)cfg->n_layers * (cfg->dim + 2 * kv_dim), sizeof(float));
    m->tap_hd = (float*)calloc((size_t)cfg->n_layers * cfg->hidden_dim, sizeof(float));
    m->tap_final = (float*)calloc((size_t)cfg->dim, sizeof(float));
    return m;
}

void port_model_free(port_model* m) {
    if (!m) return;
    for (int k = 0; k < PORT_T__COUNT; ++k) {
        for (int l = 0; l < m->c.n_layers; ++l) { free(m->t[k][l].q); free(m->t[k][l].scales); }
        free(m->t[k]);
    }
    free(m->k_cache); free(m->v_cache);
    free(m->tap_x1); free(m->tap_attn); free(m->tap_qkv); free(m->tap_hd); free(m->tap_final);
    free(m);
}

void port_model_reset(port_model* m) {
    size_t cache = (size_t)m->c.n_layers * m->c.max_seq_len * m->c.head_size * m->c.n_kv_heads;
    memset(m->k_cache, 0, cache * sizeof(float));
    memset(m->v_cache, 0, cache * sizeof(float));
}

int port_model_set_tensor(port_model* m, int kind, int layer, const void* q, const float* scales, int rows, int cols) {
    if (kind < 0 || kind >= PORT_T__COUNT || layer < 0 || layer >= m->c.n_layers) return -1;
    port_tensor* t = &m->t[kind][layer];
    free(t->q); free(t->scales);
    size_t n = (size_t)rows * cols;
    int is_norm = (kind == PORT_T_ATT_NORM || kind == PORT_T_FFN_NORM || kind == PORT_T_OUT_NORM);
    /* fp32 when no scale table comes with it (norm gains; .flm embedding rows), else the model's integer type */
    size_t es = (scales == NULL || is_norm) ? 4 : elem_size(m->c.qtype);
    t->q = malloc(n * es);
    memcpy(t->q, q, n * es);
    t->scales = NULL;
    if (scales && !is_norm) {
        size_t ns = n / (size_t)m->c.group;
        t->scales = (float*)malloc(ns * sizeof(float));
        memcpy(t->scales, scales, ns * sizeof(float));
    }
    t->rows = rows; t->cols = cols;
    return 0;
}

const float* port_tap(port_model* m, const char* name, int layer, int* n_out) {
    const port_config* c = &m->c;
    int kv_dim = c->head_size * c->n_kv_heads;
    if (!strcmp(name, "x1"))   { *n_out = c->dim; return m->tap_x1 + (size_t)layer * c->dim; }
    if (!strcmp(name, "attn")) { *n_out = c->dim; return m->tap_attn + (size_t)layer * c->dim; }
    if (!strcmp(name, "qkv"))  { *n_out = c->dim + 2 * kv_dim; return m->tap_qkv + (size_t)layer * (c->dim + 2 * kv_dim); }
    if (!strcmp(name, "hd"))   { *n_out = c->hidden_dim; return m->tap_hd + (size_t)layer * c->hidden_dim; }
    if (!strcmp(name, "final")){ *n_out = c->dim; return m->tap_final; }
    *n_out = 0;
    return NULL;
}

/* quantise rows of x (bs x n) then multiply with weight tensor t: out (bs x t->rows) */
static void qmatmul(const port_model* m, const port_tensor* t, const float* x, int bs, float* out,
                    void* qbuf, float* sbuf) {
    const int n = t->cols, gs = m->c.group, qt = m->c.qtype;
    /* Tensor::quantize (tensor.cpp:462-484) flattens [bs][n]; n % gs == 0 so groups never straddle rows */
    port_quantize(qt, qbuf, sbuf, x, (size_t)bs * n, gs);
    port_matmul(qt, out, t->q, t->scales, qbuf, sbuf, t->rows, n, bs, gs);
}

int port_forward(port_model* m, const int* tokens, int bs, int pos, float* logits_out) {
    const port_config* c = &m->c;
    const int dim = c->dim, hid = c->hidden_dim, hs = c->head_size;
    const int kv_dim = hs * c->n_kv_heads, hgs = c->n_heads / c->n_kv_heads;
    const int qkv_w = dim + 2 * kv_dim;
    const int seqlen = pos + bs;
    if (seqlen > c->max_seq_len) return -1;
    const int maxw = hid > qkv_w ? hid : qkv_w;

    float* x1  = (float*)malloc(sizeof(float) * (size_t)bs * dim);
    float* x2  = (float*)malloc(sizeof(float) * (size_t)bs * dim);
    float* qkv = (float*)malloc(sizeof(float) * (size_t)bs * qkv_w);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)bs * (maxw > c->vocab_size ? maxw : c->vocab_size));
    float* hd  = (float*)malloc(sizeof(float) * (size_t)bs * hid);
    float* h3  = (float*)malloc(sizeof(float) * (size_t)bs * hid);
    float* att = (float*)malloc(sizeof(float) * (size_t)seqlen);
    void*  qb  = malloc((size_t)bs * maxw * 2);
    float* sb  = (float*)malloc(sizeof(float) * ((size_t)bs * maxw / c->group + 1));

    /* embedding: transformer.cpp:115-122 */
    for (int i = 0; i < bs; ++i) {
        const port_tensor* e = &m->t[PORT_T_TOK_EMB][0];
        if (e->scales) {
            size_t off = (size_t)tokens[i] * dim;
            port_dequantize(c->qtype, x1 + (size_t)i * dim, (const char*)e->q + off * elem_size(c->qtype),
                            e->scales + off / c->group, (size_t)dim, c->group);
        } else {
            memcpy(x1 + (size_t)i * dim, (const float*)e->q + (size_t)tokens[i] * dim, sizeof(float) * dim);
        }
    }

    int rows = bs;     /* rows of x1 still alive (transformer.cpp:140-142 keeps only the last row) */
    float* x1p = x1;
    const float attn_scale = 1.0f / sqrtf((float)hs);   /* transformer.cpp:416 */

    for (int l = 0; l < c->n_layers; ++l) {
        /* :132-135  x2 = rmsnorm(x1); qkv = [Wq;Wk;Wv] * quantize(x2) */
        for (int i = 0; i < bs; ++i)
            port_rmsnorm(x2 + (size_t)i * dim, x1p + (size_t)i * dim, (const float*)m->t[PORT_T_ATT_NORM][l].q, (size_t)dim);
        {
            port_quantize(c->qtype, qb, sb, x2, (size_t)bs * dim, c->group);
            const port_tensor* ws[3] = { &m->t[PORT_T_WQ][l], &m->t[PORT_T_WK][l], &m->t[PORT_T_WV][l] };
            int col = 0;
            for (int k = 0; k < 3; ++k) {
                port_matmul(c->qtype, tmp, ws[k]->q, ws[k]->scales, qb, sb, ws[k]->rows, dim, bs, c->group);
                for (int i = 0; i < bs; ++i)
                    memcpy(qkv + (size_t)i * qkv_w + col, tmp + (size_t)i * ws[k]->rows, sizeof(float) * ws[k]->rows);
                col += ws[k]->rows;
            }
        }
        /* :136 execute_attn (:397-455) */
        float* kc = m->k_cache + (size_t)l * c->max_seq_len * kv_dim;
        float* vc = m->v_cache + (size_t)l * c->max_seq_len * kv_dim;
        for (int h = 0; h < c->n_kv_heads; ++h) {
            float* kh = kc + (size_t)h * c->max_seq_len * hs;   /* [max_seq][hs] for this kv head */
            float* vh = vc + (size_t)h * c->max_seq_len * hs;
            for (int i = 0; i < bs; ++i) {          /* :431-432, :439 append k (roped) and v */
                const float* row = qkv + (size_t)i * qkv_w;
                float* kdst = kh + (size_t)(pos + i) * hs;
                port_rope_v2(kdst, row + dim + h * hs, hs, pos + i);
                memcpy(vh + (size_t)(pos + i) * hs, row + dim + kv_dim + h * hs, sizeof(float) * hs);
                memcpy(qkv + (size_t)i * qkv_w + dim + h * hs, kdst, sizeof(float) * hs);   /* tap only */
            }
            for (int g = 0; g < hgs; ++g) {
                const int qh = h * hgs + g;
                for (int i = 0; i < bs; ++i) {
                    float* q = qkv + (size_t)i * qkv_w + qh * hs;
                    /* :438 + tensor.h:262-270: sequence_rope_v2 walks total_rows() = bs*hgs rows with
                     * position pos + row, so query head g of a GQA group is rotated at pos + g*bs + i
                     * (identical to pos + i when n_heads == n_kv_heads, as in LLaMA2-7B/13B). */
                    port_rope_v2(q, q, hs, pos + g * bs + i);
                    const int nctx = pos + i + 1;
                    for (int t = 0; t < nctx; ++t)                         /* :442-443 */
                        att[t] = port_dot_f32(kh + (size_t)t * hs, q, (size_t)hs) * attn_scale;
                    port_softmax_sisd(att, nctx);                          /* :444-448 */
                    /* :449 weighted_sum over all seqlen rows; rows >= nctx have weight 0 and are skipped */
                    port_weighted_sum(x2 + (size_t)i * dim + qh * hs, vh, att, nctx, hs, 1, 1e-15f);
                }
            }
        }
        memcpy(m->tap_qkv + (size_t)l * qkv_w, qkv + (size_t)(bs - 1) * qkv_w, sizeof(float) * qkv_w);
        memcpy(m->tap_attn + (size_t)l * dim, x2 + (size_t)(bs - 1) * dim, sizeof(float) * dim);
        /* :138-139  x1 += Wo * quantize(attn) */
        qmatmul(m, &m->t[PORT_T_WO][l], x2, bs, tmp, qb, sb);
        for (size_t i = 0; i < (size_t)bs * dim; ++i) x1p[i] += tmp[i];
        /* :140-142 only the last row continues after the last layer's attention */
        if (bs > 1 && l == c->n_layers - 1) { x1p = x1p + (size_t)(rows - 1) * dim; rows = 1; }
        /* :144-147  hd = swiglu(W1 q, W3 q) */
        for (int i = 0; i < rows; ++i)
            port_rmsnorm(x2 + (size_t)i * dim, x1p + (size_t)i * dim, (const float*)m->t[PORT_T_FFN_NORM][l].q, (size_t)dim);
        port_quantize(c->qtype, qb, sb, x2, (size_t)rows * dim, c->group);
        port_matmul(c->qtype, hd, m->t[PORT_T_W1][l].q, m->t[PORT_T_W1][l].scales, qb, sb, hid, dim, rows, c->group);
        port_matmul(c->qtype, h3, m->t[PORT_T_W3][l].q, m->t[PORT_T_W3][l].scales, qb, sb, hid, dim, rows, c->group);
        port_swiglu(hd, h3, (size_t)rows * hid);
        memcpy(m->tap_hd + (size_t)l * hid, hd + (size_t)(rows - 1) * hid, sizeof(float) * hid);
        /* :149-150  x1 += W2 * quantize(hd) */
        qmatmul(m, &m->t[PORT_T_W2][l], hd, rows, tmp, qb, sb);
        for (size_t i = 0; i < (size_t)rows * dim; ++i) x1p[i] += tmp[i];
        memcpy(m->tap_x1 + (size_t)l * dim, x1p + (size_t)(rows - 1) * dim, sizeof(float) * dim);
    }

    /* :154-160 final norm (in place), quantise, classifier */
    float* x = x1p + (size_t)(rows - 1) * dim;
    port_rmsnorm(x, x, (const float*)m->t[PORT_T_OUT_NORM][0].q, (size_t)dim);
    memcpy(m->tap_final, x, sizeof(float) * dim);
    qmatmul(m, &m->t[PORT_T_CLS][0], x, 1, logits_out, qb, sb);

    free(x1); free(x2); free(qkv); free(tmp); free(hd); free(h3); free(att); free(qb); free(sb);
    return 0;
}
