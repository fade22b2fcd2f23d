/* oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" shim around the UNMODIFIED reference sources (compiled in place
 * from /root/reference by oracle/build_ref.sh; outputs only under oracle/_ref/).
 * It exposes (1) the reference's leaf operators and (2) ParallelTransformer::forward()
 * so that tests can pin oracle/ref_port.c (the C restatement) and the CUDA path
 * against the reference itself.  Nothing here is product code; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg load it.
 *
 * Access to the private ParallelTransformer::forward (src/transformer/transformer.h:99)
 * is obtained with the "#define private public" idiom applied ONLY to the reference's
 * own headers (all std headers are included first so their include guards shield them).
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <stddef.h>
#include <time.h>
#include <math.h>
#include <float.h>
#include <pthread.h>
#include <semaphore.h>
#include <sched.h>
#include <unistd.h>
#include <sys/time.h>
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <span>
#include <sstream>
#include <string>
#include <string_view>
#include <thread>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

/* reference headers that transformer.h pulls in are included untouched first (#pragma once) */
#include "threadparallel.hpp"
#include "alignmem.hpp"
#include "tf_operators.h"
#include "model_loader.h"
#include "tensor.h"
#include "log.h"
#include "ftdebug.h"
/* ... then only transformer.h is seen with every member public ("class X {" -> "struct X {"
 * makes the leading implicit-private section public too; "enum class" -> "enum struct" is
 * the same thing; layout and mangling are unaffected). */
#define private public
#define protected public
#define class struct
#include "transformer.h"
#undef class
#undef private
#undef protected

#include "quant_operators.h"
#include "tf_operators.h"
#include "simd.h"

using cpuft::quant::QuantType;

static QuantType to_qt(int qt) {
    /* 0 = fp32, 1 = int16, 2 = int8 — same numbering as quant_operators.h:17-24 */
    return static_cast<QuantType>(qt);
}

extern "C" {

/* ---------------- leaf operators (quant_operators.h / tf_operators.h / simd.h) ---------------- */

void ref_quantize(int qt, void* qx, float* qs, const float* x, size_t n, int gs) {
    cpuft::quant::quantize(to_qt(qt), qx, qs, x, n, gs);
}
void ref_dequantize(int qt, float* out, const void* qx, const float* qs, size_t n, int gs) {
    cpuft::quant::dequantize(to_qt(qt), out, qx, qs, n, gs);
}
/* out[i*m + j] = W[j,:] . X[i,:]   (quant_operators.cpp:252-284, dispatch :571) */
void ref_matmul(int qt, float* out, const void* w, const float* ws, const void* x, const float* xs,
                int m, int n, int rows_x, int gs) {
    cpuft::quant::matmul(to_qt(qt), out, w, ws, x, xs, m, n, rows_x, gs);
}
void ref_rmsnorm(float* o, const float* x, const float* w, size_t n) {
    cpuft::simd::rmsnorm(o, x, w, n);       /* x86_simd.cpp:1754 */
}
void ref_rmsnorm_inplace_via_tensor(float* x, const float* w, int n) {
    /* transformer.cpp:155 calls Tensor::rmsnorm(x, x) -> 4-arg rmsnorm with o == x */
    cpuft::simd::rmsnorm(x, x, w, size_t(n));
}
void ref_swiglu(float* xo, const float* xr, size_t n) {
    cpuft::simd::swiglu(xo, xr, n);         /* x86_simd.cpp:1766 */
}
void ref_rope_v2(float* o, const float* x, int n_dims, int n_ctx, int pos) {
    cpuft::rope_v2(o, x, n_dims, n_ctx, pos, 0, 1);   /* tf_operators.cpp:355 */
}
void ref_softmax_sisd(float* x, int n) {
    cpuft::softmax_sisd(x, n);              /* tf_operators.cpp:176 */
}
void ref_weighted_sum(float* out, const float* matrix, const float* weights, int m, int n, int bs, float min_w) {
    cpuft::weighted_sum(out, matrix, weights, m, n, bs, min_w);   /* tf_operators.cpp:325 */
}
float ref_dot_f32(const float* a, const float* b, size_t n) {
    return cpuft::simd::dot_product(a, b, n);   /* x86_simd.cpp:1677 */
}
int ref_dot_i8(const int8_t* a, const int8_t* b, size_t n) {
    return cpuft::simd::dot_product(a, b, n);
}
int ref_dot_i16(const short* a, const short* b, size_t n) {
    return cpuft::simd::dot_product(a, b, n);
}
float ref_square_sum(const float* x, size_t n) {
    return cpuft::simd::square_sum(x, n);
}
float ref_array_max(const float* x, size_t n) {
    return cpuft::simd::array_max(x, n);
}
void ref_multiply(float* x, float v, size_t n) {
    cpuft::simd::multiply(x, v, n);
}
void ref_add(float* a, const float* b, size_t n) {
    cpuft::simd::add(a, b, n);
}
size_t ref_simd_size(void) {
    return cpuft::simd::get_simd_size();
}
/* Sampler (sampler.cpp:113) — argmax at temperature 0 */
int ref_sample_argmax(const float* logits, int n) {
    int best = 0;
    float bv = logits[0];
    for (int i = 1; i < n; ++i) {       /* sampler.cpp:36-46: first index of strict max */
        if (logits[i] > bv) { bv = logits[i]; best = i; }
    }
    return best;
}

/* ---------------- whole-model oracle: ParallelTransformer ---------------- */

struct RefModel {
    cpuft::ParallelTransformer pt;
    RefModel() : pt(false) {}
};

/* Mirrors ParallelTransformer::load (transformer.cpp:23-42) but lets the caller choose the
 * context cap that the reference hard-codes to 1024 at :32 (max_seq_len <= 0 keeps 1024). */
void* ref_model_load(const char* ckpt, const char* tknr, int file_type, int quant_type,
                     int num_threads, int max_batch_size, int max_seq_len) {
    auto* m = new RefModel();
    auto& pt = m->pt;
    pt._max_batch_size = max_batch_size;
    cpuft::TransformerModel tf(false);
    if (!tf.load(ckpt, tknr ? tknr : "", static_cast<cpuft::ModelFileType>(file_type))) {
        delete m;
        return nullptr;
    }
    tf.conf.max_seq_len = max_seq_len > 0 ? max_seq_len : 1024;
    pt._tkn = std::move(tf.tokenizer);
    if (tf.conf.quant_type == QuantType::NONE) {
        tf.conf.quant_type = to_qt(quant_type);
    }
    pt._sampler.build(tf.conf.vocab_size, 0);
    if (!pt._tp.init(&pt, tf, num_threads, false)) {
        delete m;
        return nullptr;
    }
    return m;
}

void ref_model_free(void* h) {
    delete static_cast<RefModel*>(h);
}

/* cfg[0..9] = dim, hidden_dim, n_layers, n_heads, n_kv_heads, head_size, vocab_size, max_seq_len, quant_type, group */
void ref_model_config(void* h, int* cfg) {
    auto& c = static_cast<RefModel*>(h)->pt._tfc;
    cfg[0] = c.dim; cfg[1] = c.hidden_dim; cfg[2] = c.n_layers; cfg[3] = c.n_heads; cfg[4] = c.n_kv_heads;
    cfg[5] = c.head_size; cfg[6] = c.vocab_size; cfg[7] = c.max_seq_len; cfg[8] = int(c.quant_type);
    cfg[9] = c.quant_group_size;
}

/* forward(tokens, pos) -> logits[vocab]   (transformer.cpp:105-161) */
int ref_forward(void* h, const int* tokens, int n_tokens, int pos, float* logits_out) {
    auto& pt = static_cast<RefModel*>(h)->pt;
    cpuft::Tensor logits;
    pt.forward(std::span<const int>(tokens, size_t(n_tokens)), pos, logits);
    memcpy(logits_out, logits.float_data(), sizeof(float) * size_t(pt._tfc.vocab_size));
    return 0;
}

/* Greedy generate through the reference's public API (transformer.cpp:76-103), temperature 0.
 * Returns number of tokens written to out (each callback delivers one sampled token). */
int ref_generate_greedy(void* h, const int* prompt, int n_prompt, int max_new, int* out, int out_cap) {
    auto& pt = static_cast<RefModel*>(h)->pt;
    std::vector<int> in(prompt, prompt + n_prompt);
    int n = 0;
    pt.generate(in, [&](std::span<const int> toks, int, bool) -> bool {
        if (n < out_cap) out[n++] = toks[0];
        return n < out_cap;
    }, max_new, 0.0f, 0.9f);
    return n;
}

/* The reference's Sampler itself (sampler.cpp:113-136) on a caller-owned logits buffer, which it rewrites in place
 * (temperature divide + softmax) exactly as generate() lets it.  *rng is the xorshift state (sampler.cpp:25-34):
 * read before, written back after, so a sequence of calls reproduces one Sampler instance. */
int ref_sampler_sample(float* logits, int n, float temperature, float topp, uint64_t* rng) {
    cpuft::Sampler s;
    s.build(n, *rng);
    cpuft::Tensor t = cpuft::Tensor::manage(logits, n);
    int tok = s.sample(t, temperature, topp);
    *rng = s._rng_state;
    return tok;
}

/* generate() with sampling through the reference's public API (transformer.cpp:76-103); the sampler is re-seeded first
 * (ParallelTransformer::load seeds it at :40). */
int ref_generate(void* h, const int* prompt, int n_prompt, int max_new, float temperature, float topp,
                 uint64_t seed, int* out, int out_cap) {
    auto& pt = static_cast<RefModel*>(h)->pt;
    pt._sampler._rng_state = seed;
    std::vector<int> in(prompt, prompt + n_prompt);
    int n = 0;
    pt.generate(in, [&](std::span<const int> toks, int, bool) -> bool {
        if (n < out_cap) out[n++] = toks[0];
        return n < out_cap;
    }, max_new, temperature, topp);
    return n;
}

/* generate(prompt text) through the reference's public API (transformer.cpp:54-75): encode, generate, decode piece by piece.
 * Writes the sampled token ids to out_tokens and the concatenated pieces to out_text; returns the number of tokens. */
int ref_generate_text(void* h, const char* prompt, int max_new, float temperature, float topp, uint64_t seed,
                      int* out_tokens, int tok_cap, char* out_text, int text_cap) {
    auto& pt = static_cast<RefModel*>(h)->pt;
    pt._sampler._rng_state = seed;
    std::string text;
    int n = 0;
    int prev = -1;
    /* the text callback does not hand out token ids; run the token-level generate with the same decode rule (:66-71) */
    auto ids = pt.encode(prompt);
    if (ids.empty()) return 0;
    pt.generate(ids, [&](std::span<const int> toks, int, bool) -> bool {
        if (n < tok_cap) out_tokens[n] = toks[0];
        ++n;
        text += pt._tkn.decode(toks[0], prev);
        prev = toks[0];
        return n < tok_cap;
    }, max_new, temperature, topp);
    int len = int(text.size()) < text_cap - 1 ? int(text.size()) : text_cap - 1;
    memcpy(out_text, text.data(), size_t(len));
    out_text[len] = 0;
    return n;
}

int ref_encode(void* h, const char* text, int* out, int cap) {
    auto v = static_cast<RefModel*>(h)->pt.encode(text);
    int n = int(v.size()) < cap ? int(v.size()) : cap;
    memcpy(out, v.data(), sizeof(int) * size_t(n));
    return int(v.size());
}

int ref_decode(void* h, const int* tokens, int n, char* out, int cap) {
    auto s = static_cast<RefModel*>(h)->pt.decode(std::span<const int>(tokens, size_t(n)));
    int len = int(s.size()) < cap - 1 ? int(s.size()) : cap - 1;
    memcpy(out, s.data(), size_t(len));
    out[len] = 0;
    return int(s.size());
}

} /* extern "C" */
