/* fl_binding.cpp — the reference-side binding of include/fastllama_b200.h, COMPILED against the reference's own headers.
 *
 * This is the code INTEGRATION.md shows a maintainer of CoderLSF/fast-llama adding: the reference's loader parses the model
 * file into cpuft::TransformerModel{conf, weights} (src/model_loaders/model_loader.h:47-113), and where
 * ParallelTransformer::parallel_global_init / parallel_thread_init (src/transformer/transformer.cpp:209-384) would copy
 * the loader's tensors into per-worker slices, the binding hands the same tensors to fl_upload; forward()
 * (transformer.cpp:105-161) becomes fl_forward.  oracle/build_ref.sh compiles this file together with the unmodified
 * reference sources into oracle/_ref/libfl_binding.so (test infrastructure; the product library is only dlopen'ed), and
 * tests/test_binding_cxx.py runs it on the GPU and compares it with the ctypes path.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "model_loader.h"
#include "tensor.h"
#include "../../include/fastllama_b200.h"

using namespace cpuft;

namespace {

struct Api {
    void* lib = nullptr;
    decltype(&fl_create) create = nullptr;
    decltype(&fl_destroy) destroy = nullptr;
    decltype(&fl_last_error) last_error = nullptr;
    decltype(&fl_upload) upload = nullptr;
    decltype(&fl_finalize) finalize = nullptr;
    decltype(&fl_forward) forward = nullptr;
    bool open(const char* path) {
        lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!lib) { fprintf(stderr, "fl_binding: dlopen(%s): %s\n", path, dlerror()); return false; }
#define FL_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(lib, #name)); if (!field) { fprintf(stderr, "fl_binding: missing %s\n", #name); return false; }
        FL_SYM(create, fl_create) FL_SYM(destroy, fl_destroy) FL_SYM(last_error, fl_last_error)
        FL_SYM(upload, fl_upload) FL_SYM(finalize, fl_finalize) FL_SYM(forward, fl_forward)
#undef FL_SYM
        return true;
    }
};

/* One tensor of the loader -> fl_upload.  Quantised tensors go as they are (row-major payload + fp32 scale per group,
 * src/components/tensor.h:473-504); an fp32 projection is quantised first with the reference's own Tensor::quantize, exactly as
 * parallel_thread_init does when the file's type differs from the requested one (transformer.cpp:288-299). */
bool upload(const Api& api, fl_engine* e, int kind, int layer, const Tensor& t, QuantType want, int group, bool quantise_fp32) {
    const int rows = t.rows() > 0 ? t.rows() : 1;
    if (t.is_quantized() || !quantise_fp32)
        return api.upload(e, kind, layer, t.int8_data(), t.is_quantized() ? t.scales() : nullptr, rows, t.columns()) == FL_OK;
    Tensor q(t.columns(), rows, want, group);
    if (!q.reserve_memory() || !q.quantize(t)) return false;
    return api.upload(e, kind, layer, q.int8_data(), q.scales(), rows, t.columns()) == FL_OK;
}

}  // namespace

extern "C" {

/* Load `ckpt` with the reference's loader, run the prompt and `n_decode` greedy steps through fl_forward.
 * logits_out: (1 + n_decode) x vocab floats.  Returns the vocabulary size, or a negative value on failure. */
int fl_binding_run(const char* lib_path, const char* ckpt, const char* tokenizer, int file_type, int quant_type,
                   const int* prompt, int n_prompt, int n_decode, float* logits_out) {
    Api api;
    if (!api.open(lib_path)) return -1;
    TransformerModel tf(false);
    if (!tf.load(ckpt, tokenizer ? tokenizer : "", static_cast<ModelFileType>(file_type))) return -2;
    const TransformerConfig& c = tf.conf;
    const TransformerWeights& w = tf.weights;
    const QuantType qt = c.quant_type != QuantType::NONE ? c.quant_type : static_cast<QuantType>(quant_type);

    fl_config fc;
    memset(&fc, 0, sizeof(fc));
    fc.dim = c.dim; fc.hidden_dim = c.hidden_dim; fc.n_layers = c.n_layers; fc.n_heads = c.n_heads; fc.n_kv_heads = c.n_kv_heads;
    fc.head_size = c.head_size > 0 ? c.head_size : c.dim / c.n_heads;
    fc.vocab_size = c.vocab_size;
    fc.max_seq_len = 1024;                                   /* ParallelTransformer::load forces 1024, transformer.cpp:32 */
    fc.quant_type = static_cast<int>(qt);                    /* QuantType numbering is shared (quant_operators.h:17-24) */
    fc.group_size = c.quant_group_size;
    fc.max_seqs = 1;
    fc.flags = 0;
    fl_engine* e = nullptr;
    if (api.create(&fc, 0, &e) != FL_OK) { fprintf(stderr, "fl_binding: %s\n", api.last_error(nullptr)); return -3; }
    const int g = c.quant_group_size;
    bool ok = upload(api, e, FL_T_TOK_EMB, 0, w.token_embedding_table, qt, g, false) &&
              upload(api, e, FL_T_OUT_NORM, 0, w.out_norm, qt, g, false) &&
              upload(api, e, FL_T_CLS, 0, w.classifier, qt, g, true);
    for (int l = 0; ok && l < c.n_layers; ++l)
        ok = upload(api, e, FL_T_ATT_NORM, l, w.attn_norm[l], qt, g, false) && upload(api, e, FL_T_FFN_NORM, l, w.ffn_norm[l], qt, g, false) &&
             upload(api, e, FL_T_WQ, l, w.attn_q[l], qt, g, true) && upload(api, e, FL_T_WK, l, w.attn_k[l], qt, g, true) &&
             upload(api, e, FL_T_WV, l, w.attn_v[l], qt, g, true) && upload(api, e, FL_T_WO, l, w.attn_o[l], qt, g, true) &&
             upload(api, e, FL_T_W1, l, w.ffn_1[l], qt, g, true) && upload(api, e, FL_T_W2, l, w.ffn_2[l], qt, g, true) &&
             upload(api, e, FL_T_W3, l, w.ffn_3[l], qt, g, true);
    if (!ok || api.finalize(e) != FL_OK) { fprintf(stderr, "fl_binding: %s\n", api.last_error(e)); api.destroy(e); return -4; }

    /* generate()'s loop (transformer.cpp:93-101) with the greedy sampler, through forward() = fl_forward */
    std::vector<int> cur(prompt, prompt + n_prompt);
    int pos = 0, rc = c.vocab_size;
    for (int step = 0; step <= n_decode; ++step) {
        float* logits = logits_out + (size_t)step * c.vocab_size;
        int32_t next = 0;
        if (api.forward(e, 0, cur.data(), (int)cur.size(), pos, logits, &next) != FL_OK) { fprintf(stderr, "fl_binding: %s\n", api.last_error(e)); rc = -5; break; }
        pos += (int)cur.size();
        cur.assign(1, next);
    }
    api.destroy(e);
    return rc;
}

}  /* extern "C" */
