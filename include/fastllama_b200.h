/* fastllama_b200.h — C-ABI of the B200-native decode engine (libfastllama_b200.so).
 *
 * Drop-in boundary for the hot path of CoderLSF/fast-llama (reference @ c7817530; file:line below are
 * relative to the reference tree).  The reference has no plugin/FFI layer; the seam this library
 * replaces is everything from ParallelTransformer::forward() downwards:
 *
 *   fl_create / fl_upload / fl_finalize   <- ParallelTransformer::load -> parallel_global_init /
 *                                            parallel_thread_init (src/transformer/transformer.cpp:23-42,
 *                                            :209-384): per-worker weight slices + KV cache + scratch
 *   fl_forward                            <- ParallelTransformer::forward(span<const int> tokens, int pos,
 *                                            Tensor& logits)  (src/transformer/transformer.h:99,
 *                                            transformer.cpp:105-161) and the six task bodies :386-505
 *   fl_generate_greedy / fl_forward_batch <- the token loop of ParallelTransformer::generate
 *                                            (transformer.cpp:76-103) with Sampler::sample_argmax
 *                                            (src/transformer/sampler.cpp:36-46) kept on the device
 *   fl_op_*                               <- the free-function operator headers the task bodies call:
 *                                            src/blas/quant_operators.h:37-82, src/blas/tf_operators.h:18-50,
 *                                            src/platforms/arch/simd.h:13-60
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * fl_status and never throws; all device memory is allocated in fl_create/fl_upload/fl_finalize (none
 * per token); one caller thread per engine (like the reference, forward() is not re-entrant).
 * There is NO CPU fallback: without a CUDA device every entry point fails with FL_ERR_CUDA.
 *
 * Results: logits are bit-identical to the reference's CPU forward() built with AVX2+FMA
 * (oracle/build_ref.sh), for INT8 and INT16 — see DESIGN.md "Exactness".
 */
#ifndef FASTLLAMA_B200_H
#define FASTLLAMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fl_engine fl_engine;   /* opaque; one per GPU */

typedef enum {
    FL_OK = 0,
    FL_ERR_INVALID = -1,      /* bad argument / shape / state */
    FL_ERR_CUDA = -2,         /* CUDA runtime error (fl_last_error has the string) */
    FL_ERR_OOM = -3,
    FL_ERR_UNSUPPORTED = -4,
    FL_ERR_NCCL = -5
} fl_status;

/* QuantType numbering of the reference (src/blas/quant_operators.h:17-24) */
typedef enum { FL_Q_NONE = 0, FL_Q_INT16 = 1, FL_Q_INT8 = 2 } fl_quant_type;

/* One entry per tensor of TransformerWeights (src/model_loaders/model_loader.h:72-92) */
typedef enum {
    FL_T_TOK_EMB = 0,   /* [vocab][dim]  fp32 (scales == NULL) or quantised (dequantised once at upload) */
    FL_T_ATT_NORM,      /* [dim] fp32, per layer */
    FL_T_WQ,            /* [dim][dim] */
    FL_T_WK,            /* [kv_dim][dim] */
    FL_T_WV,            /* [kv_dim][dim] */
    FL_T_WO,            /* [dim][dim] */
    FL_T_FFN_NORM,      /* [dim] fp32, per layer */
    FL_T_W1,            /* [hidden][dim] */
    FL_T_W2,            /* [dim][hidden] */
    FL_T_W3,            /* [hidden][dim] */
    FL_T_OUT_NORM,      /* [dim] fp32 */
    FL_T_CLS,           /* [vocab][dim] */
    FL_T__COUNT
} fl_tensor_kind;

/* TransformerConfig (src/model_loaders/model_loader.h:47-70) + engine sizing */
typedef struct {
    int32_t dim, hidden_dim, n_layers, n_heads, n_kv_heads, head_size, vocab_size;
    int32_t max_seq_len;     /* KV capacity per sequence (the reference forces 1024, transformer.cpp:32) */
    int32_t quant_type;      /* FL_Q_INT8 | FL_Q_INT16: weights AND activations */
    int32_t group_size;      /* 64 (.flm, llama2.c) or 32 (GGUF Q8_0, INT8 only) */
    int32_t max_seqs;        /* independent KV slots (request batch sharded onto this GPU); >= 1 */
    int32_t flags;           /* FL_FLAG_* */
} fl_config;

#define FL_FLAG_NO_GRAPH   1   /* launch kernels directly instead of replaying the captured CUDA graph */
#define FL_FLAG_NO_PDL     2   /* rows path: plain stream order instead of programmatic dependent launch between its kernels */
#define FL_FLAG_PROFILE    8   /* persistent kernel records per-CTA time per category (fl_profile_read) */
#define FL_FLAG_NO_MEGAKERNEL 4 /* run the step as separate kernels (one per phase) instead of the persistent decode kernel */
#define FL_FLAG_RELAXED    32  /* MEASUREMENT ONLY (7B-shaped INT8 decode): rmsnorm's sum of squares and softmax's sum as tree reductions instead
                                  of the reference's serial FP32 chains: logits agree to ~1e-3 relative, they are NOT bit-identical */
#define FL_FLAG_NO_TC      16  /* never use the tensor-core rows path (tcgen05 GEMM): prompts and sequence batches go token by token */

/* ---- lifecycle ---------------------------------------------------------------------------- */
int  fl_create(const fl_config* cfg, int device, fl_engine** out);
void fl_destroy(fl_engine* e);
const char* fl_last_error(const fl_engine* e);     /* e may be NULL: last error of an fl_op_* / fl_create call */

/* Copy one tensor in the reference's layout (row-major payload + fp32 scale per `group_size` consecutive
 * columns, src/components/tensor.h:473-504) to the device; the caller keeps ownership of the host arrays.
 * Quantised projections are re-packed into the streaming layout described in DESIGN.md. */
int  fl_upload(fl_engine* e, int kind, int layer, const void* q, const float* scales, int rows, int cols);
/* Call once after the last fl_upload: checks completeness, builds RoPE tables, captures the CUDA graph. */
int  fl_finalize(fl_engine* e);

/* ---- the hot path -------------------------------------------------------------------------- */
/* forward(): n_tokens tokens of ONE sequence (KV slot `seq_slot`) starting at position `pos`.
 * logits_out (host, vocab floats) and argmax_out (host, 1 int) may each be NULL.
 * Synchronous w.r.t. the caller (host buffers are valid on return). */
int  fl_forward(fl_engine* e, int seq_slot, const int32_t* tokens, int n_tokens, int pos,
                float* logits_out, int32_t* argmax_out);

/* One decode step for n_seqs independent sequences (slots 0..n_seqs-1): tokens[i] at pos[i] -> argmax_out[i].
 * (new; the reference has one KV cache and one position)  Up to 16 sequences share ONE persistent launch:
 * every phase is walked once per sequence, each with its own cache slot and position; per-sequence results are
 * bit-identical to fl_forward on that sequence alone.  Logits are not kept in this mode (token ids only). */
int  fl_forward_batch(fl_engine* e, int n_seqs, const int32_t* tokens, const int32_t* pos, int32_t* argmax_out);

/* n_steps greedy decode steps of sequences 0..n_seqs-1 (n_seqs <= 64) from their device-resident states, asynchronously on
 * the engine stream: per step ONE weight pass for all sequences on the tensor cores (tcgen05 group-scaled INT8 GEMM), tokens fed
 * back on the device, the step replayed as one CUDA graph.  INT8 engines only.  Results per sequence are bit-identical to
 * fl_forward on that sequence alone. */
int  fl_decode_batch_async(fl_engine* e, int n_seqs, int n_steps);
/* the first n tokens sampled for seq_slot since its last prefill (waits for the engine stream) */
int  fl_read_out_tokens(fl_engine* e, int seq_slot, int n, int32_t* out);

/* Greedy generation with the token fed back on the device (no host round trip per token):
 * prefill `prompt`, then up to max_new tokens; stops after token id 0 (transformer.cpp:93).
 * out_tokens (host) receives every sampled token, *n_out their count. */
int  fl_generate_greedy(fl_engine* e, int seq_slot, const int32_t* prompt, int n_prompt, int max_new,
                        int32_t* out_tokens, int* n_out);

/* generate() with the reference's sampler (src/transformer/transformer.cpp:76-103 + src/transformer/sampler.cpp:113-136):
 * temperature 0 is fl_generate_greedy (device-resident); otherwise each step reads the logits back and samples on the
 * host with the reference's arithmetic (temperature divide, its softmax, xorshift* coin, multinomial or top-p with
 * qsort order), seeded like Sampler::build(vocab, seed).  Same stop rule and clipping as fl_generate_greedy. */
int  fl_generate(fl_engine* e, int seq_slot, const int32_t* prompt, int n_prompt, int max_new, float temperature,
                 float topp, uint64_t seed, int32_t* out_tokens, int* n_out);

/* The sampler on its own (replaces cpuft::Sampler, src/transformer/sampler.h:13-34).  `logits` is a HOST buffer of
 * vocab_size floats and is overwritten with the probabilities, as the reference does. */
typedef struct fl_sampler fl_sampler;
int      fl_sampler_create(int vocab_size, uint64_t seed, fl_sampler** out);
void     fl_sampler_destroy(fl_sampler* s);
int      fl_sampler_sample(fl_sampler* s, float* logits, float temperature, float topp, int32_t* token_out);
uint64_t fl_sampler_state(const fl_sampler* s);

/* Device-resident decode for measurement: runs n_steps greedy steps starting from the current device
 * state of `seq_slot` (set by a previous fl_forward / fl_generate_greedy), asynchronously on the engine
 * stream.  fl_stream() exposes that stream (a cudaStream_t) so the caller can bracket it with events. */
int   fl_decode_async(fl_engine* e, int seq_slot, int n_steps);
void* fl_stream(fl_engine* e);
int   fl_sync(fl_engine* e);
/* device addresses of per-slot state for zero-copy consumers on the engine stream (e.g. an NCCL all-gather of the
 * sampled token): name in {"token","pos","argmax","out_tokens","logits","gathered"}; NULL if unknown. */
void* fl_device_ptr(fl_engine* e, const char* name, int seq_slot);
/* per-CTA SM cycles spent per category by the persistent decode kernel since the last reset:
 * out[cta*32 + k], k = 0 waiting for tagged input words (exchange latency + slowest producer), 1 activation rebuild tail
 * (quantise after the rmsnorm chain), 2 QKV, 3 Wo, 4 W1/W3, 5 W2, 6 classifier (weight-stream drains), 8 rebuild: products
 * + transpose, 9 rebuild: sum-of-squares chain, 10 attention: q/k/v fetch + RoPE + append, 11 QK^T, 12 score exchange,
 * 13 softmax, 15 PV chains, 18 argmax exchange, 19 embedding row, 7 waiting for weight stages (the stream is the limit),
 * 20 waiting for the chain token, 21 chain + hand-off; 22-31 are absolute timestamps of one traced layer (profiles/trace_layer.py).
 * Needs FL_FLAG_PROFILE.  Returns the element count (32 * n_CTAs). */
int  fl_profile_read(fl_engine* e, uint64_t* out, int cap, int reset);
/* number of kernels this engine has launched (graph nodes count once per replay) */
int64_t fl_launch_count(const fl_engine* e);
/* algorithmic bytes one decode step at context length `ctx` must read/write (weights + scales + KV), SURVEY §8d */
int64_t fl_step_bytes(const fl_engine* e, int ctx);
/* debug tap: copy an intermediate activation of the last step to the host.
 * name in {"x1","qkv","attn","hd","logits"}; returns element count or a negative fl_status. */
int  fl_tap(fl_engine* e, const char* name, float* out, int cap);

/* ---- multi-GPU (request batch sharded, weights replicated) ---------------------- */
/* Bind an NCCL communicator created by the host (ncclComm_t passed as void*), or NULL for world == 1. */
int  fl_set_comm(fl_engine* e, void* nccl_comm, int rank, int world);
/* all-gather of n_local sampled tokens per rank (ncclAllGather on the engine stream; a plain copy if world == 1).
 * local (host, n_local ints) -> all (host, world * n_local ints), synchronous; or local == NULL: gather the tokens just sampled
 * on the device for slots 0 .. n_local-1 without any host round trip - asynchronous when `all` is NULL too, the result stays
 * on the device (fl_device_ptr(e, "gathered", 0)). */
int  fl_allgather_tokens(fl_engine* e, const int32_t* local, int n_local, int32_t* all);

/* ---- per-operator entry points (known-answer / parity tests; host buffers in, host buffers out) -- */
/* quant::quantize (quant_operators.cpp:26-47) */
int  fl_op_quantize(int quant_type, int group_size, const float* x, int n, void* q_out, float* scales_out);
/* quant::matmul (quant_operators.cpp:252-284): out[i*m + j] = W[j,:] . X[i,:] */
int  fl_op_matmul_q(int quant_type, int group_size, const void* w, const float* w_scales, int m, int n,
                    const void* x, const float* x_scales, int rows_x, float* out);
/* quant::matmul (quant_operators.cpp:252-284) for up to 64 activation rows in ONE weight pass on the tensor cores (INT8;
 * tcgen05.mma.kind::i8 per quantisation group, FP32 group chain in the reference's order: bit-identical to fl_op_matmul_q).
 * w3 / w3_scales non-NULL: the fused W1/W3 pass of execute_ffn13 (transformer.cpp:468-483), out = swiglu(W X, W3 X).
 * variant: 0 (debugging knobs of the kernel otherwise). */
int  fl_op_matmul_q_tc(int group_size, const void* w, const float* w_scales, const void* w3, const float* w3_scales, int m, int n,
                       const void* x, const float* x_scales, int rows_x, float* out, int variant);
/* simd::rmsnorm (x86_simd.cpp:1754-1764) */
int  fl_op_rmsnorm(const float* x, const float* w, int n, float* out);
/* rope_v2 (tf_operators.cpp:355-402) on one head vector */
int  fl_op_rope(const float* x, int n_dims, int pos, float* out);
/* softmax_sisd (tf_operators.cpp:176-186) */
int  fl_op_softmax(const float* x, int n, float* out);
/* simd::swiglu (x86_simd.cpp:1766-1770): out = silu(a) * b */
int  fl_op_swiglu(const float* a, const float* b, int n, float* out);
/* glibc expf as restated for the device */
int  fl_op_expf(const float* x, int n, float* out);
/* execute_attn for one new token (transformer.cpp:397-455): qkv = [q | k | v] (fp32, dim + 2*kv_dim),
 * k_cache/v_cache = [n_kv_heads][pos][head_size] rows already in the cache (post-RoPE, natural order).
 * Writes out[dim] and, if non-NULL, the appended rows k_new/v_new [n_kv_heads][head_size]. */
int  fl_op_attn_decode(int n_heads, int n_kv_heads, int head_size, int pos, const float* qkv,
                       const float* k_cache, const float* v_cache, float* out, float* k_new, float* v_new);
/* Sampler::sample_argmax (sampler.cpp:36-46) */
int  fl_op_argmax(const float* logits, int n, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* FASTLLAMA_B200_H */
