/* oracle/ref_port.h — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's hot path).
 * See ref_port.c for the file:line each function follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this library; the product never links it. */
#ifndef FASTLLAMA_ORACLE_REF_PORT_H
#define FASTLLAMA_ORACLE_REF_PORT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { PORT_Q_NONE = 0, PORT_Q_INT16 = 1, PORT_Q_INT8 = 2 };   /* quant_operators.h:17-24 */

/* tensor kinds (same numbering as include/fastllama_b200.h fl_tensor_kind) */
enum {
    PORT_T_TOK_EMB = 0, PORT_T_ATT_NORM, PORT_T_WQ, PORT_T_WK, PORT_T_WV, PORT_T_WO,
    PORT_T_FFN_NORM, PORT_T_W1, PORT_T_W2, PORT_T_W3, PORT_T_OUT_NORM, PORT_T_CLS, PORT_T__COUNT
};

typedef struct {
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, head_size, vocab_size, max_seq_len;
    int qtype;  /* PORT_Q_INT16 | PORT_Q_INT8 */
    int group;  /* 64 (.flm / llama2.c) or 32 (GGUF Q8_0) */
} port_config;

/* leaf operators */
void  port_quantize(int qt, void* qx, float* qs, const float* x, size_t n, int gs);
void  port_dequantize(int qt, float* out, const void* qx, const float* qs, size_t n, int gs);
void  port_matmul(int qt, float* out, const void* w, const float* ws, const void* x, const float* xs,
                  int m, int n, int rows_x, int gs);
float port_square_sum(const float* x, size_t n);
void  port_rmsnorm(float* o, const float* x, const float* w, size_t n);
void  port_rope_v2(float* o, const float* x, int n_dims, int pos);
void  port_rope_table(float* cos_sin, int n_dims, int pos);   /* [n_dims/2][2] = (cos, sin) */
float port_dot_f32(const float* a, const float* b, size_t n);
void  port_softmax_sisd(float* x, int n);
void  port_weighted_sum(float* out, const float* matrix, const float* weights, int m, int n, int bs, float min_w);
void  port_swiglu(float* xo, const float* xr, size_t n);
float port_expf_emul(float x);      /* the double-precision algorithm the CUDA kernels run; must equal libm expf */
int   port_argmax(const float* logits, int n);
void  port_softmax(float* x, size_t n);                /* the sampler's softmax, tf_operators.cpp:188-209 */
uint32_t port_random_u32(uint64_t* state);
int   port_sample(float* logits, int n, float temperature, float topp, uint64_t* rng);   /* rewrites logits */

/* whole model */
typedef struct port_model port_model;
port_model* port_model_create(const port_config* cfg);
void        port_model_free(port_model* m);
/* q: int8/int16 payload in reference row-major layout (or fp32 when scales == NULL); copied */
int         port_model_set_tensor(port_model* m, int kind, int layer, const void* q, const float* scales,
                                  int rows, int cols);
void        port_model_reset(port_model* m);
/* ParallelTransformer::forward (transformer.cpp:105-161): n tokens of one sequence starting at pos */
int         port_forward(port_model* m, const int* tokens, int n, int pos, float* logits_out);
/* debug taps: copies of intermediate activations of the LAST forward (last row), for layer-level parity */
const float* port_tap(port_model* m, const char* name, int layer, int* n_out);

#ifdef __cplusplus
}
#endif
#endif
