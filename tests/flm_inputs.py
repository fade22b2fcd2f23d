"""Seeded inputs shared by tests/test_flm.py and tests/golden/make_flm_golden.py."""
import numpy as np

from fixtures import ModelSpec

# small enough that the golden file written by the reference converter stays ~100 KB
MICRO = ModelSpec(dim=64, hidden_dim=128, n_layers=2, n_heads=1, n_kv_heads=1, vocab_size=64)


def config_of(spec, quant_type, group_size, name):
    return dict(name=name, model_type=1, act_type=2, quant_type=quant_type, vocab_size=spec.vocab_size, dim=spec.dim,
                hidden_dim=spec.hidden_dim, n_heads=spec.n_heads, n_kv_heads=spec.n_kv_heads, n_layers=spec.n_layers,
                max_length=spec.max_seq_len, bos_token_id=1, eos_token_id=2, pad_token_id=0, rms_norm_eps=1e-5,
                rope_theta=10000.0, quant_group_size=group_size)


def micro_vocab(n):
    texts = ["<unk>", "<s>", "</s>"] + [("▁w%d" % i) if i % 3 == 0 else ("p%d" % i) for i in range(n - 3)]
    return dict(vocab_type=2, texts=texts, scores=[-float(i) for i in range(n)],
                types=[2, 2, 2] + [1] * (n - 3), special=dict(bos=1, eos=2))


def hf_tensors(spec, w):
    """(checkpoint name, tensor_type, layer, float32 array) in the order our writer emits blocks"""
    yield "model.embed_tokens.weight", 1, 0, w["tok_emb"]
    for l in range(spec.n_layers):
        p = f"model.layers.{l}."
        yield p + "self_attn.q_proj.weight", 18, l, w["wq"][l]
        yield p + "self_attn.k_proj.weight", 19, l, w["wk"][l]
        yield p + "self_attn.v_proj.weight", 20, l, w["wv"][l]
        yield p + "self_attn.o_proj.weight", 21, l, w["wo"][l]
        yield p + "mlp.gate_proj.weight", 22, l, w["w1"][l]
        yield p + "mlp.up_proj.weight", 23, l, w["w3"][l]
        yield p + "mlp.down_proj.weight", 24, l, w["w2"][l]
        yield p + "input_layernorm.weight", 17, l, w["att_norm"][l]
        yield p + "post_attention_layernorm.weight", 25, l, w["ffn_norm"][l]
    yield "model.norm.weight", 2, 0, w["out_norm"]
    yield "lm_head.weight", 3, 0, w["cls"]


def quantized_tensors(fl, spec, w, quant_type, group_size):
    """{(engine kind, layer): (payload, scales)} the way the converter stores them: fp32 embedding and norms, quantised matrices"""
    out = {}
    for _, tt, layer, arr in hf_tensors(spec, w):
        kind = fl.flm.TENSOR_TYPES[tt][0]
        arr = np.ascontiguousarray(arr, np.float32)
        if tt != 1 and arr.ndim > 1:
            out[(kind, layer)] = fl.loaders.quantize_rows(arr, quant_type, group_size)
        else:
            out[(kind, layer)] = (arr, None)
    return out
