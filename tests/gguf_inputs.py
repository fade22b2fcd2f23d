"""GGUF test files written with the `gguf` library (llama.cpp's own writer), using only the metadata keys the reference
accepts (src/model_loaders/gguf_loader.cpp:244-322).  Shared by tests/test_gguf.py and tests/golden/make_gguf_golden.py."""
import numpy as np

import flm_inputs as fi


def write_gguf(path, spec, w, q8_0, extra_key=False, vocab=None):
    import gguf
    wr = gguf.GGUFWriter(str(path), "llama")
    wr.add_name("tiny")
    wr.add_file_type(7 if q8_0 else 0)
    wr.add_context_length(spec.max_seq_len)
    wr.add_embedding_length(spec.dim)
    wr.add_block_count(spec.n_layers)
    wr.add_feed_forward_length(spec.hidden_dim)
    wr.add_head_count(spec.n_heads)
    wr.add_head_count_kv(spec.n_kv_heads)
    wr.add_rope_dimension_count(spec.head_size)
    wr.add_layer_norm_rms_eps(1e-5)
    v = vocab or fi.micro_vocab(spec.vocab_size)
    wr.add_tokenizer_model("llama")
    wr.add_token_list(v["texts"])
    wr.add_token_scores(v["scores"])
    wr.add_token_types(v["types"])
    wr.add_bos_token_id(1)
    wr.add_eos_token_id(2)
    if extra_key:
        wr.add_uint32("general.quantization_version", 2)       # written by current llama.cpp, rejected by the reference

    def put(name, a):
        a = np.ascontiguousarray(a, np.float32)
        if q8_0 and a.ndim == 2:
            wr.add_tensor(name, gguf.quants.quantize(a, gguf.GGMLQuantizationType.Q8_0),
                          raw_dtype=gguf.GGMLQuantizationType.Q8_0)
        else:
            wr.add_tensor(name, a)

    put("token_embd.weight", w["tok_emb"])
    for l in range(spec.n_layers):
        for short, key in (("attn_q", "wq"), ("attn_k", "wk"), ("attn_v", "wv"), ("attn_output", "wo"), ("ffn_gate", "w1"),
                           ("ffn_down", "w2"), ("ffn_up", "w3"), ("attn_norm", "att_norm"), ("ffn_norm", "ffn_norm")):
            put(f"blk.{l}.{short}.weight", w[key][l])
    put("output_norm.weight", w["out_norm"])
    put("output.weight", w["cls"])
    wr.write_header_to_file()
    wr.write_kv_data_to_file()
    wr.write_tensors_to_file()
    wr.close()
