"""Synthetic, seeded LLaMA2-shaped checkpoints for parity tests and the bench (no real model files exist offline).

Recipe follows SURVEY.md §8(c): N(0, dim^-1/2) projection weights, N(0, 0.05) embeddings, 1 + N(0, 0.1) norm gains,
numpy default_rng(seed).  Writers produce the on-disk formats the reference loads:
  * llama2.c legacy .bin + tokenizer.bin  (src/model_loaders/llama2c_loader.cpp:42-199, tokenizer.cpp:162-231)
  * .flm                                  (tools/convert_flm.py:465-748; flm_loader.cpp)   -> tests/flm_format.py
"""
import dataclasses
import struct
import numpy as np

from oracle_libs import (Q_INT8, Q_INT16, NP_T, port_quantize,
                         T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS)


@dataclasses.dataclass(frozen=True)
class ModelSpec:
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    max_seq_len: int = 1024
    shared_classifier: bool = False

    @property
    def head_size(self):
        return self.dim // self.n_heads

    @property
    def kv_dim(self):
        return self.head_size * self.n_kv_heads


TINY = ModelSpec(dim=512, hidden_dim=704, n_layers=2, n_heads=4, n_kv_heads=4, vocab_size=1000)       # head 128
TINY64 = ModelSpec(dim=512, hidden_dim=1024, n_layers=3, n_heads=8, n_kv_heads=8, vocab_size=1024)    # head 64
STORIES110M = ModelSpec(dim=768, hidden_dim=2048, n_layers=12, n_heads=12, n_kv_heads=12, vocab_size=32000,
                        shared_classifier=True)
LLAMA2_7B = ModelSpec(dim=4096, hidden_dim=11008, n_layers=32, n_heads=32, n_kv_heads=32, vocab_size=32000)
LLAMA2_13B = ModelSpec(dim=5120, hidden_dim=13824, n_layers=40, n_heads=40, n_kv_heads=40, vocab_size=32000)


def gen_weights(spec: ModelSpec, seed: int = 0):
    """float32 weights in the reference's row-major [out_rows][in_cols] convention."""
    rng = np.random.default_rng(seed)
    d, h, L, kv = spec.dim, spec.hidden_dim, spec.n_layers, spec.kv_dim
    sd = d ** -0.5

    def mat(*shape, s):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(s)).astype(np.float32)

    w = {
        "tok_emb": mat(spec.vocab_size, d, s=0.05),
        "att_norm": (1 + 0.1 * rng.standard_normal((L, d))).astype(np.float32),
        "wq": mat(L, d, d, s=sd), "wk": mat(L, kv, d, s=sd), "wv": mat(L, kv, d, s=sd), "wo": mat(L, d, d, s=sd),
        "ffn_norm": (1 + 0.1 * rng.standard_normal((L, d))).astype(np.float32),
        "w1": mat(L, h, d, s=sd), "w2": mat(L, d, h, s=h ** -0.5), "w3": mat(L, h, d, s=sd),
        "out_norm": (1 + 0.1 * rng.standard_normal(d)).astype(np.float32),
    }
    w["cls"] = w["tok_emb"] if spec.shared_classifier else mat(spec.vocab_size, d, s=sd)
    return w


def write_llama2c(path, spec: ModelSpec, w):
    """llama2.c legacy checkpoint: 7 int32 header then fp32 tensors (llama2c_loader.cpp:126-194).
    vocab_size > 0 means the classifier shares the embedding table (:71)."""
    hs = spec.head_size
    with open(path, "wb") as f:
        vocab = spec.vocab_size if spec.shared_classifier else -spec.vocab_size
        f.write(struct.pack("<7i", spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads,
                            vocab, spec.max_seq_len))
        for k in ("tok_emb", "att_norm", "wq", "wk", "wv", "wo", "ffn_norm", "w1", "w2", "w3", "out_norm"):
            f.write(np.ascontiguousarray(w[k], dtype=np.float32).tobytes())
        n = hs * spec.max_seq_len // 2
        f.write(np.zeros(n, np.float32).tobytes())   # freq_cis_real (parsed, unused: rope_v2 recomputes)
        f.write(np.zeros(n, np.float32).tobytes())   # freq_cis_imag
        if not spec.shared_classifier:
            f.write(np.ascontiguousarray(w["cls"], dtype=np.float32).tobytes())


def synthetic_vocab(n):
    """llama-like vocab: <unk>, <s>, </s>, 256 byte tokens, then short unique word pieces (length never a
    multiple of 8: the reference's tokenizer.bin reader loses the NUL terminator for those, tokenizer.cpp:228)."""
    toks = [b"<unk>", b"<s>", b"</s>"] + [b"<0x%02X>" % i for i in range(256)]
    i = 0
    while len(toks) < n:
        toks.append(("▁w%d" % i).encode("utf-8") if i % 3 == 0 else ("p%d" % i).encode())
        i += 1
    out = []
    for t in toks[:n]:
        if len(t) % 8 == 0:
            t = t + b"_"
        out.append(t)
    return out


def write_tokenizer_bin(path, vocab):
    with open(path, "wb") as f:
        f.write(struct.pack("<i", max(len(t) for t in vocab)))
        for i, t in enumerate(vocab):
            f.write(struct.pack("<fi", -float(i), len(t)))
            f.write(t)


def quantize_model(spec: ModelSpec, w, qt=Q_INT8, gs=64, quantize_embedding=True):
    """Quantise exactly as the reference does at load (Tensor::quantize per tensor, groups of gs along columns).
    Returns {(kind, layer): (payload ndarray, scales ndarray | None)} ready for fl_upload / port_model_set_tensor."""
    out = {}
    L = spec.n_layers

    def q(a):
        return port_quantize(qt, a, gs)

    if quantize_embedding:
        out[(T_TOK_EMB, 0)] = q(w["tok_emb"])       # llama2c_loader.cpp:83,126
    else:
        out[(T_TOK_EMB, 0)] = (np.ascontiguousarray(w["tok_emb"]), None)   # .flm keeps fp32 embeddings
    for l in range(L):
        out[(T_ATT_NORM, l)] = (np.ascontiguousarray(w["att_norm"][l]), None)
        out[(T_FFN_NORM, l)] = (np.ascontiguousarray(w["ffn_norm"][l]), None)
        for kind, name in ((T_WQ, "wq"), (T_WK, "wk"), (T_WV, "wv"), (T_WO, "wo"), (T_W1, "w1"), (T_W2, "w2"), (T_W3, "w3")):
            out[(kind, l)] = q(w[name][l])
    out[(T_OUT_NORM, 0)] = (np.ascontiguousarray(w["out_norm"]), None)
    if spec.shared_classifier and quantize_embedding:
        out[(T_CLS, 0)] = out[(T_TOK_EMB, 0)]
    else:
        out[(T_CLS, 0)] = q(w["cls"])
    return out


def prompt_tokens(spec: ModelSpec, n, seed=0):
    """BOS=1 followed by n-1 ids drawn from [3, vocab) (bypasses the tokenizer)."""
    rng = np.random.default_rng(seed)
    return np.concatenate([[1], rng.integers(3, spec.vocab_size, n - 1)]).astype(np.int32)
