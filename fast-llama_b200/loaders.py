"""Model files -> engine.  First format: the llama2.c legacy checkpoint (BASELINE.json configs[0]), read the way the
reference reads it (src/model_loaders/llama2c_loader.cpp:42-199): 7 x int32 header
(dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size [> 0: classifier shared with the embedding], max_seq_len), then
fp32 tensors tok_emb, rms_att[L], wq[L], wk[L], wv[L], wo[L], rms_ffn[L], w1[L], w2[L], w3[L], rms_final, freq_cis_real,
freq_cis_imag, [wcls].  The reference quantises every matrix (the embedding too) at load time with Tensor::quantize
(src/components/tensor.cpp:448-484 -> quant_operators.cpp:26-47); `quantize_rows` is that arithmetic in numpy, bit for bit
(IEEE float32 divisions, truncation, an all-zero group quantises to zeros)."""
import struct

import numpy as np

from .binding import (Engine, Q_INT8, Q_INT16, T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3,
                      T_OUT_NORM, T_CLS)


def quantize_rows(x, quant_type=Q_INT8, group_size=64):
    """[rows][cols] float32 -> (payload int8/int16 [rows][cols], scales float32 [rows][cols / group_size])"""
    x = np.ascontiguousarray(x, np.float32)
    rows, cols = x.shape
    if cols % group_size:
        raise ValueError(f"{cols} columns are not a multiple of the group size {group_size}")
    qf = np.float32(127.0 if quant_type == Q_INT8 else 5792.0)
    g = x.reshape(rows, cols // group_size, group_size)
    s = (np.abs(g).max(axis=2) / qf).astype(np.float32)                  # s = maxabs / QF
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.trunc(g / s[:, :, None])                                  # (T)(x / s): C truncation
    q = np.where(np.isfinite(q), q, 0.0)                                 # 0/0 -> NaN -> cvttss2si -> low bits 0
    dt = np.int8 if quant_type == Q_INT8 else np.int16
    return q.astype(np.int32).astype(dt).reshape(rows, cols), s


def read_llama2c(path):
    """-> (config dict, weights dict of float32 arrays)"""
    with open(path, "rb") as f:
        dim, hidden, L, n_heads, n_kv, vocab, max_seq = struct.unpack("<7i", f.read(28))
        shared = vocab > 0
        vocab = abs(vocab)
        hs = dim // n_heads
        kv = hs * n_kv

        def rd(*shape):
            n = int(np.prod(shape))
            a = np.frombuffer(f.read(4 * n), np.float32)
            if a.size != n:
                raise ValueError(f"{path}: truncated checkpoint")
            return a.reshape(shape)

        w = {"tok_emb": rd(vocab, dim), "att_norm": rd(L, dim), "wq": rd(L, dim, dim), "wk": rd(L, kv, dim), "wv": rd(L, kv, dim),
             "wo": rd(L, dim, dim), "ffn_norm": rd(L, dim), "w1": rd(L, hidden, dim), "w2": rd(L, dim, hidden), "w3": rd(L, hidden, dim),
             "out_norm": rd(dim)}
        rd(max_seq * hs // 2); rd(max_seq * hs // 2)                      # freq_cis_real / imag: parsed, unused (rope_v2 recomputes)
        w["cls"] = w["tok_emb"] if shared else rd(vocab, dim)
    cfg = dict(dim=dim, hidden_dim=hidden, n_layers=L, n_heads=n_heads, n_kv_heads=n_kv, vocab_size=vocab, max_seq_len=max_seq,
               shared_classifier=shared)
    return cfg, w


def engine_from_llama2c(path, quant_type=Q_INT8, group_size=64, max_seq_len=1024, device=0, **engine_kw):
    """Load a llama2.c checkpoint the way the reference's `main -c model.bin -q int8` does and return a finalized Engine."""
    cfg, w = read_llama2c(path)
    eng = Engine(cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["vocab_size"],
                 max_seq_len=max_seq_len, quant_type=quant_type, group_size=group_size, device=device, **engine_kw)
    q = lambda a: quantize_rows(a, quant_type, group_size)
    emb = q(w["tok_emb"])
    eng.upload(T_TOK_EMB, 0, *emb)
    eng.upload(T_OUT_NORM, 0, np.ascontiguousarray(w["out_norm"]), None)
    eng.upload(T_CLS, 0, *(emb if cfg["shared_classifier"] else q(w["cls"])))
    for l in range(cfg["n_layers"]):
        eng.upload(T_ATT_NORM, l, np.ascontiguousarray(w["att_norm"][l]), None)
        eng.upload(T_FFN_NORM, l, np.ascontiguousarray(w["ffn_norm"][l]), None)
        for kind, name in ((T_WQ, "wq"), (T_WK, "wk"), (T_WV, "wv"), (T_WO, "wo"), (T_W1, "w1"), (T_W2, "w2"), (T_W3, "w3")):
            eng.upload(kind, l, *q(w[name][l]))
    eng.finalize()
    return eng, cfg
