"""A synthetic sentencepiece-like vocabulary with real merges, and the texts the tokenizer tests encode."""
import numpy as np

from fixtures import ModelSpec

CONN = "▁"


def merge_vocab(n=640, seed=3):
    """<unk> <s> </s>, 256 byte pieces, single characters, then pieces built by joining two existing ones (so that the merge
    loop has work to do), with distinct-ish scores and some exact score ties.  No piece length is a multiple of 8 bytes (the
    reference's tokenizer.bin reader drops the terminator of those, tokenizer.cpp:228)."""
    r = np.random.default_rng(seed)
    texts = ["<unk>", "<s>", "</s>"] + ["<0x%02X>" % i for i in range(256)]
    scores = [0.0, 0.0, 0.0] + [0.0] * 256
    singles = list("abcdefghijklmnopqrstuvwxyzABCDE.,!?0123456789") + [CONN, "é", "ß", "中", "文", "😀"]
    for s in singles:
        texts.append(s)
        scores.append(-1000.0)
    seen = set(texts)
    pool = list(singles)
    while len(texts) < n:
        a, b = pool[int(r.integers(len(pool)))], pool[int(r.integers(len(pool)))]
        t = a + b
        if t in seen or len(t.encode("utf-8")) % 8 == 0 or len(t.encode("utf-8")) > 14 or (b.startswith(CONN) and len(b) > 1):
            continue
        seen.add(t)
        texts.append(t)
        pool.append(t)
        scores.append(-float(int(r.integers(1, 200))))              # small integer range => plenty of ties
    types = [2, 3, 3] + [6] * 256 + [1] * (n - 259)
    return dict(vocab_type=2, texts=texts, scores=scores, types=types, special=dict(bos=1, eos=2))


def spec_for(vocab):
    return ModelSpec(dim=64, hidden_dim=128, n_layers=1, n_heads=1, n_kv_heads=1, vocab_size=len(vocab["texts"]))


def sample_texts(vocab, seed=4):
    r = np.random.default_rng(seed)
    pieces = [t for t in vocab["texts"][259:]]
    out = ["a", " ", "  ", "hello world", " leading space", "trailing space ", "two  spaces", "abc.def,ghi!", "é中文😀ß",
           "unknown: ñ 世 \U0001F680 ~ #", "tab\there", "new\nline", "x" * 40, "A1b2C3d4E5", "€uro"]
    for _ in range(60):
        k = int(r.integers(1, 12))
        s = "".join(pieces[int(r.integers(len(pieces)))] for _ in range(k)).replace(CONN, " ")
        out.append(s)
    return out
