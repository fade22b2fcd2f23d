"""generate(prompt text) end to end (src/transformer/transformer.cpp:54-75): .flm file -> engine + tokenizer -> sampled ids
and text, against goldens produced by the REAL reference from the same file (tests/golden/generate_text_golden.json, written
by this file's CPU test when run with FL_WRITE_GOLDEN=1 in the build container)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import flm_inputs as fi
import tokenizer_inputs as ti
from oracle_libs import ref, Q_INT8
from fixtures import TINY, gen_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generate_text_golden.json")
CASES = [("hello world", 0.0, 0.9, 1), ("abc.def,ghi!", 0.8, 0.9, 77), (" é中 A1b2", 1.0, 1.0, 5)]
MAX_NEW = 24


def write_model(fl, path):
    spec = TINY
    vocab = ti.merge_vocab(spec.vocab_size)
    w = gen_weights(spec, seed=1)
    fl.flm.write_flm(path, fi.config_of(spec, Q_INT8, 64, "tiny"), fi.quantized_tensors(fl, spec, w, Q_INT8, 64), vocab)


def test_reference_goldens_are_current(fl, tmp_path):
    """CPU: (re)generate the goldens from the reference when it is present and check they match the committed file"""
    R = ref()
    if R is None:
        pytest.skip("oracle/_ref not built")
    p = tmp_path / "m.flm"
    write_model(fl, p)
    h = R.ref_model_load(str(p).encode(), b"", 1, Q_INT8, 2, 64, 0)
    assert h
    R.ref_generate_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_float, C.c_float, C.c_uint64, C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    out = {}
    toks = np.zeros(256, np.int32)
    text = C.create_string_buffer(1 << 14)
    T = fl.tokenizer.Tokenizer.from_flm_vocab(fl.flm.read_flm(p, tensors=False)[2])
    for prompt, temp, topp, seed in CASES:
        n = R.ref_generate_text(h, prompt.encode("utf-8"), MAX_NEW, temp, topp, seed, toks.ctypes.data_as(C.c_void_p), toks.size, text, len(text))
        ids = toks[:n].tolist()
        out[prompt] = dict(tokens=ids, text=text.value.decode("utf-8", "replace"))
        # our tokenizer alone reproduces the text from the reference's ids
        pieces, prev = [], -1
        for t in ids:
            pieces.append(T.decode_piece(t, prev)); prev = t
        assert b"".join(pieces).decode("utf-8", "replace") == out[prompt]["text"], prompt
    R.ref_model_free(h)
    if os.environ.get("FL_WRITE_GOLDEN") == "1":
        json.dump(out, open(GOLDEN, "w"), ensure_ascii=True, indent=0, sort_keys=True)
    assert json.load(open(GOLDEN)) == out


@pytest.mark.gpu
def test_generate_text_matches_reference_golden(fl, tmp_path):
    g = json.load(open(GOLDEN))
    p = tmp_path / "m.flm"
    write_model(fl, p)
    eng, cfg, vocab = fl.flm.engine_from_flm(p)
    T = fl.tokenizer.Tokenizer.from_flm_vocab(vocab)
    for prompt, temp, topp, seed in CASES:
        seen = []
        ids, toks, text = fl.generate_text(eng, T, prompt, MAX_NEW, temp, topp, seed, callback=lambda piece, n_in, n_out, ended: seen.append(piece) or True)
        assert toks.tolist() == g[prompt]["tokens"], prompt
        assert text.decode("utf-8", "replace") == g[prompt]["text"], prompt
        assert b"".join(seen) == text and ids[0] == 1
    eng.close()
