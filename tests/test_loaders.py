"""llama2.c checkpoint path (BASELINE.json configs[0]): the numpy load-time quantiser against the oracle's, and - on the GPU -
the stories110M-shaped checkpoint through the REAL reference (oracle/_ref/libref.so, its own llama2.c loader + quantiser +
forward) against our loader + engine: 200 greedy tokens and the final logits must be identical."""
import os
import tempfile

import numpy as np
import pytest

from oracle_libs import port_quantize, ref, ptr, bits, Q_INT8, Q_INT16
from fixtures import STORIES110M, TINY, gen_weights, write_llama2c, write_tokenizer_bin, synthetic_vocab, prompt_tokens


@pytest.mark.parametrize("qt,gs", [(Q_INT8, 64), (Q_INT8, 32), (Q_INT16, 64)])
def test_numpy_quantiser_equals_oracle(fl, qt, gs):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((37, 512)) * rng.lognormal(0, 2, (37, 1))).astype(np.float32)
    x[3, 64:128] = 0.0                                   # an all-zero group: 0/0
    x[5, 7] = np.float32(1e-42)                          # denormal
    q, s = fl.loaders.quantize_rows(x, qt, gs)
    qo, so = port_quantize(qt, x, gs)
    assert np.array_equal(q, qo.reshape(q.shape))
    assert np.array_equal(bits(s), bits(so.reshape(s.shape)))


def test_llama2c_reader_roundtrip(fl):
    spec = TINY
    w = gen_weights(spec, seed=5)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/m.bin", spec, w)
        cfg, r = fl.loaders.read_llama2c(d + "/m.bin")
    assert (cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["vocab_size"]) == (spec.dim, spec.hidden_dim, spec.n_layers, spec.vocab_size)
    for k in ("tok_emb", "wq", "w2", "out_norm", "cls"):
        assert np.array_equal(r[k], w[k]), k


@pytest.mark.gpu
def test_config0_stories110m_llama2c_matches_the_reference(fl):
    R = ref()
    if R is None:
        pytest.skip("oracle/_ref/libref.so not built")
    spec = STORIES110M
    w = gen_weights(spec, seed=7)
    n_new = 200
    prompt = prompt_tokens(spec, 8, seed=2)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/m.bin", spec, w)
        write_tokenizer_bin(d + "/t.bin", synthetic_vocab(spec.vocab_size))
        del w
        h = R.ref_model_load((d + "/m.bin").encode(), (d + "/t.bin").encode(), 3, Q_INT8, min(8, os.cpu_count() or 1), 64, 0)
        assert h, "the reference failed to load the llama2.c checkpoint"
        eng, cfg = fl.loaders.engine_from_llama2c(d + "/m.bin")
    logits = np.empty(spec.vocab_size, np.float32)
    R.ref_forward(h, ptr(prompt), prompt.size, 0, ptr(logits))
    want = [int(np.argmax(logits))]
    pos = prompt.size
    for _ in range(n_new):
        if want[-1] == 0:
            break
        t = np.array([want[-1]], np.int32)
        R.ref_forward(h, ptr(t), 1, pos, ptr(logits))
        want.append(int(np.argmax(logits)))
        pos += 1
    R.ref_model_free(h)
    got = eng.generate_greedy(prompt, n_new).tolist()
    assert got == want
    if want[-1] != 0:
        assert np.array_equal(bits(eng.tap("logits")[:spec.vocab_size]), bits(logits))
    eng.close()
