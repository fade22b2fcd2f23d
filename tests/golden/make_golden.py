"""Generate the committed golden vectors from the REAL reference (oracle/_ref/libref.so, built by oracle/build_ref.sh
from the untouched sources under /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (small, committed):
  tests/golden/ops_golden.npz         outputs of the reference's leaf operators on seeded inputs
  tests/golden/tiny_model_logits.npz  logits of ParallelTransformer::forward on the seeded TINY llama2.c checkpoint
Inputs are re-created from the seeds by the tests (numpy default_rng is platform independent).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_libs import ref, ptr, quantize, Q_INT8, Q_INT16  # noqa: E402
from fixtures import TINY, gen_weights, write_llama2c, write_tokenizer_bin, synthetic_vocab, prompt_tokens  # noqa: E402
import golden_inputs as gi  # noqa: E402


def sampler_golden(R):
    """Sampler::sample (sampler.cpp:113-136) and generate() with sampling (transformer.cpp:76-103) from the real reference."""
    import ctypes as C
    out = {}
    for name, logits in gi.sampler_inputs():
        for ci, (temp, topp, seed) in enumerate(gi.SAMPLER_CASES):
            rng = C.c_uint64(seed)
            toks = []
            for d in range(gi.SAMPLER_DRAWS):
                buf = logits.copy()
                toks.append(R.ref_sampler_sample(ptr(buf), buf.size, temp, topp, C.byref(rng)))
                if d == 0 and temp != 0.0:                     # the probabilities the reference's softmax left behind:
                    if ci == 0 and buf.size <= 8000:          # in full for the small inputs, as a bit checksum otherwise
                        out[f"probs_{name}_{ci}"] = buf
                    out[f"probsum_{name}_{ci}"] = gi.bits_checksum(buf)
            out[f"tokens_{name}_{ci}"] = np.array(toks, np.int32)
            out[f"state_{name}_{ci}"] = np.uint64(rng.value)
    spec = TINY
    w = gen_weights(spec, 1)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/model.bin", spec, w)
        write_tokenizer_bin(d + "/tok.bin", synthetic_vocab(spec.vocab_size))
        h = R.ref_model_load((d + "/model.bin").encode(), (d + "/tok.bin").encode(), 3, Q_INT8, 2, 64, 0)
        assert h
        prompt = prompt_tokens(spec, 6, seed=3)
        for gi_, (temp, topp, seed) in enumerate([(0.9, 0.9, 1234), (1.0, 1.0, 99), (0.6, 0.8, 5)]):
            gen = np.zeros(64, np.int32)
            n = R.ref_generate(h, ptr(prompt), prompt.size, 40, temp, topp, seed, ptr(gen), 64)
            out[f"generate_{gi_}"] = gen[:n].copy()
            out[f"generate_{gi_}_args"] = np.array([temp, topp, seed], np.float64)
        R.ref_model_free(h)
    out["prompt"] = prompt
    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)


def main():
    R = ref()
    assert R is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    out = {}
    for name, x in gi.quantize_inputs():
        for qt in (Q_INT8, Q_INT16):
            q, s = quantize(R.ref_quantize, qt, x, 64)
            out[f"quant_{name}_{qt}_q"] = q
            out[f"quant_{name}_{qt}_s"] = s
    for name, qt, gs, qw, sw, qx, sx in gi.matmul_inputs():
        m, n = qw.shape
        o = np.empty((qx.shape[0], m), np.float32)
        R.ref_matmul(qt, ptr(o), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, qx.shape[0], gs)
        out[f"matmul_{name}"] = o
    for name, x, w in gi.rmsnorm_inputs():
        o = np.empty_like(x)
        R.ref_rmsnorm(ptr(o), ptr(x), ptr(w), x.size)
        out[f"rmsnorm_{name}"] = o
    for name, x, pos in gi.rope_inputs():
        o = np.empty_like(x)
        R.ref_rope_v2(ptr(o), ptr(x), x.size, 1024, pos)
        out[f"rope_{name}"] = o
    for name, a, b in gi.dot_inputs():
        out[f"dot_{name}"] = np.float32(R.ref_dot_f32(ptr(a), ptr(b), a.size))
    for name, x in gi.softmax_inputs():
        o = x.copy()
        R.ref_softmax_sisd(ptr(o), o.size)
        out[f"softmax_{name}"] = o
    for name, V, w in gi.wsum_inputs():
        o = np.empty((w.shape[0], V.shape[1]), np.float32)
        R.ref_weighted_sum(ptr(o), ptr(V), ptr(w), V.shape[0], V.shape[1], w.shape[0], 1e-15)
        out[f"wsum_{name}"] = o
    for name, a, b in gi.swiglu_inputs():
        o = a.copy()
        R.ref_swiglu(ptr(o), ptr(b), o.size)
        out[f"swiglu_{name}"] = o
    np.savez_compressed(os.path.join(HERE, "ops_golden.npz"), **out)

    # whole model through the reference's own llama2.c loader + forward()
    seed = 1
    spec = TINY
    w = gen_weights(spec, seed)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/model.bin", spec, w)
        write_tokenizer_bin(d + "/tok.bin", synthetic_vocab(spec.vocab_size))
        h = R.ref_model_load((d + "/model.bin").encode(), (d + "/tok.bin").encode(), 3, Q_INT8, 2, 64, 0)
        assert h
        prompt = prompt_tokens(spec, 6, seed=3)
        logits = np.empty(spec.vocab_size, np.float32)
        R.ref_forward(h, ptr(prompt), prompt.size, 0, ptr(logits))
        prefill = logits.copy()
        toks, dl = [], []
        pos = prompt.size
        for _ in range(12):
            t = np.array([int(np.argmax(logits))], np.int32)
            R.ref_forward(h, ptr(t), 1, pos, ptr(logits))
            toks.append(int(t[0])); dl.append(logits.copy()); pos += 1
        gen = np.zeros(64, np.int32)
        n = R.ref_generate_greedy(h, ptr(prompt), prompt.size, 20, ptr(gen), 64)
        R.ref_model_free(h)
    np.savez_compressed(os.path.join(HERE, "tiny_model_logits.npz"), seed=seed, prompt=prompt, prefill_logits=prefill,
                        decode_tokens=np.array(toks, np.int32), decode_logits=np.stack(dl), generate_tokens=gen[:n])
    sampler_golden(R)
    print("golden vectors written:", len(out), "op arrays;", "generate ->", gen[:n].tolist())


if __name__ == "__main__":
    main()
