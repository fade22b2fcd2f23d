"""CPU, only where oracle/_ref/ was built (the container that has /root/reference; the prebuilt .so also travels to the
GPU box): the plain-C restatement against the REAL reference on fresh seeded inputs, bit for bit, and the reference built
with build.sh's -march=native AVX-512 flags against the -march=haswell build (the AVX-512 kernels are dead code)."""
import ctypes as C
import tempfile

import numpy as np
import pytest

from fixtures import (ModelSpec, TINY, TINY64, gen_weights, quantize_model, write_llama2c, write_tokenizer_bin,
                      synthetic_vocab, prompt_tokens)
from oracle_libs import port, ref, ref_native, ptr, bits, quantize, PortConfig, Q_INT8, Q_INT16

pytestmark = pytest.mark.skipif(ref() is None, reason="oracle/_ref/libref.so not built (needs /root/reference)")


def beq(a, b):
    return np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("qt", [Q_INT8, Q_INT16])
def test_quantize_and_matmul_vs_reference(qt):
    P, R = port(), ref()
    rng = np.random.default_rng(11 + qt)
    for (m, n, w) in ((64, 64, 1), (48, 704, 1), (96, 4096, 3), (16, 11008, 1)):
        W = (rng.standard_normal((m, n)) * 0.04).astype(np.float32)
        X = (rng.standard_normal((w, n)) * rng.choice([0.01, 1.0, 50.0])).astype(np.float32)
        X[0, :64] = 0.0                                    # an all-zero group (0/0 -> NaN -> 0)
        for a in (W, X):
            qp, sp = quantize(P.port_quantize, qt, a, 64)
            qr, sr = quantize(R.ref_quantize, qt, a, 64)
            assert np.array_equal(qp, qr) and beq(sp, sr)
        qw, sw = quantize(R.ref_quantize, qt, W, 64)
        qx, sx = quantize(R.ref_quantize, qt, X, 64)
        op, orf = np.empty((w, m), np.float32), np.empty((w, m), np.float32)
        P.port_matmul(qt, ptr(op), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, w, 64)
        R.ref_matmul(qt, ptr(orf), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, w, 64)
        assert beq(op, orf), (m, n, w)


def test_float_leaf_ops_vs_reference():
    P, R = port(), ref()
    rng = np.random.default_rng(12)
    for n in (64, 512, 768, 4096, 11008):
        x = (rng.standard_normal(n) * 2).astype(np.float32)
        g = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        a, b = np.empty(n, np.float32), np.empty(n, np.float32)
        P.port_rmsnorm(ptr(a), ptr(x), ptr(g), n)
        R.ref_rmsnorm(ptr(b), ptr(x), ptr(g), n)
        assert beq(a, b), n
        assert beq(np.float32(P.port_square_sum(ptr(x), n)), np.float32(R.ref_square_sum(ptr(x), n)))
        a, b = x.copy(), x.copy()
        P.port_swiglu(ptr(a), ptr(g), n)
        R.ref_swiglu(ptr(b), ptr(g), n)
        assert beq(a, b), n
    for n in (1, 3, 8, 31, 32, 33, 200, 1024):
        x = (rng.standard_normal(n) * 5).astype(np.float32)
        a, b = x.copy(), x.copy()
        P.port_softmax_sisd(ptr(a), n)
        R.ref_softmax_sisd(ptr(b), n)
        assert beq(a, b), n
    for hs in (64, 128):
        for pos in (0, 3, 77, 1023):
            x = rng.standard_normal(hs).astype(np.float32)
            a, b = np.empty(hs, np.float32), np.empty(hs, np.float32)
            P.port_rope_v2(ptr(a), ptr(x), hs, pos)
            R.ref_rope_v2(ptr(b), ptr(x), hs, 1024, pos)
            assert beq(a, b), (hs, pos)
        u, v = rng.standard_normal(hs).astype(np.float32), rng.standard_normal(hs).astype(np.float32)
        assert beq(np.float32(P.port_dot_f32(ptr(u), ptr(v), hs)), np.float32(R.ref_dot_f32(ptr(u), ptr(v), hs)))
    V = rng.standard_normal((77, 128)).astype(np.float32)
    w = rng.random((1, 77)).astype(np.float32)
    w[0, 40:] *= 1e-17
    a, b = np.empty((1, 128), np.float32), np.empty((1, 128), np.float32)
    P.port_weighted_sum(ptr(a), ptr(V), ptr(w), 77, 128, 1, 1e-15)
    R.ref_weighted_sum(ptr(b), ptr(V), ptr(w), 77, 128, 1, 1e-15)
    assert beq(a, b)


@pytest.mark.parametrize("spec", [TINY, TINY64, ModelSpec(512, 704, 2, 8, 8, 1000, shared_classifier=True)], ids=["tiny", "tiny64", "shared-cls"])
def test_forward_vs_reference_forward(spec):
    """ParallelTransformer::forward (llama2.c loader, INT8) vs port_forward: prefill + greedy decode, bit-identical logits."""
    P, R = port(), ref()
    w = gen_weights(spec, seed=21)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/m.bin", spec, w)
        write_tokenizer_bin(d + "/t.bin", synthetic_vocab(spec.vocab_size))
        h = R.ref_model_load((d + "/m.bin").encode(), (d + "/t.bin").encode(), 3, Q_INT8, 2, 64, 0)
    assert h
    qm = quantize_model(spec, w, Q_INT8, 64)
    pc = PortConfig(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.head_size, spec.vocab_size, 1024, Q_INT8, 64)
    pm = P.port_model_create(C.byref(pc))
    for (k, l), (q, s) in qm.items():
        P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None, q.shape[0] if q.ndim == 2 else 1, q.shape[-1])
    toks = prompt_tokens(spec, 7, seed=2)
    a, b = np.empty(spec.vocab_size, np.float32), np.empty(spec.vocab_size, np.float32)
    P.port_forward(pm, ptr(toks), toks.size, 0, ptr(a))
    R.ref_forward(h, ptr(toks), toks.size, 0, ptr(b))
    assert beq(a, b)
    pos = toks.size
    for _ in range(16):
        t = np.array([int(np.argmax(b))], np.int32)
        P.port_forward(pm, ptr(t), 1, pos, ptr(a))
        R.ref_forward(h, ptr(t), 1, pos, ptr(b))
        assert beq(a, b), pos
        pos += 1
    P.port_model_free(pm)
    R.ref_model_free(h)


def test_reference_chunked_prefill_equals_token_by_token_oracle():
    """The engine feeds a prompt token by token; the reference runs it as bs > 1 forwards of at most max_batch_size
    tokens (its scratch buffers are sized for that).  A 100-token prompt through the REAL reference in chunks of
    48 + 52 must leave the same logits (and the same cache: decode continues identically) as the oracle fed one token at a time.
    (Chunks stay well below the harness's max_batch_size of 64: at 63-64 tokens per forward the reference overruns its own
    scratch buffers - glibc reports "corrupted size vs. prev_size" - and every later result is garbage; DESIGN.md defect D11.)"""
    P, R = port(), ref()
    spec = TINY
    w = gen_weights(spec, seed=23)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/m.bin", spec, w)
        write_tokenizer_bin(d + "/t.bin", synthetic_vocab(spec.vocab_size))
        h = R.ref_model_load((d + "/m.bin").encode(), (d + "/t.bin").encode(), 3, Q_INT8, 2, 64, 0)
    assert h
    qm = quantize_model(spec, w, Q_INT8, 64)
    pc = PortConfig(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.head_size, spec.vocab_size, 1024, Q_INT8, 64)
    pm = P.port_model_create(C.byref(pc))
    for (k, l), (q, s) in qm.items():
        P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None, q.shape[0] if q.ndim == 2 else 1, q.shape[-1])
    toks = prompt_tokens(spec, 100, seed=4)
    a, b = np.empty(spec.vocab_size, np.float32), np.empty(spec.vocab_size, np.float32)
    for i in range(toks.size):                                   # oracle: one token per forward, like the engine
        P.port_forward(pm, ptr(toks[i:i + 1].copy()), 1, i, ptr(a))
    pos = 0
    for n in (48, 52):                                           # reference: two batched forwards
        chunk = toks[pos:pos + n].copy()
        R.ref_forward(h, ptr(chunk), n, pos, ptr(b))
        pos += n
    assert beq(a, b)
    for _ in range(6):
        t = np.array([int(np.argmax(b))], np.int32)
        P.port_forward(pm, ptr(t), 1, pos, ptr(a))
        R.ref_forward(h, ptr(t), 1, pos, ptr(b))
        assert beq(a, b), pos
        pos += 1
    P.port_model_free(pm)
    R.ref_model_free(h)


@pytest.mark.skipif(ref_native() is None, reason="libref_native.so not built (host without avx512f)")
def test_native_avx512_build_is_bit_identical_to_haswell_build():
    R, N = ref(), ref_native()
    rng = np.random.default_rng(13)
    for qt in (Q_INT8, Q_INT16):
        W = (rng.standard_normal((64, 4096)) * 0.04).astype(np.float32)
        X = rng.standard_normal((2, 4096)).astype(np.float32)
        qw, sw = quantize(R.ref_quantize, qt, W, 64)
        qx, sx = quantize(N.ref_quantize, qt, X, 64)
        qx2, sx2 = quantize(R.ref_quantize, qt, X, 64)
        assert np.array_equal(qx, qx2) and beq(sx, sx2)
        a, b = np.empty((2, 64), np.float32), np.empty((2, 64), np.float32)
        R.ref_matmul(qt, ptr(a), ptr(qw), ptr(sw), ptr(qx), ptr(sx), 64, 4096, 2, 64)
        N.ref_matmul(qt, ptr(b), ptr(qw), ptr(sw), ptr(qx), ptr(sx), 64, 4096, 2, 64)
        assert beq(a, b)
    x = rng.standard_normal(4096).astype(np.float32)
    g = rng.standard_normal(4096).astype(np.float32)
    a, b = np.empty(4096, np.float32), np.empty(4096, np.float32)
    R.ref_rmsnorm(ptr(a), ptr(x), ptr(g), 4096)
    N.ref_rmsnorm(ptr(b), ptr(x), ptr(g), 4096)
    assert beq(a, b)
