"""CPU: the plain-C restatement (oracle/ref_port.c) against the committed golden vectors, which were recorded from the
REAL reference (oracle/_ref/libref.so built from /root/reference by oracle/build_ref.sh) by tests/golden/make_golden.py.
Everything is compared bit for bit.  This is what pins the oracle on machines where /root/reference does not exist."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_inputs as gi
from fixtures import TINY, gen_weights, quantize_model
from oracle_libs import port, ptr, bits, quantize, PortConfig, Q_INT8, Q_INT16

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = np.load(os.path.join(HERE, "golden", "ops_golden.npz"))
MODEL = np.load(os.path.join(HERE, "golden", "tiny_model_logits.npz"))


def beq(a, b):
    return np.array_equal(bits(a), bits(b))


def test_quantize_golden():
    P = port()
    for name, x in gi.quantize_inputs():
        for qt in (Q_INT8, Q_INT16):
            q, s = quantize(P.port_quantize, qt, x, 64)
            assert np.array_equal(q, OPS[f"quant_{name}_{qt}_q"]), (name, qt)
            assert beq(s, OPS[f"quant_{name}_{qt}_s"]), (name, qt)


def test_matmul_golden():
    P = port()
    for name, qt, gs, qw, sw, qx, sx in gi.matmul_inputs():
        m, n = qw.shape
        o = np.empty((qx.shape[0], m), np.float32)
        P.port_matmul(qt, ptr(o), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, qx.shape[0], gs)
        assert beq(o, OPS[f"matmul_{name}"]), name


def test_rmsnorm_golden():
    P = port()
    for name, x, w in gi.rmsnorm_inputs():
        o = np.empty_like(x)
        P.port_rmsnorm(ptr(o), ptr(x), ptr(w), x.size)
        assert beq(o, OPS[f"rmsnorm_{name}"]), name


def test_rope_golden():
    P = port()
    for name, x, pos in gi.rope_inputs():
        o = np.empty_like(x)
        P.port_rope_v2(ptr(o), ptr(x), x.size, pos)
        assert beq(o, OPS[f"rope_{name}"]), name
        # the (cos, sin) table the CUDA engine uploads (rope_v2's iterated theta, glibc sincosf) is a unit rotation
        tab = np.empty(x.size, np.float32)
        P.port_rope_table(ptr(tab), x.size, pos)
        c, s = tab[0::2], tab[1::2]
        assert np.all(np.abs(c * c + s * s - 1) < 1e-6)


def test_dot_golden():
    P = port()
    for name, a, b in gi.dot_inputs():
        assert beq(np.float32(P.port_dot_f32(ptr(a), ptr(b), a.size)), OPS[f"dot_{name}"]), name


def test_softmax_golden():
    P = port()
    for name, x in gi.softmax_inputs():
        o = x.copy()
        P.port_softmax_sisd(ptr(o), o.size)
        assert beq(o, OPS[f"softmax_{name}"]), name


def test_weighted_sum_golden():
    P = port()
    for name, V, w in gi.wsum_inputs():
        o = np.empty((w.shape[0], V.shape[1]), np.float32)
        P.port_weighted_sum(ptr(o), ptr(V), ptr(w), V.shape[0], V.shape[1], w.shape[0], 1e-15)
        assert beq(o, OPS[f"wsum_{name}"]), name


def test_swiglu_golden():
    P = port()
    for name, a, b in gi.swiglu_inputs():
        o = a.copy()
        P.port_swiglu(ptr(o), ptr(b), o.size)
        assert beq(o, OPS[f"swiglu_{name}"]), name


def test_expf_emulation_equals_libm():
    """port_expf_emul is the algorithm the CUDA kernels run (exact_math.cuh); it must equal this host's glibc expf."""
    P = port()
    libm = C.CDLL("libm.so.6")
    libm.expf.restype = C.c_float
    libm.expf.argtypes = [C.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-110, 90, 60000), rng.standard_normal(40000) * 3,
                         [0.0, -0.0, 88.0, 88.72, 88.73, -87.9, -103.0, -103.97, -104.0, -200.0, 1e-30]]).astype(np.float32)
    got = np.array([P.port_expf_emul(float(v)) for v in xs], np.float32)
    want = np.array([libm.expf(float(v)) for v in xs], np.float32)
    assert beq(got, want)


def test_argmax_is_first_index_of_strict_max():
    P = port()
    x = np.random.default_rng(3).standard_normal(32000).astype(np.float32)
    assert P.port_argmax(ptr(x), x.size) == int(np.argmax(x))
    x[[17, 900]] = x.max() + 1
    assert P.port_argmax(ptr(x), x.size) == 17


def test_forward_golden_logits():
    """ParallelTransformer::forward of the real reference on the seeded TINY llama2.c checkpoint vs port_forward."""
    P = port()
    spec = TINY
    qm = quantize_model(spec, gen_weights(spec, seed=int(MODEL["seed"])), Q_INT8, 64)
    pc = PortConfig(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.head_size,
                    spec.vocab_size, 1024, Q_INT8, 64)
    pm = P.port_model_create(C.byref(pc))
    for (k, l), (q, s) in qm.items():
        assert P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None,
                                       q.shape[0] if q.ndim == 2 else 1, q.shape[-1]) == 0
    toks = MODEL["prompt"].astype(np.int32)
    logits = np.empty(spec.vocab_size, np.float32)
    P.port_forward(pm, ptr(toks), toks.size, 0, ptr(logits))
    assert beq(logits, MODEL["prefill_logits"])
    pos = toks.size
    for i, t in enumerate(MODEL["decode_tokens"]):
        assert int(np.argmax(logits)) == int(t)
        P.port_forward(pm, ptr(np.array([t], np.int32)), 1, pos, ptr(logits))
        assert beq(logits, MODEL["decode_logits"][i]), i
        pos += 1
    # the greedy token stream of ParallelTransformer::generate (stops on token 0, transformer.cpp:93)
    P.port_model_reset(pm)
    P.port_forward(pm, ptr(toks), toks.size, 0, ptr(logits))
    out = [P.port_argmax(ptr(logits), spec.vocab_size)]
    pos = toks.size
    want = MODEL["generate_tokens"].tolist()
    while len(out) < len(want) and out[-1] != 0:
        P.port_forward(pm, ptr(np.array([out[-1]], np.int32)), 1, pos, ptr(logits))
        out.append(P.port_argmax(ptr(logits), spec.vocab_size))
        pos += 1
    assert out == want[:len(out)]
    P.port_model_free(pm)


def test_prefill_equals_token_by_token():
    """The engine prefills one position at a time; per-row arithmetic of the reference's bs>1 forward does not depend
    on the other rows (MHA), so the logits must be bit-identical (DESIGN.md "Prefill")."""
    P = port()
    spec = TINY
    qm = quantize_model(spec, gen_weights(spec, seed=9), Q_INT8, 64)
    pc = PortConfig(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.head_size,
                    spec.vocab_size, 1024, Q_INT8, 64)
    a, b = P.port_model_create(C.byref(pc)), P.port_model_create(C.byref(pc))
    for pm in (a, b):
        for (k, l), (q, s) in qm.items():
            P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None, q.shape[0] if q.ndim == 2 else 1, q.shape[-1])
    toks = np.array([1, 5, 77, 300, 999, 12, 4], np.int32)
    la, lb = np.empty(spec.vocab_size, np.float32), np.empty(spec.vocab_size, np.float32)
    P.port_forward(a, ptr(toks), toks.size, 0, ptr(la))
    for i, t in enumerate(toks):
        P.port_forward(b, ptr(np.array([t], np.int32)), 1, i, ptr(lb))
    assert beq(la, lb)
    P.port_model_free(a)
    P.port_model_free(b)
