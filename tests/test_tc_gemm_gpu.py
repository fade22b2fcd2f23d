"""GPU parity of the tensor-core path (tc_gemm.cuh + batch_kernels.cuh) through the C-ABI:

* fl_op_matmul_q_tc — the tcgen05 group-scaled INT8 GEMM — against the oracle's quant::matmul
  (quant_operators.cpp:252-284) for 1..64 activation rows, group 64 and 32, ragged row counts per CTA, a K that is not a
  multiple of the 128-byte stage, and the fused W1/W3 + SwiGLU pass; bit-exact.
* fl_forward with a multi-token prompt (prompt chunks of up to 64 rows per weight pass) against the oracle's forward()
  and against the token-by-token path of the same engine (FL_FLAG_NO_TC); bit-exact logits, identical KV continuation.
* fl_forward_batch / fl_decode_batch_async (several sequences, one weight pass per step) against per-sequence decoding.
"""
import ctypes as C

import numpy as np
import pytest

from oracle_libs import port, ptr, bits, Q_INT8, port_quantize
from fixtures import ModelSpec, TINY, TINY64, gen_weights, quantize_model, prompt_tokens
from test_forward_gpu import make_port_model, make_engine, GQA

pytestmark = pytest.mark.gpu


def _rand_q(rng, rows, cols, gs):
    x = (rng.standard_normal((rows, cols)) * rng.uniform(0.2, 3.0, (rows, 1))).astype(np.float32)
    q, s = port_quantize(Q_INT8, x, gs)
    return q.reshape(rows, cols), s.reshape(rows, cols // gs)


def _oracle_matmul(w, ws, x, xs, gs):
    P = port()
    m, n = w.shape
    out = np.empty((x.shape[0], m), np.float32)
    P.port_matmul(Q_INT8, ptr(out), ptr(w), ptr(ws), ptr(x), ptr(xs), m, n, x.shape[0], gs)
    return out


SHAPES = [
    # m, n, rows, gs
    (300, 256, 8, 64),          # 2-3 rows per CTA
    (1000, 704, 5, 64),         # K not a multiple of the stage (704 = 5.5 x 128), N padded 5 -> 8
    (4096, 512, 64, 64),        # 27-28 rows per CTA, full 64 activation rows
    (12288, 1024, 33, 64),      # 83-84 rows per CTA, N padded 33 -> 64
    (32000, 512, 16, 64),       # 216-217 rows per CTA: two tiles
    (4096, 1024, 20, 32),       # Q8_0 groups: one MMA per group
    (2000, 4096, 1, 64),        # one activation row
]


@pytest.mark.parametrize("m,n,rows,gs", SHAPES, ids=[f"{s[0]}x{s[1]}-r{s[2]}-g{s[3]}" for s in SHAPES])
def test_tc_matmul_bit_exact(fl, m, n, rows, gs):
    rng = np.random.default_rng(m + n + rows)
    w, ws = _rand_q(rng, m, n, gs)
    x, xs = _rand_q(rng, rows, n, gs)
    want = _oracle_matmul(w, ws, x, xs, gs)
    got = fl.ops.matmul_q_tc(w, ws, x, xs, gs=gs)
    bad = np.argwhere(bits(got) != bits(want))
    assert bad.size == 0, (len(bad), bad[:8].tolist(), got[tuple(bad[0])], want[tuple(bad[0])])


@pytest.mark.parametrize("m,n,rows,gs", [(704, 512, 7, 64), (11008, 512, 64, 64), (2048, 768, 24, 32)],
                         ids=["704-r7", "11008-r64", "2048-r24-g32"])
def test_tc_fused_w13_swiglu_bit_exact(fl, m, n, rows, gs):
    rng = np.random.default_rng(7 * m + rows)
    w1, ws1 = _rand_q(rng, m, n, gs)
    w3, ws3 = _rand_q(rng, m, n, gs)
    x, xs = _rand_q(rng, rows, n, gs)
    a = _oracle_matmul(w1, ws1, x, xs, gs)
    b = _oracle_matmul(w3, ws3, x, xs, gs)
    P = port()
    want = a.copy()
    for i in range(rows):
        P.port_swiglu(ptr(want[i]), ptr(b[i]), m)
    got = fl.ops.matmul_q_tc(w1, ws1, x, xs, gs=gs, w3=w3, ws3=ws3)
    bad = np.argwhere(bits(got) != bits(want))
    assert bad.size == 0, (len(bad), bad[:8].tolist())


CASES = [
    ("tiny-int8", TINY, 64, True),
    ("tiny64-int8", TINY64, 64, True),
    ("tiny-int8-g32", TINY, 32, False),
    ("gqa-int8", GQA, 64, True),
]
PATHS = [("mega", 0), ("phase", 4), ("nopdl", 2)]


@pytest.mark.parametrize("path,flags", PATHS, ids=[p[0] for p in PATHS])
@pytest.mark.parametrize("name,spec,gs,qemb", CASES, ids=[c[0] for c in CASES])
def test_prompt_chunks_on_tensor_cores_bit_exact(fl, name, spec, gs, qemb, path, flags):
    """A 150-token prompt = chunks of 64 + 64 + 22 rows; then the KV cache it wrote feeds 6 decode steps."""
    w = gen_weights(spec, seed=11)
    qm = quantize_model(spec, w, Q_INT8, gs, quantize_embedding=qemb)
    pm = make_port_model(spec, qm, Q_INT8, gs)
    P = port()
    eng = make_engine(fl, spec, qm, Q_INT8, gs, flags=flags)
    ref = make_engine(fl, spec, qm, Q_INT8, gs, flags=flags | fl.FLAG_NO_TC)
    toks = prompt_tokens(spec, 150, seed=5)
    want = np.empty(spec.vocab_size, np.float32)
    # the oracle one token at a time (pinned to the reference's batched prefill in tests/test_oracle_ref.py); for GQA the
    # reference's RoPE position of query head g depends on the forward's token count, so feed it whole
    if spec.n_heads != spec.n_kv_heads:
        P.port_forward(pm, ptr(toks), toks.size, 0, ptr(want))
    else:
        for i in range(toks.size):
            t = toks[i:i + 1].copy()
            P.port_forward(pm, ptr(t), 1, i, ptr(want))
    got, am = eng.forward(toks, 0, want_argmax=True)
    slow = ref.forward(toks, 0)
    assert np.array_equal(bits(slow), bits(want)), name
    assert np.array_equal(bits(got), bits(want)), (name, np.abs(got - want).max())
    assert am == P.port_argmax(ptr(want), spec.vocab_size)
    for tap in ("attn", "hd", "x1", "final"):
        n = C.c_int(0)
        pt = P.port_tap(pm, tap.encode(), spec.n_layers - 1, C.byref(n))
        ref_tap = np.ctypeslib.as_array(pt, (n.value,)).copy()
        assert np.array_equal(bits(eng.tap(tap)), bits(ref_tap)), (name, tap)
    pos = toks.size
    for step in range(6):
        t = np.array([int(np.argmax(want))], np.int32)
        P.port_forward(pm, ptr(t), 1, pos, ptr(want))
        got = eng.forward(t, pos)
        assert np.array_equal(bits(got), bits(want)), (name, step)
        pos += 1
    P.port_model_free(pm)
    eng.close(); ref.close()


def test_generate_greedy_with_tensor_core_prefill(fl):
    spec, gs = TINY, 64
    qm = quantize_model(spec, gen_weights(spec, seed=2), Q_INT8, gs)
    prompt = prompt_tokens(spec, 70, seed=9)
    a = make_engine(fl, spec, qm, Q_INT8, gs)
    b = make_engine(fl, spec, qm, Q_INT8, gs, flags=fl.FLAG_NO_TC)
    ta, tb = a.generate_greedy(prompt, 30), b.generate_greedy(prompt, 30)
    assert ta.tolist() == tb.tolist()
    a.close(); b.close()


@pytest.mark.parametrize("n_seqs", [2, 8, 19])
@pytest.mark.parametrize("name,spec,gs", [("tiny", TINY, 64), ("tiny64-g32", TINY64, 32), ("gqa", GQA, 64)])
def test_sequences_share_one_weight_pass(fl, name, spec, gs, n_seqs):
    """fl_forward_batch and fl_decode_batch_async: every sequence's tokens equal the tokens of that sequence decoded alone."""
    qm = quantize_model(spec, gen_weights(spec, seed=4), Q_INT8, gs)
    eng = make_engine(fl, spec, qm, Q_INT8, gs, max_seqs=n_seqs)
    solo = make_engine(fl, spec, qm, Q_INT8, gs, flags=fl.FLAG_NO_TC)
    prompts = [prompt_tokens(spec, 3 + (5 * i) % 11, seed=20 + i) for i in range(n_seqs)]
    n_new = 9
    want = [solo.generate_greedy(p, n_new).tolist() for p in prompts]
    firsts = [eng.forward(p, 0, slot=i, want_logits=False, want_argmax=True) for i, p in enumerate(prompts)]
    toks = np.array(firsts, np.int32)
    pos = np.array([p.size for p in prompts], np.int32)
    outs = [[int(t)] for t in toks]
    for _ in range(4):                                   # host-driven steps
        toks = eng.forward_batch(toks, pos)
        pos += 1
        for i, t in enumerate(toks):
            outs[i].append(int(t))
    eng.decode_batch_async(n_seqs, n_new - 4)            # device-resident steps (one graph replay per step)
    for i in range(n_seqs):
        got = eng.out_tokens(n_new + 1, slot=i).tolist()
        assert got[:5] == outs[i], (name, i)
        stop = want[i].index(0) + 1 if 0 in want[i] else len(want[i])      # the solo run stops after token id 0
        assert got[:stop] == want[i][:stop], (name, i, got, want[i])
    eng.close(); solo.close()


def test_decode_past_the_context_is_rejected(fl):
    spec = TINY
    qm = quantize_model(spec, gen_weights(spec, seed=1), Q_INT8, 64)
    eng = make_engine(fl, spec, qm, Q_INT8, 64, max_seq=64, max_seqs=2)
    eng.forward(prompt_tokens(spec, 60, seed=1), 0, want_logits=False)
    eng.decode_async(4)
    with pytest.raises(fl.FlError):
        eng.decode_async(1)
    eng.forward(prompt_tokens(spec, 60, seed=1), 0, slot=1, want_logits=False)
    with pytest.raises(fl.FlError):
        eng.decode_batch_async(2, 5)
    with pytest.raises(fl.FlError):
        fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1023)
    eng.close()
