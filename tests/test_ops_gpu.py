"""GPU parity, operator level: every fl_op_* entry point of the C-ABI against the CPU oracle (oracle/ref_port.c,
itself pinned to the reference in test_oracle_port.py).  The bar is BIT-EXACT for every operator, integer and
floating point alike (DESIGN.md "Exactness")."""
import numpy as np
import pytest

from oracle_libs import port, ptr, bits, port_quantize, Q_INT8, Q_INT16, NP_T

pytestmark = pytest.mark.gpu


def beq(a, b):
    return np.array_equal(bits(a), bits(b))


@pytest.fixture(scope="module")
def ops(fl):
    return fl.ops


def test_expf_matches_glibc_restatement(ops):
    rng = np.random.default_rng(0)
    x = np.concatenate([
        rng.uniform(-110, 90, 2_000_000), rng.standard_normal(1_000_000) * 5, -np.abs(rng.standard_normal(500_000)) * 20,
        [0.0, -0.0, 88.0, 88.72, 88.73, -87.9, -103.0, -103.5, -103.97, -104.0, -200.0, np.inf, -np.inf, 1e-30, -1e-30],
    ]).astype(np.float32)
    got = ops.expf(x)
    P = port()
    want = np.array([P.port_expf_emul(float(v)) for v in x[:20000]], np.float32)
    assert beq(got[:20000], want)
    # and against the host libm for the whole vector (same glibc on the GPU box)
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    idx = rng.integers(0, x.size, 50000)
    want2 = np.array([libm.expf(float(v)) for v in x[idx]], np.float32)
    assert beq(got[idx], want2)
    tail = x[-15:]
    want3 = np.array([libm.expf(float(v)) for v in tail], np.float32)
    assert beq(got[-15:], want3)


@pytest.mark.parametrize("qt", [Q_INT8, Q_INT16])
@pytest.mark.parametrize("gs", [64, 32])
def test_quantize(ops, qt, gs):
    if gs == 32 and qt == Q_INT16:
        pytest.skip("group 32 exists only for INT8 (GGUF Q8_0)")
    rng = np.random.default_rng(1)
    for n in (64, 512, 704, 4096, 11008):
        x = (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 40.0])).astype(np.float32)
        x[n // 2: n // 2 + gs] = 0.0          # an all-zero group: 0/0 -> NaN -> 0 (quant_operators.cpp:33,42)
        x[3] = 0.0
        q, s = ops.quantize(qt, x, gs)
        qr, sr = port_quantize(qt, x, gs)
        assert np.array_equal(q, qr), (qt, gs, n)
        assert beq(s, sr), (qt, gs, n)


@pytest.mark.parametrize("qt,gs", [(Q_INT8, 64), (Q_INT8, 32), (Q_INT16, 64)])
@pytest.mark.parametrize("shape", [(96, 128, 1), (256, 4096, 1), (130, 11008, 2), (1000, 704, 3), (4096, 4096, 1)])
def test_matmul_q(ops, qt, gs, shape):
    m, n, w = shape
    rng = np.random.default_rng(m * 7 + n)
    W = (rng.standard_normal((m, n)) * 0.05).astype(np.float32)
    X = rng.standard_normal((w, n)).astype(np.float32)
    qw, sw = port_quantize(qt, W, gs)
    qx, sx = port_quantize(qt, X, gs)
    want = np.empty((w, m), np.float32)
    port().port_matmul(qt, ptr(want), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, w, gs)
    got = ops.matmul_q(qt, qw, sw, qx, sx, gs)
    assert beq(got, want), np.abs(got - want).max()


def test_matmul_q_int16_extreme_values_wrap_like_reference(ops):
    # |q| = 5792 everywhere: 64 * 5792^2 = 2 147 024 896 < 2^31 (tools/convert_flm.py:226) -- the largest legal group sum
    m, n = 8, 128
    qw = np.full((m, n), 5792, np.int16); qw[1::2] *= -1
    qx = np.full((1, n), 5792, np.int16)
    sw = np.full((m, n // 64), 1e-4, np.float32); sx = np.full((1, n // 64), 1e-4, np.float32)
    want = np.empty((1, m), np.float32)
    port().port_matmul(Q_INT16, ptr(want), ptr(qw), ptr(sw), ptr(qx), ptr(sx), m, n, 1, 64)
    assert beq(ops.matmul_q(Q_INT16, qw, sw, qx, sx, 64), want)


@pytest.mark.parametrize("n", [64, 512, 768, 4096, 5120])
def test_rmsnorm(ops, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32) * 3
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    want = np.empty(n, np.float32)
    port().port_rmsnorm(ptr(want), ptr(x), ptr(w), n)
    assert beq(ops.rmsnorm(x, w), want)


@pytest.mark.parametrize("hs", [64, 128])
def test_rope(ops, hs):
    rng = np.random.default_rng(hs)
    for pos in (0, 1, 17, 511, 1023, 2559):
        x = rng.standard_normal(hs).astype(np.float32)
        want = np.empty(hs, np.float32)
        port().port_rope_v2(ptr(want), ptr(x), hs, pos)
        assert beq(ops.rope(x, pos), want), pos


@pytest.mark.parametrize("n", [1, 2, 7, 33, 288, 544, 1024, 2560])
def test_softmax(ops, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 4).astype(np.float32)
    want = x.copy()
    port().port_softmax_sisd(ptr(want), n)
    assert beq(ops.softmax(x), want)


def test_swiglu(ops):
    rng = np.random.default_rng(5)
    a = (rng.standard_normal(11008) * 5).astype(np.float32)
    b = rng.standard_normal(11008).astype(np.float32)
    a[:4] = [0.0, -0.0, 100.0, -100.0]
    want = a.copy()
    port().port_swiglu(ptr(want), ptr(b), a.size)
    assert beq(ops.swiglu(a, b), want)


def _port_attention(n_heads, n_kv_heads, hs, pos, qkv, kc, vc):
    """execute_attn for one token with the oracle's leaf functions (mirrors oracle/ref_port.c:port_forward)."""
    P = port()
    dim, kv_dim, hgs = n_heads * hs, n_kv_heads * hs, n_heads // n_kv_heads
    out = np.zeros(dim, np.float32)
    k_new = np.zeros((n_kv_heads, hs), np.float32)
    v_new = qkv[dim + kv_dim:].reshape(n_kv_heads, hs).copy()
    scale = np.float32(1.0) / np.sqrt(np.float32(hs))
    for h in range(n_kv_heads):
        kr = np.empty(hs, np.float32)
        P.port_rope_v2(ptr(kr), ptr(np.ascontiguousarray(qkv[dim + h * hs: dim + (h + 1) * hs])), hs, pos)
        k_new[h] = kr
        K = np.concatenate([kc[h], kr[None]], 0) if pos > 0 else kr[None].copy()
        V = np.concatenate([vc[h], v_new[h][None]], 0) if pos > 0 else v_new[h][None].copy()
        K = np.ascontiguousarray(K); V = np.ascontiguousarray(V)
        for g in range(hgs):
            qh = h * hgs + g
            q = np.empty(hs, np.float32)
            P.port_rope_v2(ptr(q), ptr(np.ascontiguousarray(qkv[qh * hs:(qh + 1) * hs])), hs, pos + g)
            att = np.empty(pos + 1, np.float32)
            for t in range(pos + 1):
                att[t] = np.float32(P.port_dot_f32(ptr(K[t]), ptr(q), hs)) * scale
            P.port_softmax_sisd(ptr(att), pos + 1)
            o = np.empty(hs, np.float32)
            P.port_weighted_sum(ptr(o), ptr(V), ptr(att), pos + 1, hs, 1, 1e-15)
            out[qh * hs:(qh + 1) * hs] = o
    return out, k_new, v_new


@pytest.mark.parametrize("cfg", [(4, 4, 128), (8, 8, 64), (8, 2, 64), (4, 1, 128)])
@pytest.mark.parametrize("pos", [0, 1, 5, 63, 64, 65, 130, 300])
def test_attention_decode(ops, cfg, pos):
    n_heads, n_kv_heads, hs = cfg
    rng = np.random.default_rng(pos * 31 + n_heads)
    dim, kv_dim = n_heads * hs, n_kv_heads * hs
    qkv = rng.standard_normal(dim + 2 * kv_dim).astype(np.float32)
    kc = rng.standard_normal((n_kv_heads, pos, hs)).astype(np.float32)
    vc = rng.standard_normal((n_kv_heads, pos, hs)).astype(np.float32)
    if pos > 8:
        kc[:, 3] *= 30.0     # one dominant key -> most softmax weights fall under the 1e-15 skip threshold (tf_operators.cpp:342)
    want, wk, wv = _port_attention(n_heads, n_kv_heads, hs, pos, qkv, kc, vc)
    got, gk, gv = ops.attn_decode(n_heads, n_kv_heads, hs, pos, qkv, kc if pos else None, vc if pos else None)
    assert beq(gk, wk)
    assert beq(gv, wv)
    assert beq(got, want), np.abs(got - want).max()


def test_argmax_first_index_of_max(ops):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(32000).astype(np.float32)
    assert ops.argmax(x) == int(np.argmax(x))
    x[[17, 900, 31999]] = x.max() + 1        # ties -> first index (sampler.cpp:36-46)
    assert ops.argmax(x) == 17
    assert ops.argmax(np.full(1000, -3.0, np.float32)) == 0
