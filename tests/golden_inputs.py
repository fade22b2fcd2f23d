"""Seeded inputs shared by tests/golden/make_golden.py (which records the REAL reference's outputs for them) and
tests/test_oracle_port.py (which checks the C restatement against those recorded outputs)."""
import numpy as np

from oracle_libs import Q_INT8, Q_INT16, port_quantize


def quantize_inputs():
    rng = np.random.default_rng(100)
    for n, scale in ((64, 1.0), (4096, 0.02), (11008, 30.0), (768, 1e-4)):
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        x[1] = 0.0
        yield f"n{n}", x
    z = np.zeros(192, np.float32)
    z[70] = 1e-30
    z[130] = -2.5
    yield "zero_group", z          # group 0 is all zero: 0/0 -> NaN -> 0 in the reference (quant_operators.cpp:33,42)


def matmul_inputs():
    rng = np.random.default_rng(101)
    for qt in (Q_INT8, Q_INT16):
        for (m, n, w) in ((96, 128, 1), (64, 4096, 1), (32, 11008, 2), (40, 704, 17)):
            W = (rng.standard_normal((m, n)) * 0.05).astype(np.float32)
            X = rng.standard_normal((w, n)).astype(np.float32)
            qw, sw = port_quantize(qt, W, 64)
            qx, sx = port_quantize(qt, X, 64)
            yield f"q{qt}_{m}x{n}x{w}", qt, 64, qw, sw, qx, sx


def rmsnorm_inputs():
    rng = np.random.default_rng(102)
    for n in (64, 768, 4096, 5120):
        yield f"n{n}", (rng.standard_normal(n) * 2).astype(np.float32), (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)


def rope_inputs():
    rng = np.random.default_rng(103)
    for hs in (64, 128):
        for pos in (0, 1, 9, 100, 1023, 2500):
            yield f"hs{hs}_p{pos}", rng.standard_normal(hs).astype(np.float32), pos


def dot_inputs():
    rng = np.random.default_rng(104)
    for n in (64, 128):
        for i in range(8):
            yield f"n{n}_{i}", rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)


def softmax_inputs():
    rng = np.random.default_rng(105)
    for n in (1, 2, 7, 8, 33, 288, 1000):
        yield f"n{n}", (rng.standard_normal(n) * 3).astype(np.float32)


def wsum_inputs():
    rng = np.random.default_rng(106)
    for (m, n, bs) in ((1, 128, 1), (5, 64, 1), (300, 128, 1), (40, 128, 3)):
        V = rng.standard_normal((m, n)).astype(np.float32)
        w = rng.random((bs, m)).astype(np.float32)
        w[:, m // 2:] *= 1e-16
        yield f"{m}x{n}x{bs}", V, w


def swiglu_inputs():
    rng = np.random.default_rng(107)
    for n in (8, 2048, 11008):
        yield f"n{n}", (rng.standard_normal(n) * 4).astype(np.float32), rng.standard_normal(n).astype(np.float32)


def sampler_inputs():
    """(name, logits) for Sampler::sample: a peaked and a flat distribution at a real vocabulary size, a small one, and one
    with many exactly equal logits (so that the candidate sort sees ties)."""
    r = np.random.default_rng(77)
    peaked = (r.standard_normal(32000) * 3.0).astype(np.float32)
    flat = (r.standard_normal(32000) * 0.3).astype(np.float32)
    small = (r.standard_normal(1000) * 2.0).astype(np.float32)
    ties = (np.round(r.standard_normal(4096) * 2.0) * 0.5).astype(np.float32)
    wide = (r.standard_normal(8000) * 12.0).astype(np.float32)        # many entries more than 15 below the maximum
    return [("peaked", peaked), ("flat", flat), ("small", small), ("ties", ties), ("wide", wide)]


# (temperature, topp, seed): nucleus sampling, plain multinomial (topp outside (0,1)), greedy
SAMPLER_CASES = [(1.0, 0.9, 1), (0.7, 0.95, 2), (1.3, 0.5, 3), (1.0, 1.0, 4), (0.5, 0.0, 5), (2.0, 0.99, 6), (0.0, 0.9, 7)]
SAMPLER_DRAWS = 6


def bits_checksum(a):
    """order-sensitive checksum of the bit patterns of a float32 array (two uint64 words)"""
    b = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    k = np.arange(1, b.size + 1, dtype=np.uint64)
    return np.array([np.bitwise_xor.reduce(b * k), (b * (k | np.uint64(1))).sum(dtype=np.uint64)], np.uint64)
