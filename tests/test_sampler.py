"""Sampler parity: Sampler::sample (src/transformer/sampler.cpp:113-136) and generate() with sampling
(src/transformer/transformer.cpp:76-103).

CPU: the oracle restatement (port_sample) and the product's host sampler (fl_sampler_*, host logic in the reference too)
against golden vectors produced by the REAL reference (tests/golden/make_golden.py -> sampler_golden.npz), and against the
reference live when oracle/_ref is present.  GPU: fl_generate (device forward + host sampler) token for token against
the reference's generate() goldens and against the oracle loop on other seeds."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_inputs as gi
from oracle_libs import port, ref, ptr, bits, PortConfig, Q_INT8
from fixtures import TINY, gen_weights, quantize_model, prompt_tokens

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sampler_golden.npz")


def draws(sample_fn, logits, temp, topp, seed):
    """gi.SAMPLER_DRAWS consecutive samples from one sampler state; returns tokens, final state, first probabilities"""
    rng = C.c_uint64(seed)
    toks, first = [], None
    for d in range(gi.SAMPLER_DRAWS):
        buf = logits.copy()
        toks.append(sample_fn(ptr(buf), buf.size, temp, topp, C.byref(rng)))
        if d == 0:
            first = buf
    return toks, rng.value, first


def check_against_golden(sample_fn, who):
    g = np.load(GOLDEN)
    for name, logits in gi.sampler_inputs():
        for ci, (temp, topp, seed) in enumerate(gi.SAMPLER_CASES):
            toks, state, first = draws(sample_fn, logits, temp, topp, seed)
            assert toks == g[f"tokens_{name}_{ci}"].tolist(), (who, name, ci)
            assert state == int(g[f"state_{name}_{ci}"]), (who, name, ci)
            if temp != 0.0:
                assert np.array_equal(gi.bits_checksum(first), g[f"probsum_{name}_{ci}"]), (who, name, ci)
                if f"probs_{name}_{ci}" in g.files:
                    assert np.array_equal(bits(first), bits(g[f"probs_{name}_{ci}"])), (who, name, ci)


def test_oracle_sampler_matches_reference_golden():
    check_against_golden(port().port_sample, "port")


def test_oracle_sampler_matches_live_reference():
    R = ref()
    if R is None:
        pytest.skip("oracle/_ref not built (GPU box): covered by the golden vectors")
    P = port()
    r = np.random.default_rng(5)
    for trial in range(40):
        n = int(r.integers(2, 5000))
        logits = (r.standard_normal(n) * float(r.uniform(0.2, 8.0))).astype(np.float32)
        if trial % 4 == 0:
            logits = np.round(logits)                       # ties
        temp = float(np.float32(r.choice([0.0, 0.3, 0.8, 1.0, 1.7])))
        topp = float(np.float32(r.choice([0.0, 0.3, 0.9, 0.99, 1.0])))
        seed = int(r.integers(1, 2**62))
        a = draws(P.port_sample, logits, temp, topp, seed)
        b = draws(R.ref_sampler_sample, logits, temp, topp, seed)
        assert a[0] == b[0] and a[1] == b[1], (trial, n, temp, topp)
        assert np.array_equal(bits(a[2]), bits(b[2])), (trial, n, temp, topp)


def product_sample_fn(fl):
    """the C-ABI sampler behind the same (logits, n, temp, topp, &rng) signature; one fl_sampler per rng state"""
    L = fl.lib()
    samplers = {}

    def fn(buf, n, temp, topp, rng_ref):
        rng = rng_ref._obj
        key = id(rng)
        if key not in samplers or samplers[key][1] != rng.value:
            h = C.c_void_p()
            assert L.fl_sampler_create(n, rng.value, C.byref(h)) == 0
            samplers[key] = [h, rng.value]
        h = samplers[key][0]
        tok = C.c_int32(-1)
        assert L.fl_sampler_sample(h, buf, temp, topp, C.byref(tok)) == 0
        rng.value = L.fl_sampler_state(h)
        samplers[key][1] = rng.value
        return tok.value
    return fn


def test_product_host_sampler_matches_reference_golden(fl):
    """fl_sampler_* is host code inside the C-ABI library (no device call), so it is checked on CPU as well"""
    check_against_golden(product_sample_fn(fl), "product")


def test_product_host_sampler_matches_oracle_on_random_cases(fl):
    P = port()
    fn = product_sample_fn(fl)
    r = np.random.default_rng(6)
    for trial in range(40):
        n = int(r.integers(2, 40000))
        logits = (r.standard_normal(n) * float(r.uniform(0.2, 8.0))).astype(np.float32)
        if trial % 3 == 0:
            logits = np.round(logits * 2) / 2
        temp = float(np.float32(r.choice([0.0, 0.3, 0.8, 1.0, 1.7])))
        topp = float(np.float32(r.choice([-1.0, 0.0, 0.3, 0.9, 0.99, 1.0])))
        seed = int(r.integers(1, 2**62))
        a = draws(P.port_sample, logits, temp, topp, seed)
        b = draws(fn, logits, temp, topp, seed)
        assert a[0] == b[0] and a[1] == b[1], (trial, n, temp, topp)
        assert np.array_equal(bits(a[2]), bits(b[2])), (trial, n, temp, topp)


def test_sampler_wrapper_class(fl):
    s = fl.Sampler(1000, seed=3)
    logits = np.linspace(-3, 3, 1000).astype(np.float32)
    t0 = s.sample(logits.copy(), 0.0, 0.9)
    assert t0 == 999 and s.state == 3                       # greedy draws no random number
    t1 = s.sample(logits.copy(), 1.0, 0.9)
    assert 0 <= t1 < 1000 and s.state != 3
    s.close()


# ---------------------------------------------------------------- GPU: generate() with sampling ------------------------
def oracle_generate(P, pm, spec, prompt, max_new, temp, topp, seed):
    logits = np.empty(spec.vocab_size, np.float32)
    rng = C.c_uint64(seed)
    out, pos, cur = [], 0, prompt
    tok = -1
    while tok != 0 and pos < prompt.size + max_new:          # transformer.cpp:93-101
        P.port_forward(pm, ptr(cur), cur.size, pos, ptr(logits))
        tok = P.port_sample(ptr(logits), spec.vocab_size, temp, topp, C.byref(rng))
        out.append(tok)
        pos += cur.size
        cur = np.array([tok], np.int32)
    return out


@pytest.mark.gpu
def test_generate_with_sampling_matches_reference_golden_and_oracle(fl):
    from test_forward_gpu import make_engine, make_port_model
    g = np.load(GOLDEN)
    spec = TINY
    qm = quantize_model(spec, gen_weights(spec, seed=1), Q_INT8, 64)
    eng = make_engine(fl, spec, qm, Q_INT8, 64)
    prompt = g["prompt"].astype(np.int32)
    for i in range(3):
        temp, topp, seed = g[f"generate_{i}_args"]
        got = eng.generate(prompt, 40, float(np.float32(temp)), float(np.float32(topp)), int(seed))
        assert got.tolist() == g[f"generate_{i}"].tolist(), i
    # other seeds / settings against the oracle loop, incl. temperature 0 (the device-resident greedy path)
    pm = make_port_model(spec, qm, Q_INT8, 64)
    P = port()
    for temp, topp, seed in [(0.8, 0.9, 42), (1.5, 0.0, 43), (0.0, 0.9, 44), (1.0, 0.3, 45)]:
        P.port_model_reset(pm)
        p2 = prompt_tokens(spec, 5, seed=seed)
        want = oracle_generate(P, pm, spec, p2, 60, temp, topp, seed)
        got = eng.generate(p2, 60, temp, topp, seed)
        assert got.tolist() == want, (temp, topp, seed)
    P.port_model_free(pm)
    eng.close()
