"""GGUF loader (src/model_loaders/gguf_loader.cpp:209-488).  Test files are written with llama.cpp's own
`gguf` Python library.

CPU: Q8_0 blocks are split into payload + exactly converted fp16 scales (checked against the library's dequantiser); the
reference's strict key whitelist and tensor-name check are kept; an all-F32 file gives, through our reader + load-time
quantiser + the oracle, the logits the reference's C++ produced from the same file (golden, and live when oracle/_ref is
present).  GPU: engine_from_gguf on the F32 file equals those golden logits; on the Q8_0 file (group 32) it equals the oracle
on the tensors our reader returns (the reference itself cannot serve there: defect D5, it mis-decodes fp16 scales)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_libs import port, ref, ptr, bits, PortConfig, Q_INT8
from fixtures import TINY, gen_weights, prompt_tokens

gguf = pytest.importorskip("gguf")
from gguf_inputs import write_gguf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GGUF_GOLDEN = os.path.join(HERE, "golden", "gguf_golden.npz")


def port_model(fl, cfg, t):
    """oracle model from reader output, quantising F32 matrices at load like the reference's worker init"""
    P = port()
    gs, qt = cfg["quant_group_size"], cfg["quant_type"] or Q_INT8
    pc = PortConfig(cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"],
                    cfg["dim"] // cfg["n_heads"], cfg["vocab_size"], 1024, qt, gs)
    pm = P.port_model_create(C.byref(pc))
    for (k, l), (q, s) in t.items():
        q = np.ascontiguousarray(q)
        if s is None and q.ndim == 2 and k != fl.T_TOK_EMB:
            q, s = fl.loaders.quantize_rows(q, qt, gs)
        s = None if s is None else np.ascontiguousarray(s)
        rows = q.shape[0] if q.ndim == 2 else 1
        assert P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None, rows, q.shape[-1]) == 0
    return pm


def test_q8_0_blocks_split_exactly(fl, tmp_path):
    spec = TINY
    w = gen_weights(spec, seed=1)
    p = tmp_path / "q8.gguf"
    write_gguf(p, spec, w, q8_0=True)
    cfg, t, vocab = fl.gguf_file.read_gguf(p)
    assert (cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["vocab_size"]) == \
        (spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size)
    assert cfg["quant_type"] == Q_INT8 and cfg["quant_group_size"] == 32 and cfg["name"] == "tiny"
    assert len(vocab["texts"]) == spec.vocab_size and vocab["special"] == {"bos": 1, "eos": 2} and vocab["model"] == "llama"
    assert len(t) == 3 + 9 * spec.n_layers
    rd = gguf.GGUFReader(str(p))
    for ti in rd.tensors:
        if ti.tensor_type != gguf.GGMLQuantizationType.Q8_0:
            continue
        want = gguf.quants.dequantize(ti.data, ti.tensor_type)                   # float32 [rows][cols]
        name = ti.name
        kind = fl.gguf_file.GLOBAL_TENSORS.get(name)
        layer = 0
        if kind is None:
            _, l, short, _ = name.split(".")
            kind, layer = fl.gguf_file.LAYER_TENSORS[short], int(l)
        q, s = t[(kind, layer)]
        assert q.dtype == np.int8 and s.dtype == np.float32 and s.shape == (q.shape[0], q.shape[1] // 32)
        got = (q.astype(np.float32).reshape(q.shape[0], -1, 32) * s[:, :, None]).reshape(q.shape)
        assert np.array_equal(bits(got), bits(np.ascontiguousarray(want))), name
    assert t[(fl.T_ATT_NORM, 0)][1] is None and t[(fl.T_ATT_NORM, 0)][0].dtype == np.float32


def test_key_whitelist_and_tensor_names_like_the_reference(fl, tmp_path):
    spec = TINY
    w = gen_weights(spec, seed=1)
    p = tmp_path / "extra.gguf"
    write_gguf(p, spec, w, q8_0=False, extra_key=True)
    with pytest.raises(fl.gguf_file.GgufError, match="unknown key"):
        fl.gguf_file.read_gguf(p)
    cfg, t, _ = fl.gguf_file.read_gguf(p, strict=False)
    assert cfg["dim"] == spec.dim and cfg["quant_type"] == 0 and cfg["quant_group_size"] == 64
    data = p.read_bytes()
    bad = tmp_path / "bad.gguf"
    bad.write_bytes(data.replace(b"blk.0.attn_q.weight", b"blk.0.attn_x.weight"))
    with pytest.raises(fl.gguf_file.GgufError, match="invalid tensor name"):
        fl.gguf_file.read_gguf(bad, strict=False)
    bad.write_bytes(data[:len(data) // 2])
    with pytest.raises(fl.gguf_file.GgufError):
        fl.gguf_file.read_gguf(bad, strict=False)
    bad.write_bytes(b"GGUX" + data[4:])
    with pytest.raises(fl.gguf_file.GgufError):
        fl.gguf_file.read_gguf(bad)


def check_against_golden(forward, spec):
    g = np.load(GGUF_GOLDEN)
    prompt = g["prompt"].astype(np.int32)
    got = forward(prompt, 0)
    assert np.array_equal(bits(got), bits(g["prefill_logits"]))
    pos = prompt.size
    for i, tok in enumerate(g["decode_tokens"]):
        got = forward(np.array([tok], np.int32), pos)
        assert np.array_equal(bits(got), bits(g["decode_logits"][i])), i
        pos += 1


def test_f32_file_through_our_reader_matches_reference_golden_logits(fl, tmp_path):
    spec = TINY
    p = tmp_path / "f32.gguf"
    write_gguf(p, spec, gen_weights(spec, seed=1), q8_0=False)
    cfg, t, _ = fl.gguf_file.read_gguf(p)
    pm = port_model(fl, cfg, t)
    P = port()

    def forward(toks, pos):
        out = np.empty(spec.vocab_size, np.float32)
        P.port_forward(pm, ptr(toks), toks.size, pos, ptr(out))
        return out
    check_against_golden(forward, spec)
    P.port_model_free(pm)


def test_reference_loader_live_on_the_same_f32_file(fl, tmp_path):
    R = ref()
    if R is None:
        pytest.skip("oracle/_ref not built (GPU box): covered by tests/golden/gguf_golden.npz")
    from fixtures import TINY64
    spec = TINY64
    p = tmp_path / "f32.gguf"
    write_gguf(p, spec, gen_weights(spec, seed=9), q8_0=False)
    h = R.ref_model_load(str(p).encode(), b"", 2, Q_INT8, 2, 64, 0)
    assert h
    cfg, t, _ = fl.gguf_file.read_gguf(p)
    pm = port_model(fl, cfg, t)
    P = port()
    prompt = prompt_tokens(spec, 5, seed=2)
    a, b = np.empty(spec.vocab_size, np.float32), np.empty(spec.vocab_size, np.float32)
    R.ref_forward(h, ptr(prompt), prompt.size, 0, ptr(a))
    P.port_forward(pm, ptr(prompt), prompt.size, 0, ptr(b))
    assert np.array_equal(bits(a), bits(b))
    R.ref_model_free(h)
    P.port_model_free(pm)


@pytest.mark.gpu
def test_engine_from_gguf_f32_matches_reference_golden_logits(fl, tmp_path):
    spec = TINY
    p = tmp_path / "f32.gguf"
    write_gguf(p, spec, gen_weights(spec, seed=1), q8_0=False)
    eng, cfg, vocab = fl.gguf_file.engine_from_gguf(p)
    check_against_golden(lambda toks, pos: eng.forward(toks, pos), spec)
    eng.close()


@pytest.mark.gpu
def test_engine_from_gguf_q8_0_matches_oracle(fl, tmp_path):
    spec = TINY
    p = tmp_path / "q8.gguf"
    write_gguf(p, spec, gen_weights(spec, seed=4), q8_0=True)
    eng, cfg, _ = fl.gguf_file.engine_from_gguf(p)
    assert cfg["quant_group_size"] == 32
    _, t, _ = fl.gguf_file.read_gguf(p)
    pm = port_model(fl, cfg, t)
    P = port()
    toks = prompt_tokens(spec, 6, seed=3)
    want = np.empty(spec.vocab_size, np.float32)
    P.port_forward(pm, ptr(toks), toks.size, 0, ptr(want))
    got = eng.forward(toks, 0)
    assert np.array_equal(bits(got), bits(want))
    pos = toks.size
    for _ in range(20):
        t1 = np.array([int(np.argmax(want))], np.int32)
        P.port_forward(pm, ptr(t1), 1, pos, ptr(want))
        got = eng.forward(t1, pos)
        assert np.array_equal(bits(got), bits(want))
        pos += 1
    P.port_model_free(pm)
    eng.close()
