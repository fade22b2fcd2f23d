"""Malformed model files must end in the reader's own error type (FlmError / GgufError), never in an IndexError,
struct.error or a silent success with missing tensors - the reference's loaders return false on every such path
(flm_loader.cpp:561-634, gguf_loader.cpp:209-488)."""
import numpy as np
import pytest

import flm_inputs as fi
from oracle_libs import Q_INT8
from fixtures import gen_weights


def corruptions(data, rng, n):
    for _ in range(n):
        kind = int(rng.integers(3))
        if kind == 0:                                   # truncate
            yield data[:int(rng.integers(1, len(data)))]
        elif kind == 1:                                 # flip a byte in the first 4 KB (headers, config, tensor infos)
            b = bytearray(data)
            i = int(rng.integers(0, min(len(b), 4096)))
            b[i] ^= 1 << int(rng.integers(8))
            yield bytes(b)
        else:                                           # overwrite a 32-bit field with a huge value
            b = bytearray(data)
            i = int(rng.integers(0, min(len(b), 4096) - 4))
            b[i:i + 4] = b"\xff\xff\xff\x7f"
            yield bytes(b)


def test_flm_reader_never_escapes_with_a_foreign_exception(fl, tmp_path):
    spec = fi.MICRO
    p = tmp_path / "m.flm"
    fl.flm.write_flm(p, fi.config_of(spec, Q_INT8, 64, "m"), fi.quantized_tensors(fl, spec, gen_weights(spec, seed=3), Q_INT8, 64),
                     fi.micro_vocab(spec.vocab_size))
    data = p.read_bytes()
    want = len(fl.flm.read_flm(p)[1])
    rng = np.random.default_rng(0)
    bad = tmp_path / "bad.flm"
    outcomes = {"ok": 0, "rejected": 0}
    for blob in corruptions(data, rng, 300):
        bad.write_bytes(blob)
        try:
            cfg, t, vocab = fl.flm.read_flm(bad)
            for q, s in t.values():                     # whatever was accepted must be fully addressable
                np.asarray(q).sum()
            outcomes["ok"] += 1
        except fl.flm.FlmError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 50 and want == 3 + 9 * spec.n_layers


def test_gguf_reader_never_escapes_with_a_foreign_exception(fl, tmp_path):
    pytest.importorskip("gguf")
    from gguf_inputs import write_gguf
    spec = fi.MICRO
    p = tmp_path / "m.gguf"
    write_gguf(p, spec, gen_weights(spec, seed=3), q8_0=True)
    data = p.read_bytes()
    rng = np.random.default_rng(1)
    bad = tmp_path / "bad.gguf"
    rejected = 0
    for blob in corruptions(data, rng, 300):
        bad.write_bytes(blob)
        try:
            cfg, t, vocab = fl.gguf_file.read_gguf(bad, strict=False)
            for q, s in t.values():
                np.asarray(q).sum()
        except fl.gguf_file.GgufError:
            rejected += 1
    assert rejected > 50
