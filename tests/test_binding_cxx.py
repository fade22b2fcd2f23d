"""The reference-side binding, compiled for real: tests/cxx/fl_binding.cpp is built against the reference's own headers
(oracle/build_ref.sh -> oracle/_ref/libfl_binding.so): the reference's loader parses an .flm into cpuft::TransformerModel,
its tensors go through fl_upload as parallel_thread_init would copy them, forward() is fl_forward.  The logits it returns
must equal, bit for bit, the ctypes path on the same file and the oracle on the tensors our reader returns."""
import ctypes as C
import os

import numpy as np
import pytest

import flm_inputs as fi
from oracle_libs import port, ptr, bits, Q_INT8, Q_INT16, ORACLE
from fixtures import TINY, TINY64, gen_weights, prompt_tokens
from test_flm import port_model_from_flm

BINDING = os.path.join(ORACLE, "_ref", "libfl_binding.so")


def test_binding_source_is_compiled_where_the_reference_is_present():
    """CPU: where /root/reference exists the build recipe must have produced the binding library (so the code in
    INTEGRATION.md is code that compiles against the reference's headers)."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("no reference tree here (GPU box): the prebuilt library travels with the snapshot")
    assert os.path.exists(BINDING), "run oracle/build_ref.sh"
    lib = C.CDLL(BINDING)
    assert hasattr(lib, "fl_binding_run")


@pytest.mark.gpu
@pytest.mark.parametrize("name,spec,qt,gs", [("tiny-int8", TINY, Q_INT8, 64), ("tiny64-int16", TINY64, Q_INT16, 64), ("tiny-int8-g32", TINY, Q_INT8, 32)])
def test_reference_loader_plus_binding_equals_ctypes_path_and_oracle(fl, tmp_path, name, spec, qt, gs):
    if not os.path.exists(BINDING):
        pytest.skip("oracle/_ref/libfl_binding.so not built")
    lib = C.CDLL(BINDING)
    lib.fl_binding_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    p = tmp_path / f"{name}.flm"
    fl.flm.write_flm(p, fi.config_of(spec, qt, gs, name), fi.quantized_tensors(fl, spec, gen_weights(spec, seed=5), qt, gs),
                     fi.micro_vocab(spec.vocab_size))
    prompt = prompt_tokens(spec, 7, seed=4)
    n_dec = 5
    got = np.zeros((1 + n_dec, spec.vocab_size), np.float32)
    rc = lib.fl_binding_run(fl.lib_path().encode(), str(p).encode(), b"", 1, qt, ptr(prompt), prompt.size, n_dec, ptr(got))
    assert rc == spec.vocab_size, rc
    # the ctypes path on the same file
    eng, _, _ = fl.flm.engine_from_flm(p, quant_type=qt)
    pm, _ = port_model_from_flm(fl, p)
    P = port()
    want = np.empty(spec.vocab_size, np.float32)
    cur, pos = prompt, 0
    for step in range(1 + n_dec):
        a = eng.forward(cur, pos)
        P.port_forward(pm, ptr(cur), cur.size, pos, ptr(want))
        assert np.array_equal(bits(got[step]), bits(a)), (name, step, "binding vs ctypes")
        assert np.array_equal(bits(got[step]), bits(want)), (name, step, "binding vs oracle")
        pos += cur.size
        cur = np.array([int(np.argmax(want))], np.int32)
    P.port_model_free(pm)
    eng.close()
