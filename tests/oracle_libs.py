"""ctypes access to the CPU checkers (TEST INFRASTRUCTURE — never imported by the product).

  port()        -> oracle/libref_port.so   plain-C restatement (always available; built by oracle/Makefile)
  ref()         -> oracle/_ref/libref.so   the real reference, -march=haswell (None if never built)
  ref_native()  -> oracle/_ref/libref_native.so   reference with build.sh's -march=native flags (or None)
"""
import ctypes as C
import os
import subprocess
import functools
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")

Q_NONE, Q_INT16, Q_INT8 = 0, 1, 2
QF = {Q_INT8: 127.0, Q_INT16: 5792.0}
NP_T = {Q_INT8: np.int8, Q_INT16: np.int16}

(T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS) = range(12)

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
vp = C.c_void_p


def ptr(a):
    return a.ctypes.data_as(vp)


class PortConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "head_size",
                                       "vocab_size", "max_seq_len", "qtype", "group")]


def _build_port():
    so = os.path.join(ORACLE, "libref_port.so")
    src = os.path.join(ORACLE, "ref_port.c")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE, "libref_port.so"], stdout=subprocess.DEVNULL)
    return so


@functools.lru_cache(None)
def port():
    lib = C.CDLL(_build_port())
    lib.port_quantize.argtypes = [C.c_int, vp, vp, vp, C.c_size_t, C.c_int]
    lib.port_dequantize.argtypes = [C.c_int, vp, vp, vp, C.c_size_t, C.c_int]
    lib.port_matmul.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.port_square_sum.argtypes = [vp, C.c_size_t]
    lib.port_square_sum.restype = C.c_float
    lib.port_rmsnorm.argtypes = [vp, vp, vp, C.c_size_t]
    lib.port_rope_v2.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.port_rope_table.argtypes = [vp, C.c_int, C.c_int]
    lib.port_dot_f32.argtypes = [vp, vp, C.c_size_t]
    lib.port_dot_f32.restype = C.c_float
    lib.port_softmax_sisd.argtypes = [vp, C.c_int]
    lib.port_weighted_sum.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float]
    lib.port_swiglu.argtypes = [vp, vp, C.c_size_t]
    lib.port_expf_emul.argtypes = [C.c_float]
    lib.port_expf_emul.restype = C.c_float
    lib.port_argmax.argtypes = [vp, C.c_int]
    lib.port_softmax.argtypes = [vp, C.c_size_t]
    lib.port_sample.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_uint64)]
    lib.port_model_create.argtypes = [C.POINTER(PortConfig)]
    lib.port_model_create.restype = vp
    lib.port_model_free.argtypes = [vp]
    lib.port_model_reset.argtypes = [vp]
    lib.port_model_set_tensor.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
    lib.port_forward.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.port_tap.argtypes = [vp, C.c_char_p, C.c_int, i32p]
    lib.port_tap.restype = f32p
    return lib


def _load_ref(name):
    so = os.path.join(ORACLE, "_ref", name)
    if not os.path.exists(so):
        return None
    lib = C.CDLL(so)
    lib.ref_quantize.argtypes = [C.c_int, vp, vp, vp, C.c_size_t, C.c_int]
    lib.ref_dequantize.argtypes = [C.c_int, vp, vp, vp, C.c_size_t, C.c_int]
    lib.ref_matmul.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_rmsnorm.argtypes = [vp, vp, vp, C.c_size_t]
    lib.ref_rmsnorm_inplace_via_tensor.argtypes = [vp, vp, C.c_int]
    lib.ref_swiglu.argtypes = [vp, vp, C.c_size_t]
    lib.ref_rope_v2.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    lib.ref_softmax_sisd.argtypes = [vp, C.c_int]
    lib.ref_weighted_sum.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float]
    lib.ref_dot_f32.argtypes = [vp, vp, C.c_size_t]
    lib.ref_dot_f32.restype = C.c_float
    lib.ref_dot_i8.argtypes = [vp, vp, C.c_size_t]
    lib.ref_dot_i16.argtypes = [vp, vp, C.c_size_t]
    lib.ref_square_sum.argtypes = [vp, C.c_size_t]
    lib.ref_square_sum.restype = C.c_float
    lib.ref_array_max.argtypes = [vp, C.c_size_t]
    lib.ref_array_max.restype = C.c_float
    lib.ref_multiply.argtypes = [vp, C.c_float, C.c_size_t]
    lib.ref_add.argtypes = [vp, vp, C.c_size_t]
    lib.ref_simd_size.restype = C.c_size_t
    lib.ref_sample_argmax.argtypes = [vp, C.c_int]
    lib.ref_sampler_sample.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_uint64)]
    lib.ref_generate.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint64, vp, C.c_int]
    lib.ref_model_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_model_load.restype = vp
    lib.ref_model_free.argtypes = [vp]
    lib.ref_model_config.argtypes = [vp, i32p]
    lib.ref_forward.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    lib.ref_generate_greedy.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int]
    lib.ref_encode.argtypes = [vp, C.c_char_p, vp, C.c_int]
    lib.ref_decode.argtypes = [vp, vp, C.c_int, C.c_char_p, C.c_int]
    return lib


@functools.lru_cache(None)
def ref():
    return _load_ref("libref.so")


@functools.lru_cache(None)
def ref_native():
    return _load_ref("libref_native.so")


# ------------------------------------------------------------------ numpy conveniences
def quantize(lib_fn, qt, x, gs=64):
    """x: float32 [..., n] -> (q [..., n], scales [..., n/gs]) via the given quantize entry point."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.empty(x.shape, dtype=NP_T[qt])
    s = np.empty(x.size // gs, dtype=np.float32)
    lib_fn(qt, ptr(q), ptr(s), ptr(x), x.size, gs)
    return q, s.reshape(x.shape[:-1] + (x.shape[-1] // gs,))


def port_quantize(qt, x, gs=64):
    return quantize(port().port_quantize, qt, x, gs)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
