"""ctypes binding of include/fastllama_b200.h (the same stub a reference-side maintainer would write; see
INTEGRATION.md).  Fails loudly when the CUDA library has not been built — there is no fallback path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
Q_INT16, Q_INT8 = 1, 2
(T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS) = range(12)
FLAG_NO_GRAPH, FLAG_NO_PDL, FLAG_NO_MEGAKERNEL, FLAG_PROFILE, FLAG_NO_TC, FLAG_RELAXED = 1, 2, 4, 8, 16, 32

EXPORTED_SYMBOLS = [
    "fl_create", "fl_destroy", "fl_last_error", "fl_upload", "fl_finalize", "fl_forward", "fl_forward_batch",
    "fl_generate_greedy", "fl_generate", "fl_sampler_create", "fl_sampler_destroy", "fl_sampler_sample", "fl_sampler_state",
    "fl_decode_async", "fl_stream", "fl_device_ptr", "fl_profile_read", "fl_sync", "fl_launch_count", "fl_step_bytes", "fl_tap",
    "fl_set_comm", "fl_allgather_tokens", "fl_op_quantize", "fl_op_matmul_q", "fl_op_rmsnorm", "fl_op_rope",
    "fl_op_softmax", "fl_op_swiglu", "fl_op_expf", "fl_op_attn_decode", "fl_op_argmax",
    "fl_decode_batch_async", "fl_op_matmul_q_tc", "fl_read_out_tokens",
]


class FlError(RuntimeError):
    pass


class FlConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "head_size",
                                         "vocab_size", "max_seq_len", "quant_type", "group_size", "max_seqs", "flags")]


def lib_path():
    # FL_PROF_LIB=1 selects the build with the in-kernel profiling counters compiled in (profiles/*.py set it)
    sel = os.environ.get("FL_PROF_LIB", "")
    name = "libfastllama_b200.so" if not sel else "libfastllama_b200_prof.so" if sel == "1" else f"libfastllama_b200_{sel}.so"     # A/B builds: csrc/Makefile `variants`
    return os.path.join(HERE, name)


_lib = None


def lib():
    """Load libfastllama_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise FlError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(p)
    vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
    L.fl_create.argtypes = [C.POINTER(FlConfig), C.c_int, C.POINTER(vp)]
    L.fl_destroy.argtypes = [vp]
    L.fl_destroy.restype = None
    L.fl_last_error.argtypes = [vp]
    L.fl_last_error.restype = C.c_char_p
    L.fl_upload.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
    L.fl_finalize.argtypes = [vp]
    L.fl_forward.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
    L.fl_forward_batch.argtypes = [vp, C.c_int, vp, vp, vp]
    L.fl_generate_greedy.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, i32p]
    L.fl_decode_async.argtypes = [vp, C.c_int, C.c_int]
    L.fl_decode_batch_async.argtypes = [vp, C.c_int, C.c_int]
    L.fl_read_out_tokens.argtypes = [vp, C.c_int, C.c_int, vp]
    L.fl_op_matmul_q_tc.argtypes = [C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp, C.c_int]
    L.fl_stream.argtypes = [vp]
    L.fl_stream.restype = vp
    L.fl_device_ptr.argtypes = [vp, C.c_char_p, C.c_int]
    L.fl_device_ptr.restype = vp
    L.fl_profile_read.argtypes = [vp, vp, C.c_int, C.c_int]
    L.fl_sync.argtypes = [vp]
    L.fl_launch_count.argtypes = [vp]
    L.fl_launch_count.restype = C.c_int64
    L.fl_step_bytes.argtypes = [vp, C.c_int]
    L.fl_step_bytes.restype = C.c_int64
    L.fl_tap.argtypes = [vp, C.c_char_p, vp, C.c_int]
    L.fl_set_comm.argtypes = [vp, vp, C.c_int, C.c_int]
    L.fl_allgather_tokens.argtypes = [vp, vp, C.c_int, vp]
    L.fl_op_quantize.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, vp]
    L.fl_op_matmul_q.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]
    L.fl_op_rmsnorm.argtypes = [vp, vp, C.c_int, vp]
    L.fl_op_rope.argtypes = [vp, C.c_int, C.c_int, vp]
    L.fl_op_softmax.argtypes = [vp, C.c_int, vp]
    L.fl_op_swiglu.argtypes = [vp, vp, C.c_int, vp]
    L.fl_op_expf.argtypes = [vp, C.c_int, vp]
    L.fl_op_attn_decode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    L.fl_op_argmax.argtypes = [vp, C.c_int, vp]
    L.fl_generate.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint64, vp, i32p]
    L.fl_sampler_create.argtypes = [C.c_int, C.c_uint64, C.POINTER(vp)]
    L.fl_sampler_destroy.argtypes = [vp]
    L.fl_sampler_destroy.restype = None
    L.fl_sampler_sample.argtypes = [vp, vp, C.c_float, C.c_float, i32p]
    L.fl_sampler_state.argtypes = [vp]
    L.fl_sampler_state.restype = C.c_uint64
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc, handle=None):
    if rc < 0:
        msg = lib().fl_last_error(handle)
        raise FlError(f"fastllama_b200 error {rc}: {msg.decode() if msg else ''}")
    return rc


_NP = {Q_INT8: np.int8, Q_INT16: np.int16}


class _Ops:
    """Per-operator entry points (host buffers in/out) — used by the parity tests."""

    def quantize(self, qt, x, gs=64):
        x = np.ascontiguousarray(x, np.float32)
        q = np.empty(x.shape, _NP[qt])
        s = np.empty(x.size // gs, np.float32)
        _check(lib().fl_op_quantize(qt, gs, _p(x), x.size, _p(q), _p(s)))
        return q, s.reshape(x.shape[:-1] + (x.shape[-1] // gs,))

    def matmul_q(self, qt, w, ws, x, xs, gs=64):
        m, n = w.shape
        x = np.ascontiguousarray(x).reshape(-1, n)
        out = np.empty((x.shape[0], m), np.float32)
        _check(lib().fl_op_matmul_q(qt, gs, _p(np.ascontiguousarray(w)), _p(np.ascontiguousarray(ws, np.float32)), m, n,
                                    _p(x), _p(np.ascontiguousarray(xs, np.float32)), x.shape[0], _p(out)))
        return out

    def matmul_q_tc(self, w, ws, x, xs, gs=64, w3=None, ws3=None, variant=0):
        """quant::matmul for up to 64 activation rows in one weight pass on the tensor cores; with w3: swiglu(W x, W3 x)."""
        m, n = w.shape
        x = np.ascontiguousarray(x).reshape(-1, n)
        out = np.empty((x.shape[0], m), np.float32)
        w3c = None if w3 is None else np.ascontiguousarray(w3)
        ws3c = None if ws3 is None else np.ascontiguousarray(ws3, np.float32)
        _check(lib().fl_op_matmul_q_tc(gs, _p(np.ascontiguousarray(w)), _p(np.ascontiguousarray(ws, np.float32)), _p(w3c), _p(ws3c),
                                       m, n, _p(x), _p(np.ascontiguousarray(xs, np.float32)), x.shape[0], _p(out), variant))
        return out

    def rmsnorm(self, x, w):
        x = np.ascontiguousarray(x, np.float32); w = np.ascontiguousarray(w, np.float32)
        out = np.empty_like(x)
        _check(lib().fl_op_rmsnorm(_p(x), _p(w), x.size, _p(out)))
        return out

    def rope(self, x, pos):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        _check(lib().fl_op_rope(_p(x), x.size, pos, _p(out)))
        return out

    def softmax(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        _check(lib().fl_op_softmax(_p(x), x.size, _p(out)))
        return out

    def swiglu(self, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        out = np.empty_like(a)
        _check(lib().fl_op_swiglu(_p(a), _p(b), a.size, _p(out)))
        return out

    def expf(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(x)
        _check(lib().fl_op_expf(_p(x), x.size, _p(out)))
        return out

    def attn_decode(self, n_heads, n_kv_heads, head_size, pos, qkv, k_cache, v_cache):
        qkv = np.ascontiguousarray(qkv, np.float32)
        out = np.empty(n_heads * head_size, np.float32)
        k_new = np.empty((n_kv_heads, head_size), np.float32)
        v_new = np.empty((n_kv_heads, head_size), np.float32)
        kc = None if k_cache is None else np.ascontiguousarray(k_cache, np.float32)
        vc = None if v_cache is None else np.ascontiguousarray(v_cache, np.float32)
        _check(lib().fl_op_attn_decode(n_heads, n_kv_heads, head_size, pos, _p(qkv), _p(kc), _p(vc), _p(out), _p(k_new), _p(v_new)))
        return out, k_new, v_new

    def argmax(self, logits):
        logits = np.ascontiguousarray(logits, np.float32)
        out = np.zeros(1, np.int32)
        _check(lib().fl_op_argmax(_p(logits), logits.size, _p(out)))
        return int(out[0])


ops = _Ops()


class Engine:
    """One fl_engine (one GPU).  Mirrors ParallelTransformer's load -> forward/generate life cycle."""

    def __init__(self, dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, max_seq_len=1024,
                 quant_type=Q_INT8, group_size=64, max_seqs=1, flags=0, device=0):
        self.cfg = FlConfig(dim, hidden_dim, n_layers, n_heads, n_kv_heads, dim // n_heads, vocab_size, max_seq_len,
                            quant_type, group_size, max_seqs, flags)
        self.h = C.c_void_p()
        _check(lib().fl_create(C.byref(self.cfg), device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().fl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, kind, layer, q, scales=None):
        q = np.ascontiguousarray(q)
        rows = q.shape[0] if q.ndim == 2 else 1
        s = None if scales is None else np.ascontiguousarray(scales, np.float32)
        _check(lib().fl_upload(self.h, kind, layer, _p(q), _p(s), rows, q.shape[-1]), self.h)

    def upload_model(self, tensors):
        """tensors: {(kind, layer): (payload, scales|None)} as produced by tests/fixtures.quantize_model."""
        for (kind, layer), (q, s) in tensors.items():
            self.upload(kind, layer, q, s)
        self.finalize()

    def finalize(self):
        _check(lib().fl_finalize(self.h), self.h)

    def forward(self, tokens, pos, slot=0, want_logits=True, want_argmax=False):
        tokens = np.ascontiguousarray(tokens, np.int32)
        logits = np.empty(self.cfg.vocab_size, np.float32) if want_logits else None
        am = np.zeros(1, np.int32) if want_argmax else None
        _check(lib().fl_forward(self.h, slot, _p(tokens), tokens.size, pos, _p(logits), _p(am)), self.h)
        if want_logits and want_argmax:
            return logits, int(am[0])
        return logits if want_logits else (int(am[0]) if want_argmax else None)

    def forward_batch(self, tokens, pos):
        tokens = np.ascontiguousarray(tokens, np.int32); pos = np.ascontiguousarray(pos, np.int32)
        out = np.zeros(tokens.size, np.int32)
        _check(lib().fl_forward_batch(self.h, tokens.size, _p(tokens), _p(pos), _p(out)), self.h)
        return out

    def generate_greedy(self, prompt, max_new, slot=0):
        prompt = np.ascontiguousarray(prompt, np.int32)
        out = np.zeros(max_new + 1, np.int32)
        n = C.c_int32(0)
        _check(lib().fl_generate_greedy(self.h, slot, _p(prompt), prompt.size, max_new, _p(out), C.byref(n)), self.h)
        return out[:n.value].copy()

    def generate(self, prompt, max_new, temperature=1.0, topp=0.9, seed=0, slot=0):
        """ParallelTransformer::generate (transformer.cpp:76-103) with its sampler; temperature 0 is generate_greedy."""
        prompt = np.ascontiguousarray(prompt, np.int32)
        out = np.zeros(max_new + 1, np.int32)
        n = C.c_int32(0)
        _check(lib().fl_generate(self.h, slot, _p(prompt), prompt.size, max_new, temperature, topp, seed, _p(out),
                                 C.byref(n)), self.h)
        return out[:n.value].copy()

    def decode_async(self, n_steps, slot=0):
        _check(lib().fl_decode_async(self.h, slot, n_steps), self.h)

    def decode_batch_async(self, n_seqs, n_steps):
        """n_steps greedy steps of sequences 0..n_seqs-1 from their device-resident states: one weight pass per step."""
        _check(lib().fl_decode_batch_async(self.h, n_seqs, n_steps), self.h)

    def out_tokens(self, n, slot=0):
        """the first n tokens sampled for `slot` since its last prefill"""
        buf = np.zeros(n, np.int32)
        _check(lib().fl_read_out_tokens(self.h, slot, n, _p(buf)), self.h)
        return buf

    def sync(self):
        _check(lib().fl_sync(self.h), self.h)

    @property
    def stream(self):
        return lib().fl_stream(self.h)

    def device_ptr(self, name, slot=0):
        p = lib().fl_device_ptr(self.h, name.encode(), slot)
        if not p:
            raise FlError(f"fl_device_ptr: unknown name {name!r}")
        return p

    def profile_read(self, reset=True, events=False):
        """per-CTA counters [n_ctas][32]; with events=True also the two event logs of the traced layer ([2][4096] uint64)"""
        buf = np.zeros(32 * 1024 + 2 * 4096, np.uint64)
        n = _check(lib().fl_profile_read(self.h, _p(buf), buf.size, int(reset)), self.h)
        counters = buf[:n].reshape(-1, 32).copy()
        if events:
            return counters, buf[n:n + 2 * 4096].reshape(2, 4096).copy()
        return counters

    def launch_count(self):
        return lib().fl_launch_count(self.h)

    def step_bytes(self, ctx):
        return lib().fl_step_bytes(self.h, ctx)

    def tap(self, name):
        buf = np.empty(max(self.cfg.vocab_size, self.cfg.hidden_dim, 3 * self.cfg.dim), np.float32)
        n = _check(lib().fl_tap(self.h, name.encode(), _p(buf), buf.size), self.h)
        return buf[:n].copy()


class Sampler:
    """cpuft::Sampler (src/transformer/sampler.h:13-34): build(vocab_size, seed) / sample(logits, temperature, topp).
    Host logic, as in the reference; `logits` (float32, vocab_size) is overwritten with the probabilities."""

    def __init__(self, vocab_size, seed=0):
        self.h = C.c_void_p()
        self.n = vocab_size
        _check(lib().fl_sampler_create(vocab_size, seed, C.byref(self.h)), None)

    def sample(self, logits, temperature=1.0, topp=0.9):
        assert logits.dtype == np.float32 and logits.size == self.n and logits.flags.c_contiguous
        tok = C.c_int32(-1)
        _check(lib().fl_sampler_sample(self.h, _p(logits), temperature, topp, C.byref(tok)), None)
        return tok.value

    @property
    def state(self):
        return lib().fl_sampler_state(self.h)

    def close(self):
        if self.h:
            lib().fl_sampler_destroy(self.h)
            self.h = C.c_void_p()

    __del__ = close
