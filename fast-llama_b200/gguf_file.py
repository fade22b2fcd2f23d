"""GGUF model files (BASELINE.json config 5 ships as Q8_0 GGUF) -> engine, read the way the reference reads them
(src/model_loaders/gguf_loader.cpp:209-488):

  header   u32 'GGUF', u32 version, i64 tensor count, i64 key/value count
  kv       string key (u64 length + bytes), u32 value type, value.  The reference accepts ONLY the keys in KNOWN_KEYS and
           fails on any other (:319-322); `strict=True` (default) keeps that behaviour, `strict=False` skips unknown keys
           (files written by current llama.cpp carry more metadata than the reference tolerates).
  tensors  string name, u32 n_dims, u64 dims[n_dims] (dims[0] = columns), u32 type, u64 offset; data starts at the next
           multiple of `general.alignment` (32) after the last tensor info.  Names: token_embd / output_norm / output and
           blk.N.{attn_q, attn_k, attn_v, attn_output, attn_norm, ffn_gate, ffn_down, ffn_up, ffn_norm}.weight (:345-372);
           any other name is an error (:377-380).  Types: F32 and Q8_0 only (:399-407).
  Q8_0     blocks of {fp16 scale, int8[32]}: split into an int8 payload [rows][cols] and fp32 scales [rows][cols/32]
           (:462-470), which is the layout fl_upload takes with group_size 32.

Defect D5 of the reference (its float16_to_float32 returns the bit pattern, src/utils/utility.cpp:64, so every Q8_0 scale
it loads is garbage) is NOT reproduced: scales are converted fp16 -> fp32 exactly.  An all-F32 GGUF goes through the
reference correctly (matrices are then quantised at load with group 64, the embedding stays fp32), and that is the case
the parity test pins against the real loader."""
import struct

import numpy as np

from .binding import (Engine, Q_INT8, T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3,
                      T_OUT_NORM, T_CLS)
from .loaders import quantize_rows

MAGIC = 0x46554747
(V_UINT8, V_INT8, V_UINT16, V_INT16, V_UINT32, V_INT32, V_FLOAT32, V_BOOL, V_STRING, V_ARRAY, V_UINT64, V_INT64,
 V_FLOAT64) = range(13)
_FMT = {V_UINT8: "B", V_INT8: "b", V_UINT16: "H", V_INT16: "h", V_UINT32: "I", V_INT32: "i", V_FLOAT32: "f", V_BOOL: "B",
        V_UINT64: "Q", V_INT64: "q", V_FLOAT64: "d"}
GGML_F32, GGML_F16, GGML_Q8_0 = 0, 1, 8
FILE_TYPES = {0: "f32", 1: "f16", 7: "q8_0"}

# key -> config field (None: tokenizer / bookkeeping), exactly the set gguf_loader.cpp:244-322 accepts
KNOWN_KEYS = {
    "general.architecture": None, "general.name": "name", "general.file_type": "file_type",
    "general.quant_group_size": "quant_group_size", "general.alignment": "alignment",
    "llama.context_length": "max_seq_len", "llama.embedding_length": "dim", "llama.block_count": "n_layers",
    "llama.feed_forward_length": "hidden_dim", "llama.attention.head_count": "n_heads",
    "llama.attention.head_count_kv": "n_kv_heads", "llama.rope.dimension_count": "rope_dimension_count",
    "llama.rope.freq_base": "rope_freq_base", "llama.attention.layer_norm_rms_epsilon": "rms_norm_eps",
    "tokenizer.ggml.model": None, "tokenizer.ggml.bos_token_id": None, "tokenizer.ggml.eos_token_id": None,
    "tokenizer.ggml.padding_token_id": None, "tokenizer.ggml.tokens": None, "tokenizer.ggml.scores": None,
    "tokenizer.ggml.token_type": None,
}
GLOBAL_TENSORS = {"token_embd.weight": T_TOK_EMB, "output_norm.weight": T_OUT_NORM, "output.weight": T_CLS}
LAYER_TENSORS = {"attn_q": T_WQ, "attn_k": T_WK, "attn_v": T_WV, "attn_output": T_WO, "attn_norm": T_ATT_NORM,
                 "ffn_gate": T_W1, "ffn_down": T_W2, "ffn_up": T_W3, "ffn_norm": T_FFN_NORM}


class GgufError(ValueError):
    pass


class _Cursor:
    def __init__(self, buf):
        self.buf, self.pos = buf, 0

    def take(self, fmt):
        size = struct.calcsize("<" + fmt)
        if self.pos + size > len(self.buf):
            raise GgufError("unexpected end of file in the metadata")
        v = struct.unpack_from("<" + fmt, self.buf, self.pos)
        self.pos += size
        return v[0] if len(v) == 1 else v

    def string(self):
        n = self.take("Q")
        if self.pos + n > len(self.buf):
            raise GgufError("string runs past the end of the file")
        s = bytes(self.buf[self.pos:self.pos + n])
        self.pos += n
        return s.decode("utf-8", "surrogateescape")      # vocabulary pieces need not be valid UTF-8: keep their bytes

    def value(self, vtype):
        if vtype == V_STRING:
            return self.string()
        if vtype == V_ARRAY:
            et, n = self.take("I"), self.take("Q")
            if et == V_STRING:
                return [self.string() for _ in range(n)]
            if et not in _FMT:
                raise GgufError(f"array of value type {et} is not supported")
            dt = np.dtype("<" + {"B": "u1", "b": "i1", "H": "u2", "h": "i2", "I": "u4", "i": "i4", "f": "f4", "Q": "u8",
                                 "q": "i8", "d": "f8"}[_FMT[et]])
            if self.pos + n * dt.itemsize > len(self.buf):
                raise GgufError("array runs past the end of the file")
            a = np.frombuffer(self.buf, dt, n, self.pos)
            self.pos += n * dt.itemsize
            return a
        if vtype not in _FMT:
            raise GgufError(f"unknown value type {vtype}")
        return self.take(_FMT[vtype])


def split_q8_0(raw, rows, cols):
    """Q8_0 blocks {fp16 d, int8 q[32]} of a [rows][cols] tensor -> (int8 [rows][cols], float32 scales [rows][cols/32])"""
    if cols % 32:
        raise GgufError(f"Q8_0 tensor with {cols} columns (not a multiple of 32)")
    blocks = np.frombuffer(raw, np.dtype([("d", "<f2"), ("q", "i1", 32)]), rows * cols // 32)
    return (np.ascontiguousarray(blocks["q"]).reshape(rows, cols),
            blocks["d"].astype(np.float32).reshape(rows, cols // 32))           # fp16 -> fp32 is exact


def read_gguf(path, strict=True, tensors=True):
    """see _read_gguf; any parsing failure on a malformed file surfaces as GgufError (the reference's loader returns false)"""
    try:
        return _read_gguf(path, strict, tensors)
    except GgufError:
        raise
    except (ValueError, IndexError, OverflowError, KeyError, struct.error, MemoryError) as e:
        raise GgufError(f"{path}: malformed GGUF file ({type(e).__name__}: {e})") from e


def _read_gguf(path, strict=True, tensors=True):
    """-> (cfg dict, {(engine kind, layer): (payload, scales | None)}, vocab dict)"""
    buf = np.memmap(path, np.uint8, "r")
    c = _Cursor(buf)
    if len(buf) < 24:
        raise GgufError(f"{path}: not a GGUF file (too short)")
    magic, version, n_tensors, n_kv = c.take("IIqq")
    if magic != MAGIC or n_tensors < 1 or n_kv < 1:                              # :231-234
        raise GgufError(f"{path}: not a valid GGUF file")
    cfg = dict(version=version, alignment=32, quant_group_size=32, name="", file_type=0, n_kv_heads=0, max_seq_len=0,
               rope_dimension_count=0, rope_freq_base=10000.0, rms_norm_eps=1e-5)
    vocab = dict(special={})
    for _ in range(n_kv):
        key = c.string()
        vtype = c.take("I")
        val = c.value(vtype)
        if key not in KNOWN_KEYS:
            if strict:
                raise GgufError(f"{path}: unknown key {key!r} (the reference rejects it, gguf_loader.cpp:319-322; "
                                "pass strict=False to skip it)")
            continue
        if key == "general.architecture":
            if val != "llama":
                raise GgufError(f"{path}: unsupported architecture {val!r}")
        elif key == "general.file_type":
            if val not in FILE_TYPES:
                raise GgufError(f"{path}: unsupported file type {val}")
            cfg["file_type"] = int(val)
        elif KNOWN_KEYS[key]:
            cfg[KNOWN_KEYS[key]] = val if isinstance(val, (str, float)) else int(val)
        elif key == "tokenizer.ggml.model":
            vocab["model"] = val
        elif key == "tokenizer.ggml.tokens":
            vocab["texts"] = val
            cfg["vocab_size"] = len(val)
        elif key == "tokenizer.ggml.scores":
            vocab["scores"] = np.asarray(val, np.float32).tolist()
        elif key == "tokenizer.ggml.token_type":
            vocab["types"] = np.asarray(val, np.int64).tolist()
        else:
            vocab["special"][key.split(".")[-1].split("_")[0].replace("padding", "pad")] = int(val)
    need = ("dim", "n_layers", "hidden_dim", "n_heads", "vocab_size")
    if any(k not in cfg for k in need):
        raise GgufError(f"{path}: missing one of {need} in the metadata")
    hs = cfg["dim"] // cfg["n_heads"] if cfg["n_heads"] > 0 else 0
    if (cfg["n_heads"] < 1 or cfg["n_kv_heads"] < 1 or cfg["dim"] % cfg["n_heads"]
            or (cfg["rope_dimension_count"] > 0 and cfg["rope_dimension_count"] != hs)):      # :331-336
        raise GgufError(f"{path}: invalid dim / n_heads / n_kv_heads / rope dimension count")
    q8 = FILE_TYPES[cfg["file_type"]] == "q8_0"
    cfg["quant_type"] = Q_INT8 if q8 else 0
    if not q8:
        cfg["quant_group_size"] = 64                                             # :337-341
    out = {}
    if not tensors:
        return cfg, out, vocab
    infos, seen = [], set()
    for _ in range(n_tensors):
        name = c.string()
        kind, layer = GLOBAL_TENSORS.get(name), 0
        if kind is None:
            p = name.split(".")
            if len(p) == 4 and p[0] == "blk" and p[1].isdigit() and int(p[1]) < cfg["n_layers"] and p[3] == "weight":
                kind, layer = LAYER_TENSORS.get(p[2]), int(p[1])
        if kind is None:
            raise GgufError(f"{path}: invalid tensor name {name!r}")
        n_dims = c.take("I")
        if not 1 <= n_dims <= 2:
            raise GgufError(f"{path}: invalid shape of tensor {name}")
        dims = [c.take("Q") for _ in range(n_dims)]
        dtype, offset = c.take("I"), c.take("Q")
        if dtype not in (GGML_F32, GGML_Q8_0):
            raise GgufError(f"{path}: tensor {name}: data type {dtype} is not supported (F32 and Q8_0 only)")
        if name in seen:
            continue                                                             # ":Duplicated tensor info": first wins
        seen.add(name)
        infos.append((name, kind, layer, dims, dtype, offset))
    base = (c.pos + cfg["alignment"] - 1) // cfg["alignment"] * cfg["alignment"]
    for name, kind, layer, dims, dtype, offset in infos:
        cols, rows = dims[0], dims[1] if len(dims) > 1 else 1
        n = rows * cols
        size = n * 4 if dtype == GGML_F32 else n + n * 2 // 32
        lo = base + offset
        if lo + size > len(buf):
            raise GgufError(f"{path}: tensor {name} runs past the end of the file")
        if dtype == GGML_F32:
            a = buf[lo:lo + size].view(np.float32)
            out[(kind, layer)] = (a.reshape(rows, cols) if len(dims) > 1 else a, None)
        else:
            if cfg["quant_group_size"] != 32:
                raise GgufError(f"{path}: Q8_0 tensors need quant_group_size 32, the file says {cfg['quant_group_size']}")
            out[(kind, layer)] = split_q8_0(buf[lo:lo + size], rows, cols)
    return cfg, out, vocab


def engine_from_gguf(path, quant_type=Q_INT8, max_seq_len=1024, device=0, strict=True, **engine_kw):
    """Load a GGUF the way `main -c model.gguf [-q int8]` does and return (finalized Engine, cfg, vocab).  Q8_0 tensors
    are uploaded as they are (group 32); F32 matrices are quantised at load with `quant_type`, group 64, as the reference's
    worker initialisation does (transformer.cpp:289-304); the embedding table and the norms stay as stored."""
    cfg, t, vocab = read_gguf(path, strict=strict)
    gs = cfg["quant_group_size"]
    qt = cfg["quant_type"] or quant_type
    eng = Engine(cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["vocab_size"],
                 max_seq_len=max_seq_len, quant_type=qt, group_size=gs, device=device, **engine_kw)
    for (kind, layer), (q, s) in t.items():
        q = np.ascontiguousarray(q)
        if s is None and q.ndim == 2 and kind != T_TOK_EMB:
            q, s = quantize_rows(q, qt, gs)
        eng.upload(kind, layer, q, s)
    eng.finalize()
    return eng, cfg, vocab
