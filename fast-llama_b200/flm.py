"""The `.flm` model file (BASELINE.json configs 2-4 ship as .flm): reader, writer and engine upload.

Layout, as written by the reference's converter (tools/convert_flm.py:465-748 `FLFWriter`, :1075-1172) and read by its
loader (src/model_loaders/flm_loader.cpp:115-341 block header, :493-634 load_tensor / load_flm).  Little endian.

  file header   u32 0xFA571AEA, u8 major, u8 minor, u16 patch                                  (8 bytes)
  then blocks until end of file.  Every block starts with  u8 block_type, u8 data_type, u8 header_size, u8 header_data_size.
    base item (block_type 0): a named scalar that lives entirely in its header.  header_data_size is the value's size;
        <= 4 bytes: value at offset 4, name at 8;  8 bytes: u32 pad, value at 8, name at 16.  The name is NUL terminated and
        header_size is rounded up to 8.
    every other block: u8 name_offset, u8 name_size, u16 tail_pad, u64 data_size, [header data padded to 8], name + NUL,
        padding so that the DATA starts on the block's alignment (8, tensors 64) in the file, data, tail padding up to a
        multiple of the alignment of header_size + data_size.
  "model_config" (DICT): its data is a run of base items / STRING blocks named after the fields of the converter's
        ModelConfig (:331-352); the loader picks the names it knows (flm_loader.cpp:389-436).
  "tokenizer" (DICT): packed struct {u32 vocab_type, u32 conn_tag_pos, i32 special[8], u32 vocab_size, u32 text_size,
        {u32 index_text_pos, u32 show_text_pos, u32 type, f32 score}[vocab_size], text bytes} (:942-1016, loader :438-491).
  tensors (block_type 2): header data {u32 shape[4], u16 tensor_type, u16 layer_id, u32 n_scales}; data = payload
        immediately followed by the fp32 group scales.  One block per (tensor, layer); a layer tensor's layer-0 block
        must come before its other layers (the loader allocates on layer_id <= 0, :533-551).  The embedding table and all
        1-D tensors stay fp32; matrices are int8/int16 with `quant_group_size` groups along the columns.
"""
import struct

import numpy as np

from .binding import (Engine, Q_INT8, Q_INT16, T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3,
                      T_OUT_NORM, T_CLS)

FILE_TAG = 0xFA571AEA
B_ITEM, B_DICT, B_TENSOR, B_ARRAY, B_STRING, B_STRING_ARRAY = range(6)
D_NONE, D_INT8, D_INT16, D_INT32, D_INT64, D_UINT8, D_UINT16, D_UINT32, D_UINT64 = range(9)
D_FLOAT16, D_FLOAT32, D_FLOAT64 = 10, 11, 12
_NP_OF = {D_INT8: np.int8, D_INT16: np.int16, D_FLOAT32: np.float32}
_D_OF = {np.dtype(np.int8): D_INT8, np.dtype(np.int16): D_INT16, np.dtype(np.float32): D_FLOAT32}

# tensor_type (flm_loader.cpp:50-66) <-> engine tensor kind, and the checkpoint name the converter stores with it (:1125-1143)
TENSOR_TYPES = {
    1: (T_TOK_EMB, "model.embed_tokens.weight"), 2: (T_OUT_NORM, "model.norm.weight"), 3: (T_CLS, "lm_head.weight"),
    17: (T_ATT_NORM, "model.layers.{}.input_layernorm.weight"), 18: (T_WQ, "model.layers.{}.self_attn.q_proj.weight"),
    19: (T_WK, "model.layers.{}.self_attn.k_proj.weight"), 20: (T_WV, "model.layers.{}.self_attn.v_proj.weight"),
    21: (T_WO, "model.layers.{}.self_attn.o_proj.weight"), 22: (T_W1, "model.layers.{}.mlp.gate_proj.weight"),
    23: (T_W3, "model.layers.{}.mlp.up_proj.weight"), 24: (T_W2, "model.layers.{}.mlp.down_proj.weight"),
    25: (T_FFN_NORM, "model.layers.{}.post_attention_layernorm.weight"),
}
KIND_TO_TYPE = {kind: tt for tt, (kind, _) in TENSOR_TYPES.items()}
# the order the converter's ModelConfig serialises its fields in (:331-352); int -> int32 item, float -> float32 item
CONFIG_FIELDS = ("name", "model_type", "act_type", "quant_type", "vocab_size", "dim", "hidden_dim", "n_heads", "n_kv_heads",
                 "n_layers", "max_length", "bos_token_id", "eos_token_id", "pad_token_id", "rms_norm_eps", "rope_theta",
                 "quant_group_size")
CONFIG_DEFAULTS = dict(name="", model_type=1, act_type=2, quant_type=0, vocab_size=0, dim=0, hidden_dim=0, n_heads=0,
                       n_kv_heads=0, n_layers=0, max_length=0, bos_token_id=0, eos_token_id=0, pad_token_id=0,
                       rms_norm_eps=0.0, rope_theta=10000.0, quant_group_size=64)
CONN_TAG = "▁"


def _pad(n, align):
    return (-n) % align


def permute_qk(w, n_heads, n_kv_heads=None):
    """HF checkpoints keep each head's rotary halves apart; the reference's RoPE is interleaved, so its converter reorders
    the rows of Wq / Wk (tools/convert_flm.py:1018-1023).  Wk uses n_kv_heads."""
    h = n_kv_heads if n_kv_heads is not None and n_kv_heads != n_heads else n_heads
    return w.reshape(h, 2, w.shape[0] // h // 2, *w.shape[1:]).swapaxes(1, 2).reshape(w.shape)


# ------------------------------------------------------------------------------------------------- writer
class _Out:
    """byte sink that knows its position: in memory, or straight into a file (a 7B model is 7.4 GB)"""

    def __init__(self, f=None):
        self.parts, self.pos, self.f = [], 0, f

    def put(self, b):
        if self.f is not None:
            self.f.write(b)
        else:
            self.parts.append(b)
        self.pos += len(b)

    def bytes(self):
        return b"".join(self.parts)


def _item(name, fmt, dtype, value):
    """a named scalar (base item block)"""
    data = struct.pack("<" + fmt, value)
    nm = name.encode("utf-8") + b"\0"
    wide = len(data) > 4
    size = (16 if wide else 8) + len(nm)
    header_size = (size + 7) & ~7
    body = (b"\0" * 4 + data.ljust(8, b"\0")) if wide else data.ljust(4, b"\0")
    return struct.pack("<4B", B_ITEM, dtype, header_size, len(data)) + body + nm + b"\0" * (header_size - size)


def _block(out, name, data, block_type, data_type=D_NONE, align=8, header_data=b""):
    """a block whose data follows the header; `out.pos` is the block's file offset (it decides the head padding)"""
    nm = name.encode("utf-8")
    header_data = header_data + b"\0" * _pad(len(header_data), 8)
    name_offset = 16 + len(header_data)
    header_size = name_offset + len(nm) + 1
    header_size += _pad(out.pos + header_size, align)
    if header_size > 255:
        raise ValueError(f"block name {name!r} is too long")
    parts = data if isinstance(data, list) else [data]
    n_data = sum(len(d) for d in parts)
    tail = _pad(header_size + n_data, align)
    head = struct.pack("<6BHQ", block_type, data_type, header_size, len(header_data), name_offset, len(nm), tail, n_data)
    head += header_data + nm + b"\0"
    out.put(head + b"\0" * (header_size - len(head)))
    for d in parts:
        out.put(d)
    out.put(b"\0" * tail)


def _config_bytes(cfg):
    out = _Out()
    for k in CONFIG_FIELDS:
        v = cfg.get(k, CONFIG_DEFAULTS[k])
        # the converter picks the item type from the VALUE (ModelConfig.serialize_as_flf, :386-398): a config.json whose
        # hidden_act is "silu" leaves a string in act_type, and that is what ends up in the file
        if type(v) is str:
            _block(out, k, v.encode("utf-8") + b"\0", B_STRING, D_INT8)      # BlockDataType.CHAR == 1
        elif type(v) is float:
            out.put(_item(k, "f", D_FLOAT32, v))
        elif type(v) is int:
            out.put(_item(k, "i", D_INT32, v))
    return out.bytes()


def _tokenizer_bytes(vocab):
    """vocab = dict(vocab_type (1 bpe / 2 spm), texts [str], scores [float], types [int], special {name: id})"""
    def enc(t):
        b = t.encode("utf-8", "surrogateescape") + b"\0"
        return b + b"\0" * _pad(len(b), 8)
    toks, text = [], b""
    for t, score, typ in zip(vocab["texts"], vocab["scores"], vocab["types"]):
        index_pos = len(text)
        text += enc(t)
        show_pos = index_pos
        if t.startswith(CONN_TAG):                      # the piece is shown with a leading space instead of the tag
            show_pos = len(text)
            text += enc(" " + t[len(CONN_TAG):])
        toks.append(struct.pack("<3if", index_pos, show_pos, typ, score))
    conn_pos = len(text)
    text += enc(CONN_TAG)
    special = [-1] * 8
    for name, tid in vocab.get("special", {}).items():
        special[{"bos": 1, "eos": 2, "pad": 3}[name]] = tid
    return (struct.pack("<2I8i2I", vocab["vocab_type"], conn_pos, *special, len(vocab["texts"]), len(text))
            + b"".join(toks) + text)


def write_flm(path, cfg, tensors, vocab=None, version=(1, 0, 0)):
    """cfg: dict with CONFIG_FIELDS keys; tensors: {(engine kind, layer): (payload ndarray, scales ndarray | None)}
    (the layout quantize_rows / fl_upload use) or a callable (kind, layer) -> (payload, scales) that produces them on demand;
    vocab: see _tokenizer_bytes (None writes no tokenizer block).
    Blocks are written in the converter's order: embedding, layers 0..L-1 (q k v o gate up down, the two norms),
    output norm, classifier, each straight into the file."""
    with open(path, "wb") as f:
        _write_flm(_Out(f), cfg, tensors, vocab, version)


def _write_flm(out, cfg, tensors, vocab, version):
    out.put(struct.pack("<I2BH", FILE_TAG, *version))
    _block(out, "model_config", _config_bytes(cfg), B_DICT)
    if vocab is not None:
        _block(out, "tokenizer", _tokenizer_bytes(vocab), B_DICT)

    def tensor(kind, layer):
        q, s = tensors(kind, layer) if callable(tensors) else tensors[(kind, layer)]
        q = np.ascontiguousarray(q)
        tt = KIND_TO_TYPE[kind]
        shape = list(q.shape) + [0] * (4 - q.ndim)
        data = [memoryview(q).cast("B")]
        n_scales = 0
        if s is not None:
            s = np.ascontiguousarray(s, np.float32)
            data.append(memoryview(s).cast("B"))
            n_scales = s.size
        hd = struct.pack("<4I2HI", *shape, tt, layer, n_scales)
        _block(out, TENSOR_TYPES[tt][1].format(layer), data, B_TENSOR, _D_OF[q.dtype], 64, hd)

    tensor(T_TOK_EMB, 0)
    for l in range(cfg["n_layers"]):
        for kind in (T_WQ, T_WK, T_WV, T_WO, T_W1, T_W3, T_W2, T_ATT_NORM, T_FFN_NORM):
            tensor(kind, l)
    tensor(T_OUT_NORM, 0)
    tensor(T_CLS, 0)


# ------------------------------------------------------------------------------------------------- reader
class FlmError(ValueError):
    pass


def _read_block_header(buf, pos):
    """-> dict(type, dtype, header_size, name, data_off, data_size, size, value?, tensor fields?)"""
    if pos + 8 > len(buf):
        raise FlmError(f"block header at {pos} runs past the end of the file")
    btype, dtype, hsize, hdsize = struct.unpack_from("<4B", buf, pos)
    if hsize < 8 or pos + hsize > len(buf):
        raise FlmError(f"bad block header at {pos}")
    b = dict(type=btype, dtype=dtype, header_size=hsize)
    if btype == B_ITEM:
        off = 8 if hdsize <= 4 else 16
        raw = bytes(buf[pos + (4 if hdsize <= 4 else 8):pos + (4 if hdsize <= 4 else 8) + hdsize])
        b["name"] = bytes(buf[pos + off:pos + hsize]).split(b"\0", 1)[0].decode("utf-8")
        if dtype in (D_FLOAT32, D_FLOAT64):
            b["value"] = struct.unpack("<f" if hdsize <= 4 else "<d", raw.ljust(4 if hdsize <= 4 else 8, b"\0"))[0]
        else:
            # flm_loader.cpp:246 reads every small item as its int32 / int64 union member
            b["value"] = int.from_bytes(raw.ljust(4 if hdsize <= 4 else 8, b"\0"), "little", signed=True)
        b["size"] = hsize
        return b
    name_off, name_size, tail, data_size = struct.unpack_from("<2BHQ", buf, pos + 4)
    b["name"] = bytes(buf[pos + name_off:pos + name_off + name_size]).decode("utf-8")
    b["data_off"] = pos + hsize
    b["data_size"] = data_size
    b["size"] = hsize + data_size + tail
    if b["data_off"] + data_size > len(buf):
        raise FlmError(f"block {b['name']!r} at {pos} runs past the end of the file")
    if btype == B_TENSOR:
        *shape, tt, layer, n_scales = struct.unpack_from("<4I2HI", buf, pos + 16)
        b.update(shape=[s for s in shape if s > 0], tensor_type=tt, layer=layer, n_scales=n_scales)
    return b


def read_flm(path, tensors=True):
    """see _read_flm; any parsing failure on a malformed file surfaces as FlmError (the reference's loader returns false)"""
    try:
        return _read_flm(path, tensors)
    except FlmError:
        raise
    except (ValueError, IndexError, OverflowError, KeyError, struct.error, MemoryError) as e:
        raise FlmError(f"{path}: malformed .flm file ({type(e).__name__}: {e})") from e


def _read_flm(path, tensors=True):
    """-> (cfg dict, {(engine kind, layer): (payload, scales | None)}, vocab dict | None).  Arrays are views into one
    memory map of the file (no copy until upload)."""
    buf = np.memmap(path, np.uint8, "r")
    if len(buf) < 8 or struct.unpack_from("<I", buf, 0)[0] != FILE_TAG:
        raise FlmError(f"{path}: not an .flm file (bad tag)")
    cfg, out, vocab = dict(CONFIG_DEFAULTS), {}, None
    cfg["version"] = struct.unpack_from("<2BH", buf, 4)
    pos = 8
    while pos < len(buf):
        b = _read_block_header(buf, pos)
        if b["name"] == "model_config":
            p, end = b["data_off"], b["data_off"] + b["data_size"]
            while p < end:
                it = _read_block_header(buf, p)
                if it["type"] == B_ITEM:
                    cfg[it["name"]] = it["value"]
                elif it["type"] == B_STRING:
                    cfg[it["name"]] = bytes(buf[it["data_off"]:it["data_off"] + it["data_size"]]).split(b"\0", 1)[0].decode("utf-8")
                p += it["size"]
            if cfg["n_kv_heads"] < 1:                                        # flm_loader.cpp:424-431
                cfg["n_kv_heads"] = cfg["n_heads"]
            if cfg["n_heads"] < 1 or cfg["dim"] % cfg["n_heads"] or cfg["n_kv_heads"] > cfg["n_heads"]:
                raise FlmError(f"{path}: invalid model_config")
        elif b["name"] == "tokenizer":
            vocab = _parse_tokenizer(buf, b["data_off"])
        elif b["type"] == B_TENSOR and tensors:
            if b["tensor_type"] not in TENSOR_TYPES:
                raise FlmError(f"{path}: unsupported tensor type {b['tensor_type']} ({b['name']})")
            if b["dtype"] not in _NP_OF:
                raise FlmError(f"{path}: unsupported tensor data type {b['dtype']} ({b['name']})")
            dt = np.dtype(_NP_OF[b["dtype"]])
            n = int(np.prod(b["shape"]))
            if n * dt.itemsize + 4 * b["n_scales"] != b["data_size"]:
                raise FlmError(f"{path}: tensor {b['name']} has {b['data_size']} data bytes, its shape needs "
                               f"{n * dt.itemsize + 4 * b['n_scales']}")
            q = buf[b["data_off"]:b["data_off"] + n * dt.itemsize].view(dt).reshape(b["shape"])
            s = None
            if b["n_scales"]:
                so = b["data_off"] + n * dt.itemsize
                s = buf[so:so + 4 * b["n_scales"]].view(np.float32).reshape(b["shape"][0], -1)
            out[(TENSOR_TYPES[b["tensor_type"]][0], b["layer"])] = (q, s)
        pos += b["size"]
    return cfg, out, vocab


def _parse_tokenizer(buf, off):
    vocab_type, conn_pos, *rest = struct.unpack_from("<2I8i2I", buf, off)
    special, (n, text_size) = rest[:8], rest[8:]
    items = np.frombuffer(buf, np.dtype([("index", "<u4"), ("show", "<u4"), ("type", "<u4"), ("score", "<f4")]), n, off + 48)
    tb = off + 48 + 16 * n
    text = bytes(buf[tb:tb + text_size])

    def s(p):
        return text[p:text.index(b"\0", p)].decode("utf-8", "surrogateescape")      # pieces need not be valid UTF-8
    return dict(vocab_type=vocab_type, texts=[s(int(p)) for p in items["index"]], show=[s(int(p)) for p in items["show"]],
                scores=items["score"].tolist(), types=items["type"].tolist(), conn_tag=s(conn_pos),
                special={k: special[i] for k, i in (("bos", 1), ("eos", 2), ("pad", 3)) if special[i] >= 0})


def engine_from_flm(path, quant_type=Q_INT8, max_seq_len=1024, device=0, **engine_kw):
    """Load an .flm the way `main -c model.flm [-q int8]` does and return (finalized Engine, cfg, vocab).  An int8 / int16
    file is uploaded as stored; an f32 file (quant_type 0 in its config) has its matrices quantised at load with
    `quant_type` and the file's group size, exactly what the reference's worker initialisation does
    (transformer.cpp:289-304: tgt.quantize(src)); the embedding table and the norms stay fp32.  The reference caps the
    context at 1024 whatever the file says (transformer.cpp:32); pass max_seq_len to lift that."""
    from .loaders import quantize_rows
    cfg, t, vocab = read_flm(path)
    qt = cfg["quant_type"] or quant_type
    if qt not in (Q_INT8, Q_INT16):
        raise FlmError(f"{path}: quant_type {qt} is not supported (int8 / int16)")
    gs = cfg["quant_group_size"]
    eng = Engine(cfg["dim"], cfg["hidden_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["vocab_size"],
                 max_seq_len=max_seq_len, quant_type=qt, group_size=gs, device=device, **engine_kw)
    for (kind, layer), (q, s) in t.items():
        q = np.ascontiguousarray(q)
        if s is None and q.ndim == 2 and kind != T_TOK_EMB:
            q, s = quantize_rows(q, qt, gs)
        eng.upload(kind, layer, q, None if s is None else np.ascontiguousarray(s))
    eng.finalize()
    return eng, cfg, vocab
