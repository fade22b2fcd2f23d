"""HF LLaMA checkpoint directory -> .flm, the job of the reference's tools/convert_flm.py (`main`, :1174-1203;
`ModelConverter.load/dump`, :1030-1172), on top of flm.write_flm.

What is taken over from the converter, quirks included (they decide the bytes of the file):
  * config.json -> model_config (`ModelConfig.load`, :354-384): hidden_size -> dim, intermediate_size -> hidden_dim,
    num_attention_heads / num_key_value_heads / num_hidden_layers, max_position_embeddings -> max_length, _name_or_path -> name,
    hidden_act copied AS A STRING into act_type (so the file holds a string item "silu" where the loader reads an integer;
    the loader ignores the activation type anyway); a key that carries a field's
    own name (vocab_size, bos/eos/pad_token_id, rms_norm_eps, rope_theta) is copied only when its JSON type equals the
    field's (an integer rope_theta is ignored);
  * tokenizer.model (sentencepiece) -> vocabulary (`load_spm_vocab`, :793-836): piece, score, type UNKNOWN 0 / NORMAL 1 /
    CONTROL 2 / BYTE 3 / UNUSED 5; special ids from config.json's bos/eos/pad_token_id when >= 0 (:925-938);
  * tensors in checkpoint order (`_dump_tensors`, :1109-1172): q_proj / k_proj rows permuted for the interleaved RoPE
    (`permute_qk`), every 2-D tensor except the embedding quantised with numpy float32 arithmetic and C truncation, group 64
    (the converter hard-codes 64 whatever -g says, :1111); 1-D tensors and the embedding stay fp32.
    Byte identity with the converter holds for FP32 checkpoints (what the tests cover): every tensor is widened to float32
    before quantisation here, while `TensorLoader.quantize` (:216-243) works in the checkpoint's own dtype, so for an fp16
    (HalfStorage) checkpoint its max / 127 and division happen in float16 and payload and scales can differ from this path.
Only single-file PyTorch checkpoints (pytorch_model.bin / *.pt / consolidated.00.pth) are read, via torch.load."""
import json
import os

import numpy as np

from . import flm
from .binding import Q_INT8, Q_INT16
from .loaders import quantize_rows

_NAMES = {"model.embed_tokens.weight": (1, None), "model.norm.weight": (2, None), "lm_head.weight": (3, None),
          "input_layernorm": (17, None), "self_attn.q_proj": (18, "q"), "self_attn.k_proj": (19, "k"), "self_attn.v_proj": (20, None),
          "self_attn.o_proj": (21, None), "mlp.gate_proj": (22, None), "mlp.up_proj": (23, None), "mlp.down_proj": (24, None),
          "post_attention_layernorm": (25, None)}


def config_from_hf(conf, quant_type, group_size=64):
    """config.json dict -> the .flm model_config dict (flm.CONFIG_FIELDS)"""
    cfg = dict(flm.CONFIG_DEFAULTS)
    cfg["quant_type"], cfg["quant_group_size"] = quant_type, group_size
    renamed = {"_name_or_path": "name", "vocab_size": "vocab_size", "hidden_size": "dim", "intermediate_size": "hidden_dim",
               "num_attention_heads": "n_heads", "num_key_value_heads": "n_kv_heads", "num_hidden_layers": "n_layers",
               "hidden_act": "act_type", "max_position_embeddings": "max_length"}
    enums = ("model_type", "act_type", "quant_type")
    for k, v in conf.items():
        if k in cfg:
            if k in enums:
                if isinstance(v, str):
                    cfg[k] = {"model_type": {"NONE": 0, "LLAMA": 1}, "act_type": {"NONE": 0, "SILU": 1, "SWIGLU": 2},
                              "quant_type": {"NONE": 0, "INT16": 1, "INT8": 2, "INT4": 3}}[k][v.upper()]
            elif type(cfg[k]) is type(v):
                cfg[k] = v
        elif k in renamed:
            cfg[renamed[k]] = v          # as is: hidden_act "silu" stays a STRING in act_type (and is written as one)
    return cfg


def vocab_from_spm(path, conf=None):
    from sentencepiece import SentencePieceProcessor
    sp = SentencePieceProcessor(str(path))
    texts, scores, types = [], [], []
    for i in range(sp.vocab_size()):
        t = 1
        if sp.is_unknown(i):
            t = 0
        if sp.is_control(i):
            t = 2
        if sp.is_unused(i):
            t = 5
        if sp.is_byte(i):
            t = 3
        texts.append(sp.id_to_piece(i))
        scores.append(sp.get_score(i))
        types.append(t)
    special = {}
    for name in ("bos", "eos", "pad"):
        v = (conf or {}).get(f"{name}_token_id", -1)
        if isinstance(v, int) and v >= 0:
            special[name] = v
    return dict(vocab_type=2, texts=texts, scores=scores, types=types, special=special)


def tensors_from_state_dict(sd, cfg):
    """{HF name: float array} -> flm tensor dict, in the converter's arithmetic"""
    qt, gs = cfg["quant_type"], 64
    out = {}
    for name, a in sd.items():
        a = np.asarray(a, np.float32)
        if name in _NAMES:
            (tt, perm), layer = _NAMES[name], 0
        elif name.startswith("model.layers."):
            parts = name.split(".")
            layer = int(parts[2])
            key = ".".join(parts[3:]).removesuffix(".weight")
            if key not in _NAMES:
                continue                                          # the converter prints "Unknown tensor name" and goes on
            tt, perm = _NAMES[key]
        else:
            raise ValueError(f"Unknown tensor name:[{name}], shape:{a.shape}")
        if perm:
            a = flm.permute_qk(a, cfg["n_heads"], cfg["n_kv_heads"])
        kind = flm.TENSOR_TYPES[tt][0]
        if tt != 1 and a.ndim > 1 and qt in (Q_INT8, Q_INT16):
            out[(kind, layer)] = quantize_rows(a, qt, gs)
        else:
            out[(kind, layer)] = (np.ascontiguousarray(a), None)
    return out


def convert_hf_to_flm(model_dir, out_path, out_type="int8"):
    """`convert_flm.py -m model_dir -t out_type -o out_path` for a sentencepiece (spm) LLaMA checkpoint"""
    import torch
    qt = {"f32": 0, "int16": Q_INT16, "int8": Q_INT8}[out_type]
    with open(os.path.join(model_dir, "config.json")) as f:
        conf = json.load(f)
    cfg = config_from_hf(conf, qt)
    vocab = vocab_from_spm(os.path.join(model_dir, "tokenizer.model"), conf)
    files = [f for f in sorted(os.listdir(model_dir)) if f in ("consolidated.00.pth", "pytorch_model.bin") or f.endswith(".pt")]
    if len(files) != 1:
        raise ValueError(f"expected exactly one PyTorch checkpoint in {model_dir}, found {files}")
    sd = torch.load(os.path.join(model_dir, files[0]), map_location="cpu", weights_only=True)
    sd = {k: v.to(torch.float32).numpy() for k, v in sd.items()}
    flm.write_flm(out_path, cfg, tensors_from_state_dict(sd, cfg), vocab)
    return cfg
