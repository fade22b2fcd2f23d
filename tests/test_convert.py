"""HF checkpoint directory -> .flm (fast-llama_b200/convert.py) against the reference converter's own classes
(tools/convert_flm.py: ModelConfig.load + serialize_as_flf, Tokenizer.load + serialize_as_flf, TensorLoader.quantize, permute_qk,
FLFWriter), executed here from /root/reference (the module's head, see tests/golden/make_flm_golden.py): the two files
must be byte-identical.  Skipped where the reference tree is absent."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TOOL = "/root/reference/tools/convert_flm.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="reference tree not present")


def make_hf_dir(d, n_kv_heads=2):
    import sentencepiece as spm
    import torch
    r = np.random.default_rng(11)
    words = ["".join(r.choice(list("abcdefghijklmnop"), int(r.integers(2, 7)))) for _ in range(400)]
    (d / "corpus.txt").write_text("\n".join(" ".join(r.choice(words, 12)) for _ in range(600)))
    spm.SentencePieceTrainer.train(input=str(d / "corpus.txt"), model_prefix=str(d / "tokenizer"), vocab_size=384, model_type="bpe",
                                   byte_fallback=True, character_coverage=1.0, bos_id=1, eos_id=2, unk_id=0, pad_id=-1,
                                   minloglevel=2)
    dim, hidden, L, heads = 128, 192, 2, 4
    conf = {"_name_or_path": "tiny-llama", "architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_act": "silu",
            "hidden_size": dim, "intermediate_size": hidden, "num_attention_heads": heads, "num_key_value_heads": n_kv_heads,
            "num_hidden_layers": L, "max_position_embeddings": 2048, "rms_norm_eps": 1e-05, "rope_theta": 10000.0,
            "vocab_size": 384, "bos_token_id": 1, "eos_token_id": 2, "torch_dtype": "float16", "use_cache": True}
    (d / "config.json").write_text(json.dumps(conf))
    kv = dim // heads * n_kv_heads
    sd = {"model.embed_tokens.weight": (384, dim)}
    for l in range(L):
        p = f"model.layers.{l}."
        sd.update({p + "self_attn.q_proj.weight": (dim, dim), p + "self_attn.k_proj.weight": (kv, dim), p + "self_attn.v_proj.weight": (kv, dim),
                   p + "self_attn.o_proj.weight": (dim, dim), p + "mlp.gate_proj.weight": (hidden, dim), p + "mlp.up_proj.weight": (hidden, dim),
                   p + "mlp.down_proj.weight": (dim, hidden), p + "input_layernorm.weight": (dim,), p + "post_attention_layernorm.weight": (dim,)})
    sd.update({"model.norm.weight": (dim,), "lm_head.weight": (384, dim)})
    sd = {k: torch.from_numpy((r.standard_normal(s) * 0.05).astype(np.float32)) for k, s in sd.items()}
    torch.save(sd, d / "pytorch_model.bin")
    return conf, {k: v.numpy() for k, v in sd.items()}


def reference_convert(d, out, conf, sd, out_type):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_flm_golden import converter_head
    m = converter_head()
    qt = {"int8": m.QuantType.INT8, "int16": m.QuantType.INT16, "f32": m.QuantType.NONE}[out_type]
    c = m.ModelConfig()
    assert c.load(Path(d), qt, 64)
    t = m.Tokenizer()
    assert t.load(Path(d), "spm")
    outf = m.FLFWriter(out, True)
    outf.dump_uint32(0xFA571AEA); outf.dump_uint8(1); outf.dump_uint8(0); outf.dump_uint16(0)          # _dump_file_header
    outf.dump_block("model_config", c.serialize_as_flf(True), m.BlockType.DICT)
    outf.dump_block("tokenizer", t.serialize_as_flf(True), m.BlockType.DICT)
    dt = m.DataType(out_type) if out_type != "f32" else None
    tmap = {"input_layernorm": 17, "self_attn.q_proj": 18, "self_attn.k_proj": 19, "self_attn.v_proj": 20, "self_attn.o_proj": 21,
            "post_attention_layernorm": 25, "mlp.gate_proj": 22, "mlp.up_proj": 23, "mlp.down_proj": 24}
    for name, a in sd.items():                                                                            # _dump_tensors :1109-1172
        layer, tt = 0, {"model.embed_tokens.weight": 1, "model.norm.weight": 2, "lm_head.weight": 3}.get(name)
        if tt is None:
            layer = int(name.split(".")[2])
            tt = tmap[name.split(".", 3)[-1].rstrip(".weight")]
            if tt in (18, 19):
                a = m.permute_qk(a, c.n_heads, c.n_kv_heads)
        if dt is not None and tt != 1 and a.ndim > 1:
            q, s = m.TensorLoader.quantize(a, dt, 64)
        else:
            q, s = a.astype(np.float32), None
        outf.dump_named_tensor(name, q, s, m.TensorType(tt), layer)
    outf.ofile.close()


@pytest.mark.parametrize("out_type,n_kv", [("int8", 4), ("int16", 4), ("f32", 4), ("int8", 2)])
def test_converted_file_is_byte_identical_to_the_reference_converters(fl, tmp_path, out_type, n_kv):
    conf, sd = make_hf_dir(tmp_path, n_kv)
    ours, theirs = tmp_path / "ours.flm", tmp_path / "theirs.flm"
    cfg = fl.convert.convert_hf_to_flm(str(tmp_path), str(ours), out_type)
    reference_convert(tmp_path, str(theirs), conf, sd, out_type)
    a, b = ours.read_bytes(), theirs.read_bytes()
    assert len(a) == len(b)
    assert a == b, next(i for i, (x, y) in enumerate(zip(a, b)) if x != y)
    # and it reads back: dims from config.json, the act_type quirk preserved as a string item
    rcfg, t, vocab = fl.flm.read_flm(ours)
    assert (rcfg["dim"], rcfg["hidden_dim"], rcfg["n_heads"], rcfg["n_kv_heads"], rcfg["n_layers"], rcfg["vocab_size"]) == (128, 192, 4, n_kv, 2, 384)
    assert rcfg["act_type"] == "silu" and rcfg["name"] == "tiny-llama" and rcfg["max_length"] == 2048
    assert len(vocab["texts"]) == 384 and vocab["special"] == {"bos": 1, "eos": 2}
    assert len(t) == 3 + 9 * 2
