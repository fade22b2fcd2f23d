"""Golden .flm files and logits from the REAL reference.  Run in the build container only (reads /root/reference):

    python tests/golden/make_flm_golden.py

  tests/golden/micro_int8.flm   written by the reference CONVERTER's own classes (tools/convert_flm.py: FLFWriter,
                                ModelConfig.serialize_as_flf, Tokenizer.serialize_as_flf, TensorLoader.quantize, in the
                                order ModelConverter.dump uses, :1075-1172) from the seeded MICRO weights.  The module cannot
                                be imported whole on Python 3.12 (its ModelConverter dataclass has a mutable default), so
                                the part above that class is executed as is; ModelConverter.dump's few lines are followed
                                by hand below.
  tests/golden/flm_golden.npz   logits of the reference's C++ (load_flm + ParallelTransformer::forward, through
                                oracle/_ref/libref.so) on the TINY model written as .flm by OUR writer: pins writer ->
                                reference loader -> forward, and is what the GPU test compares the engine with.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from conftest import load_pkg  # noqa: E402
from oracle_libs import ref, ptr, Q_INT8  # noqa: E402
from fixtures import TINY, gen_weights, prompt_tokens  # noqa: E402
import flm_inputs as fi  # noqa: E402


def converter_head():
    src = open("/root/reference/tools/convert_flm.py").read().split("\n")
    cut = next(i for i, l in enumerate(src) if l.startswith("class ModelConverter")) - 1
    m = types.ModuleType("convert_flm_head")
    exec(compile("\n".join(src[:cut]), "convert_flm.py[head]", "exec"), m.__dict__)
    return m


def write_with_reference_converter(path, spec, w, vocab):
    m = converter_head()
    outf = m.FLFWriter(path, True)
    outf.dump_uint32(0xFA571AEA)                       # _dump_file_header :1092-1106
    outf.dump_uint8(1); outf.dump_uint8(0); outf.dump_uint16(0)
    c = m.ModelConfig()
    for k, v in fi.config_of(spec, Q_INT8, 64, "micro").items():
        if k in ("model_type", "act_type", "quant_type"):
            v = type(c.__dict__[k])(v)
        c.__dict__[k] = v
    outf.dump_block("model_config", c.serialize_as_flf(True), m.BlockType.DICT)
    t = m.Tokenizer()
    t.vocab_type = "spm"
    t.vocab = m.Vocab(m.VocabType.SPM, vocab["texts"], vocab["scores"], vocab["types"])
    t.special_tokens = {k + "_token_id": v for k, v in vocab["special"].items()}
    outf.dump_block("tokenizer", t.serialize_as_flf(True), m.BlockType.DICT)
    dt = m.DataType("int8")
    for name, tt, layer, arr in fi.hf_tensors(spec, w):              # _dump_tensors :1109-1172 (already permuted)
        if tt != 1 and arr.ndim > 1:
            q, s = m.TensorLoader.quantize(arr, dt, 64)
        else:
            q, s = arr.astype(np.float32), None
        outf.dump_named_tensor(name, q, s, m.TensorType(tt), layer)
    outf.ofile.close()


def main():
    fl = load_pkg()
    spec = fi.MICRO
    w = gen_weights(spec, seed=21)
    write_with_reference_converter(os.path.join(HERE, "micro_int8.flm"), spec, w, fi.micro_vocab(spec.vocab_size))
    print("micro_int8.flm:", os.path.getsize(os.path.join(HERE, "micro_int8.flm")), "bytes")

    R = ref()
    assert R is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    spec = TINY
    w = gen_weights(spec, seed=1)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        p = d + "/tiny.flm"
        fl.flm.write_flm(p, fi.config_of(spec, Q_INT8, 64, "tiny"), fi.quantized_tensors(fl, spec, w, Q_INT8, 64),
                         fi.micro_vocab(spec.vocab_size))
        h = R.ref_model_load(p.encode(), b"", 1, Q_INT8, 2, 64, 0)
        assert h, "the reference loader rejected our .flm"
        prompt = prompt_tokens(spec, 6, seed=3)
        logits = np.empty(spec.vocab_size, np.float32)
        R.ref_forward(h, ptr(prompt), prompt.size, 0, ptr(logits))
        out["prefill_logits"] = logits.copy()
        toks, dl, pos = [], [], prompt.size
        for _ in range(8):
            t = np.array([int(np.argmax(logits))], np.int32)
            R.ref_forward(h, ptr(t), 1, pos, ptr(logits))
            toks.append(int(t[0])); dl.append(logits.copy()); pos += 1
        R.ref_model_free(h)
    np.savez_compressed(os.path.join(HERE, "flm_golden.npz"), prompt=prompt, decode_tokens=np.array(toks, np.int32),
                        decode_logits=np.stack(dl), **out)
    print("flm_golden.npz written; decode tokens", toks)


if __name__ == "__main__":
    main()
