"""Golden logits from the REAL reference for a GGUF file.  Run in the build container only:

    python tests/golden/make_gguf_golden.py

tests/golden/gguf_golden.npz: logits of the reference's C++ (load_gguf + ParallelTransformer::forward, -q int8, through
oracle/_ref/libref.so) on the seeded TINY model written as an all-F32 GGUF by llama.cpp's `gguf` writer.  (The Q8_0 variant
cannot be pinned this way: the reference decodes fp16 scales wrongly, SURVEY defect D5.)"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_libs import ref, ptr, Q_INT8  # noqa: E402
from fixtures import TINY, gen_weights, prompt_tokens  # noqa: E402
from gguf_inputs import write_gguf  # noqa: E402


def main():
    R = ref()
    assert R is not None, "build oracle/_ref first (bash oracle/build_ref.sh)"
    spec = TINY
    w = gen_weights(spec, seed=1)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        p = d + "/tiny_f32.gguf"
        write_gguf(p, spec, w, q8_0=False)
        h = R.ref_model_load(p.encode(), b"", 2, Q_INT8, 2, 64, 0)
        assert h, "the reference loader rejected the GGUF file"
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
    np.savez_compressed(os.path.join(HERE, "gguf_golden.npz"), prompt=prompt, decode_tokens=np.array(toks, np.int32),
                        decode_logits=np.stack(dl), **out)
    print("gguf_golden.npz written; decode tokens", toks)


if __name__ == "__main__":
    main()
