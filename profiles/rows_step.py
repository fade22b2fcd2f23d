"""Small driver for ncu: a few launches of the tensor-core rows path at the 7B INT8 shape.
argv: mode (decode | prefill), rows (sequences or prompt tokens), steps, engine flags."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
mode = sys.argv[1] if len(sys.argv) > 1 else "decode"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n_layers = int(os.environ.get("FL_LAYERS", spec.n_layers))
import dataclasses
spec = dataclasses.replace(spec, n_layers=n_layers)
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024,
                max_seqs=rows if mode == "decode" else 1, flags=flags)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
rng = np.random.default_rng(3)
if mode == "decode":
    for i in range(rows):
        eng.forward(np.array([1 + i], np.int32), 0, slot=i, want_logits=False)      # one-token prompts: the capture is about the decode steps
    eng.decode_batch_async(rows, steps)
    eng.sync()
else:
    prompt = np.concatenate([[1], rng.integers(3, spec.vocab_size, rows - 1)]).astype(np.int32)
    for _ in range(steps):
        eng.forward(prompt, 0, want_logits=False)
eng.close()
