"""Profiling driver: LLaMA2-7B-shaped INT8 engine, a few decode steps at a chosen context length.
Used under ncu (see profiles/README.md); never a source of bench numbers."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b

ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 288
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fl = ge._pkg()
spec = shape_7b()
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, flags=flags)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
tok = np.array([5], np.int32)
eng.forward(tok, 0, want_logits=False)           # warm-up (graph capture)
eng.forward(tok, ctx - 1, want_logits=False)
t0 = time.perf_counter()
for i in range(steps):
    eng.forward(tok, ctx + i, want_logits=False)
print("steps", steps, "ctx", ctx, "ms/step (host, incl. sync)", (time.perf_counter() - t0) * 1e3 / steps)
eng.close()
