"""Config-4-style decode on one GPU: n sequences of the 7B-shaped INT8 engine advanced together by fl_forward_batch.
With the multi-sequence launch every phase is walked once per sequence inside one persistent kernel (the exchanges and serial
sections of one sequence hide behind the others' weight streaming; repeated passes over a phase's weights hit L2);
FL_NO_MULTISEQ=1 is the old behaviour (one launch per sequence).  Prints tokens/s over all sequences (CUDA-event time)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
n_max = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 160
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, max_seqs=n_max)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
stream = torch.cuda.ExternalStream(eng.stream)
for n in sorted({1, 2, 4, n_max}):
    if n > n_max:
        continue
    toks = np.arange(5, 5 + n, dtype=np.int32)
    pos = np.full(n, ctx, np.int32)
    for _ in range(3):
        toks = eng.forward_batch(toks, pos); pos += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        toks = eng.forward_batch(toks, pos); pos += 1
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    print(f"n_seqs {n}: {ms / steps:.3f} ms per step (wall {1e3 * wall / steps:.3f}), {n * steps / ms * 1e3:.1f} tokens/s, {ms / steps / n:.3f} ms per token"
          f"{'  [FL_NO_MULTISEQ]' if os.environ.get('FL_NO_MULTISEQ') else ''}", flush=True)
