"""The tensor-core rows path at the LLaMA2-7B INT8 shape on one GPU:
  * prompt chunks: fl_forward over a P-token prompt (chunks of 64 rows per weight pass) -> time to first token;
  * config-4 style decode: n sequences advanced together by fl_decode_batch_async (one weight pass per step, CUDA graph);
flags: argv[1] = engine flags (2 = no PDL, 1 = no graph).  Times are CUDA events on the engine stream."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_max = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, max_seqs=n_max, flags=flags)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
stream = torch.cuda.ExternalStream(eng.stream)
rng = np.random.default_rng(3)
wbytes = eng.step_bytes(0) - 2 * spec.n_layers * spec.kv_dim * 4


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream); fn(); e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for P in (64, 512, 1000):
    prompt = np.concatenate([[1], rng.integers(3, spec.vocab_size, P - 1)]).astype(np.int32)
    eng.forward(prompt, 0, want_logits=False)
    ms = min(timed(lambda: eng.forward(prompt, 0, want_logits=False)) for _ in range(2))
    passes = -(-P // 64)
    print(f"flags {flags}: prompt {P}: {ms:.2f} ms to first token = {P / ms * 1e3:.0f} prompt tokens/s, {ms / passes:.3f} ms per weight pass "
          f"({wbytes / (ms / passes * 1e-3) / 1e9:.0f} GB/s of weights)", flush=True)

for n in sorted({2, 8, n_max}):
    for i in range(n):
        eng.forward(np.concatenate([[1], rng.integers(3, spec.vocab_size, 31)]).astype(np.int32), 0, slot=i, want_logits=False)
    eng.decode_batch_async(n, 4); eng.sync()
    ms = timed(lambda: eng.decode_batch_async(n, steps))
    ctx = 32 + 4 + steps / 2
    bytes_step = wbytes + n * (ctx + 1) * 2 * spec.n_layers * spec.kv_dim * 4
    print(f"flags {flags}: {n} sequences: {ms / steps:.3f} ms per step, {n * steps / ms * 1e3:.1f} tokens/s per GPU, "
          f"{bytes_step / (ms / steps * 1e-3) / 1e9:.0f} GB/s algorithmic (weights once + {n} x KV at ctx {ctx:.0f})", flush=True)
eng.close()
