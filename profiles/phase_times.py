"""Per-category time inside the persistent decode kernel (FL_FLAG_PROFILE): phase_times.py [ctx] [steps] [7b | 13b]
(7b = LLaMA2-7B-shaped INT8 group 64; 13b = LLaMA2-13B-shaped INT8 group 32 = Q8_0 arithmetic)."""
import os, sys
os.environ.setdefault('FL_PROF_LIB', '1')     # the library build with the profiling counters compiled in
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 288
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
big = len(sys.argv) > 3 and sys.argv[3] == "13b"
if big:
    from fixtures import LLAMA2_13B
    spec, gs = LLAMA2_13B, 32
else:
    spec, gs = shape_7b(), 64
max_seq = max(1024, (ctx + steps + 8 + 3) // 4 * 4)
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=max_seq, group_size=gs, flags=fl.FLAG_PROFILE)
for (kind, layer), (q, s) in synth_int8_model(spec, 0, gs=gs):
    eng.upload(kind, layer, q, s)
eng.finalize()
tok = np.array([5], np.int32)
eng.forward(tok, ctx - 2, want_logits=False)
eng.profile_read(reset=True)
import time
t0 = time.perf_counter(); eng.decode_async(steps); eng.sync(); dt = time.perf_counter() - t0
pr = eng.profile_read().astype(np.float64) / steps / 1965.0     # SM cycles at 1965 MHz -> us per token per CTA
names = ["ll_wait", "build_tail", "qkv", "wo", "w13", "w2", "cls", "stage_wait", "build_pre", "build_chain", "attn_qkv_rope", "attn_qk", "attn_xchg", "attn_softmax", "pv_wait0", "pv_total", "pairbuf_wait", "drain_misc", "argmax", "embed", "stages_ahead_x1000", "wait_first16"] + ["-"] * 10
print(f"ctx {ctx}: {dt / steps * 1e3:.3f} ms/token (host clock), {steps} steps")
print(f"{'category':14s} {'mean':>9s} {'min':>9s} {'max':>9s}   (us per token, over {pr.shape[0]} CTAs)")
for k, n in enumerate(names):
    if n == '-': continue
    print(f"{n:14s} {pr[:, k].mean():9.1f} {pr[:, k].min():9.1f} {pr[:, k].max():9.1f}")
print(f"{'sum':14s} {pr[:, [i for i in range(22) if i != 20]].sum(1).mean():9.1f} {pr[:, [i for i in range(22) if i != 20]].sum(1).min():9.1f} {pr[:, [i for i in range(22) if i != 20]].sum(1).max():9.1f}")
