"""Per-category time inside the persistent kernel when fl_forward_batch advances n sequences per launch (7B-shaped INT8)."""
import os, sys
os.environ.setdefault('FL_PROF_LIB', '1')
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 160
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, max_seqs=n, flags=fl.FLAG_PROFILE)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
toks = np.arange(5, 5 + n, dtype=np.int32); pos = np.full(n, ctx, np.int32)
for _ in range(2):
    toks = eng.forward_batch(toks, pos); pos += 1
eng.profile_read(reset=True)
import time
t0 = time.perf_counter()
for _ in range(steps):
    toks = eng.forward_batch(toks, pos); pos += 1
dt = time.perf_counter() - t0
pr = eng.profile_read().astype(np.float64) / (steps * n) / 1965.0     # us per token (per sequence-step) per CTA
names = ["ll_wait", "build_tail", "qkv", "wo", "w13", "w2", "cls", "stage_wait", "build_pre", "build_chain", "attn_qkv_rope", "attn_qk", "attn_xchg", "attn_softmax", "-", "pv_total", "pairbuf_wait", "drain_misc", "argmax", "embed", "stages_ahead_x1000", "wait_first16"] + ["-"] * 10
print(f"n_seqs {n} ctx {ctx}: {dt / steps / n * 1e3:.3f} ms per token (host clock), {steps} steps")
print(f"{'category':14s} {'mean':>9s} {'min':>9s} {'max':>9s}   (us per token, over {pr.shape[0]} CTAs)")
for k, nm in enumerate(names):
    if nm == '-': continue
    print(f"{nm:14s} {pr[:, k].mean():9.1f} {pr[:, k].min():9.1f} {pr[:, k].max():9.1f}")
cols = [i for i in range(22) if i not in (14, 20)]
print(f"{'sum':14s} {pr[:, cols].sum(1).mean():9.1f} {pr[:, cols].sum(1).min():9.1f} {pr[:, cols].sum(1).max():9.1f}")
