"""What bit-exactness costs: the persistent decode kernel with FL_FLAG_RELAXED (rmsnorm's sum of squares
and softmax's sum as tree reductions instead of the reference's serial FP32 chains) against the exact kernel, 7B INT8 shape,
prompt 32 + 511 decode tokens: tokens/s of both, per-step relative logit error with both engines fed the EXACT engine's tokens,
agreement of the greedy token, the top-2 margin of the exact logits where they disagree, and the first index at which two
free-running greedy generations diverge."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as ge
from bench import Synth, shape_7b
fl = ge._pkg()
spec = shape_7b()
syn = Synth(spec, 0)
engs = {}
for name, flags in (("exact", 0), ("relaxed", fl.FLAG_RELAXED)):
    e = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, flags=flags | fl.FLAG_NO_TC)
    for (kind, layer), (q, s) in syn.items():
        e.upload(kind, layer, q, s)
    e.finalize()
    engs[name] = e
prompt = np.concatenate([[1], np.random.default_rng(7).integers(3, spec.vocab_size, 31)]).astype(np.int32)
N = 511
for name, e in engs.items():
    stream = torch.cuda.ExternalStream(e.stream)
    best = 1e9
    for _ in range(3):
        e.forward(prompt, 0, want_logits=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream); e.decode_async(N); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:8s}: {best / N:.4f} ms/token = {N / best * 1e3:.1f} tokens/s", flush=True)
# lockstep: both engines see the exact engine's tokens
a = engs["exact"].forward(prompt, 0); b = engs["relaxed"].forward(prompt, 0)
pos, rel, agree, margins = prompt.size, [], 0, []
for step in range(N):
    rel.append(float(np.abs(a - b).max() / np.abs(a).max()))
    ta, tb = int(np.argmax(a)), int(np.argmax(b))
    if ta == tb:
        agree += 1
    else:
        top2 = np.sort(a)[-2:]
        margins.append((step, float(top2[1] - top2[0]), float(np.abs(a).max())))
    t = np.array([ta], np.int32)
    a = engs["exact"].forward(t, pos); b = engs["relaxed"].forward(t, pos)
    pos += 1
print(f"lockstep over {N} steps: max relative logit error {max(rel):.3e} (median {np.median(rel):.3e}); greedy token equal in {agree}/{N} steps")
for m in margins[:10]:
    print(f"   step {m[0]}: exact top-2 margin {m[1]:.3e} (max |logit| {m[2]:.3f})")
ga = engs["exact"].generate_greedy(prompt, N); gb = engs["relaxed"].generate_greedy(prompt, N)
n = min(len(ga), len(gb))
div = next((i for i in range(n) if ga[i] != gb[i]), None)
print(f"free-running greedy generations: {'identical over %d tokens' % n if div is None else 'first divergence at token %d of %d' % (div, n)}")
