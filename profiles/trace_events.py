"""Event log of the traced layer (middle layer, last decode step) on one attention CTA (7) and one other CTA (n-3):
per warp: stage wait / ready / done, chain token receive / send, build and drain starts.  Prints a per-phase timeline."""
import os, sys
os.environ.setdefault('FL_PROF_LIB', '1')     # the library build with the profiling counters compiled in
os.environ['FL_DEBUG_SKIP'] = '16'      # bit 4: enable the event log
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 288
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, flags=fl.FLAG_PROFILE)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
eng.forward(np.array([5], np.int32), ctx - 2, want_logits=False)
eng.profile_read(reset=True)
eng.decode_async(steps); eng.sync()
_, ev = eng.profile_read(events=True)
TYPES = {1: "wait", 2: "ready", 3: "done", 4: "sb?", 5: "sb!", 6: "ROWS", 7: "DRAIN", 8: "BUILD", 9: "polled", 10: "buf", 11: "v?", 12: "v!", 13: "SCORES", 14: "SOFTMAX", 15: "ATTN_OUT", 16: "QKV_IN"}
for which, name in ((0, "CTA 7 (attention)"), (1, "CTA n-3")):
    n = int(ev[which, 0]); rows = ev[which, 1:1 + min(n, 4095)]
    t = (rows >> 24).astype(np.int64); warp = ((rows >> 20) & 15).astype(int); typ = ((rows >> 12) & 255).astype(int); arg = (rows & 0xfff).astype(int)
    order = np.argsort(t, kind="stable"); t0 = t.min()
    print(f"==== {name}: {n} events")
    # per phase summary: for each warp the sequence of events as 'time:type(arg)'
    cur = None
    for i in order:
        if typ[i] == 8 and warp[i] == 0:
            print(f"--- phase {arg[i]} build starts at {(t[i]-t0)/1e3:8.2f} us")
    for w in list(range(8)) + [9]:
        seq = [f"{(t[i]-t0)/1e3:.2f}:{TYPES.get(typ[i], typ[i])}({arg[i]})" for i in order if warp[i] == w]
        if which == 1 or w in (0, 1, 9): print(f"warp {w}: " + " ".join(seq))
