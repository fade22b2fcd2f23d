"""Per-warp timeline of the activation rebuilds of the traced layer (middle layer, last decode step), SM-cycle timestamps.
Needs the `ev` variant build (csrc/Makefile: -DFL_PROFILE -DFL_EVLOG -DFL_EVLOG_CLOCK).  usage: trace_rebuild.py [ctx] [steps]"""
import os, sys
os.environ['FL_PROF_LIB'] = 'ev'
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
NAMES = {8: "BUILD", 9: "polled", 20: "sync1", 21: "pre", 22: "sync2", 23: "chain", 24: "sync3", 25: "tail", 7: "DRAIN", 2: "ready", 3: "done", 1: "wait", 10: "buf",
         16: "QKV_IN", 13: "SCORES", 14: "SOFTMAX", 15: "ATTN_OUT"}
MHZ = 1965.0
for which, name in ((0, "CTA 7 (attention)"), (1, "CTA n-3")):
    n = int(ev[which, 0]); rows = ev[which, 1:1 + min(n, 4095)]
    t = (rows >> 24).astype(np.int64); warp = ((rows >> 20) & 15).astype(int); typ = ((rows >> 12) & 255).astype(int); arg = (rows & 0xfff).astype(int)
    order = np.argsort(t, kind="stable")
    print(f"==== {name}: {n} events")
    starts = [(t[i], arg[i]) for i in order if typ[i] == 8 and warp[i] == 0]
    for (ts, pk) in starts:
        print(f"--- phase kind {pk}: times in us after warp 0's build start")
        te = ts + 14 * MHZ           # window
        for w in range(8):
            seq = [f"{NAMES.get(typ[i], typ[i])}@{(t[i] - ts) / MHZ:.2f}" for i in order if warp[i] == w and ts - 2 * MHZ <= t[i] <= te and typ[i] in (8, 9, 20, 21, 22, 23, 24, 25, 7)]
            print(f"  warp {w}: " + " ".join(seq))
        for w in list(range(8)) + [9]:
            # drain start: pair buffer free (buf), stage wait / ready / done; chain warp (9): sb? = waits for the consumers, sb! = has them, ROWS = tile published
            names = {10: "buf", 1: "wait", 2: "ready", 3: "done", 4: "sb?", 5: "sb!", 6: "ROWS"}
            st = [f"{names[typ[i]]}@{(t[i] - ts) / MHZ:.2f}" for i in order if warp[i] == w and ts <= t[i] <= te + 10 * MHZ and typ[i] in names]
            if st: print(f"  warp {w} drain events: " + " ".join(st[:16]))
        if pk == 0:
            # attention follows the QKV drain: QKV_IN = q/k/v polled, SCORES = own scores done, SOFTMAX = scores exchanged, ATTN_OUT = softmax done (PV follows)
            an = {16: "QKV_IN", 13: "SCORES", 14: "SOFTMAX", 15: "ATTN_OUT", 11: "v?", 12: "v!"}
            for w in range(8):
                st = [f"{an[typ[i]]}@{(t[i] - ts) / MHZ:.2f}" for i in order if warp[i] == w and ts <= t[i] <= ts + 40 * MHZ and typ[i] in an]
                if st: print(f"  warp {w} attention: " + " ".join(st[:12]))
            nb = [(t[i] - ts) / MHZ for i in order if typ[i] == 8 and warp[i] == 0 and t[i] > ts]
            if nb: print(f"  next build (Wo) starts at {nb[0]:.2f}")
