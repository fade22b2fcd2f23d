"""ncu driver for the other instantiations of the persistent kernel (never a source of bench numbers):
   python profiles/prof_variant.py <7b-int8 | 7b-int16 | 13b-q8_0 | 7b-ms8> [ctx] [launches]
Every launch is ONE decode token (fl_forward of one token), so `ncu -k regex:decode_megakernel -s 2 -c 1` captures a whole token."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model
from fixtures import LLAMA2_7B, LLAMA2_13B
fl = ge._pkg()
name = sys.argv[1] if len(sys.argv) > 1 else "7b-int8"
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 288
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
spec, i16, gs, seqs = {"7b-int8": (LLAMA2_7B, False, 64, 1), "7b-int16": (LLAMA2_7B, True, 64, 1),
                       "13b-q8_0": (LLAMA2_13B, False, 32, 1), "7b-ms8": (LLAMA2_7B, False, 64, 8)}[name]
max_seq = max(1024, (ctx + 64 + 3) // 4 * 4)
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=max_seq,
                quant_type=fl.Q_INT16 if i16 else fl.Q_INT8, group_size=gs, max_seqs=seqs,
                flags=fl.FLAG_NO_TC if seqs > 1 else 0)      # several sequences through the persistent kernel (MS = true), not the rows path
for (kind, layer), (q, s) in synth_int8_model(spec, 0, int16=i16, gs=gs):
    eng.upload(kind, layer, q, s)
eng.finalize()
tok = np.array([5], np.int32)
if seqs == 1:
    eng.forward(tok, ctx - 2, want_logits=False)
    for i in range(launches):
        eng.forward(tok, ctx - 1 + i, want_logits=False)
else:
    for i in range(seqs):
        eng.forward(np.array([1 + i], np.int32), ctx - 2, slot=i, want_logits=False)
    toks = np.arange(1, seqs + 1, dtype=np.int32)
    for k in range(launches):       # fl_forward_batch: one persistent launch (MS = true) advances every sequence by one token
        toks = eng.forward_batch(toks, np.full(seqs, ctx - 1 + k, np.int32))
eng.sync()
print("done", name, ctx, launches)
eng.close()
