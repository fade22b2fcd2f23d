"""One exchange under the microscope: absolute globaltimer stamps of the middle layer of the last decode step, per CTA.
slots: 22+pk = input of phase pk complete (tags valid), 26+pk = drain of phase pk finished, 30 = attention has q/k/v,
31 = attention finished.  Prints, per phase, the skew of the producers and the latency from the LAST producer's finish to
each consumer's "input complete"."""
import os, sys
os.environ.setdefault('FL_PROF_LIB', '1')     # the library build with the profiling counters compiled in
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model, shape_7b
fl = ge._pkg()
spec = shape_7b()
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 288
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=1024, flags=fl.FLAG_PROFILE)
for (kind, layer), (q, s) in synth_int8_model(spec, 0):
    eng.upload(kind, layer, q, s)
eng.finalize()
eng.forward(np.array([5], np.int32), ctx - 2, want_logits=False)
eng.profile_read(reset=True)
eng.decode_async(steps); eng.sync()
pr = eng.profile_read().astype(np.int64)
t = pr[:, 22:32].astype(np.float64)
n_attn = spec.n_heads * 4
t[t < 1] = np.nan
base = t[:, 0].min()
t = (t - base) / 1e3
names = {0: "QKV", 1: "Wo", 2: "W13", 3: "W2"}
def st(a): return f"min {a.min():7.2f} med {np.median(a):7.2f} max {a.max():7.2f}"
for pk in range(4):
    print(f"phase {names[pk]:4s} input complete  {st(t[:, pk])}")
    if pk == 0:
        print(f"           attn has qkv    {st(t[:n_attn, 8])}")
    print(f"phase {names[pk]:4s} drain finished  {st(t[:, 4 + pk])}   (drain length {st(t[:, 4 + pk] - t[:, pk])})")
    if pk == 0:
        print(f"           attn finished   {st(t[:n_attn, 9])}   (attention length {st(t[:n_attn, 9] - t[:n_attn, 4])})")
print("latency last producer -> consumers' input complete:")
print(f"  QKV->attn   {st(t[:n_attn, 8] - t[:, 4].max())}")
print(f"  attn->Wo    {st(t[:, 1] - t[:n_attn, 9].max())}")
print(f"  Wo->W13     {st(t[:, 2] - t[:, 5].max())}")
print(f"  W13->W2     {st(t[:, 3] - t[:, 6].max())}")

for pk in range(4):
    d = t[:, 4 + pk]
    order = np.argsort(-np.nan_to_num(d))[:8]
    print(f"phase {names[pk]:4s} slowest CTAs (drain finished): " + " ".join(f"{i}:{d[i]:.1f}" for i in order) + f"   | median {np.nanmedian(d):.1f}")
    ic = t[:, pk]
    order = np.argsort(-np.nan_to_num(ic))[:6]
    print(f"           latest input complete: " + " ".join(f"{i}:{ic[i]:.1f}" for i in order) + f"   | median {np.nanmedian(ic):.1f}")
