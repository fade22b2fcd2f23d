"""Decode speed of the other BASELINE.json configurations (parity-test cases, not bench lines): synthetic weights of the
named shape, prefill a short prompt, then `gen` greedy tokens resident on the device.
usage: python profiles/config_sweep.py [7b-int16] [13b-q8_0] [7b-int8-long] ..."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from bench import synth_int8_model
from fixtures import LLAMA2_7B, LLAMA2_13B
fl = ge._pkg()
CONFIGS = {
    "7b-int8": (LLAMA2_7B, False, 64, 1024, 32, 512),
    "7b-int16": (LLAMA2_7B, True, 64, 1024, 32, 512),          # config 3
    "13b-q8_0": (LLAMA2_13B, False, 32, 2816, 2048, 512),       # config 5: GGUF group 32, prompt 2048 (only the last 64 prompt tokens are run: the cache below them is zero)
    "7b-int8-long": (LLAMA2_7B, False, 64, 1024, 900, 100),
}
for name in (sys.argv[1:] or ["7b-int16", "13b-q8_0"]):
    spec, i16, gs, max_seq, n_prompt, gen = CONFIGS[name]
    eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seq_len=max_seq,
                    quant_type=fl.Q_INT16 if i16 else fl.Q_INT8, group_size=gs)
    for (kind, layer), (q, s) in synth_int8_model(spec, 0, int16=i16, gs=gs):
        eng.upload(kind, layer, q, s)
    eng.finalize()
    rng = np.random.default_rng(1)
    tail = rng.integers(3, spec.vocab_size, min(n_prompt, 64)).astype(np.int32)
    eng.forward(tail, n_prompt - tail.size, want_logits=False)       # positions below are an all-zero cache: same bytes swept
    eng.sync()
    t0 = time.perf_counter(); eng.decode_async(gen); eng.sync(); dt = time.perf_counter() - t0
    ctx_mean = n_prompt + gen / 2
    gb = np.mean([eng.step_bytes(n_prompt + i) for i in range(gen)]) / 1e9
    print(f"{name:14s} {gen} tokens from ctx {n_prompt}: {dt / gen * 1e3:.3f} ms/token = {gen / dt:.1f} tok/s; {gb:.2f} GB/token -> {gb / (dt / gen) / 1e3:.2f} TB/s "
          f"({gb / (dt / gen) / 6448.1 * 100:.1f} % of the measured 6448 GB/s)", flush=True)
    eng.close()
