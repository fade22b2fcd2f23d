"""oracle/gqa_probe.py — TEST INFRASTRUCTURE ONLY (evidence for DESIGN.md defect D10; run in the build container).

Runs the REAL reference (oracle/_ref/libref.so) on a 1-layer model with 8 heads and n_kv_heads in {8, 4, 2, 1} for two
thread counts and compares the logits with the oracle restatement.  With n_kv_heads == n_heads the two are bit-identical;
with grouped-query attention the reference's logits depend on the thread count, because Tensor::weighted_sum
(src/components/tensor.cpp:713) computes only the first query head of every group and leaves the others' outputs to
whatever the scratch buffer held.  Each case runs in its own process (some thread counts make the reference hang)."""
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0] + "/tests")


def one(nkv, nth):
    from oracle_libs import port, ref, ptr, bits, Q_INT8
    from fixtures import ModelSpec, gen_weights, quantize_model, write_llama2c, write_tokenizer_bin, synthetic_vocab, prompt_tokens
    from test_forward_gpu import make_port_model
    R, P = ref(), port()
    spec = ModelSpec(dim=512, hidden_dim=704, n_layers=1, n_heads=8, n_kv_heads=nkv, vocab_size=1000)
    w = gen_weights(spec, seed=8)
    with tempfile.TemporaryDirectory() as d:
        write_llama2c(d + "/m.bin", spec, w)
        write_tokenizer_bin(d + "/t.bin", synthetic_vocab(spec.vocab_size))
        h = R.ref_model_load((d + "/m.bin").encode(), (d + "/t.bin").encode(), 3, Q_INT8, nth, 64, 0)
        pm = make_port_model(spec, quantize_model(spec, w, Q_INT8, 64), Q_INT8, 64)
        prompt = prompt_tokens(spec, 1, seed=2)
        a, b = np.empty(spec.vocab_size, np.float32), np.empty(spec.vocab_size, np.float32)
        R.ref_forward(h, ptr(prompt), 1, 0, ptr(a))
        P.port_forward(pm, ptr(prompt), 1, 0, ptr(b))
    print(f"n_heads=8 n_kv_heads={nkv} threads={nth}: reference == oracle: {np.array_equal(bits(a), bits(b))}, "
          f"max |diff| {np.abs(a - b).max():.6g}, reference logits checksum {int(bits(a).astype(np.uint64).sum())}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3:
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        for nkv in (8, 4, 2, 1):
            for nth in (2, 4):
                r = subprocess.run([sys.executable, __file__, str(nkv), str(nth)], capture_output=True, text=True, timeout=120)
                print("\n".join(l for l in r.stdout.splitlines() if l.startswith("n_heads")) or f"n_kv_heads={nkv} threads={nth}: failed")
