"""First-light probe of the tcgen05 GEMM (fl_op_matmul_q_tc): mismatch statistics against the oracle per shape and kernel
variant (bit 0 swaps LBO / SBO in the shared-memory descriptors).  Diagnostics only; the parity tests are tests/test_tc_gemm_gpu.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from oracle_libs import port, ptr, bits, Q_INT8, port_quantize
fl = ge._pkg()


def rq(rng, rows, cols, gs):
    x = (rng.standard_normal((rows, cols)) * rng.uniform(0.2, 3.0, (rows, 1))).astype(np.float32)
    q, s = port_quantize(Q_INT8, x, gs)
    return q.reshape(rows, cols), s.reshape(rows, cols // gs)


def run(m, n, rows, gs, variant):
    rng = np.random.default_rng(1)
    w, ws = rq(rng, m, n, gs)
    x, xs = rq(rng, rows, n, gs)
    want = np.empty((rows, m), np.float32)
    port().port_matmul(Q_INT8, ptr(want), ptr(w), ptr(ws), ptr(x), ptr(xs), m, n, rows, gs)
    got = fl.ops.matmul_q_tc(w, ws, x, xs, gs=gs, variant=variant)
    bad = np.argwhere(bits(got) != bits(want))
    print(f"m={m} n={n} rows={rows} gs={gs} variant={variant}: {len(bad)} / {got.size} mismatches", flush=True)
    if len(bad):
        for b in bad[:6]:
            print("   ", b.tolist(), got[tuple(b)], want[tuple(b)])
        rows_bad = sorted(set(bad[:, 1].tolist()))
        print("    bad weight rows (first 40):", rows_bad[:40], " bad act rows:", sorted(set(bad[:, 0].tolist()))[:20])
    return len(bad)


v = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for shape in [(300, 256, 8, 64), (4096, 512, 64, 64), (1000, 704, 5, 64), (4096, 1024, 20, 32), (32000, 512, 16, 64)]:
    run(*shape, v)
