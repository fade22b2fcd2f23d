#!/usr/bin/env python
"""bench.py — decode tokens/s of the B200 engine on BASELINE.json's configurations, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1   configs[1]: LLaMA2-7B-shaped INT8 (group 64), batch 1, prompt 32, gen 512, synthetic seeded weights.
        A "step" is one pass of the hot path over that input: the 511 decode forwards that follow the prefill + first token
        (the reference's own definition of decode speed, src/main.cpp:126,134: (total - first token) / (output tokens - 1)).
          value : decode tokens/s with everything resident in HBM (token fed back on the device, one persistent launch)
          e2e   : the same through the reference-facing call fl_forward(host token, pos) -> host logits + host argmax
        plus `other_configs`, short samples measured in the same process: configs[3] per GPU (8 sequences, one weight pass per
        step on the tensor cores), configs[2] (7B INT16) and configs[4] (13B Q8_0 group 32, 2048-token prompt -> time to first
        token on the tensor-core prompt path, then decode at ctx >= 2048).
N > 1   configs[3]: 7B INT8, 8 sequences per GPU (batch = 8 N), prompt 32, gen 256, weights replicated, the request batch sharded;
        per step ONE weight pass for the GPU's 8 sequences (fl_decode_batch_async) and one ncclAllGather of the sampled tokens
        through fl_allgather_tokens on the engine stream.  value = all ranks' tokens / max-over-ranks device time.
--impl reference: the UNMODIFIED reference's own CPU forward() (oracle/_ref/libref.so) on the host cores, on the SAME model
        shape at full depth: the bench model is written as a 32-layer int8 .flm (the reference's loader accepts our files,
        tests/test_flm.py), loaded by the reference, prefilled, and real decode tokens are timed at contexts spread over the
        benchmark's range.  One step = one decode token (a bounded sample of the 511-token step).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))

PROMPT, GEN = 32, 512
SEQS_PER_GPU, GEN4 = 8, 256
METRIC = "decode tokens/s LLaMA2-7B INT8 batch=1 seq=128->512"
UNIT = "tokens/s"


def shape_7b():
    from fixtures import LLAMA2_7B
    return LLAMA2_7B


# ----------------------------------------------------------------------------------------------------------
class Synth:
    """Random INT8 (or INT16) payloads + fp32 group scales with realistic magnitudes (no 27 GB float model is materialised).
    Every layer gets its own HBM copy; rolling rows makes the contents differ without regenerating 200 MB per layer."""

    def __init__(self, spec, seed=0, int16=False, gs=64):
        from oracle_libs import (T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS)
        self.spec, self.K = spec, dict(emb=T_TOK_EMB, an=T_ATT_NORM, fn=T_FFN_NORM, on=T_OUT_NORM, cls=T_CLS)
        rng = np.random.default_rng(seed)
        d, h, kv = spec.dim, spec.hidden_dim, spec.kv_dim

        def qmat(rows, cols, sd):
            amp, lim, dt = (1800.0, 5792, np.int16) if int16 else (40.0, 127, np.int8)
            q = np.clip(np.rint(rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(amp)), -lim, lim).astype(dt)
            s = (np.float32(sd / amp) * (0.75 + 0.5 * rng.random((rows, cols // gs), dtype=np.float32))).astype(np.float32)
            return q, s

        self.base = {T_WQ: qmat(d, d, d ** -0.5), T_WK: qmat(kv, d, d ** -0.5), T_WV: qmat(kv, d, d ** -0.5), T_WO: qmat(d, d, d ** -0.5),
                     T_W1: qmat(h, d, d ** -0.5), T_W3: qmat(h, d, d ** -0.5), T_W2: qmat(d, h, h ** -0.5)}
        self.emb = rng.standard_normal((spec.vocab_size, d), dtype=np.float32) * np.float32(0.05)
        self.cls = qmat(spec.vocab_size, d, d ** -0.5)
        self.norms = [((1 + 0.1 * rng.standard_normal(d)).astype(np.float32), (1 + 0.1 * rng.standard_normal(d)).astype(np.float32))
                      for _ in range(spec.n_layers)]

    def get(self, kind, layer):
        K = self.K
        if kind == K["emb"]:
            return self.emb, None
        if kind == K["on"]:
            return np.ones(self.spec.dim, np.float32), None
        if kind == K["cls"]:
            return self.cls
        if kind == K["an"]:
            return self.norms[layer][0], None
        if kind == K["fn"]:
            return self.norms[layer][1], None
        q, s = self.base[kind]
        return np.roll(q, layer * 7, axis=0), np.roll(s, layer * 7, axis=0)

    def items(self):
        K = self.K
        yield (K["emb"], 0), self.get(K["emb"], 0)
        yield (K["on"], 0), self.get(K["on"], 0)
        yield (K["cls"], 0), self.get(K["cls"], 0)
        for l in range(self.spec.n_layers):
            yield (K["an"], l), self.get(K["an"], l)
            yield (K["fn"], l), self.get(K["fn"], l)
            for k in self.base:
                yield (k, l), self.get(k, l)


def synth_int8_model(spec, seed=0, int16=False, gs=64):
    return Synth(spec, seed, int16, gs).items()


def bench_config(world):
    spec = shape_7b()
    if world == 1:
        return {"workload": "LLaMA2-7B INT8 (group 64) batch=1 prompt=32 gen=512, synthetic seeded weights (BASELINE configs[1])",
                "dim": spec.dim, "hidden_dim": spec.hidden_dim, "n_layers": spec.n_layers, "vocab": spec.vocab_size,
                "kv": "fp32", "step": f"{GEN - 1} decode forwards after prefill+first token",
                "l2": "inputs larger than L2 (7.0 GB of weights streamed per token vs 126 MB L2)", "parallelism": "single GPU"}
    return {"workload": f"LLaMA2-7B INT8 (group 64) batch={SEQS_PER_GPU * world} = {SEQS_PER_GPU} sequences/GPU x {world} GPUs, prompt=32 gen={GEN4}, "
                        "synthetic seeded weights (BASELINE configs[3])",
            "dim": spec.dim, "hidden_dim": spec.hidden_dim, "n_layers": spec.n_layers, "vocab": spec.vocab_size, "kv": "fp32",
            "step": f"{GEN4 - 1} decode steps of {SEQS_PER_GPU} sequences per GPU after their prefills; per step one weight pass (tcgen05 GEMM) "
                    "and one ncclAllGather of the sampled tokens (fl_allgather_tokens)",
            "l2": "inputs larger than L2 (7.0 GB of weights streamed per step vs 126 MB L2)",
            "parallelism": f"dp{world} (weights replicated, request batch sharded, {SEQS_PER_GPU} sequences per GPU)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class _StdoutToStderr:
    """the reference's loader printf()s to stdout; the bench line must be the only thing there"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ----------------------------------------------------------------------------------------------------------
def reference_decode(n_tokens, n_warm, threads=None):
    """The UNMODIFIED reference (oracle/_ref/libref.so: its own loader, quantised matmul, attention, thread pool) on the full
    32-layer 7B INT8 model of the benchmark: written here as an .flm with our writer, loaded by the reference's load_flm
    (src/model_loaders/flm_loader.cpp:561), prompt of 32 tokens prefilled in two batched forwards, then n_warm + n_tokens single-token
    forwards timed one by one at positions spread over the benchmark's decode range (33 .. 543).  KV rows the sample never
    wrote are zero; timing does not depend on their values.  Returns per-token seconds of the timed tokens."""
    from oracle_libs import ref, ptr, Q_INT8
    import importlib.util
    R = ref()
    if R is None:
        return None
    spec = shape_7b()
    cores = threads or os.cpu_count() or 1
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    fl = ge._pkg()
    from flm_inputs import config_of, micro_vocab
    t0 = time.perf_counter()
    syn = Synth(spec, seed=0)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "bench7b.flm")
        fl.flm.write_flm(path, config_of(spec, Q_INT8, 64, "bench7b"), syn.get, micro_vocab(spec.vocab_size))
        fbytes = os.path.getsize(path)
        del syn
        with _StdoutToStderr():
            h = R.ref_model_load(path.encode(), b"", 1, Q_INT8, cores, 64, 0)
    assert h, "the reference failed to load the benchmark .flm"
    t_load = time.perf_counter() - t0
    logits = np.empty(spec.vocab_size, np.float32)
    prompt = np.concatenate([[1], np.random.default_rng(7).integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32)
    t0 = time.perf_counter()
    R.ref_forward(h, ptr(prompt[:16].copy()), 16, 0, ptr(logits))
    R.ref_forward(h, ptr(prompt[16:].copy()), 16, 16, ptr(logits))
    t_prefill = time.perf_counter() - t0
    total = n_warm + n_tokens
    positions = np.unique(np.linspace(PROMPT, PROMPT + GEN - 2, total).astype(int)) if total > 1 else np.array([PROMPT])
    while positions.size < total:
        positions = np.append(positions, positions[-1] + 1)
    times = []
    for p in positions[:total]:
        t = np.array([int(np.argmax(logits))], np.int32)
        t1 = time.perf_counter()
        R.ref_forward(h, ptr(t), 1, int(p), ptr(logits))
        times.append(time.perf_counter() - t1)
    R.ref_model_free(h)
    timed = times[n_warm:]
    return {"per_token_s": timed, "cores": cores, "load_s": t_load, "prefill_s": t_prefill, "file_bytes": fbytes,
            "positions": [int(p) for p in positions[n_warm:total]]}


def cpu_baseline_dict(r):
    v = len(r["per_token_s"]) / sum(r["per_token_s"])
    return {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "sample": f"unmodified reference (oracle/_ref/libref.so, AVX2+FMA build, {r['cores']} threads) on the FULL 32-layer 7B INT8 model "
                      f"({r['file_bytes'] / 1e9:.1f} GB .flm written by our writer, loaded by the reference's load_flm): 32-token prompt prefilled, then "
                      f"{len(r['per_token_s'])} real single-token forwards timed at positions {r['positions'][0]}..{r['positions'][-1]} "
                      f"(mean {1e3 * sum(r['per_token_s']) / len(r['per_token_s']):.0f} ms/token, min {1e3 * min(r['per_token_s']):.0f}); no extrapolation"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    n_tok = max(1, args.steps)
    r = reference_decode(n_tok, max(1, min(args.warmup, 3)))
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref.so was never built (no /root/reference at build time)"}), flush=True)
        return
    base = cpu_baseline_dict(r)
    v = base["value"]
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "impl": "reference", "config": bench_config(1),
            "step_is": "one decode token of the reference at full depth (bounded sample of the 511-token step); value = tokens / measured seconds",
            "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "load_s": r["load_s"], "prefill_s": r["prefill_s"], "wall_s": time.perf_counter() - t0}
    if world > 1:
        line["note"] = ("the reference has one KV cache and one position (no multi-sequence batching, SURVEY D6): its throughput for a batch of "
                        "sequences is this single-sequence figure, sequences run one after the other")
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _events(torch, stream, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def _make_engine(fl, spec, local_rank, seed, **kw):
    eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, device=local_rank, **kw)
    int16 = kw.get("quant_type", fl.Q_INT8) == fl.Q_INT16
    for (kind, layer), (q, s) in Synth(spec, seed, int16=int16, gs=kw.get("group_size", 64)).items():
        eng.upload(kind, layer, q, s)
    eng.finalize()
    return eng


def other_configs(fl, torch, eng, local_rank, peak):
    """Short samples of the other BASELINE configurations, same process, CUDA events on the engine stream."""
    from fixtures import LLAMA2_13B
    out = []
    spec = shape_7b()
    rng = np.random.default_rng(11)
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    wbytes = eng.step_bytes(0) - 2 * spec.n_layers * spec.kv_dim * 4
    kv_tok = 2 * spec.n_layers * spec.kv_dim * 4
    # ---- configs[3] on one GPU: 8 sequences, one weight pass per step
    try:
        n, steps = SEQS_PER_GPU, 96
        for i in range(n):
            eng.forward(np.concatenate([[1], rng.integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32), 0, slot=i, want_logits=False)
        eng.decode_batch_async(n, 4); eng.sync()
        ms = _events(torch, stream, lambda: eng.decode_batch_async(n, steps))
        ctx = PROMPT + 4 + steps / 2
        b = wbytes + n * (ctx + 1) * kv_tok
        ach = b / (ms / steps * 1e-3) / 1e9
        out.append({"workload": f"configs[3] per GPU: 7B INT8, {n} sequences, prompt 32, one weight pass per step (tcgen05 GEMM, CUDA graph)",
                    "value": n * steps / ms * 1e3, "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
                    "roofline": {"achieved": ach, "peak": peak, "frac": ach / peak, "unit": "GB/s",
                                 "bytes_per_step": b, "note": f"weights + scales once + {n} x KV at mean ctx {ctx:.0f}"}})
    except Exception as ex:
        out.append({"workload": "configs[3] per GPU", "error": repr(ex)})
    return out


def other_engines(fl, torch, local_rank, peak):
    """configs[2] and configs[4]: own engines (created after the main engine was closed)."""
    from fixtures import LLAMA2_13B
    out = []
    rng = np.random.default_rng(12)
    # ---- configs[2]: 7B INT16 (int16 x int16 products in wrapping int32, CUDA cores: no tensor-core type)
    try:
        spec = shape_7b()
        eng = _make_engine(fl, spec, local_rank, 1, max_seq_len=1024, quant_type=fl.Q_INT16, group_size=64)
        stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
        eng.forward(np.concatenate([[1], rng.integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32), 0, want_logits=False)
        eng.decode_async(8); eng.sync()
        steps = 96
        ms = _events(torch, stream, lambda: eng.decode_async(steps))
        b = float(np.mean([eng.step_bytes(PROMPT + 9 + i) for i in range(steps)]))
        ach = b / (ms / steps * 1e-3) / 1e9
        out.append({"workload": "configs[2]: LLaMA2-7B INT16 (group 64) batch=1, decode from ctx 41 (persistent kernel)",
                    "value": steps / ms * 1e3, "unit": UNIT, "steps": steps, "ms_per_token": ms / steps,
                    "roofline": {"achieved": ach, "peak": peak, "frac": ach / peak, "unit": "GB/s", "bytes_per_token": b}})
        eng.close()
    except Exception as ex:
        out.append({"workload": "configs[2]: 7B INT16", "error": repr(ex)})
    # ---- configs[4]: 13B Q8_0 (group 32), prompt 2048 on the tensor-core prompt path, then decode
    try:
        spec = LLAMA2_13B
        P = 2048
        eng = _make_engine(fl, spec, local_rank, 2, max_seq_len=2048 + 512, quant_type=fl.Q_INT8, group_size=32)
        stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
        prompt = np.concatenate([[1], rng.integers(3, spec.vocab_size, P - 1)]).astype(np.int32)
        eng.forward(prompt[:128], 0, want_logits=False)            # warm the prompt path
        t0 = time.perf_counter()
        ttft_ms = _events(torch, stream, lambda: eng.forward(prompt, 0, want_logits=False))
        ttft_wall = time.perf_counter() - t0
        eng.decode_async(4); eng.sync()
        steps = 64
        ms = _events(torch, stream, lambda: eng.decode_async(steps))
        b = float(np.mean([eng.step_bytes(P + 5 + i) for i in range(steps)]))
        ach = b / (ms / steps * 1e-3) / 1e9
        out.append({"workload": "configs[4]: LLaMA2-13B Q8_0 (INT8 group 32) batch=1 prompt=2048, decode from ctx 2053 (persistent kernel)",
                    "value": steps / ms * 1e3, "unit": UNIT, "steps": steps, "ms_per_token": ms / steps,
                    "ttft_ms": ttft_ms, "ttft_note": f"2048-token prompt = 32 weight passes of 64 rows (tcgen05 GEMM); {P / ttft_ms * 1e3:.0f} prompt tokens/s; host wall {ttft_wall * 1e3:.0f} ms",
                    "roofline": {"achieved": ach, "peak": peak, "frac": ach / peak, "unit": "GB/s", "bytes_per_token": b}})
        eng.close()
    except Exception as ex:
        out.append({"workload": "configs[4]: 13B Q8_0", "error": repr(ex)})
    return out


def run_single(args, local_rank):
    import torch
    import __graft_entry__ as ge
    fl = ge._pkg()
    torch.cuda.set_device(local_rank)
    spec = shape_7b()
    t_load = time.perf_counter()
    eng = _make_engine(fl, spec, local_rank, 0, max_seq_len=1024, quant_type=fl.Q_INT8, group_size=64, max_seqs=SEQS_PER_GPU)
    t_load = time.perf_counter() - t_load
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    prompt = np.concatenate([[1], np.random.default_rng(7).integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32)
    n_dec = GEN - 1

    def one_step():
        """prefill + first token (untimed), then n_dec decode forwards (timed on the engine stream)."""
        eng.forward(prompt, 0, want_logits=False, want_argmax=True)
        l0 = eng.launch_count()
        ms = _events(torch, stream, lambda: eng.decode_async(n_dec))
        return ms, eng.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = [], 0
    for _ in range(args.steps):
        t, l = one_step()
        ms.append(t)
        launches += l
    clocks = sampler.stop()
    total_ms = float(sum(ms))
    value = args.steps * n_dec / (total_ms / 1e3)
    ttft_ms = min(_events(torch, stream, lambda: eng.forward(prompt, 0, want_logits=False)) for _ in range(3))

    # ---- e2e: host token in, host logits out, host argmax; every copy inside the timed region
    def e2e_step():
        logits = eng.forward(prompt, 0)
        tok = int(np.argmax(logits))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pos = PROMPT
        for _ in range(n_dec):
            logits = eng.forward(np.array([tok], np.int32), pos)
            tok = int(np.argmax(logits))
            pos += 1
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    e2e_step()
    e2e_t = float(np.mean([e2e_step() for _ in range(max(1, min(args.steps, 2)))]))
    e2e_value = n_dec / e2e_t

    peak, peak_src = _peaks()
    mean_ctx = PROMPT + 1 + (n_dec - 1) / 2.0
    step_bytes = float(np.mean([eng.step_bytes(PROMPT + 1 + i) for i in range(n_dec)]))
    ms_per_token = total_ms / (args.steps * n_dec)
    achieved = step_bytes / (ms_per_token * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "scope": f"decode_megakernel, ONE persistent launch = {n_dec} tokens (every phase of every layer of every token); algorithmic "
                         f"bytes per token = INT8 weights + fp32 group scales + fp32 KV read/write at mean ctx {mean_ctx:.0f} = {step_bytes / 1e9:.3f} GB "
                         "(SURVEY 8d); achieved = bytes per token / measured time per token (CUDA events on the engine stream)",
                "traffic": 7.350e9, "traffic_note": "dram read 7.332 GB + write 0.018 GB per token from ncu --set full of a 1-token launch at ctx 288 (profiles/r02/ncu_full_r02_mega_7b_int8_final.csv; L2 -> SM 8.69 GB: 7.18 GB of bulk copies + 1.5 GB of polls and K rows)",
                "peak_source": peak_src, "frac_of_8TBs": achieved / 8000.0}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": bench_config(1),
            "ms_per_token": ms_per_token, "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n_dec, "d2h_bytes_per_step": 4 * spec.vocab_size * n_dec,
                    "api": "fl_forward(host token, pos) -> host logits, host argmax"},
            "roofline": roofline, "load_s": t_load,
            "prefill": {"ttft_ms": ttft_ms, "note": "32-token prompt: one weight pass on the tensor-core prompt path (tcgen05 group-scaled INT8 GEMM) + first token"}}
    if not args.no_other_configs:
        oc = other_configs(fl, torch, eng, local_rank, peak)
        eng.close()
        eng = None
        oc += other_engines(fl, torch, local_rank, peak)
        line["other_configs"] = oc
    if eng is not None:
        eng.close()
    if not args.no_cpu_baseline:
        try:
            r = reference_decode(8, 2)
            line["cpu_baseline"] = cpu_baseline_dict(r) if r else {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": "oracle/_ref/libref.so not built"}
        except Exception as ex:   # the baseline is a report, never the measured path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def nccl_comm_for_engine(torch, dist, rank, world, local_rank):
    """An ncclComm_t of our own for fl_set_comm (the C-ABI takes the raw communicator): unique id from rank 0, broadcast through
    torch.distributed, ncclCommInitRank through ctypes on the libnccl the process already has loaded."""
    nccl = C.CDLL("libnccl.so.2")

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_byte * 128)]
    uid = UniqueId()
    if rank == 0:
        rc = nccl.ncclGetUniqueId(C.byref(uid))
        assert rc == 0, f"ncclGetUniqueId: {rc}"
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone().cuda(local_rank)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().numpy().tobytes())
    C.memmove(C.byref(uid), raw, 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank: {rc}"
    return nccl, comm


def run_multi(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    fl = ge._pkg()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nccl, comm = nccl_comm_for_engine(torch, dist, rank, world, local_rank)
    spec = shape_7b()
    n = SEQS_PER_GPU
    t_load = time.perf_counter()
    eng = _make_engine(fl, spec, local_rank, rank, max_seq_len=1024, quant_type=fl.Q_INT8, group_size=64, max_seqs=n)
    t_load = time.perf_counter() - t_load
    rc = fl.lib().fl_set_comm(eng.h, comm, rank, world)
    assert rc == 0, "fl_set_comm failed"
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    rng = np.random.default_rng(7 + rank)
    prompts = [np.concatenate([[1], rng.integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32) for _ in range(n)]
    n_dec = GEN4 - 1

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        for i, p in enumerate(prompts):
            eng.forward(p, 0, slot=i, want_logits=False)
        barrier()
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n_dec):
            eng.decode_batch_async(n, 1)
            rc = fl.lib().fl_allgather_tokens(eng.h, None, n, None)      # device-resident tokens, ncclAllGather on the engine stream
            assert rc == 0
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), eng.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = [], 0
    for _ in range(args.steps):
        t, l = one_step()
        ms.append(t)
        launches += l
    clocks = sampler.stop()
    tt = torch.tensor([float(sum(ms))], device=f"cuda:{local_rank}")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = world * n * args.steps * n_dec / (total_ms / 1e3)

    # ---- e2e: host tokens + positions in, host argmax out, host all-gather result out, every step
    def e2e_step():
        toks = np.array([eng.forward(p, 0, slot=i, want_logits=False, want_argmax=True) for i, p in enumerate(prompts)], np.int32)
        pos = np.full(n, PROMPT, np.int32)
        allt = np.zeros(n * world, np.int32)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_dec):
            toks = eng.forward_batch(toks, pos)
            pos += 1
            rc = fl.lib().fl_allgather_tokens(eng.h, toks.ctypes.data_as(C.c_void_p), n, allt.ctypes.data_as(C.c_void_p))
            assert rc == 0
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    e2e_t = float(np.mean([e2e_step() for _ in range(2)]))
    tt = torch.tensor([e2e_t], device=f"cuda:{local_rank}")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * n * n_dec / float(tt.item())

    peak, peak_src = _peaks()
    wbytes = eng.step_bytes(0) - 2 * spec.n_layers * spec.kv_dim * 4
    mean_ctx = PROMPT + 1 + (n_dec - 1) / 2.0
    step_bytes = wbytes + n * (mean_ctx + 1) * 2 * spec.n_layers * spec.kv_dim * 4
    ms_per_step = total_ms / (args.steps * n_dec)
    achieved = step_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "scope": f"per GPU and decode step: the step's kernels (qgemm_kernel = tcgen05 group-scaled INT8 GEMM for the five projections, attention, "
                         f"rmsnorm/quantise, argmax); algorithmic bytes per step = weights + scales ONCE + {n} sequences' fp32 KV at mean ctx {mean_ctx:.0f} = "
                         f"{step_bytes / 1e9:.3f} GB (SURVEY 8d); achieved = bytes / measured time per step (CUDA events on the engine stream, max over ranks)",
                "traffic": None, "peak_source": peak_src, "frac_of_8TBs": achieved / 8000.0}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": bench_config(world),
            "per_gpu_value": value / world, "ms_per_decode_step": ms_per_step, "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (4 * n * 2 + 4 * n) * n_dec, "d2h_bytes_per_step": (4 * n + 4 * n * world) * n_dec,
                    "api": "fl_forward_batch(host tokens, host positions) -> host argmax; fl_allgather_tokens(host local) -> host all"},
            "roofline": roofline, "load_s": t_load,
            "note": "N > 1 measures BASELINE configs[3] (8 sequences per GPU); the N = 1 line measures configs[1] (batch 1) and carries the 8-sequence "
                    "single-GPU figure under other_configs: compare per_gpu_value with that for the scaling of this configuration"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif world == 1:
        run_single(args, local_rank)
    else:
        run_multi(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
