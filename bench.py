#!/usr/bin/env python
"""bench.py — decode tokens/s of the B200 engine on BASELINE.json's headline config, with roofline and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1): configs[1] = LLaMA2-7B-shaped INT8 (group 64), batch 1, prompt 32, gen 512, synthetic seeded weights.
A "step" is one pass of the hot path over that input: the 511 decode forwards that follow the prefill + first token
(the reference's own definition of decode speed, src/main.cpp:126,134: (total - first token) / (output tokens - 1)).
  value : decode tokens/s with everything resident in HBM (token fed back on the device, CUDA-graph replay)
  e2e   : the same through the reference-facing call fl_forward(host token, pos) -> host logits + host argmax
N>1: one process per GPU, weights replicated, one sequence per rank, NCCL all-gather of the sampled tokens per step.
--impl reference: the reference's own CPU forward() (oracle/_ref/libref.so) on the host cores, bounded sample.
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
METRIC = "decode tokens/s LLaMA2-7B INT8 batch=1 seq=128->512"
UNIT = "tokens/s"


def shape_7b():
    from fixtures import LLAMA2_7B
    return LLAMA2_7B


# ----------------------------------------------------------------------------------------------------------
def synth_int8_model(spec, seed=0, int16=False, gs=64):
    """Random INT8 (or INT16) payloads + fp32 group scales with realistic magnitudes (no 27 GB float model is materialised)."""
    from oracle_libs import (T_TOK_EMB, T_ATT_NORM, T_WQ, T_WK, T_WV, T_WO, T_FFN_NORM, T_W1, T_W2, T_W3, T_OUT_NORM, T_CLS)
    rng = np.random.default_rng(seed)
    d, h, kv = spec.dim, spec.hidden_dim, spec.kv_dim

    def qmat(rows, cols, sd):
        amp, lim, dt = (1800.0, 5792, np.int16) if int16 else (40.0, 127, np.int8)
        q = np.clip(np.rint(rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(amp)), -lim, lim).astype(dt)
        s = (np.float32(sd / amp) * (0.75 + 0.5 * rng.random((rows, cols // gs), dtype=np.float32))).astype(np.float32)
        return q, s

    base = {T_WQ: qmat(d, d, d ** -0.5), T_WK: qmat(kv, d, d ** -0.5), T_WV: qmat(kv, d, d ** -0.5), T_WO: qmat(d, d, d ** -0.5),
            T_W1: qmat(h, d, d ** -0.5), T_W3: qmat(h, d, d ** -0.5), T_W2: qmat(d, h, h ** -0.5)}
    yield (T_TOK_EMB, 0), ((rng.standard_normal((spec.vocab_size, d), dtype=np.float32) * np.float32(0.05)), None)
    yield (T_OUT_NORM, 0), (np.ones(d, np.float32), None)
    yield (T_CLS, 0), qmat(spec.vocab_size, d, d ** -0.5)
    for l in range(spec.n_layers):
        yield (T_ATT_NORM, l), ((1 + 0.1 * rng.standard_normal(d)).astype(np.float32), None)
        yield (T_FFN_NORM, l), ((1 + 0.1 * rng.standard_normal(d)).astype(np.float32), None)
        for k, (q, s) in base.items():
            # every layer gets its own HBM copy; rolling rows makes the contents differ without regenerating 200 MB
            yield (k, l), (np.roll(q, l * 7, axis=0), np.roll(s, l * 7, axis=0))


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


# ----------------------------------------------------------------------------------------------------------
def cpu_reference_baseline(max_seconds=40.0, threads=None):
    """The reference's own CPU forward() (oracle/_ref/libref.so) on this host: ONE 3-layer slice of the 7B shape (same dim /
    hidden / vocab, classifier shared with the embedding), min over a few decode tokens at ctx ~ PROMPT.  Decode time is
    proportional to the weight bytes swept (>94 % of it is quant::matmul), so the slice is scaled by
    (32 + c) / (3 + c) with c = vocab*dim / params-per-layer = 0.65 (the classifier in units of a layer).
    Falls back to the C restatement (1 core) if the reference library is absent."""
    from oracle_libs import ref, port, ptr, Q_INT8, PortConfig
    from fixtures import ModelSpec, gen_weights, write_llama2c, write_tokenizer_bin, synthetic_vocab, quantize_model
    spec7 = shape_7b()
    cores = threads or os.cpu_count() or 1
    R = ref()
    n_tok, L = 5, 3
    if R is not None:
        spec = ModelSpec(spec7.dim, spec7.hidden_dim, L, spec7.n_heads, spec7.n_kv_heads, spec7.vocab_size, 1024, True)
        one = gen_weights(ModelSpec(spec7.dim, spec7.hidden_dim, 1, spec7.n_heads, spec7.n_kv_heads, spec7.vocab_size, 1024, True), seed=1)
        w = {k: (np.broadcast_to(v[0], (L,) + v.shape[1:]) if k not in ("tok_emb", "out_norm", "cls") else v) for k, v in one.items()}
        with tempfile.TemporaryDirectory() as d:
            write_llama2c(d + "/m.bin", spec, w)
            write_tokenizer_bin(d + "/t.bin", synthetic_vocab(spec.vocab_size))
            del w, one
            h = R.ref_model_load((d + "/m.bin").encode(), (d + "/t.bin").encode(), 3, Q_INT8, cores, 64, 0)
        assert h, "reference failed to load the synthetic slice"
        logits = np.empty(spec.vocab_size, np.float32)
        prompt = np.arange(1, PROMPT + 1, dtype=np.int32)
        R.ref_forward(h, ptr(prompt), PROMPT, 0, ptr(logits))
        best = 1e9
        for i in range(n_tok):
            t = np.array([int(np.argmax(logits))], np.int32)
            t0 = time.perf_counter()
            R.ref_forward(h, ptr(t), 1, PROMPT + i, ptr(logits))
            best = min(best, time.perf_counter() - t0)
        R.ref_model_free(h)
        per_layer_params = 4 * spec7.dim * spec7.dim + 3 * spec7.dim * spec7.hidden_dim
        c = spec7.vocab_size * spec7.dim / per_layer_params
        per_token = best * (spec7.n_layers + c) / (L + c)
        return {"value": 1.0 / per_token, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"reference forward() (unmodified sources, AVX2+FMA build, {cores} threads) on a {L}-layer slice of the 7B shape: "
                          f"min of {n_tok} decode tokens at ctx {PROMPT} = {best * 1e3:.1f} ms, scaled by weight bytes to 32 layers "
                          f"((32+{c:.2f})/({L}+{c:.2f}))"}
    # port fallback: one layer's worth of matmul work on one core
    P = port()
    spec = ModelSpec(spec7.dim, spec7.hidden_dim, 1, spec7.n_heads, spec7.n_kv_heads, 2048, 1024, True)
    qm = quantize_model(spec, gen_weights(spec, 1), Q_INT8, 64)
    pc = PortConfig(spec.dim, spec.hidden_dim, 1, spec.n_heads, spec.n_kv_heads, spec.head_size, spec.vocab_size, 1024, Q_INT8, 64)
    pm = P.port_model_create(C.byref(pc))
    for (k, l), (q, s) in qm.items():
        P.port_model_set_tensor(pm, k, l, ptr(q), ptr(s) if s is not None else None, q.shape[0] if q.ndim == 2 else 1, q.shape[-1])
    logits = np.empty(spec.vocab_size, np.float32)
    t0 = time.perf_counter()
    for i in range(2):
        P.port_forward(pm, ptr(np.array([5 + i], np.int32)), 1, i, ptr(logits))
    per_layer = (time.perf_counter() - t0) / 2
    P.port_model_free(pm)
    per_token = per_layer * (spec7.n_layers + 0.65)
    return {"value": 1.0 / per_token, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "C restatement, one 7B-shaped layer (vocab 2048), 2 tokens, scaled to 32 layers + lm_head"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    spec = shape_7b()
    t0 = time.perf_counter()
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 2))):
        base = cpu_reference_baseline()
        vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (GEN - 1) / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "impl": "reference",
            "config": {"workload": "LLaMA2-7B INT8 .flm-shaped batch=1 prompt=32 gen=512 (reference CPU forward, bounded sample)",
                       "dim": spec.dim, "hidden_dim": spec.hidden_dim, "n_layers": spec.n_layers, "vocab": spec.vocab_size},
            "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    fl = ge._pkg()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    spec = shape_7b()
    max_seq = 1024
    eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size,
                    max_seq_len=max_seq, quant_type=fl.Q_INT8, group_size=64, max_seqs=1, device=local_rank)
    t_load = time.perf_counter()
    for (kind, layer), (q, s) in synth_int8_model(spec, seed=rank):
        eng.upload(kind, layer, q, s)
    eng.finalize()
    t_load = time.perf_counter() - t_load
    stream = torch.cuda.ExternalStream(eng.stream, device=local_rank)
    prompt = np.concatenate([[1], np.random.default_rng(7 + rank).integers(3, spec.vocab_size, PROMPT - 1)]).astype(np.int32)
    n_dec = GEN - 1
    class _DevInt:      # zero-copy torch view of the engine's sampled-token word (fl_device_ptr)
        def __init__(self, p):
            self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (int(p), False), "version": 2}
    tok_buf = torch.as_tensor(_DevInt(eng.device_ptr("argmax")), device=f"cuda:{local_rank}")
    gathered = torch.zeros(world, dtype=torch.int32, device=f"cuda:{local_rank}") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(timed):
        """prefill + first token (untimed), then n_dec decode forwards (timed on the engine stream)."""
        eng.forward(prompt, 0, want_logits=False, want_argmax=True)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        ev0.record(stream)
        if world == 1:
            eng.decode_async(n_dec)
        else:
            # batch > 1 across GPUs: one NCCL all-gather of the sampled tokens per decode step
            for _ in range(n_dec):
                eng.decode_async(1)
                with torch.cuda.stream(stream):
                    fl.shard.gather_tokens(tok_buf, out=gathered)
        ev1.record(stream)
        barrier()
        return ev0.elapsed_time(ev1), eng.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        one_step(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = [], 0
    for _ in range(args.steps):
        t, l = one_step(True)
        ms.append(t)
        launches += l
    clocks = sampler.stop()
    total_ms = float(sum(ms))
    if world > 1:
        tt = torch.tensor([total_ms], device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * args.steps * n_dec / (total_ms / 1e3)

    # ---- e2e: host token in, host logits out, host argmax; every copy inside the timed region
    def e2e_step():
        logits = eng.forward(prompt, 0)
        tok = int(np.argmax(logits))
        barrier()
        t0 = time.perf_counter()
        pos = PROMPT
        for _ in range(n_dec):
            logits = eng.forward(np.array([tok], np.int32), pos)
            tok = int(np.argmax(logits))
            pos += 1
        torch.cuda.synchronize()
        return time.perf_counter() - t0
    e2e_step()
    e2e_s = [e2e_step() for _ in range(max(1, min(args.steps, 2)))]
    e2e_t = float(np.mean(e2e_s))
    if world > 1:
        tt = torch.tensor([e2e_t], device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = float(tt.item())
    e2e_value = world * n_dec / e2e_t

    # ---- roofline of the whole decode step and of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    mean_ctx = PROMPT + 1 + (n_dec - 1) / 2.0
    step_bytes = float(np.mean([eng.step_bytes(PROMPT + 1 + i) for i in range(n_dec)]))
    ms_per_token = total_ms / (args.steps * n_dec)
    achieved = step_bytes / (ms_per_token * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "scope": f"decode_megakernel, {'ONE persistent launch' if world == 1 else 'one launch per token (+ token all-gather),'} = {n_dec} tokens (every phase of every layer of every token); algorithmic "
                         f"bytes per token = INT8 weights + fp32 group scales + fp32 KV read/write at mean ctx {mean_ctx:.0f} = {step_bytes / 1e9:.3f} GB "
                         "(SURVEY 8d); achieved = bytes per token / measured time per token (CUDA events on the engine stream)",
                "traffic": 7.383e9, "traffic_note": "dram read 7.353 GB + write 0.030 GB per token from ncu --set full of a 1-token launch at ctx 288 (profiles/r01/ncu_full_megakernel_v6.csv)",
                "peak_source": peak_src, "frac_of_8TBs": achieved / 8000.0}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": {"workload": "LLaMA2-7B INT8 (group 64) batch=1/GPU prompt=32 gen=512, synthetic seeded weights",
                       "dim": spec.dim, "hidden_dim": spec.hidden_dim, "n_layers": spec.n_layers, "vocab": spec.vocab_size,
                       "kv": "fp32", "step": f"{n_dec} decode forwards after prefill+first token",
                       "l2": "inputs larger than L2 (7.0 GB of weights streamed per token vs 126 MB L2)",
                       "parallelism": f"dp{world} (weights replicated, 1 sequence per GPU, token all-gather per step)" if world > 1 else "single GPU"},
            "ms_per_token": ms_per_token, "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * n_dec, "d2h_bytes_per_step": 4 * spec.vocab_size * n_dec,
                    "api": "fl_forward(host token, pos) -> host logits, host argmax"},
            "roofline": roofline, "load_s": t_load}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_baseline()
        except Exception as ex:   # the baseline is a report, never the measured path
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
