"""Worker of tests/test_nccl_gpu.py (run under torch.distributed.run, one rank per GPU): binds a raw ncclComm_t to the engine
(fl_set_comm) and checks fl_allgather_tokens in both forms — host tokens in / host tokens out, and the device-resident form
that gathers the tokens the engine has just sampled — against the tokens every rank computes for itself."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from bench import nccl_comm_for_engine
    from oracle_libs import Q_INT8
    from fixtures import TINY, gen_weights, quantize_model, prompt_tokens
    fl = ge._pkg()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nccl, comm = nccl_comm_for_engine(torch, dist, rank, world, local)
    spec, n = TINY, 3
    qm = quantize_model(spec, gen_weights(spec, seed=1), Q_INT8, 64)       # the same weights on every rank (replicated)
    eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seqs=n, device=local)
    eng.upload_model(qm)
    assert fl.lib().fl_set_comm(eng.h, comm, rank, world) == 0
    # host form
    mine = np.arange(n, dtype=np.int32) + 100 * rank
    allt = np.zeros(n * world, np.int32)
    assert fl.lib().fl_allgather_tokens(eng.h, mine.ctypes.data_as(C.c_void_p), n, allt.ctypes.data_as(C.c_void_p)) == 0
    want = np.concatenate([np.arange(n, dtype=np.int32) + 100 * r for r in range(world)])
    assert allt.tolist() == want.tolist(), (rank, allt)
    # device-resident form: every rank decodes its own shard of the request batch; all ranks end up with all tokens
    def shard(r):
        return [prompt_tokens(spec, 4 + i + r, seed=10 * r + i) for i in range(n)]
    firsts = [eng.forward(p, 0, slot=i, want_logits=False, want_argmax=True) for i, p in enumerate(shard(rank))]
    eng.decode_batch_async(n, 2)
    got = np.zeros(n * world, np.int32)
    assert fl.lib().fl_allgather_tokens(eng.h, None, n, got.ctypes.data_as(C.c_void_p)) == 0
    # what the other ranks must have sampled: recompute their sequences here, one at a time
    solo = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, device=local, flags=fl.FLAG_NO_TC)
    solo.upload_model(qm)
    exp = []
    for r in range(world):
        for p in shard(r):
            exp.append(int(solo.generate_greedy(p, 2)[2]))
    assert got.tolist() == exp, (rank, got.tolist(), exp)
    assert firsts is not None
    eng.close(); solo.close()
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: nccl worker ok", flush=True)


if __name__ == "__main__":
    main()
