"""N > 1 host logic on CPU: two gloo processes shard a request batch, 'decode' with a deterministic stand-in for the
engine (token = f(sequence, step)), all-gather the sampled tokens every step and must agree on the full token matrix and
on which sequences have ended (token id 0, transformer.cpp:93)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_token(seq, step):
    v = (seq * 7919 + step * 104729) % 97
    return 0 if v == 13 else 3 + v          # 0 ends a sequence now and then


def _worker(rank, world, port, batch, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    sh = ge._pkg().shard
    mine = sh.local_sequences(batch, world, rank)
    assert all(sh.owner_of(s, batch, world)[0] == rank for s in mine)
    assert [sh.owner_of(s, batch, world)[1] for s in mine] == list(range(len(mine)))
    done = torch.zeros(batch, dtype=torch.bool)
    hist = []
    for step in range(steps):
        local = torch.tensor([_fake_token(s, step) for s in mine], dtype=torch.int32)
        allt = sh.gather_tokens(local)
        assert allt.shape == (batch,)
        done = sh.finished_mask(allt, done)
        hist.append(allt.clone())
    q.put((rank, torch.stack(hist).numpy(), done.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_on_tokens_and_stops():
    world, batch, steps = 2, 8, 12
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([[_fake_token(s, t) for s in range(batch)] for t in range(steps)], np.int32)
    for rank, hist, done in res:
        assert np.array_equal(hist, want), rank
        assert np.array_equal(done, (want == 0).any(0)), rank


def test_single_rank_needs_no_collective():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    sh = ge._pkg().shard
    t = torch.tensor([5, 6], dtype=torch.int32)
    assert sh.gather_tokens(t) is t
    assert sh.local_sequences(4, 1, 0) == [0, 1, 2, 3]
    try:
        sh.local_sequences(5, 2, 0)
    except ValueError:
        pass
    else:
        raise AssertionError("uneven batch must be rejected")
