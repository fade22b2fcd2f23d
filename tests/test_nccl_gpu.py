"""fl_set_comm / fl_allgather_tokens (weights replicated, request batch sharded, one all-gather of the sampled
tokens per step).  With one GPU: the world == 1 path.  With two or more: two ranks under torch.distributed.run, a raw
ncclComm_t bound to each engine, host and device-resident forms (tests/nccl_worker.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_libs import Q_INT8
from fixtures import TINY, gen_weights, quantize_model, prompt_tokens

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_allgather_single_rank_is_a_copy(fl):
    spec = TINY
    qm = quantize_model(spec, gen_weights(spec, seed=1), Q_INT8, 64)
    eng = fl.Engine(spec.dim, spec.hidden_dim, spec.n_layers, spec.n_heads, spec.n_kv_heads, spec.vocab_size, max_seqs=2)
    eng.upload_model(qm)
    assert fl.lib().fl_set_comm(eng.h, None, 0, 1) == 0
    mine = np.array([7, 9], np.int32)
    out = np.zeros(2, np.int32)
    assert fl.lib().fl_allgather_tokens(eng.h, mine.ctypes.data_as(C.c_void_p), 2, out.ctypes.data_as(C.c_void_p)) == 0
    assert out.tolist() == [7, 9]
    toks = [eng.forward(prompt_tokens(spec, 5 + i, seed=i), 0, slot=i, want_logits=False, want_argmax=True) for i in range(2)]
    assert fl.lib().fl_allgather_tokens(eng.h, None, 2, out.ctypes.data_as(C.c_void_p)) == 0
    assert out.tolist() == toks
    assert fl.lib().fl_set_comm(eng.h, None, 0, 2) < 0          # world > 1 needs a communicator
    eng.close()


def test_allgather_two_ranks_over_nccl(fl):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(HERE, "nccl_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("nccl worker ok") == 2, r.stdout[-2000:]
