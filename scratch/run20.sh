timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -k "kv_slots or forward_batch" 2>&1 | tail -2
timeout 200 python profiles/multiseq_bench.py 8 24 160 2>&1 | grep "n_seqs [28]:"
FL_DEBUG_SKIP=32 timeout 200 python profiles/multiseq_bench.py 8 24 160 2>&1 | grep "n_seqs [28]:"
