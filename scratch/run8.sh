timeout 300 python -m pytest tests/test_forward_gpu.py -x -q -k "kv_slots or forward_batch" 2>&1 | tail -8
timeout 300 python profiles/multiseq_bench.py 8 32 160 2>&1 | tail -5
FL_NO_MULTISEQ=1 timeout 300 python profiles/multiseq_bench.py 8 16 160 2>&1 | tail -5
