timeout 60 profiles/micro/pv_bench
(timeout 150 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python profiles/phase_times.py 288 64 2>&1 | grep -E "ctx|pv_|softmax|attn_|sum"
