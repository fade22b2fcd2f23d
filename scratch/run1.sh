(timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for w in 3 6 99; do echo "== window $w"; FL_WINDOW=$w timeout 100 python profiles/phase_times.py 288 64 2>&1 | grep -E "ctx|ll_wait|attn_qkv|stage_wait|wait_first"; done
