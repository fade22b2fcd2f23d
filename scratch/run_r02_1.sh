set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python profiles/tc_probe.py 0 > gpurun_out/tc_probe_v0.log 2>&1; echo "probe0 rc=$?"
tail -30 gpurun_out/tc_probe_v0.log
timeout 300 python profiles/tc_probe.py 1 > gpurun_out/tc_probe_v1.log 2>&1; echo "probe1 rc=$?"
tail -30 gpurun_out/tc_probe_v1.log
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/test_tc.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/test_tc.log
