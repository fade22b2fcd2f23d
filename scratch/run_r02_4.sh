set -x
timeout 900 python -m pytest tests/test_tc_gemm_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/test_tc.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/test_tc.log
timeout 900 python profiles/rows_bench.py 0 8 64 > gpurun_out/rows_bench_pdl.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/rows_bench_pdl.log
export FL_LAYERS=4
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rows_decode8.csv python profiles/rows_step.py decode 8 5 1 > gpurun_out/ncu_rows_decode8.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rows_prefill64.csv python profiles/rows_step.py prefill 64 3 1 > gpurun_out/ncu_rows_prefill64.log 2>&1; echo rc=$?
