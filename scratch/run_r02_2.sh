set -x
timeout 900 python profiles/rows_bench.py 0 8 64 > gpurun_out/rows_bench_pdl.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/rows_bench_pdl.log
timeout 900 python profiles/rows_bench.py 2 8 64 > gpurun_out/rows_bench_nopdl.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/rows_bench_nopdl.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/test_all.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/test_all.log
