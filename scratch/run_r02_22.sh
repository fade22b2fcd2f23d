for v in "" nolc; do
  echo "== variant '$v'"
  FL_PROF_LIB=$v timeout 600 python profiles/config_sweep.py 13b-q8_0 7b-int8-long > gpurun_out/config_sweep_$v.log 2>&1; cat gpurun_out/config_sweep_$v.log
done
timeout 600 python -m pytest tests/test_forward_gpu.py -q -m gpu --timeout 400 -x -k "long_context" > gpurun_out/test_long.log 2>&1; rc=$?; tail -3 gpurun_out/test_long.log; echo "long rc=$rc"
