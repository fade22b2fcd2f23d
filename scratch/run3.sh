timeout 400 python profiles/config_sweep.py 7b-int16 7b-int8-long 13b-q8_0 2>&1 | tee gpurun_out/config_sweep.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_forward_gpu.py -x -q -k "tiny-int8 and mega" 2>&1 | tail -8 | tee gpurun_out/sanitizer.log
