python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_r01_final.log; cat gpurun_out/pytest_gpu_r01_final.log
