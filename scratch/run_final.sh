(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
(time timeout 600 python bench.py --steps 3 --warmup 3) > gpurun_out/bench.log 2>&1; tail -c 400 gpurun_out/bench.log
timeout 100 python profiles/phase_times.py 288 64 > gpurun_out/phase288.log 2>&1; cat gpurun_out/phase288.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"decode_megakernel|set_state" -c 40 --csv --log-file gpurun_out/launches_r01.csv python profiles/prof_step.py 288 8 > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:decode_megakernel -s 3 -c 1 -o gpurun_out/prof_r01_final python profiles/prof_step.py 288 2 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
