timeout 300 python profiles/phase_times.py 288 64 2>&1 | tail -26 > gpurun_out/phase_times_v5_gate.log; cat gpurun_out/phase_times_v5_gate.log
