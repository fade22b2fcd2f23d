set -x
export FL_PROF_LIB=1
FL_DEBUG_SKIP=1 timeout 600 python profiles/phase_times.py 288 64 > gpurun_out/phase_times_skip1.log 2>&1; tail -30 gpurun_out/phase_times_skip1.log
timeout 600 python profiles/phase_times.py 288 64 > gpurun_out/phase_times_v7.log 2>&1; tail -30 gpurun_out/phase_times_v7.log
unset FL_PROF_LIB
# 2-GPU checks are a separate call
