set -x
export FL_LAYERS=4
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rows_decode8.csv python profiles/rows_step.py decode 8 3 1 > gpurun_out/ncu_rows_decode8.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rows_prefill64.csv python profiles/rows_step.py prefill 64 2 1 > gpurun_out/ncu_rows_prefill64.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_rows_decode8.log
