timeout 180 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; tail -2 gpurun_out/smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu --timeout 400 -x -k "long_context" > gpurun_out/test_long.log 2>&1; rc=$?; tail -5 gpurun_out/test_long.log; echo "long rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v_.json 2> gpurun_out/bench_v_.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_v_.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'])"
timeout 600 python profiles/config_sweep.py 13b-q8_0 7b-int8-long 7b-int16 > gpurun_out/config_sweep_v20.log 2>&1; cat gpurun_out/config_sweep_v20.log
