timeout 600 python profiles/config_sweep.py 13b-q8_0 7b-int8-long > gpurun_out/config_sweep_uu6.log 2>&1; cat gpurun_out/config_sweep_uu6.log
timeout 900 python -m pytest tests -q -m gpu --timeout 400 -x > gpurun_out/test_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_all.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v_.json 2> gpurun_out/bench_v_.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_v_.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'])"
