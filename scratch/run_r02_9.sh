set -x
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/test_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/test_all.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['prefill'])"
FL_PROF_LIB=1 timeout 600 python profiles/phase_times.py 288 64 > gpurun_out/phase_times_v9.log 2>&1; head -24 gpurun_out/phase_times_v9.log
