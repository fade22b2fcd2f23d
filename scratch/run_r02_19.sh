timeout 180 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; tail -2 gpurun_out/smoke.log; echo "smoke rc=$rc"
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > gpurun_out/test_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/test_all.log
for v in "" ""; do
  FL_PROF_LIB=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v_$v.json 2> gpurun_out/bench_v_$v.err; echo "variant '$v' rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_v_$v.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'])"
done
FL_PROF_LIB=1 timeout 300 python profiles/phase_times.py 288 64 > gpurun_out/phase_times_v19.log 2>&1; head -26 gpurun_out/phase_times_v19.log
