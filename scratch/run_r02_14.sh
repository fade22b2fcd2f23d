for v in "" hg ns cph2 ""; do
  FL_PROF_LIB=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v_$v.json 2> gpurun_out/bench_v_$v.err; echo "variant '$v' rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_v_$v.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'])"
done
FL_PROF_LIB=1 timeout 300 python profiles/phase_times.py 288 64 > gpurun_out/phase_times_v14.log 2>&1; head -26 gpurun_out/phase_times_v14.log
