for v in "" w7 w7nr nr ""; do
  FL_PROF_LIB=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_v_$v.json 2> gpurun_out/bench_v_$v.err; echo "variant '$v' rc=$?"; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_v_$v.json') if l.startswith('{')][-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'])"
done
