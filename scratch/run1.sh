python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<EOP
import json
d=json.loads([l for l in open("gpurun_out/bench.log") if l.startswith("{")][-1])
print("BENCH", d["value"], d["ms_per_token"], d["e2e"]["value"], d["roofline"]["frac"])
EOP
