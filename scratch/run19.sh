timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -1 gpurun_out/bench_v6.json | cut -c1-400
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
