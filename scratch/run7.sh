timeout 900 python bench.py > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; tail -1 gpurun_out/bench_v5.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_v5.json 2>/dev/null; tail -1 gpurun_out/bench_ref_v5.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
