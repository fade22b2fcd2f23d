set -x
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/test_all.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/test_all.log
timeout 900 python profiles/relaxed_ab.py > gpurun_out/relaxed_ab.log 2>&1; echo "rc=$?"; tail -20 gpurun_out/relaxed_ab.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json
