timeout 600 python -m pytest tests/test_full_size_gpu.py tests/test_forward_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
