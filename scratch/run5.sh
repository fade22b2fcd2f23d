timeout 900 python -m pytest tests/test_full_size_gpu.py -x -q --durations=5 2>&1 | tail -15
