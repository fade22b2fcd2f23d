(time timeout 400 python -m pytest tests/test_loaders.py -x -q -m gpu) 2>&1 | tail -8
