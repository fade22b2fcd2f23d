python -m pytest tests/test_gguf.py tests/test_flm.py tests/test_sampler.py -m gpu -x -q 2>&1 | tail -6
