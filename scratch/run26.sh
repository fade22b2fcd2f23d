timeout 25 python -m pytest tests/test_flm.py tests/test_generate_text.py -m gpu -x -q 2>&1 | tail -2
