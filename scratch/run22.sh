timeout 200 python -m pytest tests/test_generate_text.py -m gpu -x -q 2>&1 | tail -3
