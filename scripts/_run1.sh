python -m pytest tests/test_matching.py -q -m gpu 2>&1 | tail -15
