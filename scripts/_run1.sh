python -m pytest tests/test_glue.py tests/test_pose.py -q -m gpu 2>&1 | tail -15
