python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r2_tests.log
tail -12 gpurun_out/r2_tests.log
