python -m pytest tests/test_matching.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2u_tests.log
scripts/ab_matching.sh mq mp2 mp1 mq mp2 2>&1 | tee gpurun_out/r2u_matching.txt
