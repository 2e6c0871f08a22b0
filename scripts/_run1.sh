python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2i_tests.log; cat gpurun_out/r2i_tests.log
scripts/ab_variants.sh base2 tailnopin tail tail_mc tail_c4 2>&1 | tee gpurun_out/r2i_ab.txt
