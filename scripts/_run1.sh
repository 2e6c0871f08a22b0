python -m pytest tests/test_matching.py tests/test_cabi.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2k_tests.log
python scripts/bench_matching.py > gpurun_out/r2k_bench_matching.json 2> gpurun_out/r2k_bench_matching.err; cat gpurun_out/r2k_bench_matching.json
ncu --set full --clock-control none --import-source on -k regex:"match_" -s 24 -c 3 -o gpurun_out/prof_r2k_matching -f python scripts/bench_matching.py > gpurun_out/r2k_ncu.log 2>&1; tail -2 gpurun_out/r2k_ncu.log | cut -c1-200
