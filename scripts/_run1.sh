python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_tests.log
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --tiles > gpurun_out/r2_bench_tiles.json 2> gpurun_out/r2_bench_tiles.err
tail -8 gpurun_out/r2_tests.log; cat gpurun_out/r2_bench.json; cat gpurun_out/r2_bench_tiles.json; tail -3 gpurun_out/r2_bench.err
