set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "not tiles and not False" 2>&1 | tail -40 > gpurun_out/r2a_tests.log
python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --tiles > gpurun_out/r2a_bench_tiles.json 2> gpurun_out/r2a_bench_tiles.err
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
