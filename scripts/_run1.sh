python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2l_tests.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 600 gpurun_out/r2l_bench.json
bash scripts/ncu_stream.sh r2l
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2l_launches.log 2>&1; tail -3 gpurun_out/r2l_launches.csv | cut -c1-200
