python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2s_tests.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -c 300 gpurun_out/r2s_bench.json
bash scripts/ncu_stream.sh r2s
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2s_launches.log 2>&1; grep -c vsl_ gpurun_out/r2s_launches.csv
python scripts/measure_variants.py --bench-only 2>/dev/null | tee gpurun_out/r2s_modes.txt | cut -c1-110
