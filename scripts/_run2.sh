set -x
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"vsl_prep|vsl_stream" -s 4 -c 2 -o gpurun_out/prof_r2e python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
tail -3 gpurun_out/r2e_ncu.log
