set -x
python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vsl_stream -s 2 -c 1 -o gpurun_out/prof_r2d python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
