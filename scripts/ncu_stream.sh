#!/bin/bash
# ncu --set full of the preparation + streaming launches of one bench step: scripts/ncu_stream.sh <tag> [variant.so]
tag=$1; lib=${2:-ppea_depth_b200/libppea_vsl.so}
PPEA_LIB=$lib python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_plain.log 2>&1 &&
PPEA_LIB=$lib ncu --set full --clock-control none --import-source on -k regex:"vsl_prep|vsl_stream|vsl_smooth" -s 6 -c 3 -o gpurun_out/prof_${tag} -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log | cut -c1-300
