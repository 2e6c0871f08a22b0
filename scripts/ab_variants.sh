#!/bin/bash
# A/B of library variants on one box: scripts/ab_variants.sh name1 name2 ...  (build/variants/<name>.so), two alternating passes
mkdir -p gpurun_out
for pass in 1 2; do
  for v in "$@"; do
    PPEA_LIB=build/variants/$v.so python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ab_${v}_${pass}.json 2> gpurun_out/ab_${v}_${pass}.err
    python - <<P
import json
d=json.loads(open("gpurun_out/ab_${v}_${pass}.json").read().strip().splitlines()[-1])
st={k:round(x,4) for k,x in d["roofline"]["stage_ms"].items() if "unused" not in k}
print("$v pass $pass ms/step %.4f"%d["ms_per_step"], st, d.get("loss_check",{}).get("ok"))
P
  done
done
