#!/bin/bash
mkdir -p gpurun_out
run() {  # name lib args...
  n=$1; lib=$2; shift 2
  PPEA_LIB=$lib python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ab2_$n.json 2> gpurun_out/ab2_$n.err
  python - <<P
import json
d=json.loads(open("gpurun_out/ab2_$n.json").read().strip().splitlines()[-1])
st={k:round(x,4) for k,x in d["roofline"]["stage_ms"].items() if "unused" not in k}
print("$n $* ms/step %.4f"%d["ms_per_step"], st, d.get("loss_check",{}).get("ok"))
P
}
for pass in 1 2; do
  for v in orig new; do
    lib=build/variants/orig.so; [ $v = new ] && lib=ppea_depth_b200/libppea_vsl.so
    run ${v}_mono_$pass $lib
    run ${v}_multi_$pass $lib --path multi
    run ${v}_multi_fa_$pass $lib --path multi --float-atomics
    run ${v}_cs_multi_$pass $lib --path multi --workload cityscapes
    run ${v}_mono_fa_$pass $lib --float-atomics
  done
done
