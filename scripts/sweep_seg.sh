#!/bin/bash
# tuning: streaming-step segment length sweep (PPEA_STREAM_SEG_ROWS overrides the per-launch choice)
for seg in "$@"; do
  PPEA_STREAM_SEG_ROWS=$seg python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/seg_${seg}.json 2> gpurun_out/seg_${seg}.err
  python - <<P
import json
d=json.loads(open("gpurun_out/seg_${seg}.json").read().strip().splitlines()[-1])
st={k:round(x,4) for k,x in d["roofline"]["stage_ms"].items() if "unused" not in k}
print("seg $seg ms/step %.4f"%d["ms_per_step"], st, d.get("loss_check",{}).get("ok"))
P
done
