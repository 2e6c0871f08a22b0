PPEA_LIB=build/variants/ring5.so python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2n_tests.log
one() { # lib rows
  PPEA_PREP_SEG_ROWS=$2 PPEA_LIB=build/variants/$1.so python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/r2n_$1_$2.json 2>/dev/null
  python - <<P | tee -a gpurun_out/r2n_ab.txt
import json
d=json.loads(open("gpurun_out/r2n_$1_$2.json").read().strip().splitlines()[-1])
print("$1 rows $2 ms/step %.4f prep %.4f stream %.4f ok %s"%(d["ms_per_step"], d["roofline"]["stage_ms"]["vsl_prep_kernel"], d["roofline"]["stage_ms"]["vsl_stream_kernel"], d["loss_check"]["ok"]))
P
}
rm -f gpurun_out/r2n_ab.txt
one regs 0
for l in ring5 ring6 ring4; do for r in 0 16 20 24 32; do one $l $r; done; done
one regs 0
