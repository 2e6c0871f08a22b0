python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2p_tests.log
for pass in 1 2; do
for m in "" "--float-atomics" "--path multi" "--path multi --float-atomics"; do
python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-extras $m 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('[$m]', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if 'unused' not in k}, d['loss_check']['ok'], d['config']['deterministic_backward'], round(d['roofline']['step_frac_of_peak'],4))"
done; done 2>&1 | tee gpurun_out/r2p_det_ab.txt
