python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo rc=$?
tail -c 600 gpurun_out/r2_bench_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_full.json'))
print('value',d['value'],'ms',d['ms_per_step'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['repeats_ms_per_step'],'f32',d['e2e']['float32_frames']['value'],d['e2e']['float32_frames']['repeats_ms_per_step'])
print('roof',d['roofline']['frac'],d['roofline']['kernel_ms'],d['roofline']['stage_ms'])
print('cpu',d['cpu_baseline'])
print('loss_check',d['loss_check'])
print('sweep96',d.get('sharded_sweep96'))
print('eager',d.get('torch_cuda_eager'))
print('clocks',d['clocks'])
PY
