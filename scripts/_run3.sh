python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if v>0.005})"
python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --tiles 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if v>0.005})"
python -m pytest tests/test_gpu_parity.py -x -q -k "not tiles and not False" 2>&1 | tail -2
