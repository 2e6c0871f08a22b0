for v in build/variants/lib_prep3.so build/variants/lib_prep4.so build/variants/lib_prep5.so; do
  echo "== variant: $v"
  PPEA_LIB=$v python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['ms_per_step'], {k:round(v,4) for k,v in d['roofline']['stage_ms'].items() if v>0.005})"
done
