#!/bin/bash
# A/B of library variants on the matching bench: scripts/ab_matching.sh name1 name2 ...
for v in "$@"; do
  PPEA_LIB=build/variants/$v.so python scripts/bench_matching.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'ms %.4f'%d['ms'], 'planar %.4f'%d['ms_planar_kernel'], 'same', d['quad_equals_planar_bitwise'], 'frac %.4f'%d['roofline']['frac'])"
done
