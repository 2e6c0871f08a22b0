#!/bin/bash
for v in "$@"; do
  PPEA_LIB=build/variants/$v.so python scripts/bench_decoder.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'fwd %.4f dx %.4f dw %.4f'%(d['fused_forward_ms'],d['grad_x_ms'],d['grad_weight_bias_ms']))"
done
