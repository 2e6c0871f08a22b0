#!/bin/bash
# scripts/build_variant.sh <name> [nvcc flags...]: builds build/variants/<name>.so from the tree with extra -D flags (A/B runs: scripts/ab_variants.sh)
name=$1; shift
mkdir -p build/variants
python - "$name" "$@" <<'P'
import subprocess, sys
from ppea_depth_b200 import _cabi
out = "build/variants/%s.so" % sys.argv[1]
cmd = _cabi.nvcc_command(out=out, extra=["-Xptxas", "-v"] + sys.argv[2:])
r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
if r.returncode != 0:
    print(r.stdout[-4000:]); sys.exit(1)
import re
txt = r.stdout
# registers / spills of the two launches of the streaming step
for m in re.finditer(r"Compiling entry function '(_ZN4ppea(?:17vsl_stream_kernelILb1ELb0ELb0|15vsl_prep_kernel|22vsl_smooth_tail_kernel|21match_features_kernel|26match_features_quad_kernel|19disp_head_dw_kernel)[^']*)'.*?\n(?:.*\n){0,3}?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt):
    print(sys.argv[1], m.group(1)[8:40], "regs", m.group(5), "spill", m.group(3), m.group(4))
print("built", out)
P
