"""Small fused forward+backward cases for compute-sanitizer (ragged sizes, all modes)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gpu_helpers import run_cuda
from oracle import vsl_oracle as O
from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
for (B, H, W, S, multi, det) in [(2, 33, 47, 3, False, False), (1, 40, 72, 2, True, False), (2, 32, 64, 4, False, True), (1, 19, 35, 1, False, False)]:
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=3)
    inputs, outputs = make_batch(cfg)
    for s in range(1, S):
        outputs[("disp", s)] = outputs[("disp", s)][..., :H >> s, :W >> s].contiguous()
        inputs[("color", 0, s)] = inputs[("color", 0, s)][..., :H >> s, :W >> s].contiguous()
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    losses, grads, maps = run_cuda(inputs, outputs, opt, multi, None if multi else make_noise(cfg, S), deterministic=det)
    print((B, H, W, S, multi, det), float(losses["loss"]), {k: float(v.abs().max()) for k, v in list(grads.items())[:2]})
print("done")
