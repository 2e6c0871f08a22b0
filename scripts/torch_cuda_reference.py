"""Context number (not part of bench.py): the reference's own ATen op sequence (oracle/vsl_oracle.py, the same
calls as trainer.py:871-918 + 1032-1160) executed by PyTorch on the SAME B200 -- what a user of the reference runs
today -- next to the fused path.  KITTI 12x3x192x640, 4 scales, mono, fwd+bwd, inputs resident, device noise."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vsl_oracle as O
from ppea_depth_b200.synth import SynthConfig, make_batch

B, H, W, S = 12, 192, 640, 4
inputs, outputs = make_batch(SynthConfig(batch=B, height=H, width=W, num_scales=S))
dev = "cuda"
ins = {k: v.to(dev) for k, v in inputs.items()}
base = {k: v.to(dev) for k, v in outputs.items()}
opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)

def step():
    outs = dict(base)
    for s in range(S):
        outs[("disp", s)] = base[("disp", s)].detach().requires_grad_(True)
    for f in (-1, 1):
        outs[("cam_T_cam", 0, f)] = base[("cam_T_cam", 0, f)].detach().requires_grad_(True)
    noise = [torch.randn(B, 1, H, W, device=dev) for _ in range(S)]
    losses, _ = O.view_synthesis_losses(ins, outs, opt, False, noise)
    losses["loss"].backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("torch-CUDA eager reference path: %.3f ms/step = %.1f Mpix/s (peak mem %.0f MB)" % (ms, B * H * W / ms / 1e3, torch.cuda.max_memory_allocated() / 1e6))
