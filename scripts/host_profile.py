"""Host-side cost of one public-API step (generate_images_pred + compute_losses + backward) with resident inputs:
wall time per step, and a cProfile of where the Python time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from types import SimpleNamespace
from ppea_depth_b200.loss import ViewSynthesisLoss
from ppea_depth_b200.synth import SynthConfig, make_batch
B, H, W, S = 12, 192, 640, 4
inputs, outputs = make_batch(SynthConfig(batch=B, height=H, width=W, num_scales=S))
dev = "cuda"
ins = {k: v.to(dev) for k, v in inputs.items()}
base = {k: v.to(dev) for k, v in outputs.items()}
opt = SimpleNamespace(sclm=S - 1, v1_multiscale=False, height=H, width=W, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1],
                      disable_automasking=False, no_ssim=False, selec_reproj=True, disable_motion_masking=False,
                      no_matching_augmentation=False, batch_size=B, disparity_smoothness=1e-3)
mod = ViewSynthesisLoss(opt, noise_mode="device")
def step(sync=False):
    outs = dict(base)
    for s in range(S):
        outs[("disp", s)] = base[("disp", s)].detach().requires_grad_(True)
    for f in (-1, 1):
        outs[("cam_T_cam", 0, f)] = base[("cam_T_cam", 0, f)].detach().requires_grad_(True)
    mod.generate_images_pred(ins, outs, False)
    losses, _ = mod.compute_losses(ins, outs, False)
    losses["loss"].backward()
    if sync:
        return float(losses["loss"].item())
for _ in range(10): step()
n = 200
for sync in (False, True):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): step(sync)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("public API, resident inputs, device noise, %s: %.4f ms/step (wall), %.0f Mpix/s" % (
        "loss read back every step" if sync else "no readback", (t1 - t0) / n * 1e3, B * H * W / ((t1 - t0) / n) / 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
