"""Round-1 side measurements on one GPU: other workloads / paths / modes of bench.py, and the resident
throughput of the public autograd API (no CUDA graph, no H2D) to expose host overhead."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

def bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "10", "--no-cpu-baseline", "--no-e2e", *args],
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    st = {k: round(v, 4) for k, v in d["roofline"]["stage_ms"].items() if "unused" not in k}
    print("bench %-48s ms/step %.4f  Mpix/s %.0f  frac_step %.4f  %s" % (" ".join(args), d["ms_per_step"], d["value"], d["roofline"]["step_frac_of_peak"], st), flush=True)

MODES = ([], ["--path", "multi"], ["--float-atomics"], ["--path", "multi", "--float-atomics"], ["--workload", "cityscapes"],
         ["--workload", "cityscapes", "--path", "multi"], ["--workload", "hires"], ["--workload", "sweep96"])
for a in MODES:
    bench(*a)
    if "--pair" in sys.argv:
        bench(*a, "--no-fused")
if "--bench-only" in sys.argv:
    sys.exit(0)

# public API, inputs resident
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
def step():
    outs = dict(base)
    for s in range(S):
        outs[("disp", s)] = base[("disp", s)].detach().requires_grad_(True)
    for f in (-1, 1):
        outs[("cam_T_cam", 0, f)] = base[("cam_T_cam", 0, f)].detach().requires_grad_(True)
    mod.generate_images_pred(ins, outs, False)
    losses, _ = mod.compute_losses(ins, outs, False)
    losses["loss"].backward()
for _ in range(10): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 200
for _ in range(n): step()
torch.cuda.synchronize(); t1 = time.perf_counter()
print("public API, resident inputs, device noise: %.4f ms/step (wall), %.0f Mpix/s" % ((t1 - t0) / n * 1e3, B * H * W / ((t1 - t0) / n) / 1e6))
t0 = time.perf_counter()
for _ in range(n):
    outs = dict(base)
t_host = (time.perf_counter() - t0) / n
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
