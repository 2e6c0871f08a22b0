import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import vsl_oracle as O
from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
from gpu_helpers import run_cuda, forced_from
import emul_harness as E

def report(B, H, W, seed):
    S = 1
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=seed)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    opt = O.default_opt(sclm=0, height=H, width=W, batch_size=B, disparity_smoothness=0.0)
    losses, grads, maps = run_cuda(inputs, outputs, opt, False, noise)
    fw = E.forward_scale(inputs, outputs, opt, 0, False, noise)
    fw["sel"] = (maps[0]["src_idx"][:, 0] | (maps[0]["mask"][:, 0] << 2)).to(torch.uint8).contiguous()
    g_r = 1.0 / (float(maps[0]["mask"].sum()) + 1e-7)
    gd, gP = E.backward_scale(fw, opt, g_r, 0.0)
    d = (grads[("disp", 0)] - gd).abs()
    bad = d > 1e-4 * gd.abs().max()
    nz = torch.nonzero(bad)
    print("case", (B, H, W), "max rel %.3e nbad %d" % (float(d.max() / gd.abs().max()), int(bad.sum())))
    if len(nz):
        print("   bbox b %d..%d y %d..%d x %d..%d" % (nz[:, 0].min(), nz[:, 0].max(), nz[:, 2].min(), nz[:, 2].max(), nz[:, 3].min(), nz[:, 3].max()))
        for t in nz[:12].tolist():
            b, _, y, x = t
            print("    ", t, "gpu %.4e emul %.4e sel %d" % (float(grads[("disp", 0)][b, 0, y, x]), float(gd[b, 0, y, x]), int(fw["sel"][b, y, x])))

report(2, 40, 72, 13)
report(2, 48, 64, 13)
report(2, 40, 64, 13)
report(2, 48, 72, 13)
report(1, 40, 72, 13)
report(2, 64, 96, 13)
