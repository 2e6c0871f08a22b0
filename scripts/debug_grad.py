import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import vsl_oracle as O
from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
from gpu_helpers import run_cuda, forced_from
import emul_harness as E

def report(B, H, W, S, seed, multi=False):
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=seed)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    losses, grads, maps = run_cuda(inputs, outputs, opt, multi, None if multi else noise)
    forced = forced_from(maps)
    _, g64, _ = O.run_fwd_bwd(inputs, outputs, opt, multi, noise, dtype=torch.float64, forced=forced)
    _, g32, _ = O.run_fwd_bwd(inputs, outputs, opt, multi, noise, forced=forced)
    opt0 = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B, disparity_smoothness=0.0)
    _, g64ns, _ = O.run_fwd_bwd(inputs, outputs, opt0, multi, noise, dtype=torch.float64, forced=forced)
    print("case", (B, H, W, S), "multi", multi)
    for k, ref in g64.items():
        d = (grads[k].double() - ref)
        e = d.abs()
        idx = torch.nonzero(e == e.max())[0].tolist()
        print("  ", k, "max err %.3e rel %.3e at %s ; l2 rel %.3e ; oracle32 rel max %.3e" % (
            float(e.max()), float(e.max() / ref.abs().max()), idx, float(d.norm() / ref.norm()),
            float((g32[k].double() - ref).abs().max() / ref.abs().max())))
    # emulator gradient (no smoothness) vs GPU minus smoothness part
    for s in range(S):
        fw = E.forward_scale(inputs, outputs, opt, s, multi, noise)
        # use GPU selection in the emulator backward
        fw["sel"] = (maps[s]["src_idx"][:, 0] | (maps[s]["mask"][:, 0] << 2)).to(torch.uint8).contiguous()
        msum = float(maps[s]["mask"].sum()) if not multi else None
        if multi:
            m = outputs["consistency_mask"].unsqueeze(1) * (1 - outputs["augmentation_mask"][:B]); msum = float(m.sum())
        g_r = (1.0 / S) / (msum + 1e-7); g_c = (1.0 / S) / (B * H * W) if multi else 0.0
        gd, gP = E.backward_scale(fw, opt, g_r, g_c)
        smooth_part = (g64[("disp", s)] - g64ns[("disp", s)])
        d = grads[("disp", s)].double() - smooth_part - gd.double()
        e = d.abs(); idx = torch.nonzero(e == e.max())[0].tolist()
        print("   s=%d GPU-vs-emul max %.3e (rel %.3e) at %s ; emul-vs-64 rel max %.3e" % (
            s, float(e.max()), float(e.max() / gd.abs().max()), idx,
            float((gd.double() - g64ns[("disp", s)]).abs().max() / g64ns[("disp", s)].abs().max())))

report(2, 64, 96, 4, 1)
report(2, 40, 72, 1, 13)
report(2, 64, 96, 4, 1, multi=True)
