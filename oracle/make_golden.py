"""ORACLE tooling (test infrastructure): generates tests/golden/*.pt by running the
REFERENCE's own code (oracle/ref_import.py -> /root/reference/ppeadepth) on small
synthetic batches.  Run in the build container (the GPU box has no /root/reference):

    python -m oracle.make_golden

Each fixture holds the inputs (so it does not depend on the generator), the
reference's losses, gradients, its own per-scale depth / sampling grid / warped
images (outputs[("depth",0,s)], ("sample",f,s), ("color",f,s) after
generate_images_pred), and the per-pixel maps of the restatement
(oracle/vsl_oracle.py), which tests/test_oracle.py pins to the reference first.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import as R            # noqa: E402
from oracle import vsl_oracle as O            # noqa: E402
from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (synth kwargs, opt overrides, is_multi)
CASES = {
    "mono_s4_48x96": (dict(batch=1, height=48, width=96, num_scales=4, seed=11), dict(), False),
    "multi_s4_48x96": (dict(batch=1, height=48, width=96, num_scales=4, seed=12), dict(), True),
    "mono_s1_40x72_ragged": (dict(batch=2, height=40, width=72, num_scales=1, seed=13), dict(), False),
    "mono_s2_noselec": (dict(batch=2, height=32, width=64, num_scales=2, seed=14), dict(selec_reproj=False), False),
    "mono_s2_nossim": (dict(batch=2, height=32, width=64, num_scales=2, seed=15), dict(no_ssim=True), False),
    "mono_s2_noautomask": (dict(batch=2, height=32, width=64, num_scales=2, seed=16), dict(disable_automasking=True), False),
    "multi_s2_nomotion_noaug": (dict(batch=2, height=32, width=64, num_scales=2, seed=17),
                                dict(disable_motion_masking=True, no_matching_augmentation=True), True),
    "mono_s3_v1_multiscale": (dict(batch=1, height=48, width=96, num_scales=3, seed=18, v1_multiscale=True),
                              dict(v1_multiscale=True), False),
    "mono_s1_identity_pose": (dict(batch=2, height=32, width=64, num_scales=1, seed=19, identity_pose=True), dict(), False),
}


def make_case(name):
    skw, okw, is_multi = CASES[name]
    cfg = SynthConfig(**skw)
    inputs, outputs = make_batch(cfg)
    S = cfg.num_scales
    opt = O.default_opt(sclm=S - 1, height=cfg.height, width=cfg.width, batch_size=cfg.batch, **okw)
    noise = make_noise(cfg, S)
    losses, grads, (tr, ins, outs) = R.run_reference(inputs, outputs, opt, is_multi, None if is_multi else noise)
    _, _, maps = O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise, want_maps=True)
    ref_maps = {}
    for s in range(S):
        ref_maps[s] = dict(depth=outs[("depth", 0, s)].detach())
        if s == S - 1:     # the reference's grids / warped images of the coarsest scale only (fixture size)
            ref_maps[s]["sample"] = [outs[("sample", f, s)].detach() for f in opt.frame_ids[1:]]
            ref_maps[s]["warped"] = [outs[("color", f, s)].detach() for f in opt.frame_ids[1:]]
    keep_out = {k: v for k, v in outputs.items()
                if not (isinstance(k, tuple) and k[0] in ("axisangle", "translation"))}
    return dict(name=name, synth=skw, opt=vars(opt), is_multi=is_multi, inputs=inputs, outputs=keep_out, noise=noise,
                ref_losses=losses, ref_grads=grads, ref_maps=ref_maps,
                oracle_maps={s: dict(r=maps[s]["r"], ident=maps[s]["ident"], mask=maps[s]["mask"].to(torch.uint8),
                                     src_idx=maps[s]["src_idx"].to(torch.uint8)) for s in maps})


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name in CASES:
        fx = make_case(name)
        path = os.path.join(GOLDEN_DIR, name + ".pt")
        # store images as uint8 (they are exact k/255) and maps in fp32 to keep the fixtures small
        packed = dict(fx)
        packed["inputs"] = {k: ((v * 255.0).round().to(torch.uint8) if k[0] == "color" else v) for k, v in fx["inputs"].items()}
        torch.save(packed, path)
        print(name, "%.1f KB" % (os.path.getsize(path) / 1024), {k: float(v) for k, v in fx["ref_losses"].items() if k == "loss"})


if __name__ == "__main__":
    main()
