"""-m gpu, round 2: what the benchmark actually times and what a trainer actually calls.

  * the CUDA-graph plan (runner.FusedPlan capture / replay -- the path bench.py's `value` comes from) against the autograd
    path and the oracle after EVERY replay, over alternating input sets;
  * BASELINE.json configs[2] (Cityscapes 24x192x512) and configs[3] (hi-res 8x320x1024, deterministic backward) at full
    size against the oracle;
  * the reference's own `process_batch` sequence (trainer.py:436-461: mono -> key copy -> compute_matching_mask -> multi
    -> in-place `+=` of the loss dictionaries -> ONE backward) on the installed methods;
  * the per-source byproducts ("sample" / "color", trainer.py:909-918) against the reference's own tensors;
  * frames that are not k/255 (planar fp32 gathers instead of the packed RGBA8 ones);
  * the cached step plans of the host layer (ring reuse, forward without backward).
"""
from types import SimpleNamespace

import pytest
import torch

from conftest import load_golden
from gpu_helpers import FeedNoise, check_against_oracle, run_cuda
from oracle import vsl_oracle as O
from ppea_depth_b200 import functional as Fn
from ppea_depth_b200.functional import VslConfig
from ppea_depth_b200.loss import ViewSynthesisLoss, install
from ppea_depth_b200.runner import FusedPlan
from ppea_depth_b200.synth import CITYSCAPES_K, SynthConfig, make_batch, make_noise

pytestmark = pytest.mark.gpu


def _plan_inputs(inputs, outputs, noise, S, multi, dev="cuda"):
    d = lambda x: x.to(dev)
    kw = dict(noise=[d(z) for z in noise]) if not multi else dict(
        cons_mask=d(outputs["consistency_mask"]), aug_mask=d(outputs["augmentation_mask"]),
        mono_depth=[d(outputs[("mono_depth", 0, s)]) for s in range(S)])
    args = ([d(outputs[("disp", s)]) for s in range(S)], [d(outputs[("cam_T_cam", 0, -1)]), d(outputs[("cam_T_cam", 0, 1)])],
            d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])],
            d(inputs[("K", 0)]), d(inputs[("inv_K", 0)]), [d(inputs[("color", 0, s)]) for s in range(S)])
    return args, kw


@pytest.mark.parametrize("multi,det,fused", [(False, False, None), (True, False, None), (False, True, None), (False, False, "tiles")])
def test_graph_replay_plan_matches_autograd_path_and_oracle(multi, det, fused):
    """bench.py times FusedPlan.capture()/replay() with PPEA_F_RAW_PREZEROED (the gradient finish re-zeroes the coarse
    raw fields between replays).  Two input sets share ONE plan (inputs refreshed with copy_, as a training loop does);
    after each of 4 replays the losses and every gradient must equal the autograd path's (same kernels, fresh buffers)
    and satisfy the oracle's tolerances -- a field that was not re-zeroed would show up at the second replay."""
    B, H, W, S = 2, 64, 96, 4
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    sets = []
    for seed in (61, 62):
        cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=seed)
        inputs, outputs = make_batch(cfg)
        sets.append((inputs, outputs, make_noise(cfg, S)))
    args, kw = _plan_inputs(*sets[0], S, multi)
    plan = FusedPlan(VslConfig(is_multi=multi, deterministic=det, want_loss_px=True), *args, fused=fused, **kw)
    plan.capture()
    refs = [run_cuda(i, o, opt, multi, None if multi else n, deterministic=det, fused=fused) for (i, o, n) in sets]
    for rep in range(4):
        inputs, outputs, noise = sets[rep % 2]
        a2, k2 = _plan_inputs(inputs, outputs, noise, S, multi)
        for dst, src in zip(plan.disps, a2[0]):
            dst.copy_(src)
        for dst, src in zip(plan.T, a2[1]):
            dst.copy_(src)
        plan.tgt.copy_(a2[2])
        for dst, src in zip(plan.src, a2[3]):
            dst.copy_(src)
        for dst, src in zip(plan.colors, a2[6]):
            dst.copy_(src)
        if multi:
            plan.cons_mask.copy_(k2["cons_mask"]); plan.aug_mask.copy_(k2["aug_mask"].reshape(-1)[:B])
            for dst, src in zip(plan.mono_depth, k2["mono_depth"]):
                dst.copy_(src)
        else:
            for dst, src in zip(plan.noise, k2["noise"]):
                dst.copy_(src)
        plan.replay()
        torch.cuda.synchronize()
        l_ref, g_ref, m_ref = refs[rep % 2]
        got = plan.losses.cpu()
        assert abs(float(got[0]) - float(l_ref["loss"])) <= 1e-6 * abs(float(l_ref["loss"])), (rep, float(got[0]), float(l_ref["loss"]))
        losses = {"loss": got[0]}
        for s in range(S):
            losses["loss/%d" % s], losses["reproj_loss/%d" % s] = got[1 + 4 * s], got[2 + 4 * s]
            if multi:
                losses["consistency_loss/%d" % s] = got[3 + 4 * s]
        grads = {("disp", s): plan.grad_disp[s].cpu() for s in range(S)}
        if not multi:
            grads[("cam_T_cam", 0, -1)], grads[("cam_T_cam", 0, 1)] = plan.grad_T[0].cpu(), plan.grad_T[1].cpu()
        for k, g in grads.items():
            scale = float(g_ref[k].abs().max())
            tol = 0.0 if det else 2e-6 * scale + 1e-12          # atomics: summation order; deterministic: bit for bit
            assert float((g - g_ref[k]).abs().max()) <= tol, (rep, k)
        maps = {s: dict(depth=plan.depth[s].cpu(), r=plan.loss_px[s].cpu(), mask=((plan.sel[s].cpu() >> 2) & 1).unsqueeze(1),
                        src_idx=(plan.sel[s].cpu() & 3).unsqueeze(1)) for s in range(S)}
        for s in range(S):
            assert torch.equal(maps[s]["mask"], m_ref[s]["mask"]) and torch.equal(maps[s]["src_idx"], m_ref[s]["src_idx"]), (rep, s)
        check_against_oracle(inputs, outputs, opt, multi, noise, losses, grads, maps)


FULL = {
    "cityscapes": dict(batch=24, height=192, width=512, num_scales=4, intrinsics=CITYSCAPES_K),
    "hires": dict(batch=8, height=320, width=1024, num_scales=4),
}


@pytest.mark.parametrize("name,det,multi", [("cityscapes", False, False), ("cityscapes", False, True), ("hires", True, False)])
def test_full_size_configs_against_oracle(name, det, multi):
    """BASELINE.json configs[2] / configs[3] at full size: per-pixel maps, selection, reduced losses and gradients against
    the oracle (a few seconds of CPU each), plus the UNFORCED loss to 1e-5."""
    kw = FULL[name]
    cfg = SynthConfig(seed=71, **kw)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, 4)
    opt = O.default_opt(sclm=3, height=kw["height"], width=kw["width"], batch_size=kw["batch"])
    losses, grads, maps = run_cuda(inputs, outputs, opt, multi, None if multi else noise, deterministic=det)
    report = {}
    n_flip = check_against_oracle(inputs, outputs, opt, multi, noise, losses, grads, maps, report=report)
    report.pop("_m64")
    # (pose gradients are sums over every pixel, knife-edge samples included: "kink" = at least as close to the float64
    # gradient as the reference's own fp32 is, factor 2)
    ok = lambda k, v: v.startswith("fp32") or (k[0] == "cam_T_cam" and v == "kink")
    assert all(ok(k, v) for k, v in report.items()), {k: v for k, v in report.items() if not ok(k, v)}
    l32, _, _ = O.run_fwd_bwd(inputs, outputs, opt, multi, noise)
    assert abs(float(losses["loss"]) - float(l32["loss"])) <= 1e-5 * float(l32["loss"]), (n_flip, float(losses["loss"]), float(l32["loss"]))


def _opt_ns(B, H, W, S):
    return SimpleNamespace(sclm=S - 1, v1_multiscale=False, height=H, width=W, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1],
                           disable_automasking=False, no_ssim=False, selec_reproj=True, disable_motion_masking=False,
                           no_matching_augmentation=False, batch_size=B, disparity_smoothness=1e-3)


def test_process_batch_sequence_on_installed_methods():
    """trainer.py:436-461 verbatim on a Trainer stand-in whose loss methods were rebound by install(): mono pass, copy of the
    depth / disp keys to mono_*, consistency_mask *= compute_matching_mask, multi pass, `losses[key] += val` IN PLACE on the
    entries of the multi dictionary, one backward.  Against the oracle's mono + multi sums and gradients."""
    B, H, W, S = 2, 64, 96, 4
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=83)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)

    class Trainer:        # (stands for ppeadepth.trainer.Trainer: only the attributes process_batch touches)
        freeze_tp = False

        def __init__(self):
            self.opt = _opt_ns(B, H, W, S)

    install(Trainer)
    tr = Trainer()
    ins, outs = O.clone_batch(inputs, outputs, device="cuda")
    mono_outputs = {k: v for k, v in outs.items() if not (isinstance(k, tuple) and k[0] == "mono_depth") and k not in ("consistency_mask",)}
    outs["lowest_cost"] = (1.0 / outs[("mono_depth", 0, 0)][:, 0].detach()) * (0.8 + 0.4 * torch.rand(B, H, W, device="cuda"))
    cons0 = outs["consistency_mask"].clone()
    with FeedNoise(noise):
        tr.generate_images_pred(ins, mono_outputs)
        mono_losses, _ = tr.compute_losses(ins, mono_outputs, is_multi=False)
    for key in list(mono_outputs.keys()):
        _key = list(key) if isinstance(key, tuple) else [key]
        if _key[0] in ["depth", "disp"]:
            _key[0] = "mono_" + key[0]
            outs[tuple(_key)] = mono_outputs[key]
    outs["consistency_mask"] = (outs["consistency_mask"] * tr.compute_matching_mask(outs))
    tr.generate_images_pred(ins, outs, is_multi=True)
    losses, _ = tr.compute_losses(ins, outs, is_multi=True)
    for key, val in mono_losses.items():
        losses[key] += val
    losses["loss"].backward()
    torch.cuda.synchronize()

    # the oracle, same sequence on the CPU
    o_in, o_out = O.clone_batch(inputs, outputs, device="cpu")
    l_mono, g_mono, m_mono = O.run_fwd_bwd(inputs, outputs, tr.opt, False, noise, want_maps=True)
    mono_depth = {s: m_mono[s]["depth"] for s in range(S)}
    out2 = dict(outputs)
    for s in range(S):
        out2[("mono_depth", 0, s)] = mono_depth[s]
    lc = outs["lowest_cost"].cpu()
    md = mono_depth[0][:, 0]
    mm = (((1.0 / lc - md) / md) < 1.0) & (((md - 1.0 / lc) / (1.0 / lc)) < 1.0)       # trainer.py:859-869
    out2["consistency_mask"] = cons0.cpu() * mm
    l_multi, g_multi, _ = O.run_fwd_bwd(inputs, out2, tr.opt, True, noise)
    for k in ("loss",) + tuple("loss/%d" % s for s in range(S)) + tuple("reproj_loss/%d" % s for s in range(S)):
        want = float(l_multi[k]) + float(l_mono[k])
        assert abs(float(losses[k]) - want) <= 2e-5 * abs(want), (k, float(losses[k]), want)
    for s in range(S):
        got = outs[("disp", s)].grad.cpu()
        want = g_multi[("disp", s)] + g_mono[("disp", s)]
        assert float((got - want).abs().max()) <= 2e-4 * float(want.abs().max()), s
    for f in (-1, 1):      # T is detached on the multi path: the pose gradient is the mono pass's alone
        got = outs[("cam_T_cam", 0, f)].grad.cpu()
        assert float((got - g_mono[("cam_T_cam", 0, f)]).abs().max()) <= 2e-4 * float(g_mono[("cam_T_cam", 0, f)].abs().max()), f


@pytest.mark.parametrize("name", ["mono_s4_48x96", "multi_s2_nomotion_noaug"])
def test_materialized_warps_match_the_reference(name):
    """("sample", f, s) and ("color", f, s) (trainer.py:909-918) through install(..., materialize_warps=True) against the
    tensors the reference itself produced (stored by oracle/make_golden.py for the coarsest scale), uint8 frames included."""
    fx = load_golden(name)
    opt, multi = fx["opt"], fx["is_multi"]
    S = opt.sclm + 1
    for as_u8 in (False, True):
        ins, outs = O.clone_batch(fx["inputs"], fx["outputs"], device="cuda")
        if as_u8:
            ins = {k: ((v * 255).round().to(torch.uint8) if k[0] == "color" else v) for k, v in ins.items()}
        mod = ViewSynthesisLoss(opt, materialize_warps=True)
        with FeedNoise(fx["noise"] if not multi else []):
            mod.generate_images_pred(ins, outs, multi)
        s = S - 1
        for i, f in enumerate(opt.frame_ids[1:]):
            grid, warped = fx["ref_maps"][s]["sample"][i], fx["ref_maps"][s]["warped"][i]
            assert float((outs[("sample", f, s)].cpu() - grid).abs().max()) <= 2e-5          # normalised coordinates
            assert float((outs[("color", f, s)].detach().cpu() - warped).abs().max()) <= 2e-4   # ~1e-2 px at image gradients <= 1/px
            ci = outs[("color_identity", f, s)]
            assert ci.dtype == torch.float32 and float((ci.cpu() - fx["inputs"][("color", f, 0)]).abs().max()) == 0.0


@pytest.mark.parametrize("multi", [False, True])
def test_frames_that_are_not_bytes_take_the_planar_gathers(multi):
    """The streaming kernel packs the source frames to RGBA8 when every value is exactly k/255 (ToTensor frames) and
    otherwise gathers from the planar fp32 frames: same parity bar on un-quantised frames, and the two gathers agree on
    quantised ones to fp32 rounding (the tile kernel, which always reads fp32, is the witness)."""
    B, H, W, S = 2, 64, 96, 4
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=91)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    rough = dict(inputs)
    g = torch.Generator().manual_seed(5)
    for k in inputs:
        if k[0] == "color":
            rough[k] = (inputs[k] + 0.003 * torch.rand(inputs[k].shape, generator=g)).clamp(0, 1)     # no longer k/255
    losses, grads, maps = run_cuda(rough, outputs, opt, multi, None if multi else noise)
    check_against_oracle(rough, outputs, opt, multi, noise, losses, grads, maps)
    l_t, g_t, m_t = run_cuda(rough, outputs, opt, multi, None if multi else noise, fused="tiles")
    for s in range(S):       # same fp32 frames, same arithmetic, same summation order of the windows: the maps agree bit for bit
        assert torch.equal(maps[s]["depth"], m_t[s]["depth"])
        same = (maps[s]["src_idx"] == m_t[s]["src_idx"]) & (maps[s]["mask"] == m_t[s]["mask"])
        assert int((~same).sum()) <= max(2, same.numel() // 5000)
        assert float((maps[s]["r"] - m_t[s]["r"])[same].abs().max()) <= 1e-6
    # quantised frames: packed gathers vs the tile kernel's fp32 gathers
    l_p, g_p, m_p = run_cuda(inputs, outputs, opt, multi, None if multi else noise)
    l_q, g_q, m_q = run_cuda(inputs, outputs, opt, multi, None if multi else noise, fused="tiles")
    for k in l_q:
        assert abs(float(l_p[k]) - float(l_q[k])) <= 2e-6 * abs(float(l_q[k])) + 1e-12, k


def test_step_plan_cache_ring_and_unfinished_steps():
    """Cached plans (functional._StepPlan): outputs live in a ring of two buffer sets per (shape, flags); a forward whose
    backward never ran must not leak its un-normalised gradient fields into the next use of the slot."""
    B, H, W, S = 1, 32, 64, 3
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=97)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    opt = O.default_opt(sclm=S - 1, height=H, width=W, batch_size=B)
    Fn._PLANS.clear()
    ref_l, ref_g, _ = run_cuda(inputs, outputs, opt, False, noise)          # builds the plan, uses slot 0
    assert len(Fn._PLANS) == 1
    mod = ViewSynthesisLoss(opt, keep_maps=True, deterministic=False)          # (same plan key as run_cuda's module)
    depth_ptrs = []
    for it in range(4):              # forwards WITHOUT backward: every slot is left dirty
        ins, outs = O.clone_batch(inputs, outputs, device="cuda")
        with FeedNoise(noise):
            mod.generate_images_pred(ins, outs, False)
            mod.compute_losses(ins, outs, False)
        depth_ptrs.append(outs[("depth", 0, 0)].data_ptr())
    assert len(set(depth_ptrs)) == 2 and depth_ptrs[0] == depth_ptrs[2] and depth_ptrs[0] != depth_ptrs[1]
    l2, g2, _ = run_cuda(inputs, outputs, opt, False, noise)                 # a complete step on a previously dirty slot
    assert len(Fn._PLANS) == 1
    assert float(l2["loss"]) == float(ref_l["loss"])
    for k in ref_g:
        assert float((g2[k] - ref_g[k]).abs().max()) <= 2e-6 * float(ref_g[k].abs().max()) + 1e-12, k
    # plan_cache=False: fresh buffers every call, same numbers
    ins, outs = O.clone_batch(inputs, outputs, device="cuda")
    mod = ViewSynthesisLoss(opt, plan_cache=False, deterministic=False)
    with FeedNoise(noise):
        mod.generate_images_pred(ins, outs, False)
        losses, _ = mod.compute_losses(ins, outs, False)
    losses["loss"].backward()
    assert float(losses["loss"]) == float(ref_l["loss"]) and len(Fn._PLANS) == 1


def test_shapes_are_checked_before_any_launch():
    """K / inv_K / T are indexed as ptr + 16 b by the kernels: a batch-1 matrix is expanded (the reference modules rely on
    matmul broadcasting), anything else raises on the host."""
    B, H, W, S = 2, 32, 64, 1
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=3)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, S)
    opt = O.default_opt(sclm=0, height=H, width=W, batch_size=B)
    ref_l, _, _ = run_cuda(inputs, outputs, opt, False, noise)
    one = dict(inputs)
    one[("K", 0)], one[("inv_K", 0)] = inputs[("K", 0)][:1], inputs[("inv_K", 0)][:1]        # (all items share the intrinsics)
    l1, _, _ = run_cuda(one, outputs, opt, False, noise)
    assert float(l1["loss"]) == float(ref_l["loss"])
    bad = dict(inputs)
    bad[("K", 0)] = torch.cat([inputs[("K", 0)], inputs[("K", 0)][:1]])                       # batch 3 against a batch of 2
    with pytest.raises(ValueError):
        run_cuda(bad, outputs, opt, False, noise)
    bad = dict(inputs)
    bad[("color", -1, 0)] = inputs[("color", -1, 0)][..., :-2]
    with pytest.raises(ValueError):
        run_cuda(bad, outputs, opt, False, noise)
    with pytest.raises(ValueError):
        run_cuda(inputs, outputs, opt, False, [z[..., :-1] for z in noise])


@pytest.mark.parametrize("multi", [False, True])
def test_loss_pct_branch_reports_the_mask_fraction(multi):
    """`--loss_pct` (trainer.py:1116-1123): percent = reprojection_loss_mask.sum() / (batch_size * height * width) per scale,
    here read from the sums the forward already reduced; checked against the mask the kernel itself reports (sel bit 2 on the
    mono path, 1 - consistency target mask on the multi path via the oracle)."""
    B, H, W, S = 2, 40, 72, 3
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=S, seed=5)
    inputs, outputs = make_batch(cfg)
    opt = _opt_ns(B, H, W, S)
    opt.loss_pct, opt.debug = True, False
    mod = ViewSynthesisLoss(opt, noise_mode="reference", keep_maps=True)
    ins = {k: v.cuda() for k, v in inputs.items()}
    outs = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in outputs.items()}
    for s in range(S):
        outs[("disp", s)].requires_grad_(True)
    torch.manual_seed(3)
    mod.generate_images_pred(ins, outs, multi)
    losses, _ = mod.compute_losses(ins, outs, multi)
    noise = None
    if not multi:
        torch.manual_seed(3)
        noise = [torch.randn(B, 1, H, W) for _ in range(S)]
    _, ref_maps = O.view_synthesis_losses(inputs, outputs, opt, is_multi=multi, noise=noise, want_maps=True)
    for s in range(S):
        pct = float(outs[("loss_pct", "m" if multi else "t", s)])
        want = float(ref_maps[s]["mask"].sum()) / (B * H * W)
        assert abs(pct - want) <= 2.0 / (B * H * W) + 1e-6, (s, pct, want)      # (near-tie automask flips: at most a pixel or two)
