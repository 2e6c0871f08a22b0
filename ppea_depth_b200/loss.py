"""Drop-in replacements of the four loss methods of the reference Trainer.

Same names, signatures, dict-key conventions and return values as
/root/reference/ppeadepth/trainer.py:
  generate_images_pred(self, inputs, outputs, is_multi=False)      :871-918
  compute_reprojection_loss(self, pred, target)                    :995-1007
  compute_loss_masks(reprojection_loss, identity_reprojection_loss):1009-1027 (static)
  compute_losses(self, inputs, outputs, is_multi=False)            :1032-1160

``install(Trainer)`` rebinds them on the reference's Trainer class, so
`trainer.py`, `process_batch` and the launch command are used unchanged:

    import ppeadepth.trainer as T, ppea_depth_b200
    ppea_depth_b200.install(T.Trainer)

`generate_images_pred` runs the fused forward (it must: the reference reads
outputs[("depth", 0, s)] right after it, trainer.py:443-451, 466) and parks the
result in ``outputs``; `compute_losses` turns it into the reference's loss
dict.  The warped images / sampling grids (("color", f, s), ("sample", f, s))
that the reference also stores are never read again by the training step
(SURVEY.md §3.2); they are produced only with ``materialize_warps=True``.

The tie-break noise of the automask (trainer.py:1084-1087) is drawn exactly as
the reference draws it -- ``torch.randn`` on the CPU default generator, one
(B,1,H,W) draw per scale, scale-ascending -- so seeding reproduces the
reference's masks; ``noise_mode="device"`` draws on the GPU instead (no H2D
copy, different stream of numbers).
"""
from __future__ import annotations

import torch

from . import functional as Fn
from .functional import VslConfig
from .images import FrameCache

_RESULT_KEY = "_ppea_vsl_result"


def _opt(self, name, default):
    return getattr(self.opt, name, default)


def _config(self, is_multi, **kw):
    o = self.opt
    return VslConfig(
        is_multi=bool(is_multi),
        automask=True,     # the reference never passes identity=None to compute_loss_masks (trainer.py:1088-1091)
        selec_reproj=bool(_opt(self, "selec_reproj", True)),
        no_ssim=bool(_opt(self, "no_ssim", False)),
        motion_mask=not _opt(self, "disable_motion_masking", False),
        match_aug=not _opt(self, "no_matching_augmentation", False),
        deterministic=bool(getattr(self, "ppea_deterministic", False)),
        min_depth=float(o.min_depth), max_depth=float(o.max_depth),
        disparity_smoothness=float(_opt(self, "disparity_smoothness", 1e-3)),
        want_loss_px=bool(getattr(self, "ppea_keep_maps", False)),
        fused=getattr(self, "ppea_fused", None),
        **kw)


def _draw_noise(self, n, shape, device):
    """The tie-break draws of `n` scales.  "reference": torch.randn on the CPU default generator, one (B,1,H,W) draw per
    scale, scale-ascending, exactly as trainer.py:1086-1087 (then copied to the device); "device": ONE draw of all
    scales on the GPU (no host RNG, no H2D copy; a different stream of numbers)."""
    mode = getattr(self, "ppea_noise_mode", "reference")
    if mode == "device":
        return list(torch.randn((n,) + tuple(shape), device=device).unbind(0))
    return [torch.randn(shape).to(device, non_blocking=True) for _ in range(n)]


def _run_fused(self, inputs, outputs, is_multi):
    o = self.opt
    S = o.sclm + 1
    if len(o.frame_ids) != 3:
        raise NotImplementedError("the fused path implements the reference's two-source configuration "
                                  "(frame_ids [0,-1,1]; selec_reproj hard-codes it, trainer.py:1078-1083)")
    f0, f1 = o.frame_ids[1], o.frame_ids[2]
    T = (outputs[("cam_T_cam", 0, f0)], outputs[("cam_T_cam", 0, f1)])
    dev = outputs[("disp", 0)].device
    groups = []
    if _opt(self, "v1_multiscale", False):
        # every scale has its own image resolution (trainer.py:883-884): one call per scale
        for s in range(S):
            groups.append((s, 1, s))
    else:
        groups.append((0, S, 0))                     # all scales share source_scale 0 (trainer.py:886-888)
    results = []
    frame = FrameCache()          # uint8 frames are expanded on the device, once each (images.py)
    for first, n, ss in groups:
        scales = range(first, first + n)
        disps = [outputs[("disp", s)] for s in scales]
        colors = [frame(inputs[("color", 0, s)]) for s in scales]
        tgt = frame(inputs[("color", 0, ss)])
        kw = {}
        cfg = _config(self, is_multi, first_scale=first, total_scales=S, plan_cache=bool(getattr(self, "ppea_plan_cache", True)))
        if is_multi:
            kw["mono_depth"] = [outputs[("mono_depth", 0, s)] for s in scales]
            kw["cons_mask"] = outputs.get("consistency_mask")
            kw["aug_mask"] = outputs.get("augmentation_mask")
            if kw["aug_mask"] is not None:
                kw["aug_mask"] = kw["aug_mask"][:o.batch_size]
            if not _opt(self, "disable_automasking", False) and getattr(self, "ppea_noise_mode", "reference") == "reference":
                # the reference draws the tie-break noise on the multi path too, before discarding the automask
                # (trainer.py:1084-1087 run ahead of the is_multi override :1101): keep the CPU generator in step with it
                for _ in scales:
                    torch.randn((tgt.shape[0], 1, tgt.shape[2], tgt.shape[3]))
        elif not _opt(self, "disable_automasking", False):    # the flag only removes the noise (trainer.py:1084-1087)
            kw["noise"] = _draw_noise(self, len(scales), (tgt.shape[0], 1, tgt.shape[2], tgt.shape[3]), dev)
        res = Fn.view_synthesis_loss(disps, T, tgt, (frame(inputs[("color", f0, ss)]), frame(inputs[("color", f1, ss)])),
                                     inputs[("K", ss)], inputs[("inv_K", ss)], colors, cfg, **kw)
        results.append((first, n, res))
    return results


def generate_images_pred(self, inputs, outputs, is_multi=False):
    """Fused stand-in for trainer.py:871-918: writes outputs[("depth", 0, s)] for every
    scale (and, on request, the warped images) and parks the fused result for compute_losses."""
    results = _run_fused(self, inputs, outputs, is_multi)
    for first, n, res in results:
        for i in range(n):
            outputs[("depth", 0, first + i)] = res.depth[i]
    outputs[(_RESULT_KEY, bool(is_multi))] = results
    if getattr(self, "ppea_materialize_warps", False):
        _materialize_warps(self, inputs, outputs, is_multi)


def _materialize_warps(self, inputs, outputs, is_multi):
    """The reference's per-source byproducts (trainer.py:904-918) -- ("sample", f, s), ("color", f, s) and
    ("color_identity", f, s) -- through the piecewise operators.  Nothing on the training step reads them (SURVEY.md
    §3.2); they exist for callers that log or inspect the warps (``materialize_warps=True``)."""
    o = self.opt
    frame = FrameCache()
    for s in range(o.sclm + 1):
        ss = s if _opt(self, "v1_multiscale", False) else 0
        depth = outputs[("depth", 0, s)]
        H, W = depth.shape[-2], depth.shape[-1]
        for f in o.frame_ids[1:]:
            T = outputs[("cam_T_cam", 0, f)]
            if is_multi:
                T = T.detach()
            cam = Fn.backproject_depth(depth, inputs[("inv_K", ss)], H, W)
            pix, _ = Fn.project_3d(cam, inputs[("K", ss)], T, H, W)
            src = frame(inputs[("color", f, ss)])
            outputs[("sample", f, s)] = pix
            outputs[("color", f, s)] = Fn.grid_sample_border(src, pix)
            if not _opt(self, "disable_automasking", False):
                outputs[("color_identity", f, s)] = src
        if is_multi:
            # logging byproduct of the multi path (trainer.py:1134-1138; its only reader in the reference is commented out):
            # 1 / (mono_depth * consistency_mask + multi_depth * (1 - consistency_mask)), a few elementwise launches
            keep = torch.ones_like(depth)
            if not _opt(self, "disable_motion_masking", False) and outputs.get("consistency_mask") is not None:
                keep = keep * outputs["consistency_mask"].unsqueeze(1)
            if not _opt(self, "no_matching_augmentation", False) and outputs.get("augmentation_mask") is not None:
                keep = keep * (1 - outputs["augmentation_mask"][:o.batch_size])
            cm = 1 - keep
            outputs["consistency_target/{}".format(s)] = 1 / (outputs[("mono_depth", 0, s)].detach() * cm + depth.detach() * (1 - cm))


def compute_reprojection_loss(self, pred, target):
    """0.85 * mean_c SSIM + 0.15 * mean_c |target - pred|  -> (B,1,H,W); trainer.py:995-1007."""
    return Fn.reprojection_loss(pred, target, bool(_opt(self, "no_ssim", False)))


def compute_loss_masks(reprojection_loss, identity_reprojection_loss):
    """argmin([reproj, identity]) == 0 as float; trainer.py:1009-1027.  A (B,1,H,W) comparison --
    inside the fused path it is a predicate in the forward kernel's epilogue."""
    if identity_reprojection_loss is None:
        return torch.ones_like(reprojection_loss)
    return (reprojection_loss <= identity_reprojection_loss).float()


def compute_losses(self, inputs, outputs, is_multi=False):
    """Loss dict of trainer.py:1032-1160: "loss", "loss/{s}", "reproj_loss/{s}" and (multi)
    "consistency_loss/{s}" as 0-dim tensors attached to the fused autograd node."""
    results = outputs.pop((_RESULT_KEY, bool(is_multi)), None)
    if results is None:
        results = _run_fused(self, inputs, outputs, is_multi)
        for first, n, res in results:
            for i in range(n):
                outputs[("depth", 0, first + i)] = res.depth[i]
    losses = {}
    total = None
    for first, n, res in results:
        # one `select` view per dictionary entry (not unbind: the reference updates the entries in place,
        # `losses[key] += val`, trainer.py:459-461, which autograd forbids on the views of a multi-output view op)
        v = res.losses
        for i in range(n):
            s = first + i
            k = 1 + 4 * i               # [loss/s, reproj_loss/s, consistency_loss/s, smooth/s]  (ppea_vsl.h PPEA_LOSSES_PER_SCALE)
            losses["reproj_loss/{}".format(s)] = v[k + 1]
            if is_multi:
                losses["consistency_loss/{}".format(s)] = v[k + 2]
            losses["loss/{}".format(s)] = v[k]
        total = v[0] if total is None else total + v[0]     # each call already divides by sclm+1
        if _opt(self, "loss_pct", False):
            _log_mask_fraction(self, outputs, res, first, n, is_multi)
    losses["loss"] = total
    if getattr(self, "ppea_keep_maps", False):
        outputs[("ppea_maps", bool(is_multi))] = results
    return losses, []


def _log_mask_fraction(self, outputs, res, first, n, is_multi):
    """The `--loss_pct` branch of compute_losses (trainer.py:1116-1123): the fraction of pixels the reprojection mask keeps,
    per scale, from the sums the forward already reduced (row[1] = sum(mask), ppea_vsl.h) -- no extra pass over the mask.
    Kept on the device in outputs[("loss_pct", mode, scale)]; printed with opt.debug and sent to wandb every 50th step by the
    main process, exactly where the reference does."""
    import sys
    o = self.opt
    stride = res.sums.numel() // n
    for i in range(n):
        scale = first + i
        percent = res.sums[i * stride + 1] / (o.batch_size * o.height * o.width)
        mode = "m" if is_multi else "t"
        outputs[("loss_pct", mode, scale)] = percent
        if _opt(self, "debug", False):
            print(percent)
        if getattr(self, "step", 1) % 50 == 0 and getattr(self, "is_main", False) and "wandb" in sys.modules:
            sys.modules["wandb"].log({"Train/pp_{}_{}".format(mode, scale): percent}, step=self.step)


def compute_matching_mask(self, outputs):
    """Trainer.compute_matching_mask (trainer.py:859-869) in one launch: where the cost volume's best depth and the teacher
    network disagree by less than a factor of two either way.  Returns the (B,H,W) bool mask."""
    import ctypes
    from . import _cabi as C
    mono = outputs[("mono_depth", 0, 0)]
    if not mono.is_cuda:
        raise RuntimeError("ppea_depth_b200 has no CPU path: mono_depth must be a CUDA tensor")
    mono = mono.detach().contiguous().float()
    lowest = outputs["lowest_cost"]
    if not torch.is_tensor(lowest):
        lowest = torch.as_tensor(lowest)
    lowest = lowest.to(mono.device).detach().contiguous().float()
    B, _, H, W = mono.shape
    if lowest.numel() != B * H * W:
        raise ValueError("compute_matching_mask: lowest_cost must be (B,H,W) at the resolution of mono_depth")
    with torch.cuda.device(mono.device):
        mask = torch.empty(B, H, W, device=mono.device, dtype=torch.bool)
        C.check(C.lib().ppea_matching_mask(mono.data_ptr(), lowest.data_ptr(), mask.data_ptr(), ctypes.c_size_t(mask.numel()),
                                           torch.cuda.current_stream().cuda_stream))
    return mask


def install(trainer_cls, deterministic=True, noise_mode="reference", fused=None, plan_cache=True, materialize_warps=False):
    """Rebinds the reference Trainer's loss methods to the fused implementation.

    deterministic      bit-reproducible gradients (64-bit fixed-point accumulation of the coarse-scale fields; the default: on the
                       fused step it is also the faster accumulation, 0.374 against 0.381 ms at the KITTI shape); False: float
                       atomics, whose summation order varies from run to run like the reference's own CUDA grid_sample backward
    noise_mode         "reference": the automask's tie-break noise from the CPU generator exactly as the reference draws it;
                       "device": one torch.randn on the GPU per call
    fused              None: the fused training step whenever gradients are needed; "tiles" / False: see VslConfig.fused
    plan_cache         reuse pre-built parameter blocks and a ring of output buffers per (shape, flags): the outputs of a
                       call (depth maps, loss tensors) stay valid until the second next call of the same kind
    materialize_warps  also produce ("sample" | "color" | "color_identity", f, s) as the reference does (trainer.py:909-918)
    """
    trainer_cls.generate_images_pred = generate_images_pred
    trainer_cls.compute_reprojection_loss = compute_reprojection_loss
    trainer_cls.compute_loss_masks = staticmethod(compute_loss_masks)
    trainer_cls.compute_losses = compute_losses
    trainer_cls.compute_matching_mask = compute_matching_mask
    trainer_cls.ppea_deterministic = deterministic
    trainer_cls.ppea_noise_mode = noise_mode
    trainer_cls.ppea_fused = fused     # None: single-launch training step whenever it applies (functional.VslConfig.fused)
    trainer_cls.ppea_plan_cache = plan_cache
    trainer_cls.ppea_materialize_warps = materialize_warps
    return trainer_cls


class ViewSynthesisLoss:
    """Stand-alone holder of the four methods for callers without the reference Trainer
    (tests, bench.py): ``ViewSynthesisLoss(opt).generate_images_pred(inputs, outputs)`` etc."""

    def __init__(self, opt, deterministic=True, noise_mode="reference", keep_maps=False, fused=None, plan_cache=True,
                 materialize_warps=False):
        self.opt = opt
        self.ppea_fused = fused
        self.ppea_plan_cache = plan_cache
        self.ppea_materialize_warps = materialize_warps
        self.ppea_deterministic = deterministic
        self.ppea_noise_mode = noise_mode
        self.ppea_keep_maps = keep_maps

    generate_images_pred = generate_images_pred
    compute_reprojection_loss = compute_reprojection_loss
    compute_loss_masks = staticmethod(compute_loss_masks)
    compute_losses = compute_losses
