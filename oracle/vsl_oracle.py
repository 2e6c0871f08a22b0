"""ORACLE (test infrastructure, not product code).

CPU restatement of PPEA-Depth's self-supervised view-synthesis loss path in
plain PyTorch ops, written from the behaviour of
  /root/reference/ppeadepth/layers.py:14-23, 138-257   and
  /root/reference/ppeadepth/trainer.py:871-918, 995-1160.
Only `tests/`, `__graft_entry__.smoke()` and bench.py's `cpu_baseline` /
`--impl reference` legs may import this module; the product path
(`ppea_depth_b200`) never does.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md §4/§8c), so this restatement is pinned against the reference's OWN
code executed in the build container (`oracle/ref_import.py`,
`oracle/make_golden.py`) and against the fixtures those scripts committed
under `tests/golden/`; `tests/test_oracle.py` re-checks both.

The arithmetic that the reference delegates to ATen (bilinear interpolate,
grid_sample, avg_pool2d, reflection pad) is *also* restated index-by-index, on the
kernels' own scalar arithmetic, by `tests/emul/vsl_emul.cpp` (driven by
`tests/test_emul.py`); this file uses the ATen ops so that it costs what the
reference costs on a CPU (it doubles as the timed CPU baseline).

Works in fp32 (the reference's precision) or fp64 (margin analysis).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F


def default_opt(**kw):
    """Flags the loss path reads (options.py).  `selec_reproj`
    is store_false in the reference => default True (options.py:428-430)."""
    o = dict(sclm=0, v1_multiscale=False, height=192, width=640, min_depth=0.1,
             max_depth=100.0, frame_ids=[0, -1, 1], disable_automasking=False,
             no_ssim=False, selec_reproj=True, disable_motion_masking=False,
             no_matching_augmentation=False, batch_size=12,
             disparity_smoothness=1e-3)
    o.update(kw)
    return SimpleNamespace(**o)


# --------------------------------------------------------------------------
# per-op restatements
# --------------------------------------------------------------------------
def images_from_u8(img_u8):
    """torchvision ToTensor on a uint8 image (datasets/mono_dataset.py:62, :106; torchvision functional.to_tensor:
    ``img.to(float32).div(255)``): the float frames the loss reads are exactly k/255."""
    return img_u8.to(torch.float32).div(255)


def compute_matching_mask(mono_depth, lowest_cost):
    """Trainer.compute_matching_mask, trainer.py:859-869 (mono_depth (B,1,H,W), lowest_cost (B,H,W)) -> (B,H,W) bool."""
    matching_depth = 1 / lowest_cost.unsqueeze(1)
    mask = ((matching_depth - mono_depth) / mono_depth) < 1.0
    mask = mask * (((mono_depth - matching_depth) / matching_depth) < 1.0)
    return mask[:, 0]


def disp_to_depth(disp, min_depth, max_depth):
    # layers.py:14-23 -- scaled = 1/max + (1/min - 1/max) * disp ; depth = 1/scaled
    lo = 1.0 / max_depth
    hi = 1.0 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1.0 / scaled


def upsample_disp(disp, height, width):
    # trainer.py:886-887
    return F.interpolate(disp, [height, width], mode="bilinear", align_corners=False)


def pixel_grid(batch, height, width, dtype, device):
    # layers.py:147-161 -- homogeneous pixel coordinates (x, y, 1) as (B,3,HW)
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype, device=device),
                            torch.arange(width, dtype=dtype, device=device), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype, device=device)], 0)
    return pix.unsqueeze(0).expand(batch, 3, height * width)


def backproject(depth, inv_K, height, width):
    # layers.py:163-168 -- cam = depth * (inv_K[:3,:3] @ pix), homogeneous 1 appended
    B = depth.shape[0]
    pix = pixel_grid(B, height, width, depth.dtype, depth.device)
    rays = torch.matmul(inv_K[:, :3, :3], pix)
    cam = depth.reshape(B, 1, -1) * rays
    return torch.cat([cam, torch.ones_like(cam[:, :1])], 1)


def project(points, K, T, height, width, eps=1e-7):
    # layers.py:184-199 (dc=False on the loss path, trainer.py:243)
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    c = torch.matmul(P, points)
    uv = c[:, :2] / (c[:, 2:3] + eps)
    uv = uv.reshape(B, 2, height, width).permute(0, 2, 3, 1)
    u = uv[..., 0] / (width - 1)
    v = uv[..., 1] / (height - 1)
    return (torch.stack([u, v], -1) - 0.5) * 2


def warp(src, grid):
    # trainer.py:911-914
    return F.grid_sample(src, grid, padding_mode="border", align_corners=True)


_C1 = 0.01 ** 2
_C2 = 0.03 ** 2


def ssim(x, y):
    # layers.py:243-257 -- 3x3 reflect-padded box moments, (1 - n/d)/2 clamped
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + _C1) * (2 * sigma_xy + _C2)
    d = (mu_x ** 2 + mu_y ** 2 + _C1) * (sigma_x + sigma_y + _C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def photometric(pred, target, no_ssim=False):
    # trainer.py:995-1007
    l1 = torch.abs(target - pred).mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def smoothness(disp, img):
    # layers.py:210-223
    dx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    dy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    ix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    iy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    return (dx * torch.exp(-ix)).mean() + (dy * torch.exp(-iy)).mean()


def normalised_smoothness(disp, img):
    # trainer.py:1147-1149
    mean_disp = disp.mean(2, True).mean(3, True)
    return smoothness(disp / (mean_disp + 1e-7), img)


# --------------------------------------------------------------------------
# the path: generate_images_pred + compute_losses for all scales
# --------------------------------------------------------------------------
def view_synthesis_losses(inputs, outputs, opt, is_multi=False, noise=None, want_maps=False, forced=None):
    """Returns (losses, maps).

    `noise`: list with one (B,1,H,W) standard-normal tensor per scale (the
    reference draws it with torch.randn on the CPU generator, trainer.py:1086);
    None => drawn here in the same order.
    `maps[s]` (when want_maps): per-pixel tensors r (post selec_reproj), ident
    (min identity loss + noise), mask, src_idx (0/1 source that receives the
    gradient, 2 = none), depth, warped images.
    `forced`: optional {s: (mask, src_idx)} -- evaluate the loss with a GIVEN selection instead of
    the oracle's own min/argmin decisions (both are piecewise constant, so autograd is unaffected);
    used to compare gradients between implementations whose fp32 decisions differ in a few
    near-tie pixels.
    """
    S = opt.sclm + 1
    srcs = list(opt.frame_ids[1:])
    losses, maps = {}, {}
    total = 0
    for s in range(S):
        ss = s if opt.v1_multiscale else 0
        Hs, Ws = opt.height >> ss, opt.width >> ss
        disp = outputs[("disp", s)]
        disp_up = disp if opt.v1_multiscale else upsample_disp(disp, opt.height, opt.width)
        _, depth = disp_to_depth(disp_up, opt.min_depth, opt.max_depth)      # trainer.py:890
        target = inputs[("color", 0, ss)]
        cam = backproject(depth, inputs[("inv_K", ss)], Hs, Ws)              # trainer.py:904
        warped, per_src, grids = [], [], []
        for f in srcs:
            T = outputs[("cam_T_cam", 0, f)]
            if is_multi:
                T = T.detach()                                               # trainer.py:900-902
            grid = project(cam, inputs[("K", ss)], T, Hs, Ws)
            w = warp(inputs[("color", f, ss)], grid)
            warped.append(w)
            grids.append(grid.detach())
            per_src.append(photometric(w, target, opt.no_ssim))
        per_src = torch.cat(per_src, 1)                                       # (B,nsrc,H,W)
        ident = torch.cat([photometric(inputs[("color", f, ss)], target, opt.no_ssim) for f in srcs], 1)
        ident = ident.min(1, keepdim=True)[0]                                 # trainer.py:1069

        r, amin = per_src.min(1, keepdim=True)                                # trainer.py:1076
        src_idx = amin.clone()
        if opt.selec_reproj:                                                  # trainer.py:1077-1083
            assert len(srcs) == 2, "selec_reproj hard-codes frames -1/+1 in the reference"
            dark_a = (warped[0].sum(1, keepdim=True) < 0.1).detach()
            dark_b = (warped[1].sum(1, keepdim=True) < 0.1).detach()
            r = torch.where(dark_a, per_src[:, 1:2], r)
            r = torch.where(dark_b, per_src[:, 0:1], r)
            r = torch.where(dark_a & dark_b, torch.zeros_like(r), r)
            src_idx = torch.where(dark_a, torch.ones_like(src_idx), src_idx)
            src_idx = torch.where(dark_b, torch.zeros_like(src_idx), src_idx)
            src_idx = torch.where(dark_a & dark_b, torch.full_like(src_idx, 2), src_idx)

        # trainer.py:1084-1091: `disable_automasking` only removes the tie-break noise -- the identity
        # loss is still handed to compute_loss_masks (never None), so the argmin mask is always applied.
        if not opt.disable_automasking:
            z = noise[s] if noise is not None else torch.randn(ident.shape)
            ident = ident + z.to(ident) * 0.00001
        mask = (r <= ident).to(r.dtype)              # == (argmin(cat[r,ident])==0), first-min tie rule

        if forced is not None:
            f_mask, f_src = forced[s]
            f_src = f_src.reshape(r.shape).to(torch.int64)
            r = torch.where(f_src == 0, per_src[:, 0:1], torch.where(f_src == 1, per_src[:, 1:2], torch.zeros_like(r)))
            src_idx = f_src
            mask = f_mask.reshape(r.shape).to(r.dtype)

        if is_multi:                                                          # trainer.py:1101-1109
            mask = torch.ones_like(mask)
            if not opt.disable_motion_masking:
                mask = mask * outputs["consistency_mask"].unsqueeze(1)
            if not opt.no_matching_augmentation:
                mask = mask * (1 - outputs["augmentation_mask"][:opt.batch_size])
            cons_mask = (1 - mask)

        reproj = (r * mask).sum() / (mask.sum() + 1e-7)                       # trainer.py:1113-1114
        loss = reproj
        losses["reproj_loss/{}".format(s)] = reproj
        if is_multi:                                                          # trainer.py:1127-1139
            mono_depth = outputs[("mono_depth", 0, s)].detach()
            cons = (torch.abs(depth - mono_depth) * cons_mask).mean()
            losses["consistency_loss/{}".format(s)] = cons
            loss = loss + cons
        smooth = normalised_smoothness(disp, inputs[("color", 0, s)])          # trainer.py:1147-1149
        losses["smooth_loss/{}".format(s)] = smooth
        loss = loss + opt.disparity_smoothness * smooth / (2 ** s)
        losses["loss/{}".format(s)] = loss
        total = total + loss
        if want_maps:
            maps[s] = dict(r=r.detach(), ident=ident.detach(), mask=mask.detach(),
                           src_idx=src_idx.detach(), depth=depth.detach(),
                           warped=[w.detach() for w in warped], per_src=per_src.detach(), grid=grids)
    losses["loss"] = total / S                                                # trainer.py:1157-1158
    return losses, maps


def clone_batch(inputs, outputs, dtype=torch.float32, device="cpu", requires_grad=True):
    """Deep-copies a synthetic batch; disp and poses become autograd leaves."""
    ins = {k: v.detach().to(device=device, dtype=dtype).clone() for k, v in inputs.items()}
    outs = {}
    for k, v in outputs.items():
        t = v.detach().to(device=device, dtype=dtype).clone()
        if requires_grad and isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam"):
            t.requires_grad_(True)
        outs[k] = t
    return ins, outs


def run_fwd_bwd(inputs, outputs, opt, is_multi=False, noise=None, dtype=torch.float32, want_maps=False, forced=None):
    """One timed unit of the metric: forward, then backward to
    every disp_s and (mono path) every T_f."""
    ins, outs = clone_batch(inputs, outputs, dtype=dtype)
    if noise is not None:
        noise = [z.to(dtype) for z in noise]
    losses, maps = view_synthesis_losses(ins, outs, opt, is_multi, noise, want_maps, forced)
    losses["loss"].backward()
    grads = {k: v.grad for k, v in outs.items()
             if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam") and v.grad is not None}
    return {k: v.detach() for k, v in losses.items()}, grads, maps
