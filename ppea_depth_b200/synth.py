"""Synthetic KITTI/CityScapes-shaped batches for the view-synthesis loss path.

The reference trains on dataset images (uint8 -> ToTensor, so exact k/255
values, /root/reference/ppeadepth/datasets/mono_dataset.py:96-112) with
normalised KITTI intrinsics scaled per pyramid level
(kitti_dataset.py:26-29, mono_dataset.py:172-182).  There is no dataset on the
build or GPU box, so tests and bench.py use this generator:

* ``("color", f, 0)``: smooth random base + fine noise, quantised to k/255;
  the +-1 frames are small shifts of frame 0 (so the automask is mixed) with
  dark rectangles (so the ``selec_reproj`` dark-pixel rule fires);
* ``("color", 0, s)``: area-downsampled pyramid of frame 0;
* ``("K", s)`` / ``("inv_K", s)``: normalised K scaled by (W>>s, H>>s), pinv;
* ``("disp", s)``: smooth sigmoid-range maps, scaled so that the induced flow
  is a few pixels (as it is in real training, where depth is metres);
* ``("cam_T_cam", 0, f)``: small axis-angle / translation poses;
* multi-frame extras: consistency / augmentation masks and a mono depth.

Everything is generated on the CPU with an explicit ``torch.Generator`` so a
seed pins the batch; callers move the dicts to the GPU themselves.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

KITTI_K = ((0.58, 0.0, 0.5, 0.0), (0.0, 1.92, 0.5, 0.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 1.0))
# Plausible normalised CityScapes intrinsics (the reference reads per-frame
# *_cam.txt files that are not available offline).
CITYSCAPES_K = ((1.10, 0.0, 0.53, 0.0), (0.0, 2.95, 0.5, 0.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 1.0))


@dataclass
class SynthConfig:
    batch: int = 12
    height: int = 192
    width: int = 640
    num_scales: int = 4
    frame_ids: tuple = (0, -1, 1)
    intrinsics: tuple = KITTI_K
    seed: int = 0
    disp_lo: float = 0.01
    disp_hi: float = 0.09
    rot_std: float = 0.01
    trans_std: float = 0.05
    dark_frac: float = 0.02
    identity_pose: bool = False
    v1_multiscale: bool = False


def _smooth_field(gen, b, c, h, w, cell=8):
    ch, cw = max(h // cell, 2), max(w // cell, 2)
    base = torch.rand(b, c, ch, cw, generator=gen)
    return F.interpolate(base, size=(h, w), mode="bilinear", align_corners=False)


def _quantise(img):
    return torch.round(img.clamp(0.0, 1.0) * 255.0) / 255.0


def axisangle_to_matrix(axisangle, translation, invert=False):
    """Rodrigues rotation + translation -> (B,4,4); same maths as
    /root/reference/ppeadepth/layers.py:26-100 (kept in PyTorch there too)."""
    from .layers import transformation_from_parameters
    return transformation_from_parameters(axisangle, translation, invert)


def _flow_field(disp_full, K, inv_K, T, H, W):
    """Pixel displacement induced by (disp, K, T): the same geometry as
    layers.py:163-199, in float64, used only to synthesise consistent frames."""
    B = disp_full.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64),
                            torch.arange(W, dtype=torch.float64), indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(1, 3, -1)
    depth = 1.0 / (0.01 + 9.99 * disp_full.double().reshape(B, 1, -1))
    cam = depth * torch.matmul(inv_K.double()[:, :3, :3], pix)
    cam = torch.cat([cam, torch.ones_like(cam[:, :1])], 1)
    c = torch.matmul(torch.matmul(K.double(), T.double())[:, :3, :], cam)
    u = (c[:, 0] / (c[:, 2] + 1e-7)).reshape(B, H, W)
    v = (c[:, 1] / (c[:, 2] + 1e-7)).reshape(B, H, W)
    return u - xs, v - ys


def make_batch(cfg: SynthConfig):
    """Returns (inputs, outputs) dicts shaped like the reference Trainer's."""
    g = torch.Generator().manual_seed(cfg.seed)
    B, H, W, S = cfg.batch, cfg.height, cfg.width, cfg.num_scales
    inputs, outputs = {}, {}

    for s in range(S):
        K = torch.tensor(cfg.intrinsics, dtype=torch.float64)
        K[0, :] *= W // (2 ** s)
        K[1, :] *= H // (2 ** s)
        inv_K = torch.linalg.pinv(K)
        inputs[("K", s)] = K.float().unsqueeze(0).repeat(B, 1, 1).contiguous()
        inputs[("inv_K", s)] = inv_K.float().unsqueeze(0).repeat(B, 1, 1).contiguous()

    # disparity pyramid: one smooth field, consistent across scales (as the
    # outputs of one decoder are), plus a small per-scale perturbation
    field = (_smooth_field(g, B, 1, H, W, cell=16) + 0.03 * torch.rand(B, 1, H, W, generator=g)).clamp(0, 1)
    for s in range(S):
        k = 2 ** s
        f_s = F.avg_pool2d(field, k) if k > 1 else field
        f_s = (f_s + 0.02 * s * (torch.rand(f_s.shape, generator=g) - 0.5)).clamp(0, 1)
        outputs[("disp", s)] = (cfg.disp_lo + (cfg.disp_hi - cfg.disp_lo) * f_s).contiguous()

    for f in cfg.frame_ids[1:]:
        if cfg.identity_pose:
            aa = torch.zeros(B, 1, 3)
            tr = torch.zeros(B, 1, 3)
        else:
            aa = cfg.rot_std * torch.randn(B, 1, 3, generator=g)
            tr = cfg.trans_std * torch.randn(B, 1, 3, generator=g)
        outputs[("axisangle", 0, f)] = aa
        outputs[("translation", 0, f)] = tr
        outputs[("cam_T_cam", 0, f)] = axisangle_to_matrix(aa, tr, invert=(f < 0)).contiguous()

    # frame 0 texture; the +-1 frames are that texture displaced by (minus) the
    # flow the scale-0 geometry predicts, so warping them back roughly
    # re-aligns with frame 0; "static" blobs keep frame 0 (identity loss wins
    # there => mixed automask); dark rectangles trigger selec_reproj.
    base = 0.15 + 0.7 * _smooth_field(g, B, 3, H, W)
    texture = base + 0.05 * torch.randn(B, 3, H, W, generator=g)
    inputs[("color", 0, 0)] = _quantise(texture)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64),
                            torch.arange(W, dtype=torch.float64), indexing="ij")
    static = (_smooth_field(g, B, 1, H, W, cell=24) > 0.62).float()
    rh, rw = max(H // 10, 2), max(W // 10, 2)
    n_rect = max(1, int(round(cfg.dark_frac * 50))) if cfg.dark_frac > 0 else 0
    shared_rects = [(int(torch.randint(0, H - rh + 1, (1,), generator=g)),
                     int(torch.randint(0, W - rw + 1, (1,), generator=g))) for _ in range(B)]
    for f in cfg.frame_ids[1:]:
        du, dv = _flow_field(outputs[("disp", 0)], inputs[("K", 0)], inputs[("inv_K", 0)],
                             outputs[("cam_T_cam", 0, f)], H, W)
        gx = ((xs - du) / (W - 1) - 0.5) * 2
        gy = ((ys - dv) / (H - 1) - 0.5) * 2
        grid = torch.stack([gx, gy], -1).float()
        moved = F.grid_sample(texture, grid, padding_mode="border", align_corners=True)
        src = static * texture + (1 - static) * moved
        src = src + 0.01 * torch.randn(B, 3, H, W, generator=g)
        for b in range(B):
            rects = [shared_rects[b]] if n_rect else []
            for _ in range(n_rect):
                rects.append((int(torch.randint(0, H - rh + 1, (1,), generator=g)),
                              int(torch.randint(0, W - rw + 1, (1,), generator=g))))
            for (y0, x0) in rects:
                src[b, :, y0:y0 + rh, x0:x0 + rw] = 0.0
        inputs[("color", f, 0)] = _quantise(src)
    for s in range(1, S):
        for f in cfg.frame_ids:
            if f != 0 and not cfg.v1_multiscale:
                continue
            inputs[("color", f, s)] = _quantise(F.avg_pool2d(inputs[("color", f, 0)], 2 ** s))

    # multi-frame extras (trainer.py:1101-1141, networks/repdepth.py:559-577)
    outputs["consistency_mask"] = (torch.rand(B, H, W, generator=g) < 0.7).float()
    outputs["augmentation_mask"] = (torch.rand(B, 1, 1, 1, generator=g) < 0.5).float()
    for s in range(S):
        hh, ww = (H // (2 ** s), W // (2 ** s)) if cfg.v1_multiscale else (H, W)
        md = 0.5 + 9.5 * _smooth_field(g, B, 1, hh, ww)
        outputs[("mono_depth", 0, s)] = md.contiguous()
    return inputs, outputs


def make_noise(cfg: SynthConfig, n_draws: int, seed: int | None = None):
    """The reference draws torch.randn(B,1,H,W) on the CPU default generator
    once per scale per compute_losses call (trainer.py:1084-1087).  Tests feed
    the same tensors to the oracle and to the CUDA path."""
    g = torch.Generator().manual_seed(cfg.seed + 1000 if seed is None else seed)
    out = []
    for s in range(n_draws):
        hh, ww = ((cfg.height >> s, cfg.width >> s) if cfg.v1_multiscale
                  else (cfg.height, cfg.width))
        out.append(torch.randn(cfg.batch, 1, hh, ww, generator=g))
    return out


def algorithmic_bytes(batch, height, width, num_scales, is_multi=False, deterministic=False):
    """Compulsory HBM traffic of one fwd+bwd pass (SURVEY.md §8d / BASELINE.md §3):
    per full-res pixel at scale s, mono/atomics: 86 + 24/4^s bytes."""
    n = batch * height * width
    per_px = 0.0
    for s in range(num_scales):
        base = 98.0 if is_multi else 86.0
        if deterministic:
            base += 8.0
        per_px += base + 24.0 / (4 ** s)
    return per_px * n
