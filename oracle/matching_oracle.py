"""ORACLE (test infrastructure, not product code).

CPU restatement of the plane-sweep cost volume of the multi-frame encoder,
`match_features` (/root/reference/ppeadepth/networks/replk_matching_adapter.py:261-340; the same code
sits in replk_matching.py:127-206 and resnet_encoder.py:164-246): for every batch item the lookup
features are warped into the current frame at `num_depth_bins` hypothesised depths
(BackprojectDepth + Project3D + F.grid_sample(padding_mode="zeros", align_corners=True)), the
channel-mean L1 difference to the current features is masked at the borders, averaged over the lookup
frames, and depth bins that never landed inside the image are set to the per-pixel maximum.

Pinned by running the reference's own method (unbound, on a stand-in `self`) in the build container:
`run_reference_match_features` / tests/golden/matching_*.pt / tests/test_matching.py.
"""
from __future__ import annotations

import types

import numpy as np
import torch
import torch.nn.functional as F

from . import vsl_oracle as O


def depth_bins(min_depth_bin, max_depth_bin, num_depth_bins, binning="linear"):
    # replk_matching_adapter.py:134-155
    if binning == "inverse":
        return (1 / np.linspace(1 / max_depth_bin, 1 / min_depth_bin, num_depth_bins)[::-1]).astype(np.float32)
    if binning == "linear":
        return np.linspace(min_depth_bin, max_depth_bin, num_depth_bins).astype(np.float32)
    raise NotImplementedError(binning)


def match_features(current_feats, lookup_feats, relative_poses, K, invK, bins, set_missing_to_max=True):
    """current_feats (B,C,h,w), lookup_feats (B,F,C,h,w), relative_poses (B,F,4,4), K/invK (B,4,4) of the
    matching scale, bins (D,) -> (cost_volume (B,D,h,w), missing_mask (B,D,h,w))."""
    B, C, h, w = current_feats.shape
    D = len(bins)
    bins_t = torch.as_tensor(np.asarray(bins), dtype=current_feats.dtype)
    warp_depths = bins_t.view(D, 1, 1, 1).expand(D, 1, h, w)
    vols, masks = [], []
    for b in range(B):
        cost = torch.zeros(D, h, w, dtype=current_feats.dtype)
        counts = torch.zeros(D, h, w, dtype=current_feats.dtype)
        world = O.backproject(warp_depths, invK[b:b + 1].expand(D, 4, 4), h, w)            # :281
        for f in range(lookup_feats.shape[1]):
            pose = relative_poses[b, f]
            if float(pose.sum()) == 0:                                                        # :289-291
                continue
            feat = lookup_feats[b, f].unsqueeze(0).expand(D, C, h, w)
            pix = O.project(world, K[b:b + 1].expand(D, 4, 4), pose.unsqueeze(0).expand(D, 4, 4), h, w)
            warped = F.grid_sample(feat, pix, padding_mode="zeros", mode="bilinear", align_corners=True)
            x_vals = (pix[..., 0] / 2 + 0.5) * (w - 1)                                         # :302-304
            y_vals = (pix[..., 1] / 2 + 0.5) * (h - 1)
            edge = ((x_vals >= 2.0) * (x_vals <= w - 2) * (y_vals >= 2.0) * (y_vals <= h - 2)).to(current_feats.dtype)
            cur_mask = torch.zeros_like(edge)
            cur_mask[:, 2:-2, 2:-2] = 1.0                                                     # :310-312
            edge = edge * cur_mask
            diffs = torch.abs(warped - current_feats[b:b + 1]).mean(1) * edge                 # :314-315
            cost = cost + diffs
            counts = counts + (diffs > 0).to(cost.dtype)
        cost = cost / (counts + 1e-7)                                                         # :321
        missing = (cost == 0).to(cost.dtype)
        if set_missing_to_max:                                                                # :325-328
            cost = cost * (1 - missing) + cost.max(0)[0].unsqueeze(0) * missing
        vols.append(cost)
        masks.append(missing)
    return torch.stack(vols, 0), torch.stack(masks, 0)


def cost_volume_tail(cost_volume, missing_mask, num_bins_threshold=None, mask_volume=True):
    """replk_matching_adapter.py:380-387 (compute_confidence_mask) and :439-453 (forward): confidence mask, min / argmin of the
    volume with its zeros set to 100, volume masked by the confidence.  Returns (confidence, mins, argmin, masked volume)."""
    D = cost_volume.shape[1]
    thr = D if num_bins_threshold is None else num_bins_threshold
    confidence = (((cost_volume * (1 - missing_mask)) > 0).sum(1) == thr).float()            # :380-387, :439-440
    viz = cost_volume.clone()
    viz[viz == 0] = 100                                                                      # :443-444
    mins, argmin = torch.min(viz, 1)                                                         # :445
    out = cost_volume * confidence.unsqueeze(1) if mask_volume else cost_volume.clone()      # :449
    return confidence, mins, argmin, out


def run_reference_match_features(current_feats, lookup_feats, relative_poses, K, invK, bins, set_missing_to_max=True):
    """The reference's own `RepLKMatchingAdapter.match_features`, unbound, on a stand-in self."""
    from . import ref_import as R
    R.load_reference()
    import ppeadepth.networks.replk_matching_adapter as M
    from ppeadepth.layers import BackprojectDepth, Project3D
    B, C, h, w = current_feats.shape
    D = len(bins)
    me = types.SimpleNamespace(num_depth_bins=D, matching_height=h, matching_width=w, set_missing_to_max=set_missing_to_max,
                               backprojector=BackprojectDepth(D, h, w), projector=Project3D(D, h, w))
    wd = [torch.ones((1, h, w)) * float(d) for d in bins]
    me.warp_depths = torch.stack(wd, 0).float()
    with torch.no_grad():
        return M.RepLKMatchingAdapter.match_features(me, current_feats, lookup_feats, relative_poses, K, invK)


def synthetic_case(B=2, Fr=1, C=16, h=12, w=20, D=8, seed=0, zero_pose_item=None, min_bin=0.5, max_bin=12.0):
    """Smooth random features, small relative poses, normalised-KITTI intrinsics of the matching scale."""
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(B, C, max(h // 4, 2), max(w // 4, 2), generator=g), size=(h, w), mode="bilinear",
                         align_corners=False)
    cur = (base + 0.05 * torch.randn(B, C, h, w, generator=g)).contiguous()
    look = torch.stack([(torch.roll(base, shifts=(0, 1 + f), dims=(2, 3)) + 0.05 * torch.randn(B, C, h, w, generator=g))
                        for f in range(Fr)], 1).contiguous()
    Kn = torch.tensor([[0.58 * w, 0, 0.5 * w, 0], [0, 1.92 * h, 0.5 * h, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
    K = Kn.float().unsqueeze(0).repeat(B, 1, 1)
    invK = torch.linalg.pinv(Kn).float().unsqueeze(0).repeat(B, 1, 1)
    from ppea_depth_b200.layers import transformation_from_parameters
    poses = torch.stack([transformation_from_parameters(0.02 * torch.randn(B, 1, 3, generator=g),
                                                        0.3 * torch.randn(B, 1, 3, generator=g)) for _ in range(Fr)], 1)
    if zero_pose_item is not None:
        poses[zero_pose_item] = 0.0          # "missing lookup frame" convention of the reference (:289-291)
    return cur, look, poses.contiguous(), K, invK, depth_bins(min_bin, max_bin, D)


# ---------------------------------------------------------------------------------------------------------------
# match_features_dyn (replk_matching_adapter.py:163-258): the variant selected when the encoder is given a teacher
# depth (:400, :439-442).  Same plane sweep, plus
#   * an occlusion map of the lookup image (sum of RGB < 0.15, nearest-resized to the matching resolution :166; the
#     reference hard-codes [48, 128] and 96 bins -- the restatement takes the sizes from the tensors), projected into
#     every layer of the volume; where it exceeds pool_th -- and the item is not augmented (aug_mask == 0) -- the warped
#     features are set to 1 (set_1) or replaced by the 3-D max-pool of the un-occluded neighbourhood (pool, radius pool_r);
#   * the frames combined by minimum (cv_min: zeros count as 1 before, ones are zeroed after) instead of the average.
# No caller inside the reference passes a teacher depth (repdepth.py:603-609 does not), so this is API surface only.
# ---------------------------------------------------------------------------------------------------------------
def occlusion_map(lookup_images, h, w):
    """(N,3,H,W) images -> (N,h,w) float: 1 where the nearest-resized darkness test fires (:166, `occ_batch[b] > 0`)."""
    occ = F.interpolate((lookup_images.sum(1).unsqueeze(1) < 0.15).float(), [h, w])
    return (occ[:, 0] > 0).float()


def match_features_dyn(current_feats, lookup_feats, relative_poses, K, invK, bins, lookup_images, cv_min, aug_mask, set_1, pool,
                       pool_r, pool_th, set_missing_to_max=True):
    B, C, h, w = current_feats.shape
    D = len(bins)
    dt = current_feats.dtype
    bins_t = torch.as_tensor(np.asarray(bins), dtype=dt)
    warp_depths = bins_t.view(D, 1, 1, 1).expand(D, 1, h, w)
    occ_batch = occlusion_map(lookup_images, h, w)                                             # (N,h,w), indexed by batch_idx
    vols, masks = [], []
    for b in range(B):
        cost = torch.ones(D, h, w, dtype=dt) if cv_min else torch.zeros(D, h, w, dtype=dt)
        counts = torch.zeros(D, h, w, dtype=dt)
        world = O.backproject(warp_depths, invK[b:b + 1].expand(D, 4, 4), h, w)
        for f in range(lookup_feats.shape[1]):
            pose = relative_poses[b, f]
            if float(pose.sum()) == 0:
                continue
            feat = lookup_feats[b, f].unsqueeze(0).expand(D, C, h, w)
            pix = O.project(world, K[b:b + 1].expand(D, 4, 4), pose.unsqueeze(0).expand(D, 4, 4), h, w)
            warped = F.grid_sample(feat, pix, padding_mode="zeros", mode="bilinear", align_corners=True).clone()
            if float(aug_mask.reshape(B, -1)[b, 0]) == 0 and (set_1 or pool):                   # :197-209
                occ = occ_batch[b].view(1, 1, h, w).expand(D, 1, h, w)
                m = F.grid_sample(occ, pix, padding_mode="zeros", mode="bilinear", align_corners=True) > pool_th
                m = m.expand(D, C, h, w)
                if set_1:
                    warped[m] = 1.0
                elif pool:
                    x = warped.clone()
                    x[m] = 0
                    x = F.max_pool3d(x.permute(1, 0, 2, 3), pool_r * 2 + 1, stride=1, padding=pool_r).permute(1, 0, 2, 3)
                    warped[m] = x[m]
            x_vals = (pix[..., 0] / 2 + 0.5) * (w - 1)
            y_vals = (pix[..., 1] / 2 + 0.5) * (h - 1)
            edge = ((x_vals >= 2.0) * (x_vals <= w - 2) * (y_vals >= 2.0) * (y_vals <= h - 2)).to(dt)
            cur_mask = torch.zeros_like(edge)
            cur_mask[:, 2:-2, 2:-2] = 1.0
            diffs = torch.abs(warped - current_feats[b:b + 1]).mean(1) * (edge * cur_mask)
            if cv_min:                                                                          # :238-240
                diffs = torch.where(diffs == 0, torch.ones_like(diffs), diffs)
                cost = torch.minimum(diffs, cost)
            else:
                cost = cost + diffs
                counts = counts + (diffs > 0).to(dt)
        if cv_min:                                                                              # :245-246
            cost = torch.where(cost == 1, torch.zeros_like(cost), cost)
        else:
            cost = cost / (counts + 1e-7)
        missing = (cost == 0).to(dt)
        if set_missing_to_max:
            cost = cost * (1 - missing) + cost.max(0)[0].unsqueeze(0) * missing
        vols.append(cost)
        masks.append(missing)
    return torch.stack(vols, 0), torch.stack(masks, 0)


def run_reference_match_features_dyn(current_feats, lookup_feats, relative_poses, K, invK, bins, lookup_images, cv_min, aug_mask, set_1,
                                     pool, pool_r, pool_th, set_missing_to_max=True):
    """The reference's own `RepLKMatchingAdapter.match_features_dyn`, unbound, on a stand-in self.  The method hard-codes the
    matching resolution 48x128 and 96 depth bins (:166, :199)."""
    from . import ref_import as R
    R.load_reference()
    import ppeadepth.networks.replk_matching_adapter as M
    from ppeadepth.layers import BackprojectDepth, Project3D
    B, C, h, w = current_feats.shape
    D = len(bins)
    assert (h, w, D) == (48, 128, 96), "the reference method only works at its hard-coded sizes"
    me = types.SimpleNamespace(num_depth_bins=D, matching_height=h, matching_width=w, set_missing_to_max=set_missing_to_max,
                               backprojector=BackprojectDepth(D, h, w), projector=Project3D(D, h, w))
    me.warp_depths = torch.stack([torch.ones((1, h, w)) * float(d) for d in bins], 0).float()
    with torch.no_grad():
        return M.RepLKMatchingAdapter.match_features_dyn(me, current_feats, lookup_feats, relative_poses, K, invK, lookup_images,
                                                         cv_min, aug_mask, set_1, pool, pool_r, pool_th)


def synthetic_dyn_case(B=1, C=4, h=48, w=128, D=96, seed=0, occluded=0.12, augmented_item=None):
    """synthetic_case + a lookup image with black (occluded) rectangles and the (B,1,1,1) augmentation mask."""
    cur, look, poses, K, invK, bins = synthetic_case(B=B, Fr=1, C=C, h=h, w=w, D=D, seed=seed, min_bin=0.5, max_bin=20.0)
    g = torch.Generator().manual_seed(1000 + seed)
    img = torch.round((0.2 + 0.6 * torch.rand(B, 3, 4 * h, 4 * w, generator=g)) * 255) / 255      # (exact k/255: stored as uint8)
    n = max(1, int(occluded * 40))
    for b in range(B):
        for _ in range(n):
            y0 = int(torch.randint(0, 4 * h - 24, (1,), generator=g))
            x0 = int(torch.randint(0, 4 * w - 40, (1,), generator=g))
            img[b, :, y0:y0 + 24, x0:x0 + 40] = 3.0 / 255
    aug = torch.zeros(B, 1, 1, 1)
    if augmented_item is not None:
        aug[augmented_item] = 1.0
    return cur, look, poses, K, invK, bins, img, aug
