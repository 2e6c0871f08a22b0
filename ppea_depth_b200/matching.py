"""Plane-sweep cost volume of the multi-frame encoder.

Drop-in for `match_features` of the reference's matching encoders
(/root/reference/ppeadepth/networks/replk_matching_adapter.py:261-340, replk_matching.py:127-206,
resnet_encoder.py:164-246): same arguments, same two outputs, one CUDA launch instead of a Python loop over the
batch with ~15 launches and (D,C,h,w) temporaries per item.  `install_matching(EncoderClass)` rebinds the method; the
encoder keeps providing `num_depth_bins`, `warp_depths`, `set_missing_to_max` exactly as the reference builds them
(`compute_depth_bins`, :134-155).  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _cabi as C


def _f32c(t, name):
    if not t.is_cuda:
        raise RuntimeError("ppea_depth_b200 has no CPU path: %s must be a CUDA tensor" % name)
    return t.detach().contiguous().float()


_WORKSPACES = {}      # (device, bytes) -> scratch of the channel-quad feature copies (ppea_match_features_ws), reused across calls


def _match_workspace(device, B, F, Cn, h, w):
    need = C.lib().ppea_match_workspace_bytes(B, F, Cn, h, w)
    if need == 0:
        return None, 0
    key = (device.index, need)
    ws = _WORKSPACES.get(key)
    if ws is None:
        if len(_WORKSPACES) > 8:
            _WORKSPACES.clear()
        ws = _WORKSPACES[key] = torch.empty(need, device=device, dtype=torch.uint8)
    return ws, need


def match_features(current_feats, lookup_feats, relative_poses, K, invK, depth_bins, set_missing_to_max=True, eps=1e-7, planar=False):
    """current_feats (B,C,h,w), lookup_feats (B,F,C,h,w), relative_poses (B,F,4,4), K / invK (B,4,4) of the matching scale,
    depth_bins (D,) hypothesised depths -> (cost_volume (B,D,h,w), missing_mask (B,D,h,w)), both float32.  With C % 4 == 0 the
    features are re-laid channel-quad first (128-bit gathers, include/ppea_vsl.h ppea_match_features_ws); `planar=True` forces the
    planar kernel (bit-identical results; tests and A/B runs)."""
    cur = _f32c(current_feats, "current_feats")
    look = _f32c(lookup_feats, "lookup_feats")
    poses = _f32c(relative_poses, "relative_poses")
    K, invK = _f32c(K, "K"), _f32c(invK, "invK")
    B, Cn, h, w = cur.shape
    F = look.shape[1]
    if look.shape != (B, F, Cn, h, w) or poses.shape != (B, F, 4, 4) or K.shape != (B, 4, 4) or invK.shape != (B, 4, 4):
        raise ValueError("match_features: inconsistent shapes")
    bins = torch.as_tensor(depth_bins, dtype=torch.float32).reshape(-1).to(cur.device).contiguous()
    D = bins.numel()
    with torch.cuda.device(cur.device):
        cost = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        missing = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        ws, ws_bytes = (None, 0) if planar else _match_workspace(cur.device, B, F, Cn, h, w)
        C.check(C.lib().ppea_match_features_ws(cur.data_ptr(), look.data_ptr(), poses.data_ptr(), K.data_ptr(), invK.data_ptr(),
                                               bins.data_ptr(), cost.data_ptr(), missing.data_ptr(), B, F, Cn, h, w, D,
                                               1 if set_missing_to_max else 0, float(eps), ws.data_ptr() if ws is not None else None, ws_bytes,
                                               torch.cuda.current_stream().cuda_stream))
    return cost, missing


def match_features_dyn(current_feats, lookup_feats, relative_poses, K, invK, depth_bins, lookup_images, cv_min, aug_mask, set_1, pool,
                       pool_r, pool_th, set_missing_to_max=True, eps=1e-7):
    """`match_features_dyn` (replk_matching_adapter.py:163-258): match_features with the occlusion handling of the dynamic-scene
    stage -- `lookup_images` (N,3,H,W) give the occlusion map (sum of RGB < 0.15, nearest-resized to the matching resolution; the
    reference hard-codes 48x128, :166), `aug_mask` (B,1,1,1) switches it off per item, `set_1` / `pool` (radius `pool_r`, threshold
    `pool_th`) choose the replacement, `cv_min` the combination of the lookup frames.  Same two outputs as match_features."""
    import torch.nn.functional as Fnn
    cur = _f32c(current_feats, "current_feats")
    look = _f32c(lookup_feats, "lookup_feats")
    poses = _f32c(relative_poses, "relative_poses")
    K, invK = _f32c(K, "K"), _f32c(invK, "invK")
    B, Cn, h, w = cur.shape
    F = look.shape[1]
    if look.shape != (B, F, Cn, h, w) or poses.shape != (B, F, 4, 4) or K.shape != (B, 4, 4) or invK.shape != (B, 4, 4):
        raise ValueError("match_features_dyn: inconsistent shapes")
    if int(pool_r) > 2:
        raise ValueError("match_features_dyn: pool radius <= 2")
    bins = torch.as_tensor(depth_bins, dtype=torch.float32).reshape(-1).to(cur.device).contiguous()
    D = bins.numel()
    occ = aug = None
    if set_1 or pool:
        imgs = _f32c(lookup_images, "lookup_images")
        if imgs.shape[0] < B:
            raise ValueError("match_features_dyn: one lookup image per batch item expected")
        occ = (Fnn.interpolate((imgs.sum(1).unsqueeze(1) < 0.15).float(), [h, w])[:, 0] > 0).float().contiguous()      # :166, :198
        aug = _f32c(aug_mask, "aug_mask").reshape(aug_mask.shape[0], -1)[:B, 0].contiguous()                            # aug_mask[b][0][0][0]
    with torch.cuda.device(cur.device):
        cost = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        missing = torch.empty(B, D, h, w, device=cur.device, dtype=torch.float32)
        C.check(C.lib().ppea_match_features_dyn(cur.data_ptr(), look.data_ptr(), poses.data_ptr(), K.data_ptr(), invK.data_ptr(), bins.data_ptr(),
                                                occ.data_ptr() if occ is not None else None, aug.data_ptr() if aug is not None else None,
                                                cost.data_ptr(), missing.data_ptr(), B, F, Cn, h, w, D, 1 if set_missing_to_max else 0,
                                                1 if cv_min else 0, 1 if set_1 else 0, 1 if pool else 0, int(pool_r), float(pool_th), float(eps),
                                                torch.cuda.current_stream().cuda_stream))
    return cost, missing


def cost_volume_tail(cost_volume, missing_mask=None, num_bins_threshold=None, mask_volume=True):
    """The rest of the reference's matching block in one sweep (replk_matching_adapter.py:380-387, :439-453):
    confidence_mask = compute_confidence_mask(cost_volume * (1 - missing_mask)); mins, argmin = torch.min over the bins of
    the volume with its zeros set to 100; cost_volume *= confidence_mask (in place, when `mask_volume`).
    Returns (confidence_mask (B,h,w), mins (B,h,w), argmin (B,h,w) int64)."""
    if not cost_volume.is_cuda or cost_volume.dtype != torch.float32 or not cost_volume.is_contiguous():
        raise RuntimeError("cost_volume_tail: contiguous float32 CUDA volume expected (there is no CPU path)")
    B, D, h, w = cost_volume.shape
    miss = _f32c(missing_mask, "missing_mask") if missing_mask is not None else None
    thr = D if num_bins_threshold is None else int(num_bins_threshold)
    with torch.cuda.device(cost_volume.device):
        conf = torch.empty(B, h, w, device=cost_volume.device, dtype=torch.float32)
        mins = torch.empty(B, h, w, device=cost_volume.device, dtype=torch.float32)
        argmin = torch.empty(B, h, w, device=cost_volume.device, dtype=torch.int64)
        C.check(C.lib().ppea_match_tail(cost_volume.data_ptr(), miss.data_ptr() if miss is not None else None, conf.data_ptr(),
                                        mins.data_ptr(), argmin.data_ptr(), B, D, h, w, thr, 1 if mask_volume else 0,
                                        torch.cuda.current_stream().cuda_stream))
    return conf, mins, argmin


def compute_confidence_mask_method(self, cost_volume, num_bins_threshold=None):
    """Bound-method form of compute_confidence_mask (replk_matching_adapter.py:380-387)."""
    vol = _f32c(cost_volume, "cost_volume")
    thr = self.num_depth_bins if num_bins_threshold is None else num_bins_threshold
    B, D, h, w = vol.shape
    with torch.cuda.device(vol.device):
        conf = torch.empty(B, h, w, device=vol.device, dtype=torch.float32)
        C.check(C.lib().ppea_match_tail(vol.data_ptr(), None, conf.data_ptr(), None, None, B, D, h, w, int(thr), 0,
                                        torch.cuda.current_stream().cuda_stream))
    return conf


def match_features_method(self, current_feats, lookup_feats, relative_poses, K, invK):
    """Bound-method form with the reference's signature; reads `self.warp_depths` ((D,1,h,w): one depth per bin) and
    `self.set_missing_to_max` like the reference method."""
    bins = self.warp_depths.reshape(self.warp_depths.shape[0], -1)[:, 0]
    return match_features(current_feats, lookup_feats, relative_poses, K, invK, bins, bool(self.set_missing_to_max))


def match_features_dyn_method(self, current_feats, lookup_feats, relative_poses, K, invK, lookup_images, cv_min, aug_mask, set_1, pool,
                              pool_r, pool_th):
    """Bound-method form with the reference's signature (replk_matching_adapter.py:163)."""
    bins = self.warp_depths.reshape(self.warp_depths.shape[0], -1)[:, 0]
    return match_features_dyn(current_feats, lookup_feats, relative_poses, K, invK, bins, lookup_images, cv_min, aug_mask, set_1, pool,
                              pool_r, pool_th, bool(self.set_missing_to_max))


def install_matching(encoder_cls):
    """Rebinds `match_features`, `match_features_dyn` (where the class has it) and `compute_confidence_mask` of a reference matching
    encoder class (RepLKMatchingAdapter, RepLKMatching, ResnetEncoderMatching)."""
    encoder_cls.match_features = match_features_method
    if hasattr(encoder_cls, "match_features_dyn"):
        encoder_cls.match_features_dyn = match_features_dyn_method
    encoder_cls.compute_confidence_mask = compute_confidence_mask_method
    return encoder_cls
