"""Glue between the multi-frame encoder and the loss (remainder), kept on the device.

  matching_glue        repdepth.py:615-620 (the two nearest-neighbour upsamples) + trainer.py:450-451 / :859-869 (consistency
                       mask times the matching mask) in ONE launch, which also reduces what DepthBins.update needs
  DeviceDepthBins      drop-in for trainer.DepthBins (trainer.py:41-69): same attributes / methods, the running extrema stay
                       device tensors and `update` is one launch (no host max(), no .item())
  zero_missing_poses   repdepth.py:502-505 without its per-item host sync

There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C


def _cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError("ppea_depth_b200 has no CPU path: %s must be a CUDA tensor" % name)
    return t.detach().contiguous().float()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def matching_glue(lowest_cost, confidence_mask, mono_depth, want_extrema=True):
    """lowest_cost, confidence_mask (B,h,w); mono_depth (B,1,H,W)  ->  (outputs["lowest_cost"] (B,H,W),
    outputs["consistency_mask"] (B,H,W) already multiplied by compute_matching_mask, extrema): `extrema` is an opaque
    device buffer for DeviceDepthBins.update_from (None when not wanted)."""
    mono = _cuda_f32(mono_depth, "mono_depth")
    if not torch.is_tensor(lowest_cost):
        lowest_cost = torch.as_tensor(lowest_cost)
    lc = lowest_cost.to(mono.device).detach().contiguous().float()
    conf = _cuda_f32(confidence_mask, "confidence_mask")
    B, _, H, W = mono.shape
    h, w = lc.shape[-2], lc.shape[-1]
    if lc.numel() != B * h * w or conf.numel() != B * h * w:
        raise ValueError("matching_glue: lowest_cost / confidence_mask must be (B,h,w)")
    with torch.cuda.device(mono.device):
        up = torch.empty(B, H, W, device=mono.device, dtype=torch.float32)
        cons = torch.empty(B, H, W, device=mono.device, dtype=torch.float32)
        scratch = torch.empty(2 * B, device=mono.device, dtype=torch.int32) if want_extrema else None
        C.check(C.lib().ppea_matching_glue(lc.data_ptr(), conf.data_ptr(), mono.data_ptr(), up.data_ptr(), cons.data_ptr(),
                                           scratch.data_ptr() if scratch is not None else None, B, h, w, H, W, _stream()))
    return up, cons, scratch


class DeviceDepthBins:
    """trainer.DepthBins (trainer.py:41-69) with its state on the device: `min_depth` / `max_depth` are 1-element CUDA tensors,
    `update(mono_depth)` costs two launches and no host synchronisation, `update_from(extrema)` one (the extrema were reduced
    by matching_glue on its way through mono_depth)."""

    def __init__(self, opt_min_depth, device="cuda"):
        self.opt_min_depth = float(opt_min_depth)
        self.min_depth = torch.tensor([0.1], device=device)
        self.max_depth = torch.tensor([10.0], device=device)
        self.updated = False

    def update_from(self, extrema):
        self.updated = True
        with torch.cuda.device(self.min_depth.device):
            C.check(C.lib().ppea_depth_bins_update(extrema.data_ptr(), extrema.numel() // 2, self.opt_min_depth, self.min_depth.data_ptr(),
                                                   self.max_depth.data_ptr(), _stream()))

    def update(self, mono_depth):
        mono = _cuda_f32(mono_depth, "mono_depth")
        B = mono.shape[0]
        n = mono.numel() // B
        ones = torch.ones(B, 1, 1, device=mono.device)
        # (the glue kernel is the reduction: a 1x1 "cost volume" of ones keeps its other products trivial)
        _, _, extrema = matching_glue(ones, ones, mono.reshape(B, 1, 1, n))
        self.update_from(extrema)

    def load(self, min_depth, max_depth):
        self.min_depth = torch.as_tensor(min_depth, dtype=torch.float32).reshape(1).to(self.min_depth.device)
        self.max_depth = torch.as_tensor(max_depth, dtype=torch.float32).reshape(1).to(self.max_depth.device)

    def compute(self):
        return self.min_depth.float(), self.max_depth.float()


def zero_missing_poses(pose, pose_feats):
    """In place: pose[b] *= 0 for the batch items whose pose features are all zero (repdepth.py:502-505)."""
    if not pose.is_cuda or pose.dtype != torch.float32 or not pose.is_contiguous():
        raise RuntimeError("zero_missing_poses: contiguous float32 CUDA pose expected (there is no CPU path)")
    feats = _cuda_f32(pose_feats, "pose_feats")
    B = pose.shape[0]
    if feats.shape[0] != B:
        raise ValueError("zero_missing_poses: one feature block per batch item")
    with torch.cuda.device(pose.device):
        C.check(C.lib().ppea_zero_missing_poses(feats.data_ptr(), ctypes.c_size_t(feats.numel() // B), pose.data_ptr(), pose.numel() // B, B, _stream()))
    return pose
