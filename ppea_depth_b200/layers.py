"""Drop-in counterparts of /root/reference/ppeadepth/layers.py for the loss path.

Same names, constructor/forward signatures and return shapes as the reference
(`disp_to_depth` :14-23, `transformation_from_parameters` :26-42,
`BackprojectDepth` :138-168, `Project3D` :171-199, `upsample` :204-207,
`get_smooth_loss` :210-223, `SSIM` :226-257), so `trainer.py` and the
cost-volume encoders bind to them unchanged.  The module classes and
`get_smooth_loss` run hand-written sm_100a kernels through the C ABI
(`include/ppea_vsl.h`); `transformation_from_parameters` takes the fused pose
kernels for CUDA tensors (csrc/pose.cu) and stays PyTorch tensor algebra on the
host side (dataset / test tooling).  There is no CPU fallback: a CPU
tensor raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid disparity -> (scaled_disp, depth); reference layers.py:14-23."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled_disp = lo + (hi - lo) * disp
    return scaled_disp, 1 / scaled_disp


def get_translation_matrix(translation_vector):
    """(B,1,3)/(B,3) translation -> (B,4,4) homogeneous matrix; layers.py:45-59."""
    t = translation_vector.contiguous().view(-1, 3, 1)
    eye = torch.eye(4, device=t.device, dtype=t.dtype).expand(t.shape[0], 4, 4)
    top = torch.cat([eye[:, :3, :3], t], 2)
    return torch.cat([top, eye[:, 3:, :]], 1)


def rot_from_axisangle(vec):
    """(B,1,3) axis-angle -> (B,4,4) rotation (Rodrigues); layers.py:62-100."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    rows = [
        torch.cat([x * xC + ca, xyC - zs, zxC + ys, zero], 2),
        torch.cat([xyC + zs, y * yC + ca, yzC - xs, zero], 2),
        torch.cat([zxC - ys, yzC + xs, z * zC + ca, zero], 2),
        torch.cat([zero, zero, zero, one], 2),
    ]
    return torch.cat(rows, 1)


class _PoseToMatrix(torch.autograd.Function):
    """One launch each way (csrc/pose.cu) instead of ~30 + ~60 tiny ones."""

    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        import ctypes  # noqa: F401
        from . import _cabi as C
        aa = axisangle.detach().reshape(-1, 3).contiguous().float()
        tr = translation.detach().reshape(-1, 3).contiguous().float()
        B = aa.shape[0]
        with torch.cuda.device(aa.device):
            T = torch.empty(B, 4, 4, device=aa.device, dtype=torch.float32)
            C.check(C.lib().ppea_pose_to_matrix_forward(aa.data_ptr(), tr.data_ptr(), int(bool(invert)), T.data_ptr(), B,
                                                        torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(aa, tr)
        ctx.invert = bool(invert)
        ctx.shapes = (axisangle.shape, translation.shape)
        return T

    @staticmethod
    def backward(ctx, grad_T):
        from . import _cabi as C
        aa, tr = ctx.saved_tensors
        B = aa.shape[0]
        g = grad_T.contiguous().float()
        with torch.cuda.device(aa.device):
            g_aa, g_tr = torch.empty_like(aa), torch.empty_like(tr)
            C.check(C.lib().ppea_pose_to_matrix_backward(aa.data_ptr(), tr.data_ptr(), int(ctx.invert), g.data_ptr(), g_aa.data_ptr(),
                                                         g_tr.data_ptr(), B, torch.cuda.current_stream().cuda_stream))
        return g_aa.reshape(ctx.shapes[0]), g_tr.reshape(ctx.shapes[1]), None


def transformation_from_parameters(axisangle, translation, invert=False):
    """Pose-net output -> 4x4 camera transform; layers.py:26-42.  The fused loss returns dL/dT (B,4,4); autograd carries it
    through this function.  CUDA tensors take the fused kernels (`ppea_pose_to_matrix_forward/backward`); CPU tensors -- the
    dataset / test-tooling side, where the reference also runs it on the host -- the same tensor algebra in PyTorch."""
    if axisangle.is_cuda:
        return _PoseToMatrix.apply(axisangle, translation, invert)
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        return torch.matmul(R.transpose(1, 2), get_translation_matrix(-t))
    return torch.matmul(get_translation_matrix(t), R)


def upsample(x):
    """Nearest x2 upsample used by the decoders; layers.py:204-207."""
    return F.interpolate(x, scale_factor=2, mode="nearest")


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError(
                "ppea_depth_b200 runs sm_100a CUDA kernels only (no CPU fallback); "
                "got a tensor on %s" % t.device)


class BackprojectDepth(nn.Module):
    """Depth image -> homogeneous point cloud (B,4,H*W); layers.py:138-168."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        from . import functional as Fn
        _require_cuda(depth, inv_K)
        return Fn.backproject_depth(depth, inv_K, self.height, self.width)


class Project3D(nn.Module):
    """Points -> normalised sampling grid (B,H,W,2) through K @ T;
    layers.py:171-199.  With ``dc=True`` also returns the projected depth."""

    def __init__(self, batch_size: int, height: int, width: int, dc=False, eps=1e-7):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        self.dc, self.eps = dc, eps

    def forward(self, points, K, T):
        from . import functional as Fn
        _require_cuda(points, K, T)
        pix, z = Fn.project_3d(points, K, T, self.height, self.width, self.eps)
        if self.dc:
            return pix, z
        return pix


class SSIM(nn.Module):
    """3x3 reflect-padded SSIM dissimilarity map in [0,1]; layers.py:226-257."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y):
        from . import functional as Fn
        _require_cuda(x, y)
        return Fn.ssim(x, y)


def get_smooth_loss(disp, img):
    """Edge-aware first-order smoothness of `disp`; layers.py:210-223."""
    from . import functional as Fn
    _require_cuda(disp, img)
    return Fn.smooth_loss(disp, img)
