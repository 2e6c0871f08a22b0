"""ORACLE (test infrastructure, not product code).

CPU restatement of `transformation_from_parameters` (/root/reference/ppeadepth/layers.py:26-42) with its helpers
`get_translation_matrix` (:45-59) and `rot_from_axisangle` (:62-100), in the same torch ops (any dtype, so that autograd in
float64 gives the reference gradients).  Pinned by tests/golden/pose_transform.pt, which `make_golden()` below produced by
calling the reference's own function (`ppeadepth.layers` imports with torch alone), and against the reference live where it
is mounted (tests/test_pose.py).
"""
from __future__ import annotations

import os

import torch


def rot_from_axisangle(vec):
    angle = torch.norm(vec, 2, 2, True)                       # :67
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0].unsqueeze(1), axis[..., 1].unsqueeze(1), axis[..., 2].unsqueeze(1)
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), dtype=vec.dtype)
    rot[:, 0, 0] = torch.squeeze(x * xC + ca)                 # :87-96
    rot[:, 0, 1] = torch.squeeze(xyC - zs)
    rot[:, 0, 2] = torch.squeeze(zxC + ys)
    rot[:, 1, 0] = torch.squeeze(xyC + zs)
    rot[:, 1, 1] = torch.squeeze(y * yC + ca)
    rot[:, 1, 2] = torch.squeeze(yzC - xs)
    rot[:, 2, 0] = torch.squeeze(zxC - ys)
    rot[:, 2, 1] = torch.squeeze(yzC + xs)
    rot[:, 2, 2] = torch.squeeze(z * zC + ca)
    rot[:, 3, 3] = 1
    return rot


def get_translation_matrix(translation_vector):
    T = torch.zeros(translation_vector.shape[0], 4, 4, dtype=translation_vector.dtype)
    t = translation_vector.contiguous().view(-1, 3, 1)
    T[:, 0, 0] = 1
    T[:, 1, 1] = 1
    T[:, 2, 2] = 1
    T[:, 3, 3] = 1
    T[:, :3, 3, None] = t
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def synthetic_poses(B=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    aa = 0.3 * torch.randn(B, 1, 3, generator=g)
    tr = torch.randn(B, 1, 3, generator=g)
    aa[0] = 0.0                      # zero rotation: norm() at the origin
    aa[1] *= 1e-4                    # tiny rotation
    return aa, tr


def make_golden():
    """python -c "from oracle import pose_oracle as P; P.make_golden()"  (build container: needs /root/reference)"""
    from . import ref_import as R
    R.load_reference()
    import ppeadepth.layers as L
    aa, tr = synthetic_poses()
    out = {"axisangle": aa, "translation": tr}
    for inv in (False, True):
        a, t = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        M = L.transformation_from_parameters(a, t, inv)
        w = torch.linspace(0.5, 2.0, M.numel()).reshape(M.shape)
        (M * w).sum().backward()
        out["inv" if inv else "fwd"] = {"M": M.detach(), "w": w, "g_axisangle": a.grad.clone(), "g_translation": t.grad.clone()}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pose_transform.pt")
    torch.save(out, path)
    return path
