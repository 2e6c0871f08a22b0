"""TEST TOOLING: builds tests/emul/vsl_emul.cpp (the CPU driver around the
scalar device functions of ppea_depth_b200/csrc/vsl_math.cuh) and exposes it
through ctypes, so the per-pixel arithmetic the CUDA kernels share can be
checked against the oracle in the GPU-less build container."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "vsl_emul.cpp")
OUT = os.path.join(HERE, "emul", "libvsl_emul.so")

F_MULTI, F_AUTOMASK, F_SELEC, F_NO_SSIM, F_DET, F_MOTION, F_AUG, F_POSE = (1 << i for i in range(8))


def build():
    deps = [SRC, os.path.join(ROOT, "ppea_depth_b200", "csrc", "vsl_math.cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "ppea_depth_b200", "csrc"), "-o", OUT, SRC])
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def flags_from_opt(opt, is_multi):
    fl = 0
    if is_multi:
        fl |= F_MULTI
    fl |= F_AUTOMASK          # always on; opt.disable_automasking only drops the noise (trainer.py:1084-1091)
    if opt.selec_reproj:
        fl |= F_SELEC
    if opt.no_ssim:
        fl |= F_NO_SSIM
    if not opt.disable_motion_masking:
        fl |= F_MOTION
    if not opt.no_matching_augmentation:
        fl |= F_AUG
    return fl


def forward_scale(inputs, outputs, opt, s, is_multi, noise):
    """Runs the emulator's forward for pyramid scale s; returns dict of maps."""
    B, H, W = opt.batch_size, opt.height, opt.width
    disp = outputs[("disp", s)].detach().float().contiguous()
    hs, ws = disp.shape[-2:]
    fl = flags_from_opt(opt, is_multi)
    f0, f1 = opt.frame_ids[1:]
    c = lambda t: t.detach().float().contiguous()
    tgt, s0, s1 = c(inputs[("color", 0, 0)]), c(inputs[("color", f0, 0)]), c(inputs[("color", f1, 0)])
    K, iK = c(inputs[("K", 0)]), c(inputs[("inv_K", 0)])
    T0, T1 = c(outputs[("cam_T_cam", 0, f0)]), c(outputs[("cam_T_cam", 0, f1)])
    nz = c(noise[s]) if (noise is not None and not opt.disable_automasking and not is_multi) else None
    cm = c(outputs["consistency_mask"]) if is_multi else None
    am = c(outputs["augmentation_mask"]).reshape(-1) if is_multi else None
    md = c(outputs[("mono_depth", 0, s)]) if is_multi else None
    depth = torch.empty(B, 1, H, W)
    loss_px = torch.empty(B, 1, H, W)
    sel = torch.empty(B, H, W, dtype=torch.uint8)
    w0, w1 = torch.empty(B, 3, H, W), torch.empty(B, 3, H, W)
    g0, g1 = torch.empty(B, H, W, 2), torch.empty(B, H, W, 2)
    sums = np.zeros(3, dtype=np.float64)
    lo = 1.0 / opt.max_depth
    rng = 1.0 / opt.min_depth - lo
    rc = lib().emul_vsl_forward(
        B, H, W, hs, ws, ctypes.c_uint(fl), ctypes.c_float(lo), ctypes.c_float(rng), ctypes.c_float(1e-7),
        _p(disp), _p(tgt), _p(s0), _p(s1), _p(K), _p(iK), _p(T0), _p(T1), _p(nz), _p(cm), _p(am), _p(md),
        _p(depth), _p(loss_px), _p(sel), _p(w0), _p(w1), _p(g0), _p(g1),
        sums.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    keep = (disp, tgt, s0, s1, K, iK, T0, T1, cm, am, md)
    return dict(depth=depth, loss_px=loss_px, sel=sel, warped=[w0, w1], grid=[g0, g1], sums=sums,
                flags=fl, _keep=keep, lo=lo, rng=rng, hs=hs, ws=ws)


def backward_scale(fw, opt, g_r, g_c):
    """Emulator backward of g_r*sum(r*mask) + g_c*sum(|depth-mono|*(1-mask)) for the
    scale whose forward result is `fw`; returns (grad_disp, grad_P[B,2,3,4])."""
    B, H, W = opt.batch_size, opt.height, opt.width
    disp, tgt, s0, s1, K, iK, T0, T1, cm, am, md = fw["_keep"]
    gd = torch.empty(B, 1, fw["hs"], fw["ws"])
    gP = np.zeros((B, 2, 3, 4), dtype=np.float64)
    rc = lib().emul_vsl_backward(
        B, H, W, fw["hs"], fw["ws"], ctypes.c_uint(fw["flags"]), ctypes.c_float(fw["lo"]), ctypes.c_float(fw["rng"]),
        ctypes.c_float(1e-7), _p(disp), _p(tgt), _p(s0), _p(s1), _p(K), _p(iK), _p(T0), _p(T1), _p(cm), _p(am),
        _p(md), _p(fw["sel"]), ctypes.c_float(g_r), ctypes.c_float(g_c), _p(gd),
        gP.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return gd, torch.from_numpy(gP)
