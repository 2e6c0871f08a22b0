"""torch.autograd.Function wrappers around the C ABI (include/ppea_vsl.h).

`view_synthesis_loss` is the fused replacement of the loop bodies of
`Trainer.generate_images_pred` + `Trainer.compute_losses`
(/root/reference/ppeadepth/trainer.py:871-918, 1032-1160) for all pyramid
scales of one call: one forward launch sequence, one backward launch sequence,
gradients to every ``disp_s`` and (mono path) both poses.  The remaining
functions are the piecewise operators behind the reference's nn.Modules.

PyTorch is plumbing here (device memory, streams, autograd graph); all
arithmetic runs in libppea_vsl.so.  CPU tensors raise -- there is no fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _cabi as C


def _f32c(t, name):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ppea_depth_b200: %s must be a CUDA tensor (no CPU fallback), got %s" % (name, t.device))
    if t.dtype != torch.float32:
        t = t.float()     # under autocast the loss path stays fp32, as grid_sample's autocast policy does
    return t.contiguous()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _mat44(m, batch, name):
    """(B,4,4) contiguous fp32; a batch-1 matrix is expanded (the reference relies on matmul broadcasting)."""
    if m.dim() != 3 or tuple(m.shape[-2:]) != (4, 4) or m.shape[0] not in (1, batch):
        raise ValueError("%s must be (%d,4,4) or (1,4,4), got %s" % (name, batch, tuple(m.shape)))
    if m.shape[0] != batch:
        m = m.expand(batch, 4, 4)
    return m.contiguous()


@dataclass
class VslConfig:
    """Flags and constants of one call (the options the reference path reads)."""
    is_multi: bool = False
    automask: bool = True            # identity-reprojection automask (the reference always applies it on the mono path)
    selec_reproj: bool = True        # opt.selec_reproj (default True)
    no_ssim: bool = False
    motion_mask: bool = True         # not opt.disable_motion_masking
    match_aug: bool = True           # not opt.no_matching_augmentation
    deterministic: bool = False
    min_depth: float = 0.1
    max_depth: float = 100.0
    eps: float = 1e-7
    disparity_smoothness: float = 1e-3
    first_scale: int = 0
    total_scales: Optional[int] = None
    want_loss_px: bool = False
    plan_cache: bool = True          # training steps have static shapes: reuse one pre-built parameter block and a ring of
                                     # output / workspace / gradient buffers per (shape, flags) instead of allocating ~20
                                     # tensors and refilling the C struct every call (see _StepPlan).  Outputs of a call
                                     # stay valid until the second next call with the same shapes and flags.
    fused: object = None             # fused training step; None / True = whenever some input requires grad: the
                                     # warp-streaming kernel (vsl_stream.cu); "tiles" = the shared-memory tile kernel
                                     # (vsl_fused.cu); False = the forward + backward kernel pair

    def use_fused(self, needs_grad):
        return bool(needs_grad) and (self.fused is None or bool(self.fused))

    def flags(self, grad_pose):
        f = 0
        if self.is_multi:
            f |= C.F_MULTI
        if self.automask:
            f |= C.F_AUTOMASK
        if self.selec_reproj:
            f |= C.F_SELEC_REPROJ
        if self.no_ssim:
            f |= C.F_NO_SSIM
        if self.deterministic:
            f |= C.F_DETERMINISTIC
        if self.motion_mask:
            f |= C.F_MOTION_MASK
        if self.match_aug:
            f |= C.F_MATCH_AUG
        if grad_pose:
            f |= C.F_GRAD_POSE
        if self.fused == "tiles":
            f |= C.F_FUSED_TILES
        return f


@dataclass
class VslResult:
    losses: torch.Tensor                  # (1 + 4*S,): [loss, (loss/s, reproj_loss/s, consistency_loss/s, smooth/s) * S]
    depth: List[torch.Tensor]             # per scale (B,1,H,W)  == outputs[("depth", 0, s)]
    sel: List[torch.Tensor]               # per scale (B,H,W) uint8 selection map (PPEA_SEL_*)
    loss_px: List[Optional[torch.Tensor]] # per scale (B,1,H,W) per-pixel reprojection loss (if requested)
    sums: torch.Tensor = field(default=None)

    @property
    def loss(self):
        return self.losses[0]

    def scale_loss(self, s):
        return self.losses[1 + C.LOSSES_PER_SCALE * s]

    def reproj_loss(self, s):
        return self.losses[1 + C.LOSSES_PER_SCALE * s + 1]

    def consistency_loss(self, s):
        return self.losses[1 + C.LOSSES_PER_SCALE * s + 2]

    def smooth_loss(self, s):
        return self.losses[1 + C.LOSSES_PER_SCALE * s + 3]

    def automask(self, s):
        return ((self.sel[s] & C.SEL_AUTOMASK) != 0)

    def source_index(self, s):
        return (self.sel[s] & C.SEL_SRC_MASK)


class _Bundle:
    """Non-differentiable inputs + configuration of one fused call."""
    __slots__ = ("cfg", "tgt", "src", "K", "inv_K", "cons_mask", "aug_mask", "colors", "noise", "mono_depth",
                 "B", "H", "W", "S", "result")


def _fill_params(bundle, flags, T, disps, depth, loss_px, sel, grad_disp, sums, losses, workspace):
    cfg = bundle.cfg
    p = C.PpeaVslParams()
    p.struct_size = ctypes.sizeof(C.PpeaVslParams)
    p.flags = flags
    p.batch, p.height, p.width = bundle.B, bundle.H, bundle.W
    p.num_scales = bundle.S
    p.first_scale = cfg.first_scale
    p.total_scales = cfg.total_scales if cfg.total_scales is not None else bundle.S
    lo = 1.0 / cfg.max_depth
    p.disp_lo = lo
    p.disp_range = 1.0 / cfg.min_depth - lo
    p.eps = cfg.eps
    p.disparity_smoothness = cfg.disparity_smoothness
    p.tgt = bundle.tgt.data_ptr()
    p.src[0], p.src[1] = bundle.src[0].data_ptr(), bundle.src[1].data_ptr()
    p.K, p.inv_K = bundle.K.data_ptr(), bundle.inv_K.data_ptr()
    p.T[0], p.T[1] = T[0].data_ptr(), T[1].data_ptr()
    p.cons_mask = bundle.cons_mask.data_ptr() if bundle.cons_mask is not None else None
    p.aug_mask = bundle.aug_mask.data_ptr() if bundle.aug_mask is not None else None
    for s in range(bundle.S):
        sc = p.scales[s]
        sc.disp_h, sc.disp_w = disps[s].shape[-2], disps[s].shape[-1]
        sc.disp = disps[s].data_ptr()
        sc.color = bundle.colors[s].data_ptr()
        sc.noise = bundle.noise[s].data_ptr() if bundle.noise is not None else None
        sc.mono_depth = bundle.mono_depth[s].data_ptr() if bundle.mono_depth is not None else None
        sc.depth = depth[s].data_ptr()
        sc.loss_px = loss_px[s].data_ptr() if loss_px[s] is not None else None
        sc.sel = sel[s].data_ptr()
        sc.grad_disp = grad_disp[s].data_ptr() if grad_disp is not None else None   # forward: NULL => no pre-zeroing
    p.sums = sums.data_ptr()
    p.losses = losses.data_ptr()
    p.workspace = workspace.data_ptr()
    p.workspace_bytes = workspace.numel() * workspace.element_size()
    return p


def _fused_struct(ws):
    f = C.PpeaVslFused()
    f.struct_size = ctypes.sizeof(C.PpeaVslFused)
    f.workspace = ws.data_ptr()
    f.workspace_bytes = ws.numel() * 4
    return f


class _FusedViewSynthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bundle, T0, T1, *disps):
        lib = C.lib()
        dev = bundle.tgt.device
        B, H, W, S = bundle.B, bundle.H, bundle.W, bundle.S
        T = (_f32c(T0, "T[0]"), _f32c(T1, "T[1]"))
        disps = tuple(_f32c(d, "disp") for d in disps)
        with torch.cuda.device(dev):
            depth = [torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) for _ in range(S)]
            sel = [torch.empty(B, H, W, device=dev, dtype=torch.uint8) for _ in range(S)]
            loss_px = [torch.empty(B, 1, H, W, device=dev, dtype=torch.float32) if bundle.cfg.want_loss_px else None
                       for _ in range(S)]
            sums = torch.empty(lib.ppea_vsl_sums_floats(B, S), device=dev, dtype=torch.float32)
            losses = torch.empty(1 + C.LOSSES_PER_SCALE * S, device=dev, dtype=torch.float32)
            ws = torch.empty(lib.ppea_vsl_workspace_bytes(B, H, W, S) // 4, device=dev, dtype=torch.float32)
            fused = bundle.cfg.use_fused(any(ctx.needs_input_grad))
            fws = None
            if fused:
                # one launch evaluates the loss and the un-normalised gradient fields; backward() only rescales them
                grad_pose = bool(ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
                flags = bundle.cfg.flags(grad_pose=grad_pose)
                p = _fill_params(bundle, flags, T, disps, depth, loss_px, sel, None, sums, losses, ws)
                fws = torch.empty(max(lib.ppea_vsl_fused_workspace_bytes(ctypes.byref(p)) // 4, 4), device=dev, dtype=torch.float32)
                C.check(lib.ppea_vsl_fused_forward(ctypes.byref(p), ctypes.byref(_fused_struct(fws)), _stream()))
                ctx.fused_flags = flags
            else:
                flags = bundle.cfg.flags(grad_pose=False)
                p = _fill_params(bundle, flags, T, disps, depth, loss_px, sel, None, sums, losses, ws)
                C.check(lib.ppea_vsl_forward(ctypes.byref(p), _stream()))
        ctx.bundle = bundle
        ctx.fused_ws = fws
        ctx.saved = (T, disps, depth, sel, loss_px, sums)
        ctx.mark_non_differentiable(*depth, *sel, *[t for t in loss_px if t is not None], sums)
        outs = [losses] + depth + sel + [sums] + [t for t in loss_px if t is not None]
        return tuple(outs)

    @staticmethod
    def backward(ctx, grad_losses, *unused):
        lib = C.lib()
        bundle = ctx.bundle
        T, disps, depth, sel, loss_px, sums = ctx.saved
        dev = bundle.tgt.device
        B, H, W, S = bundle.B, bundle.H, bundle.W, bundle.S
        grad_pose = (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and not bundle.cfg.is_multi
        with torch.cuda.device(dev):
            if grad_losses is None:
                grad_losses = torch.zeros(1 + C.LOSSES_PER_SCALE * S, device=dev, dtype=torch.float32)
            grad_losses = grad_losses.contiguous().float()
            if ctx.fused_ws is not None:
                flags = ctx.fused_flags
                grad_pose = bool(flags & C.F_GRAD_POSE)
            else:
                flags = bundle.cfg.flags(grad_pose=grad_pose)
            grad_disp = [torch.empty_like(d) for d in disps]
            gT = [torch.empty(B, 4, 4, device=dev, dtype=torch.float32) for _ in range(2)] if grad_pose else None
            if ctx.fused_ws is not None:
                scratch_fwd = torch.empty(4, device=dev, dtype=torch.float32)   # forward workspace is not used by backward
                p = _fill_params(bundle, flags, T, disps, depth, loss_px, sel, grad_disp, sums, grad_losses, scratch_fwd)
                g = C.PpeaVslGrads()
                g.struct_size = ctypes.sizeof(C.PpeaVslGrads)
                g.grad_losses = grad_losses.data_ptr()
                if grad_pose:
                    g.grad_T[0], g.grad_T[1] = gT[0].data_ptr(), gT[1].data_ptr()
                C.check(lib.ppea_vsl_fused_backward(ctypes.byref(p), ctypes.byref(g), ctypes.byref(_fused_struct(ctx.fused_ws)), _stream()))
                return (None, gT[0] if grad_pose else None, gT[1] if grad_pose else None, *grad_disp)
            ws_bytes = lib.ppea_vsl_backward_workspace_bytes(B, H, W, S, flags)
            ws = torch.empty(max(ws_bytes // 4, 4), device=dev, dtype=torch.float32)
            scratch_fwd = torch.empty(4, device=dev, dtype=torch.float32)   # forward workspace is not used by backward
            p = _fill_params(bundle, flags, T, disps, depth, loss_px, sel, grad_disp, sums, grad_losses, scratch_fwd)
            g = C.PpeaVslGrads()
            g.struct_size = ctypes.sizeof(C.PpeaVslGrads)
            g.grad_losses = grad_losses.data_ptr()
            if grad_pose:
                g.grad_T[0], g.grad_T[1] = gT[0].data_ptr(), gT[1].data_ptr()
            g.workspace = ws.data_ptr()
            g.workspace_bytes = ws.numel() * 4
            C.check(lib.ppea_vsl_backward(ctypes.byref(p), ctypes.byref(g), _stream()))
        return (None, gT[0] if grad_pose else None, gT[1] if grad_pose else None, *grad_disp)


# ---------------------------------------------------------------------------
# Cached execution plans of the fused training step.  The reference's training step has static shapes
# (batch_size is baked into its modules, layers.py:142-161; ragged batches are dropped, trainer.py:215-218), so
# everything `_FusedViewSynthLoss` does per call on the host -- ~20 tensor allocations, ~70 ctypes field writes,
# ten tensors through the autograd engine -- is hoisted into a plan keyed by (device, shapes, flags):
# a ring of RING buffer sets, each with its parameter block already pointing at its outputs.  A call rebinds the
# ~20 input pointers and launches; only the `losses` vector goes through autograd.
# ---------------------------------------------------------------------------
_PLAN_RING = 2
_PLANS = {}


class _PlanSlot:
    __slots__ = ("depth", "sel", "loss_px", "sums", "losses", "ws", "fws", "grad_disp", "gT", "p", "g", "f", "pending", "zero_gl")


class _StepPlan:
    def __init__(self, bundle, flags, disp_shapes):
        lib = C.lib()
        dev = bundle.tgt.device
        B, H, W, S = bundle.B, bundle.H, bundle.W, bundle.S
        self.flags = flags | C.F_RAW_PREZEROED      # the slot's workspace is zeroed once here and kept clean by every backward
        self.slots = []
        self.next = 0
        f32 = dict(device=dev, dtype=torch.float32)
        cfg = bundle.cfg
        with torch.cuda.device(dev):
            for _ in range(_PLAN_RING):
                sl = _PlanSlot()
                sl.depth = [torch.empty(B, 1, H, W, **f32) for _ in range(S)]
                sl.sel = [torch.empty(B, H, W, device=dev, dtype=torch.uint8) for _ in range(S)]
                sl.loss_px = [torch.empty(B, 1, H, W, **f32) if cfg.want_loss_px else None for _ in range(S)]
                sl.sums = torch.empty(lib.ppea_vsl_sums_floats(B, S), **f32)
                sl.losses = torch.empty(1 + C.LOSSES_PER_SCALE * S, **f32)
                sl.ws = torch.empty(max(lib.ppea_vsl_workspace_bytes(B, H, W, S) // 4, 4), **f32)
                sl.grad_disp = [torch.empty(shape, **f32) for shape in disp_shapes]
                sl.gT = [torch.empty(B, 4, 4, **f32) for _ in range(2)] if (flags & C.F_GRAD_POSE) else None
                sl.zero_gl = None
                p = C.PpeaVslParams()
                p.struct_size = ctypes.sizeof(C.PpeaVslParams)
                p.flags = self.flags
                p.batch, p.height, p.width = B, H, W
                p.num_scales = S
                p.first_scale = cfg.first_scale
                p.total_scales = cfg.total_scales if cfg.total_scales is not None else S
                lo = 1.0 / cfg.max_depth
                p.disp_lo, p.disp_range, p.eps = lo, 1.0 / cfg.min_depth - lo, cfg.eps
                p.disparity_smoothness = cfg.disparity_smoothness
                for s in range(S):
                    sc = p.scales[s]
                    sc.disp_h, sc.disp_w = disp_shapes[s][-2], disp_shapes[s][-1]
                    sc.depth = sl.depth[s].data_ptr()
                    sc.loss_px = sl.loss_px[s].data_ptr() if sl.loss_px[s] is not None else None
                    sc.sel = sl.sel[s].data_ptr()
                    sc.grad_disp = sl.grad_disp[s].data_ptr()
                p.sums, p.losses = sl.sums.data_ptr(), sl.losses.data_ptr()
                p.workspace, p.workspace_bytes = sl.ws.data_ptr(), sl.ws.numel() * 4
                # the fused workspace is sized from a fully described call: bind this call's inputs first
                self._bind(p, bundle, None, None)
                sl.fws = torch.empty(max(lib.ppea_vsl_fused_workspace_bytes(ctypes.byref(p)) // 4, 4), **f32)
                sl.fws.zero_()
                sl.p = p
                sl.f = _fused_struct(sl.fws)
                g = C.PpeaVslGrads()
                g.struct_size = ctypes.sizeof(C.PpeaVslGrads)
                if sl.gT is not None:
                    g.grad_T[0], g.grad_T[1] = sl.gT[0].data_ptr(), sl.gT[1].data_ptr()
                sl.g = g
                sl.pending = False
                self.slots.append(sl)

    @staticmethod
    def _bind(p, b, T, disps):
        p.tgt = b.tgt.data_ptr()
        p.src[0], p.src[1] = b.src[0].data_ptr(), b.src[1].data_ptr()
        p.K, p.inv_K = b.K.data_ptr(), b.inv_K.data_ptr()
        if T is not None:
            p.T[0], p.T[1] = T[0].data_ptr(), T[1].data_ptr()
        p.cons_mask = b.cons_mask.data_ptr() if b.cons_mask is not None else None
        p.aug_mask = b.aug_mask.data_ptr() if b.aug_mask is not None else None
        for s in range(b.S):
            sc = p.scales[s]
            if disps is not None:
                sc.disp = disps[s].data_ptr()
            sc.color = b.colors[s].data_ptr()
            sc.noise = b.noise[s].data_ptr() if b.noise is not None else None
            sc.mono_depth = b.mono_depth[s].data_ptr() if b.mono_depth is not None else None

    def take(self):
        sl = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if sl.pending:                 # this slot's last forward never saw its backward: its raw gradient fields are dirty
            sl.fws.zero_()
        return sl


def _plan_for(bundle, flags, disps):
    shapes = tuple(tuple(d.shape) for d in disps)
    cfg = bundle.cfg
    key = (bundle.tgt.device.index, bundle.B, bundle.H, bundle.W, shapes, flags, cfg.want_loss_px, cfg.first_scale, cfg.total_scales,
           cfg.min_depth, cfg.max_depth, cfg.eps, cfg.disparity_smoothness, bundle.noise is not None)
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= 16:
            _PLANS.pop(next(iter(_PLANS)))
        plan = _PLANS[key] = _StepPlan(bundle, flags, shapes)
    return plan


class _PlannedFusedLoss(torch.autograd.Function):
    """The fused training step through a cached plan: one tensor (the loss vector) crosses autograd; the other
    products are handed back through ``bundle.result``."""

    @staticmethod
    def forward(ctx, bundle, T0, T1, *disps):
        lib = C.lib()
        T = (_f32c(T0, "T[0]"), _f32c(T1, "T[1]"))
        disps = tuple(_f32c(d, "disp") for d in disps)
        grad_pose = bool(ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and not bundle.cfg.is_multi
        plan = _plan_for(bundle, bundle.cfg.flags(grad_pose=grad_pose), disps)
        sl = plan.take()
        plan._bind(sl.p, bundle, T, disps)
        with torch.cuda.device(bundle.tgt.device):
            C.check(lib.ppea_vsl_fused_forward(ctypes.byref(sl.p), ctypes.byref(sl.f), _stream()))
        sl.pending = True
        ctx.slot = sl
        ctx.keep = (bundle, T, disps)          # the backward reads the inputs again (recompute-free, but disp / colour / frames)
        ctx.grad_pose = grad_pose
        bundle.result = (sl.depth, sl.sel, sl.loss_px, sl.sums)
        return sl.losses

    @staticmethod
    def backward(ctx, grad_losses):
        lib = C.lib()
        sl = ctx.slot
        bundle = ctx.keep[0]
        with torch.cuda.device(bundle.tgt.device):
            if grad_losses is None:
                if sl.zero_gl is None:
                    sl.zero_gl = torch.zeros_like(sl.losses)
                grad_losses = sl.zero_gl
            if grad_losses.dtype != torch.float32 or not grad_losses.is_contiguous():
                grad_losses = grad_losses.contiguous().float()
            sl.g.grad_losses = grad_losses.data_ptr()
            C.check(lib.ppea_vsl_fused_backward(ctypes.byref(sl.p), ctypes.byref(sl.g), ctypes.byref(sl.f), _stream()))
        sl.pending = False
        gT = sl.gT if ctx.grad_pose else (None, None)
        return (None, gT[0], gT[1], *sl.grad_disp)


def view_synthesis_loss(disps, T, tgt, src, K, inv_K, colors, cfg: VslConfig, noise=None, cons_mask=None,
                        aug_mask=None, mono_depth=None) -> VslResult:
    """Fused view-synthesis loss over ``len(disps)`` pyramid scales.

    disps      list of (B,1,h_s,w_s) disparity maps, outputs[("disp", s)]            (differentiable)
    T          pair of (B,4,4) poses outputs[("cam_T_cam", 0, f)], f = frame_ids[1:]  (differentiable on the mono path)
    tgt, src   (B,3,H,W) target frame and the pair of source frames at source_scale 0
    K, inv_K   (B,4,4) intrinsics of source_scale 0
    colors     list of (B,3,h_s,w_s) target pyramid inputs[("color", 0, s)] for the smoothness term
    noise      list of (B,1,H,W) standard-normal draws (mono path with automasking; trainer.py:1086)
    cons_mask, aug_mask, mono_depth   multi-frame extras (trainer.py:1101-1141)
    """
    S = len(disps)
    if not 1 <= S <= C.MAX_SCALES:
        raise ValueError("1..%d scales per call" % C.MAX_SCALES)
    b = _Bundle()
    b.cfg = cfg
    b.tgt = _f32c(tgt, "tgt")
    b.src = (_f32c(src[0], "src[0]"), _f32c(src[1], "src[1]"))
    b.K, b.inv_K = _f32c(K, "K"), _f32c(inv_K, "inv_K")
    b.B, _, b.H, b.W = b.tgt.shape
    b.S = S
    b.colors = [_f32c(c, "color") for c in colors]
    b.noise = b.mono_depth = b.cons_mask = b.aug_mask = None
    if cfg.is_multi:
        if mono_depth is None:
            raise ValueError("is_multi needs mono_depth")
        b.mono_depth = [_f32c(m.detach(), "mono_depth") for m in mono_depth]
        if cfg.motion_mask:
            b.cons_mask = _f32c(cons_mask.detach(), "consistency_mask")
        if cfg.match_aug:
            b.aug_mask = _f32c(aug_mask.detach().reshape(-1)[:b.B], "augmentation_mask")
    elif cfg.automask and noise is not None:                 # None: opt.disable_automasking (no tie-break noise)
        b.noise = [_f32c(z, "noise") for z in noise]
    for s in range(S):
        if disps[s].shape[0] != b.B or tuple(b.colors[s].shape[-2:]) != tuple(disps[s].shape[-2:]):
            raise ValueError("scale %d: disp %s / color %s mismatch" % (s, tuple(disps[s].shape), tuple(b.colors[s].shape)))
    T0, T1 = T
    if cfg.is_multi:
        T0, T1 = T0.detach(), T1.detach()                     # trainer.py:900-902
    # The kernels index K / inv_K / T as ptr + 16 * b and every (B,*,H,W) map by b: shapes are checked here, and a
    # batch-1 matrix (the reference modules broadcast it through torch.matmul) is expanded to the batch.
    b.K, b.inv_K = _mat44(b.K, b.B, "K"), _mat44(b.inv_K, b.B, "inv_K")
    T0, T1 = _mat44(T0, b.B, "T[0]"), _mat44(T1, b.B, "T[1]")
    for name, t in (("src[0]", b.src[0]), ("src[1]", b.src[1])):
        if tuple(t.shape) != tuple(b.tgt.shape):
            raise ValueError("%s %s does not match the target frame %s" % (name, tuple(t.shape), tuple(b.tgt.shape)))
    full = (b.B, 1, b.H, b.W)
    for name, lst in (("noise", b.noise), ("mono_depth", b.mono_depth)):
        for t in (lst or ()):
            if tuple(t.shape) != full:
                raise ValueError("%s %s is not %s" % (name, tuple(t.shape), full))
    if b.cons_mask is not None and b.cons_mask.numel() != b.B * b.H * b.W:
        raise ValueError("consistency_mask %s is not (%d,%d,%d)" % (tuple(b.cons_mask.shape), b.B, b.H, b.W))
    if b.aug_mask is not None and b.aug_mask.numel() != b.B:
        raise ValueError("augmentation_mask needs one entry per batch item")
    needs_grad = torch.is_grad_enabled() and (T0.requires_grad or T1.requires_grad or any(d.requires_grad for d in disps))
    if cfg.plan_cache and cfg.use_fused(needs_grad):
        losses = _PlannedFusedLoss.apply(b, T0, T1, *disps)
        depth, sel, lp, sums = b.result
        return VslResult(losses=losses, depth=list(depth), sel=list(sel), loss_px=list(lp), sums=sums)
    outs = _FusedViewSynthLoss.apply(b, T0, T1, *disps)
    losses = outs[0]
    depth = list(outs[1:1 + S])
    sel = list(outs[1 + S:1 + 2 * S])
    sums = outs[1 + 2 * S]
    lp = list(outs[2 + 2 * S:]) if cfg.want_loss_px else [None] * S
    return VslResult(losses=losses, depth=depth, sel=sel, loss_px=lp, sums=sums)


# ---------------------------------------------------------------------------
# piecewise operators (the reference's nn.Module / function API)
# ---------------------------------------------------------------------------
def _no_grad_for(t, what):
    if t.requires_grad:
        raise NotImplementedError("ppea_depth_b200: gradient wrt %s is not implemented (it is data on the "
                                  "reference's loss path)" % what)


class _Ssim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        x, y = _f32c(x, "x"), _f32c(y, "y")
        B, Cn, H, W = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            C.check(C.lib().ppea_ssim_forward(C.ptr(x), C.ptr(y), C.ptr(out), B * Cn, H, W, _stream()))
        ctx.save_for_backward(x, y)
        return out

    @staticmethod
    def backward(ctx, go):
        x, y = ctx.saved_tensors
        B, Cn, H, W = x.shape
        go = go.contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            C.check(C.lib().ppea_ssim_backward(C.ptr(x), C.ptr(y), C.ptr(go), C.ptr(gx), C.ptr(gy), B * Cn, H, W, _stream()))
        return gx, gy


def ssim(x, y):
    """SSIM.forward, layers.py:243-257."""
    return _Ssim.apply(x, y)


class _Reprojection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, no_ssim):
        pred, target = _f32c(pred, "pred"), _f32c(target, "target")
        B, Cn, H, W = pred.shape
        if Cn != 3:
            raise ValueError("compute_reprojection_loss expects 3-channel images")
        out = torch.empty(B, 1, H, W, device=pred.device, dtype=torch.float32)
        with torch.cuda.device(pred.device):
            C.check(C.lib().ppea_reprojection_forward(C.ptr(pred), C.ptr(target), C.ptr(out), B, H, W, int(no_ssim), _stream()))
        ctx.save_for_backward(pred, target)
        ctx.no_ssim = int(no_ssim)
        return out

    @staticmethod
    def backward(ctx, go):
        pred, target = ctx.saved_tensors
        B, _, H, W = pred.shape
        go = go.contiguous().float()
        gp = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            C.check(C.lib().ppea_reprojection_backward(C.ptr(pred), C.ptr(target), C.ptr(go), C.ptr(gp), B, H, W, ctx.no_ssim, _stream()))
        return gp, None, None


def reprojection_loss(pred, target, no_ssim=False):
    """Trainer.compute_reprojection_loss, trainer.py:995-1007."""
    _no_grad_for(target, "the target image")
    return _Reprojection.apply(pred, target, no_ssim)


class _Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, inv_K, H, W):
        depth, inv_K = _f32c(depth, "depth"), _f32c(inv_K, "inv_K")
        B = depth.shape[0]
        if depth.numel() != B * H * W:
            raise ValueError("depth %s does not match (%d,1,%d,%d)" % (tuple(depth.shape), B, H, W))
        cam = torch.empty(B, 4, H * W, device=depth.device, dtype=torch.float32)
        with torch.cuda.device(depth.device):
            C.check(C.lib().ppea_backproject_forward(C.ptr(depth), C.ptr(inv_K), C.ptr(cam), B, H, W, _stream()))
        ctx.save_for_backward(inv_K)
        ctx.shape = (tuple(depth.shape), H, W)
        return cam

    @staticmethod
    def backward(ctx, gcam):
        (inv_K,) = ctx.saved_tensors
        shape, H, W = ctx.shape
        gcam = gcam.contiguous().float()
        gd = torch.empty(shape, device=gcam.device, dtype=torch.float32)
        with torch.cuda.device(gcam.device):
            C.check(C.lib().ppea_backproject_backward(C.ptr(gcam), C.ptr(inv_K), C.ptr(gd), shape[0], H, W, _stream()))
        return gd, None, None, None


def backproject_depth(depth, inv_K, height, width):
    """BackprojectDepth.forward, layers.py:163-168 -> (B,4,H*W)."""
    _no_grad_for(inv_K, "inv_K")
    return _Backproject.apply(depth, inv_K, height, width)


class _Project3D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, K, T, H, W, eps):
        points, K, T = _f32c(points, "points"), _f32c(K, "K"), _f32c(T, "T")
        B = points.shape[0]
        pix = torch.empty(B, H, W, 2, device=points.device, dtype=torch.float32)
        z = torch.empty(B, 1, H, W, device=points.device, dtype=torch.float32)
        with torch.cuda.device(points.device):
            C.check(C.lib().ppea_project3d_forward(C.ptr(points), C.ptr(K), C.ptr(T), C.ptr(pix), C.ptr(z), B, H, W, eps, _stream()))
        ctx.save_for_backward(points, K, T)
        ctx.dims = (H, W, eps)
        return pix, z

    @staticmethod
    def backward(ctx, gpix, gz):
        points, K, T = ctx.saved_tensors
        H, W, eps = ctx.dims
        B = points.shape[0]
        lib = C.lib()
        gpix = gpix.contiguous().float() if gpix is not None else torch.zeros(B, H, W, 2, device=points.device)
        gz = gz.contiguous().float() if gz is not None else None
        gpts = torch.empty_like(points)
        gT = torch.empty(B, 4, 4, device=points.device, dtype=torch.float32)
        part = torch.empty(max(lib.ppea_project3d_partials_bytes(B, H, W) // 4, 1), device=points.device, dtype=torch.float32)
        with torch.cuda.device(points.device):
            C.check(lib.ppea_project3d_backward(C.ptr(points), C.ptr(K), C.ptr(T), C.ptr(gpix), C.ptr(gz), C.ptr(gpts),
                                               C.ptr(gT), C.ptr(part), B, H, W, eps, _stream()))
        return gpts, None, gT, None, None, None


def project_3d(points, K, T, height, width, eps=1e-7):
    """Project3D.forward, layers.py:184-199 -> (pix (B,H,W,2), z (B,1,H,W))."""
    _no_grad_for(K, "K")
    return _Project3D.apply(points, K, T, height, width, float(eps))


class _GridSampleBorder(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, grid):
        src, grid = _f32c(src, "src"), _f32c(grid, "grid")
        B, Cn, H, W = src.shape
        oh, ow = grid.shape[1], grid.shape[2]
        out = torch.empty(B, Cn, oh, ow, device=src.device, dtype=torch.float32)
        with torch.cuda.device(src.device):
            C.check(C.lib().ppea_warp_forward(C.ptr(src), C.ptr(grid), C.ptr(out), B, Cn, H, W, oh, ow, _stream()))
        ctx.save_for_backward(src, grid)
        return out

    @staticmethod
    def backward(ctx, go):
        src, grid = ctx.saved_tensors
        B, Cn, H, W = src.shape
        oh, ow = grid.shape[1], grid.shape[2]
        go = go.contiguous().float()
        gg = torch.empty_like(grid)
        with torch.cuda.device(src.device):
            C.check(C.lib().ppea_warp_backward(C.ptr(src), C.ptr(grid), C.ptr(go), C.ptr(gg), B, Cn, H, W, oh, ow, _stream()))
        return None, gg


def grid_sample_border(src, grid):
    """F.grid_sample(src, grid, padding_mode="border", align_corners=True) as called at trainer.py:911-914."""
    _no_grad_for(src, "the source image")
    return _GridSampleBorder.apply(src, grid)


class _Smooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp, img):
        disp, img = _f32c(disp, "disp"), _f32c(img, "img")
        B, _, H, W = disp.shape
        lib = C.lib()
        out = torch.empty((), device=disp.device, dtype=torch.float32)
        ws = torch.empty(lib.ppea_smooth_workspace_bytes(B, H, W) // 4, device=disp.device, dtype=torch.float32)
        with torch.cuda.device(disp.device):
            C.check(lib.ppea_smooth_forward(C.ptr(disp), C.ptr(img), C.ptr(out), C.ptr(ws), B, H, W, _stream()))
        ctx.save_for_backward(disp, img)
        return out

    @staticmethod
    def backward(ctx, go):
        disp, img = ctx.saved_tensors
        B, _, H, W = disp.shape
        go = go.contiguous().float()
        gd = torch.empty_like(disp)
        with torch.cuda.device(disp.device):
            C.check(C.lib().ppea_smooth_backward(C.ptr(disp), C.ptr(img), C.ptr(go), C.ptr(gd), B, H, W, _stream()))
        return gd, None


def smooth_loss(disp, img):
    """get_smooth_loss, layers.py:210-223 (1-channel disp, 3-channel image)."""
    _no_grad_for(img, "the image")
    if disp.shape[1] != 1 or img.shape[1] != 3:
        raise ValueError("get_smooth_loss expects (B,1,H,W) disp and (B,3,H,W) image")
    return _Smooth.apply(disp, img)
