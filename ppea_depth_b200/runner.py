"""Fixed-shape execution plan for the fused loss: buffers, parameter block and (optionally) a
CUDA graph are built once and replayed every step.

The training step has static shapes (the reference even bakes batch_size into its modules,
/root/reference/ppeadepth/layers.py:142-161, and drops ragged batches, trainer.py:215-218), so
everything the autograd wrapper does per call -- output allocation, filling the C parameter
struct, eight kernel launches -- can be hoisted: ``FusedPlan.capture()`` records
forward + backward into one CUDA graph whose replay costs a single launch on the host.
Inputs are read from the tensors bound at construction; refresh them with ``copy_`` (their
addresses are baked into the graph).
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Optional, Sequence

import torch

from . import _cabi as C
from .functional import VslConfig, _f32c


class FusedPlan:
    def __init__(self, cfg: VslConfig, disps: Sequence[torch.Tensor], T: Sequence[torch.Tensor], tgt: torch.Tensor,
                 src: Sequence[torch.Tensor], K: torch.Tensor, inv_K: torch.Tensor, colors: Sequence[torch.Tensor],
                 noise: Optional[Sequence[torch.Tensor]] = None, cons_mask: Optional[torch.Tensor] = None,
                 aug_mask: Optional[torch.Tensor] = None, mono_depth: Optional[Sequence[torch.Tensor]] = None,
                 grad_pose: Optional[bool] = None, fused: Optional[bool] = None):
        lib = C.lib()
        self.cfg = cfg
        self.tgt = _f32c(tgt, "tgt")
        self.device = self.tgt.device
        self.B, _, self.H, self.W = self.tgt.shape
        self.S = len(disps)
        B, H, W, S, dev = self.B, self.H, self.W, self.S, self.device
        self.disps = [_f32c(d.detach(), "disp") for d in disps]
        self.T = [_f32c(t.detach(), "T") for t in T]
        self.src = [_f32c(s, "src") for s in src]
        self.K, self.inv_K = _f32c(K, "K"), _f32c(inv_K, "inv_K")
        self.colors = [_f32c(c, "color") for c in colors]
        self.noise = [_f32c(z, "noise") for z in noise] if (noise is not None and not cfg.is_multi) else None
        self.cons_mask = _f32c(cons_mask, "cons_mask") if (cfg.is_multi and cfg.motion_mask) else None
        self.aug_mask = _f32c(aug_mask.reshape(-1)[:B], "aug_mask") if (cfg.is_multi and cfg.match_aug) else None
        self.mono_depth = [_f32c(m, "mono_depth") for m in mono_depth] if cfg.is_multi else None
        self.grad_pose = (not cfg.is_multi) if grad_pose is None else bool(grad_pose)
        # single-launch training step (vsl_fused.cu) unless the forward + backward kernel pair is asked for
        self.fused = True if fused is None else bool(fused)
        if fused == "tiles":          # the shared-memory tile kernel of the fused step instead of the streaming kernel
            self.cfg = cfg = dataclasses.replace(cfg, fused="tiles")
        self.tiles = self.fused and cfg.fused == "tiles"
        f32 = dict(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            self.depth = [torch.empty(B, 1, H, W, **f32) for _ in range(S)]
            self.sel = [torch.empty(B, H, W, device=dev, dtype=torch.uint8) for _ in range(S)]
            self.loss_px = [torch.empty(B, 1, H, W, **f32) if cfg.want_loss_px else None for _ in range(S)]
            self.sums = torch.empty(lib.ppea_vsl_sums_floats(B, S), **f32)
            self.losses = torch.empty(1 + C.LOSSES_PER_SCALE * S, **f32)
            self.grad_losses = torch.zeros(1 + C.LOSSES_PER_SCALE * S, **f32)
            self.grad_losses[0] = 1.0                     # d loss / d loss
            self.grad_disp = [torch.empty_like(d) for d in self.disps]
            self.grad_T = [torch.empty(B, 4, 4, **f32) for _ in range(2)] if self.grad_pose else None
            self.flags_fwd = cfg.flags(grad_pose=False)
            self.flags_bwd = cfg.flags(grad_pose=self.grad_pose)
            if self.fused:
                # the plan owns the fused workspace: zero-filled once here, kept clean by every backward
                self.flags_bwd |= C.F_RAW_PREZEROED
                self.flags_fwd = self.flags_bwd
            elif not cfg.deterministic:   # the forward zero-fills the persistent grad buffers on the fly
                self.flags_bwd |= C.F_GRAD_PREZEROED
            self.ws_fwd = torch.empty(max(lib.ppea_vsl_workspace_bytes(B, H, W, S) // 4, 4), **f32)
            self.ws_bwd = torch.empty(max(lib.ppea_vsl_backward_workspace_bytes(B, H, W, S, self.flags_bwd) // 4, 4), **f32)
        self._p_fwd = self._params(self.flags_fwd)
        self._p_bwd = self._params(self.flags_bwd)
        g = C.PpeaVslGrads()
        g.struct_size = ctypes.sizeof(C.PpeaVslGrads)
        g.grad_losses = self.grad_losses.data_ptr()
        if self.grad_pose:
            g.grad_T[0], g.grad_T[1] = self.grad_T[0].data_ptr(), self.grad_T[1].data_ptr()
        g.workspace = self.ws_bwd.data_ptr()
        g.workspace_bytes = self.ws_bwd.numel() * 4
        self._g = g
        self._lib = lib
        self.graph = None
        self._f = None
        self._raw_pending = False
        if self.fused:
            with torch.cuda.device(dev):
                self.ws_fused = torch.zeros(max(lib.ppea_vsl_fused_workspace_bytes(ctypes.byref(self._p_fwd)) // 4, 4), **f32)
            f = C.PpeaVslFused()
            f.struct_size = ctypes.sizeof(C.PpeaVslFused)
            f.workspace = self.ws_fused.data_ptr()
            f.workspace_bytes = self.ws_fused.numel() * 4
            self._f = f

    # kernels launched by one forward / backward call (for bench.py's gpu_launches)
    @property
    def launches_forward(self):
        if self.fused and not self.tiles:
            return 4        # preparation (packed sources, identity loss, target window sums), streaming step, smoothness (in its shadow), finish
        return 2            # fused forward / fused step (tiles + smoothness CTAs), finish

    @property
    def launches_backward(self):
        if self.fused:
            return 1     # gradient finish (combines the un-normalised fields; its last B CTAs weight the pose sums)
        n = 1 + (1 if self.grad_pose else 0)     # fused backward (tiles + smoothness CTAs), pose finish
        if self.cfg.deterministic:               # + smoothness backward + one upsample gather per coarse scale
            n += 1 + sum(1 for d in self.disps if tuple(d.shape[-2:]) != (self.H, self.W))
        return n                                 # (the non-deterministic path also enqueues S memset nodes)

    def _params(self, flags):
        cfg = self.cfg
        p = C.PpeaVslParams()
        p.struct_size = ctypes.sizeof(C.PpeaVslParams)
        p.flags = flags
        p.batch, p.height, p.width = self.B, self.H, self.W
        p.num_scales = self.S
        p.first_scale = cfg.first_scale
        p.total_scales = cfg.total_scales if cfg.total_scales is not None else self.S
        lo = 1.0 / cfg.max_depth
        p.disp_lo, p.disp_range, p.eps = lo, 1.0 / cfg.min_depth - lo, cfg.eps
        p.disparity_smoothness = cfg.disparity_smoothness
        p.tgt = self.tgt.data_ptr()
        p.src[0], p.src[1] = self.src[0].data_ptr(), self.src[1].data_ptr()
        p.K, p.inv_K = self.K.data_ptr(), self.inv_K.data_ptr()
        p.T[0], p.T[1] = self.T[0].data_ptr(), self.T[1].data_ptr()
        p.cons_mask = self.cons_mask.data_ptr() if self.cons_mask is not None else None
        p.aug_mask = self.aug_mask.data_ptr() if self.aug_mask is not None else None
        for s in range(self.S):
            sc = p.scales[s]
            sc.disp_h, sc.disp_w = self.disps[s].shape[-2], self.disps[s].shape[-1]
            sc.disp = self.disps[s].data_ptr()
            sc.color = self.colors[s].data_ptr()
            sc.noise = self.noise[s].data_ptr() if self.noise is not None else None
            sc.mono_depth = self.mono_depth[s].data_ptr() if self.mono_depth is not None else None
            sc.depth = self.depth[s].data_ptr()
            sc.loss_px = self.loss_px[s].data_ptr() if self.loss_px[s] is not None else None
            sc.sel = self.sel[s].data_ptr()
            sc.grad_disp = self.grad_disp[s].data_ptr()
        p.sums = self.sums.data_ptr()
        p.losses = self.losses.data_ptr()
        p.workspace = self.ws_fwd.data_ptr()
        p.workspace_bytes = self.ws_fwd.numel() * 4
        return p

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward(self):
        if self.fused:
            if self._raw_pending:         # a forward without its backward left the raw coarse-scale fields dirty
                self.ws_fused.zero_()
            self._raw_pending = True
            C.check(self._lib.ppea_vsl_fused_forward(ctypes.byref(self._p_fwd), ctypes.byref(self._f), self._stream()))
        else:
            C.check(self._lib.ppea_vsl_forward(ctypes.byref(self._p_fwd), self._stream()))
        return self.losses

    def backward(self):
        if self.fused:
            self._raw_pending = False
            C.check(self._lib.ppea_vsl_fused_backward(ctypes.byref(self._p_bwd), ctypes.byref(self._g), ctypes.byref(self._f),
                                                      self._stream()))
        else:
            C.check(self._lib.ppea_vsl_backward(ctypes.byref(self._p_bwd), ctypes.byref(self._g), self._stream()))
        return self.grad_disp, self.grad_T

    def step(self):
        self.forward()
        self.backward()

    def capture(self, backward=True):
        """Records forward (+ backward) into a CUDA graph; ``replay()`` afterwards."""
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):         # warm-up outside capture: function attributes, lazy module load
                self.forward()
                if backward:
                    self.backward()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.forward()
                if backward:
                    self.backward()
        self.graph = graph
        return graph

    def replay(self):
        self.graph.replay()

    # ---- per-stage device timing (bench.py's roofline leg) -------------------
    FWD_STAGES = ("fwd_unused0", "vsl_forward_kernel", "fwd_unused1", "finish")
    BWD_STAGES = ("grad_init", "vsl_backward_kernel", "upsample_gather", "pose_finish")
    FUSED_FWD_STAGES = ("grad_raw_zero", "vsl_fused_kernel", "fwd_unused1", "finish")
    STREAM_FWD_STAGES = ("grad_raw_zero", "vsl_prep_kernel", "vsl_stream_kernel", "finish")
    FUSED_BWD_STAGES = ("bwd_unused0", "vsl_grad_finish_kernel", "bwd_unused1", "bwd_unused2")

    def enable_trace(self):
        """Asks the library to record a CUDA event before/after every stage of forward and
        backward (include/ppea_vsl.h: trace_events).  Affects eager calls made after this."""
        lib = self._lib
        self._ev = []
        for p in (self._p_fwd, self._p_bwd):
            arr = (ctypes.c_void_p * C.TRACE_EVENTS)()
            for i in range(C.TRACE_EVENTS):
                arr[i] = lib.ppea_event_create()
                if not arr[i]:
                    raise RuntimeError("cudaEventCreate failed")
            p.trace_events = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
            self._ev.append(arr)

    def disable_trace(self):
        for p, arr in zip((self._p_fwd, self._p_bwd), getattr(self, "_ev", [])):
            p.trace_events = None
            for i in range(C.TRACE_EVENTS):
                self._lib.ppea_event_destroy(arr[i])
        self._ev = []

    def trace_ms(self):
        """Stage durations (ms) of the most recent traced forward + backward."""
        out = {}
        ms = ctypes.c_float()
        stages = (self.FUSED_FWD_STAGES, self.FUSED_BWD_STAGES) if self.fused else (self.FWD_STAGES, self.BWD_STAGES)
        if self.fused and not self.tiles:
            stages = (self.STREAM_FWD_STAGES, self.FUSED_BWD_STAGES)
        for arr, names in zip(self._ev, stages):
            for i, name in enumerate(names):
                C.check(self._lib.ppea_event_elapsed_ms(arr[i], arr[i + 1], ctypes.byref(ms)))
                out[name] = ms.value
        return out
