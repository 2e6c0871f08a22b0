"""Decoder tail: the disparity head of DepthDecoderV2 as one launch each way.

Reference: ``self.outputs[("disp", 0)] = self.sigmoid(self.disp_convs[0](x))`` (networks/depth_decoder_v2.py:123-129, :239),
``Conv3x3`` = ``nn.ReflectionPad2d(1)`` + ``nn.Conv2d(C, 1, 3)`` (layers.py:119-135), then ``disp_to_depth``
(layers.py:14-23) on the loss side (trainer.py:888).

  disp_head(x, weight, bias)               sigmoid(conv3x3_reflect(x)) with autograd (grad_x, grad_weight, grad_bias)
  disp_head_with_depth(x, w, b, lo, hi)    the same plus depth = disp_to_depth(disp, lo, hi)[1] from the same launch (no grad
                                           through depth: the loss differentiates disp itself)
  FusedDispHead(conv3x3)                   module that wraps the reference's Conv3x3 (same parameters, same state_dict keys)
  install_decoder(decoder)                 rebinds a built DepthDecoderV2: disp_convs[0] -> FusedDispHead, sigmoid -> identity

There is no CPU path: CPU tensors raise.  The dense ConvBlocks of the decoder stay cuDNN.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _cabi as C


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(x, weight, bias):
    if not (x.is_cuda and weight.is_cuda and bias.is_cuda):
        raise RuntimeError("ppea_depth_b200 has no CPU path: disp_head needs CUDA tensors")
    if x.dim() != 4 or tuple(weight.shape) != (1, x.shape[1], 3, 3) or bias.numel() != 1:
        raise ValueError("disp_head: x (B,C,H,W), weight (1,C,3,3), bias (1) expected, got %s %s %s" %
                         (tuple(x.shape), tuple(weight.shape), tuple(bias.shape)))
    if x.shape[1] > 256 or x.shape[2] < 2 or x.shape[3] < 2:
        raise ValueError("disp_head: C <= 256 and H, W >= 2 (ReflectionPad2d(1)) required")


def _forward(x, weight, bias, depth_range):
    _check(x, weight, bias)
    x = x.detach().contiguous().float()
    w = weight.detach().contiguous().float()
    b = bias.detach().contiguous().float()
    B, Cn, H, W = x.shape
    with torch.cuda.device(x.device):
        disp = torch.empty(B, 1, H, W, device=x.device, dtype=torch.float32)
        depth = torch.empty_like(disp) if depth_range is not None else None
        lo, hi = depth_range if depth_range is not None else (0.0, 0.0)
        C.check(C.lib().ppea_disp_head_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), disp.data_ptr(),
                                               depth.data_ptr() if depth is not None else None, B, Cn, H, W, float(lo), float(hi), _stream()))
    return x, w, disp, depth


class _DispHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, depth_range):
        xc, wc, disp, depth = _forward(x, weight, bias, depth_range)
        ctx.save_for_backward(xc, wc, disp)
        if depth is None:
            return disp
        ctx.mark_non_differentiable(depth)
        return disp, depth

    @staticmethod
    def backward(ctx, grad_disp, *unused):
        x, w, disp = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        B, Cn, H, W = x.shape
        g = grad_disp.contiguous().float()
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x) if need_x else None
            gw = torch.empty_like(w) if (need_w or need_b) else None
            gb = torch.empty(1, device=x.device, dtype=torch.float32) if (need_w or need_b) else None
            ws = None
            if gw is not None:
                ws = torch.empty(C.lib().ppea_disp_head_workspace_bytes(B, Cn, H, W) // 4, device=x.device, dtype=torch.float32)
            C.check(C.lib().ppea_disp_head_backward(x.data_ptr(), w.data_ptr(), disp.data_ptr(), g.data_ptr(),
                                                    gx.data_ptr() if gx is not None else None, gw.data_ptr() if gw is not None else None,
                                                    gb.data_ptr() if gb is not None else None, ws.data_ptr() if ws is not None else None,
                                                    B, Cn, H, W, _stream()))
        return gx, (gw if need_w else None), (gb if need_b else None), None


def disp_head(x, weight, bias):
    """sigmoid(Conv2d(C,1,3)(ReflectionPad2d(1)(x))) -- depth_decoder_v2.py:239 -- as one launch, differentiable."""
    return _DispHead.apply(x, weight, bias.reshape(1), None)


def disp_head_with_depth(x, weight, bias, min_depth, max_depth):
    """(disp, depth): depth = disp_to_depth(disp, min_depth, max_depth)[1] (layers.py:14-23) from the same launch."""
    return _DispHead.apply(x, weight, bias.reshape(1), (float(min_depth), float(max_depth)))


class FusedDispHead(nn.Module):
    """Takes the place of ``disp_convs[0]`` (a reference ``Conv3x3``) inside DepthDecoderV2 and already applies the sigmoid.
    The wrapped module keeps its parameters under the same names (``pad``, ``conv.weight``, ``conv.bias``), so checkpoints
    of the reference load unchanged."""

    def __init__(self, conv3x3):
        super().__init__()
        if not isinstance(getattr(conv3x3, "pad", None), nn.ReflectionPad2d):
            raise ValueError("FusedDispHead: the wrapped Conv3x3 must use reflection padding (layers.py:126-127)")
        c = conv3x3.conv
        if c.out_channels != 1 or tuple(c.kernel_size) != (3, 3) or tuple(c.stride) != (1, 1) or tuple(c.padding) != (0, 0) or c.bias is None:
            raise ValueError("FusedDispHead: Conv2d(C, 1, 3) with bias expected (depth_decoder_v2.py:125)")
        self.pad = conv3x3.pad
        self.conv = c

    def forward(self, x):
        return disp_head(x, self.conv.weight, self.conv.bias)


def install_decoder(decoder):
    """Rebind a built DepthDecoderV2: its last two statements (conv + sigmoid, depth_decoder_v2.py:239) become one launch.
    ``decoder.disp_convs[0]`` must be the reference's Conv3x3; ``decoder.sigmoid`` becomes the identity (the head applies it)."""
    head = decoder.disp_convs[0]
    if not isinstance(head, FusedDispHead):
        decoder.disp_convs[0] = FusedDispHead(head)
        decoder.sigmoid = nn.Identity()
    return decoder
