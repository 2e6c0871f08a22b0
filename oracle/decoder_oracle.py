"""ORACLE (test infrastructure, not product code): the disparity head of DepthDecoderV2 on the CPU.

Restates, in the ATen ops the reference calls,
    self.outputs[("disp", 0)] = self.sigmoid(self.disp_convs[0](x))          networks/depth_decoder_v2.py:239
    Conv3x3.forward: self.conv(self.pad(x)), ReflectionPad2d(1) + Conv2d(C,1,3)   layers.py:119-135
    disp_to_depth                                                              layers.py:14-23
Pinned to tests/golden/decoder_*.pt, which `oracle/make_golden_decoder.py` produced by running the reference's own
DepthDecoderV2 / Conv3x3 (and autograd for the gradients), and to the reference live where it is mounted
(tests/test_decoder.py).  Only tests/, __graft_entry__.smoke() and bench legs that say so may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def disp_head(x, weight, bias, dtype=torch.float32):
    """sigmoid(conv2d(reflection_pad(x, 1), weight, bias)); x (B,C,H,W), weight (1,C,3,3), bias (1)."""
    x, weight, bias = x.to(dtype), weight.to(dtype), bias.to(dtype)
    out = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), weight, bias.reshape(1))     # layers.py:131-134
    return torch.sigmoid(out)                                                              # depth_decoder_v2.py:239


def disp_to_depth(disp, min_depth, max_depth):
    min_disp = 1 / max_depth                                                               # layers.py:19-22
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    return scaled_disp, 1 / scaled_disp


def disp_head_grads(x, weight, bias, grad_disp, dtype=torch.float32):
    """(grad_x, grad_weight, grad_bias) of sum(disp * grad_disp) by autograd over the restatement."""
    x = x.detach().to(dtype).requires_grad_(True)
    w = weight.detach().to(dtype).requires_grad_(True)
    b = bias.detach().to(dtype).reshape(1).requires_grad_(True)
    (disp_head(x, w, b, dtype) * grad_disp.to(dtype)).sum().backward()
    return x.grad, w.grad, b.grad


def synthetic_head_case(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    weight = torch.randn(1, C, 3, 3, generator=g) * (1.5 / (3.0 * C ** 0.5))
    bias = torch.randn(1, generator=g) * 0.1
    grad = torch.randn(B, 1, H, W, generator=g)
    return x, weight, bias, grad


def run_reference_conv3x3(x, weight, bias, grad_disp):
    """The reference's own Conv3x3 (ppeadepth/layers.py) + nn.Sigmoid + autograd."""
    from . import ref_import
    ref_import.load_reference()
    from ppeadepth.layers import Conv3x3
    conv = Conv3x3(x.shape[1], 1)
    with torch.no_grad():
        conv.conv.weight.copy_(weight)
        conv.conv.bias.copy_(bias.reshape(1))
    xr = x.clone().requires_grad_(True)
    disp = torch.nn.Sigmoid()(conv(xr))
    (disp * grad_disp).sum().backward()
    return disp.detach(), xr.grad, conv.conv.weight.grad, conv.conv.bias.grad


def run_reference_decoder(feats, seed, grad_seed):
    """The reference's own DepthDecoderV2 (dc=False) on the encoder features `feats`: returns the input of its disparity head
    (captured by a hook), the head's parameters, ("disp", 0) and the gradients autograd sends to the head's input / parameters."""
    import numpy as np
    from . import ref_import
    ref_import.load_reference()
    from ppeadepth.networks.depth_decoder_v2 import DepthDecoderV2
    torch.manual_seed(seed)
    dec = DepthDecoderV2(np.array([f.shape[1] for f in feats]))
    grabbed = {}

    def hook(mod, inp):
        inp[0].retain_grad()
        grabbed["x"] = inp[0]
    dec.disp_convs[0].register_forward_pre_hook(hook)
    disp = dec([f.clone() for f in feats])[("disp", 0)]
    grad = torch.randn(disp.shape, generator=torch.Generator().manual_seed(grad_seed))
    (disp * grad).sum().backward()
    head = dec.disp_convs[0].conv
    return dict(x=grabbed["x"].detach().clone(), weight=head.weight.detach().clone(), bias=head.bias.detach().clone(), grad_disp=grad,
                disp=disp.detach().clone(), grad_x=grabbed["x"].grad.clone(), grad_weight=head.weight.grad.clone(),
                grad_bias=head.bias.grad.clone())
