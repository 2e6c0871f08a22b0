"""uint8 colour frames on the loss path (input format).

The loss reads the un-augmented ``("color", f, s)`` frames, which the reference builds on the CPU as
torchvision ``ToTensor`` of a uint8 PIL image (/root/reference/ppeadepth/datasets/mono_dataset.py:62, :106):
``float32(k) / 255``.  A dataset that hands the planar uint8 frames over instead moves a quarter of the
bytes across PCIe; `images_to_float` expands them on the device (`ppea_images_u8_to_f32`, one correctly
rounded division per value -- bit-identical to ToTensor), and the loss methods installed by
`ppea_depth_b200.install` accept uint8 CUDA tensors wherever the reference passes float ones.
There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C


def images_to_float(img: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """uint8 CUDA tensor (any shape, planar like the float frames) -> float32 tensor ``img / 255`` of the same
    shape; float32 tensors pass through untouched.  `out`: optional preallocated contiguous float32 CUDA tensor with as
    many elements (an input pipeline's device staging buffer)."""
    if img.dtype == torch.float32:
        return img
    if img.dtype != torch.uint8:
        raise TypeError("colour frames are float32 or uint8, got %s" % img.dtype)
    if not img.is_cuda:
        raise RuntimeError("ppea_depth_b200 has no CPU path: move the uint8 frames to the GPU first")
    src = img.contiguous()
    with torch.cuda.device(src.device):
        if out is None:
            out = torch.empty(src.shape, device=src.device, dtype=torch.float32)
        elif out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or out.numel() != src.numel():
            raise ValueError("images_to_float: `out` must be a contiguous float32 CUDA tensor with %d elements" % src.numel())
        C.check(C.lib().ppea_images_u8_to_f32(src.data_ptr(), out.data_ptr(), ctypes.c_size_t(src.numel()),
                                              torch.cuda.current_stream().cuda_stream))
    return out


class FrameCache:
    """Expands every distinct uint8 frame of one step once (the target frame is read as the warp target and as the
    smoothness image of scale 0)."""

    def __init__(self):
        self._done = {}

    def __call__(self, img):
        if img.dtype == torch.float32:
            return img
        key = (img.data_ptr(), tuple(img.shape))
        if key not in self._done:
            self._done[key] = images_to_float(img)
        return self._done[key]
