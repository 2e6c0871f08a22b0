"""Input format on the device (remainder): the dataset's LANCZOS pyramid and packed RGBx frames.

Reference: ``MonoDataset.preprocess`` (datasets/mono_dataset.py:96-112) -- every colour frame is resized on the CPU workers with
``transforms.Resize((height // 2**i, width // 2**i), interpolation=Image.LANCZOS)`` (:79-85), scale ``i`` from scale ``i - 1``
(scale 0 from the raw frame, key ``-1``), then ``ToTensor``.

  resize_lanczos_u8(frames_u8, (h, w))   PIL's 8-bit LANCZOS resize of uint8 CUDA planes (N,C,H,W), bit-identical to PIL
  ImagePyramid(height, width, num_scales)(raw_u8)   the dataset's chain: {scale: (N,3,h_s,w_s) uint8}; ``as_float=True`` also
                                         applies ToTensor's /255 (images_to_float), i.e. the tensors the reference's loader yields
  pack_rgbx(frames_u8)                   (N,3,H,W) planar or (N,H,W,3) interleaved uint8 -> (N,H,W) int32 words r | g<<8 | b<<16

There is no CPU path: CPU tensors raise (the tap tables are built on the host by the library, as PIL builds them).
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi as C
from .images import images_to_float

_TABLES = {}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def lanczos_table(in_size, out_size, device):
    """(bounds (out,2) int32, coeffs (out,ksize) int32, ksize) of one resize direction on `device` (cached)."""
    key = (int(in_size), int(out_size), str(device))
    if key not in _TABLES:
        ksize = C.lib().ppea_lanczos_ksize(int(in_size), int(out_size))
        if ksize <= 0:
            raise ValueError("lanczos_table: bad sizes %s -> %s" % (in_size, out_size))
        bounds = torch.empty(out_size, 2, dtype=torch.int32)
        coeffs = torch.empty(out_size, ksize, dtype=torch.int32)
        got = C.lib().ppea_lanczos_table(int(in_size), int(out_size), bounds.data_ptr(), coeffs.data_ptr())
        if got != ksize:
            C.check(got if got < 0 else -1)
        _TABLES[key] = (bounds.to(device), coeffs.to(device), ksize)
    return _TABLES[key]


def resize_lanczos_u8(frames, size):
    """PIL ``Image.resize(size[::-1], LANCZOS)`` of every plane of `frames` (..., H, W) uint8 CUDA -> (..., size[0], size[1])."""
    if not frames.is_cuda or frames.dtype != torch.uint8:
        raise RuntimeError("resize_lanczos_u8: uint8 CUDA planes expected (there is no CPU path)")
    src = frames.contiguous()
    in_h, in_w = src.shape[-2:]
    out_h, out_w = int(size[0]), int(size[1])
    n_planes = src.numel() // (in_h * in_w)
    with torch.cuda.device(src.device):
        dst = torch.empty(*src.shape[:-2], out_h, out_w, device=src.device, dtype=torch.uint8)
        bx, cx, kx = lanczos_table(in_w, out_w, src.device) if out_w != in_w else (None, None, 0)
        by, cy, ky = lanczos_table(in_h, out_h, src.device) if out_h != in_h else (None, None, 0)
        tmp = torch.empty(n_planes * in_h * out_w, device=src.device, dtype=torch.uint8) if (kx and ky) else None
        p = lambda t: t.data_ptr() if t is not None else None
        C.check(C.lib().ppea_resize_lanczos_u8(src.data_ptr(), dst.data_ptr(), p(tmp), ctypes.c_size_t(n_planes), in_h, in_w, out_h, out_w,
                                               p(bx), p(cx), kx, p(by), p(cy), ky, _stream()))
    return dst


class ImagePyramid:
    """The resize chain of MonoDataset (mono_dataset.py:79-85, :101-104) on the device."""

    def __init__(self, height, width, num_scales=4):
        self.sizes = [(height // (2 ** i), width // (2 ** i)) for i in range(num_scales)]

    def __call__(self, raw_u8, as_float=False):
        out = {}
        cur = raw_u8
        for i, size in enumerate(self.sizes):
            cur = resize_lanczos_u8(cur, size)            # inputs[(n, im, i)] = self.resize[i](inputs[(n, im, i - 1)])
            out[i] = images_to_float(cur) if as_float else cur
        return out


def pack_rgbx(frames):
    """(N,3,H,W) planar or (N,H,W,3) interleaved uint8 CUDA frames -> (N,H,W) int32 words r | g << 8 | b << 16."""
    if not frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4:
        raise RuntimeError("pack_rgbx: 4-D uint8 CUDA frames expected (there is no CPU path)")
    src = frames.contiguous()
    interleaved = src.shape[-1] == 3 and src.shape[1] != 3
    if not interleaved and src.shape[1] != 3:
        raise ValueError("pack_rgbx: (N,3,H,W) or (N,H,W,3) expected, got %s" % (tuple(src.shape),))
    N = src.shape[0]
    H, W = (src.shape[1], src.shape[2]) if interleaved else (src.shape[2], src.shape[3])
    with torch.cuda.device(src.device):
        dst = torch.empty(N, H, W, device=src.device, dtype=torch.int32)
        C.check(C.lib().ppea_pack_rgbx_u8(src.data_ptr(), dst.data_ptr(), ctypes.c_size_t(N), H, W, int(interleaved), _stream()))
    return dst
