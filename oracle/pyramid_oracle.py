"""ORACLE (test infrastructure, not product code): PIL's 8-bit LANCZOS resize and the dataset's pyramid chain on the CPU.

Restates libImaging/Resample.c of Pillow (the algorithm behind `transforms.Resize(size, interpolation=Image.LANCZOS)` on a PIL
image, which is what MonoDataset.preprocess calls, datasets/mono_dataset.py:79-85, :96-112) in numpy integer arithmetic:
precompute_coeffs (window, double-precision Lanczos-3 weights normalised to 1), normalize_coeffs_8bpc (22-bit fixed point, round
half away from zero), ImagingResampleHorizontal_8bpc / Vertical_8bpc (int32 accumulation from 2^21, >> 22, clip to 0..255),
horizontal pass first.  Pillow is a third-party dependency of the reference (requirements: pillow); the restatement is pinned
(i) to Pillow itself, live, wherever it is importable (tests/test_pyramid.py) and (ii) to tests/golden/pyramid_*.pt, produced by
`oracle/make_golden_pyramid.py` running the reference's own MonoDataset.preprocess.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3)
    return 0.0


def lanczos_table(in_size, out_size):
    """(bounds (out,2) int32, coeffs (out,ksize) int32): Resample.c precompute_coeffs + normalize_coeffs_8bpc."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coeffs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)            # (C truncation; the operand is > -1 whenever it is negative... clamp)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        for x, w in enumerate(k):
            coeffs[xx, x] = int(-0.5 + w * (1 << PRECISION_BITS)) if w < 0 else int(0.5 + w * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coeffs


def _pass(img, bounds, coeffs, axis):
    """One resampling pass of uint8 planes (..., H, W) along `axis` (-1 horizontal, -2 vertical)."""
    src = np.moveaxis(img.astype(np.int64), axis, -1)
    out = np.empty(src.shape[:-1] + (bounds.shape[0],), np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, xmax = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + (src[..., xmin:xmin + xmax] * coeffs[xx, :xmax].astype(np.int64)).sum(-1)
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, -1, axis)


def resize_lanczos_u8(img, size):
    """img (..., H, W) uint8 numpy -> (..., size[0], size[1]): ImagingResample, horizontal pass first, skipped when unchanged."""
    out_h, out_w = size
    cur = np.asarray(img)
    if out_w != cur.shape[-1]:
        cur = _pass(cur, *lanczos_table(cur.shape[-1], out_w), axis=-1)
    if out_h != cur.shape[-2]:
        cur = _pass(cur, *lanczos_table(cur.shape[-2], out_h), axis=-2)
    return cur


def pyramid(raw, height, width, num_scales):
    """The chain of MonoDataset.preprocess: scale i from scale i - 1 (mono_dataset.py:101-104)."""
    out, cur = {}, raw
    for i in range(num_scales):
        cur = resize_lanczos_u8(cur, (height // 2 ** i, width // 2 ** i))
        out[i] = cur
    return out


def pil_resize(img_chw, size):
    """Pillow itself: (3,H,W) uint8 numpy -> (3,h,w) through Image.resize(LANCZOS)."""
    from PIL import Image
    im = Image.fromarray(np.ascontiguousarray(np.moveaxis(img_chw, 0, -1)))
    return np.moveaxis(np.asarray(im.resize((size[1], size[0]), Image.LANCZOS)), -1, 0)
