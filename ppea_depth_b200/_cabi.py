"""ctypes binding of libppea_vsl.so (C ABI in include/ppea_vsl.h).

The library is plain CUDA C++ with `extern "C"` entry points -- no libtorch,
no pybind.  PyTorch is only the owner of device memory and streams: tensors
are passed as raw device pointers and every call is enqueued on
``torch.cuda.current_stream()``.  There is no CPU fallback: if the shared
library is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = (os.environ.get("PPEA_LIB") and os.path.abspath(os.environ["PPEA_LIB"])) or os.path.join(PKG_DIR, "libppea_vsl.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")
SOURCES = ("api.cu", "vsl_fwd.cu", "vsl_bwd.cu", "vsl_fused.cu", "vsl_stream.cu", "smooth.cu", "ops.cu", "matching.cu", "pose.cu", "decoder_tail.cu", "pyramid.cu")
HEADERS = ("vsl_common.cuh", "vsl_math.cuh", "vsl_gather.cuh", "smooth.cuh")

ABI_VERSION = 11
TRACE_EVENTS = 5
MAX_SCALES = 4
SUMS_PER_SCALE = 8
LOSSES_PER_SCALE = 4

F_MULTI = 1 << 0
F_AUTOMASK = 1 << 1
F_SELEC_REPROJ = 1 << 2
F_NO_SSIM = 1 << 3
F_DETERMINISTIC = 1 << 4
F_MOTION_MASK = 1 << 5
F_MATCH_AUG = 1 << 6
F_GRAD_POSE = 1 << 7
F_GRAD_PREZEROED = 1 << 8
F_RAW_PREZEROED = 1 << 9
F_FUSED_TILES = 1 << 10

SEL_SRC_MASK = 3
SEL_AUTOMASK = 4

c_float_p = ctypes.c_void_p   # raw device pointers


class PpeaVslScale(ctypes.Structure):
    _fields_ = [
        ("disp_h", ctypes.c_int32), ("disp_w", ctypes.c_int32),
        ("disp", c_float_p), ("color", c_float_p), ("noise", c_float_p), ("mono_depth", c_float_p),
        ("depth", c_float_p), ("loss_px", c_float_p), ("sel", ctypes.c_void_p), ("grad_disp", c_float_p),
    ]


class PpeaVslParams(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("num_scales", ctypes.c_int32), ("first_scale", ctypes.c_int32), ("total_scales", ctypes.c_int32),
        ("disp_lo", ctypes.c_float), ("disp_range", ctypes.c_float), ("eps", ctypes.c_float),
        ("disparity_smoothness", ctypes.c_float),
        ("tgt", c_float_p), ("src", c_float_p * 2), ("K", c_float_p), ("inv_K", c_float_p), ("T", c_float_p * 2),
        ("cons_mask", c_float_p), ("aug_mask", c_float_p),
        ("scales", PpeaVslScale * MAX_SCALES),
        ("sums", c_float_p), ("losses", c_float_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("trace_events", ctypes.POINTER(ctypes.c_void_p)),
    ]


class PpeaVslGrads(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("grad_losses", c_float_p), ("grad_T", c_float_p * 2),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


class PpeaVslFused(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
    ]


_I, _P, _F, _SZ, _U = ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint32

# name -> (restype, argtypes); every symbol include/ppea_vsl.h declares
SIGNATURES = {
    "ppea_abi_version": (_I, []),
    "ppea_strerror": (ctypes.c_char_p, [_I]),
    "ppea_event_create": (_P, []),
    "ppea_event_destroy": (None, [_P]),
    "ppea_event_record": (_I, [_P, _P]),
    "ppea_event_elapsed_ms": (_I, [_P, _P, ctypes.POINTER(ctypes.c_float)]),
    "ppea_vsl_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "ppea_vsl_backward_workspace_bytes": (_SZ, [_I, _I, _I, _I, _U]),
    "ppea_vsl_sums_floats": (_SZ, [_I, _I]),
    "ppea_vsl_forward": (_I, [ctypes.POINTER(PpeaVslParams), _P]),
    "ppea_vsl_backward": (_I, [ctypes.POINTER(PpeaVslParams), ctypes.POINTER(PpeaVslGrads), _P]),
    "ppea_vsl_fused_workspace_bytes": (_SZ, [ctypes.POINTER(PpeaVslParams)]),
    "ppea_vsl_fused_forward": (_I, [ctypes.POINTER(PpeaVslParams), ctypes.POINTER(PpeaVslFused), _P]),
    "ppea_vsl_fused_backward": (_I, [ctypes.POINTER(PpeaVslParams), ctypes.POINTER(PpeaVslGrads), ctypes.POINTER(PpeaVslFused), _P]),
    "ppea_ssim_forward": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "ppea_ssim_backward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ppea_reprojection_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "ppea_reprojection_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ppea_backproject_forward": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "ppea_backproject_backward": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "ppea_project3d_forward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "ppea_project3d_partials_bytes": (_SZ, [_I, _I, _I]),
    "ppea_project3d_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P]),
    "ppea_warp_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ppea_warp_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ppea_smooth_workspace_bytes": (_SZ, [_I, _I, _I]),
    "ppea_smooth_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ppea_smooth_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ppea_images_u8_to_f32": (_I, [_P, _P, _SZ, _P]),
    "ppea_pose_to_matrix_forward": (_I, [_P, _P, _I, _P, _I, _P]),
    "ppea_pose_to_matrix_backward": (_I, [_P, _P, _I, _P, _P, _P, _I, _P]),
    "ppea_matching_mask": (_I, [_P, _P, _P, _SZ, _P]),
    "ppea_matching_glue": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ppea_depth_bins_update": (_I, [_P, _I, _F, _P, _P, _P]),
    "ppea_zero_missing_poses": (_I, [_P, _SZ, _P, _I, _I, _P]),
    "ppea_disp_head_forward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P]),
    "ppea_disp_head_workspace_bytes": (_SZ, [_I, _I, _I, _I]),
    "ppea_disp_head_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ppea_lanczos_ksize": (_I, [_I, _I]),
    "ppea_lanczos_table": (_I, [_I, _I, _P, _P]),
    "ppea_resize_lanczos_u8": (_I, [_P, _P, _P, _SZ, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P]),
    "ppea_pack_rgbx_u8": (_I, [_P, _P, _SZ, _I, _I, _I, _P]),
    "ppea_match_tail": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "ppea_match_features": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P]),
    "ppea_match_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "ppea_match_features_ws": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _SZ, _P]),
    "ppea_match_features_dyn": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _F, _P]),
}


def nvcc_command(out=LIB_PATH, extra=()):
    """The in-tree build: nvcc cross-compiles sm_100a without a GPU."""
    return (["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
             "--compiler-options", "-fPIC", "-shared", "-I", INCLUDE_DIR, "-o", out]
            + list(extra) + [os.path.join(CSRC_DIR, s) for s in SOURCES])


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC_DIR, s) for s in SOURCES + HEADERS] + [os.path.join(INCLUDE_DIR, "ppea_vsl.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    cmd = nvcc_command(extra=["-Xptxas", "-v"] if verbose else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        sys.stderr.write(res.stdout)
    return LIB_PATH


_LIB = None


def lib():
    """Loads libppea_vsl.so (must have been built in-tree: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ppea_depth_b200: %s is missing -- build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype, fn.argtypes = res, args
    ver = handle.ppea_abi_version()
    if ver != ABI_VERSION:
        raise RuntimeError("libppea_vsl.so ABI %d != binding ABI %d; rebuild" % (ver, ABI_VERSION))
    assert ctypes.sizeof(PpeaVslParams) % 8 == 0
    _LIB = handle
    return _LIB


def check(rc):
    if rc != 0:
        raise RuntimeError("libppea_vsl: %s (code %d)" % (lib().ppea_strerror(rc).decode(), rc))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())
