"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/ppea_vsl.h declares,
agrees with the ctypes mirror on struct layout, and rejects bad arguments on the host (no launch)."""
import ctypes
import os
import re
import subprocess

import pytest

from ppea_depth_b200 import _cabi as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ppea_vsl.h")


@pytest.fixture(scope="module")
def lib():
    C.build()
    return C.lib()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppea_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
        assert s in C.SIGNATURES, "ctypes binding misses " + s
    assert sorted(C.SIGNATURES) == syms


def test_abi_version_and_strerror(lib):
    assert lib.ppea_abi_version() == C.ABI_VERSION
    assert b"NULL" in lib.ppea_strerror(-1)
    assert b"shape" in lib.ppea_strerror(-2)
    assert lib.ppea_strerror(0) == b"success"


def test_struct_layout_matches_header(lib, tmp_path):
    """sizeof/offsetof as the C compiler sees them == the ctypes mirror."""
    prog = tmp_path / "layout.c"
    prog.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "ppea_vsl.h"
int main(void) {
  printf("%zu %zu %zu\\n", sizeof(PpeaVslParams), sizeof(PpeaVslScale), sizeof(PpeaVslGrads));
  printf("%zu %zu %zu %zu %zu %zu\\n", offsetof(PpeaVslParams, tgt), offsetof(PpeaVslParams, scales), offsetof(PpeaVslParams, sums),
         offsetof(PpeaVslParams, workspace_bytes), offsetof(PpeaVslParams, trace_events), offsetof(PpeaVslScale, grad_disp));
  printf("%zu %zu\\n", offsetof(PpeaVslGrads, grad_T), offsetof(PpeaVslGrads, workspace_bytes));
  printf("%zu %zu %zu\\n", sizeof(PpeaVslFused), offsetof(PpeaVslFused, workspace), offsetof(PpeaVslFused, workspace_bytes));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(prog)])
    out = subprocess.check_output([str(exe)], text=True).split()
    got = [int(x) for x in out]
    P, S, G = C.PpeaVslParams, C.PpeaVslScale, C.PpeaVslGrads
    want = [ctypes.sizeof(P), ctypes.sizeof(S), ctypes.sizeof(G), P.tgt.offset, P.scales.offset, P.sums.offset,
            P.workspace_bytes.offset, P.trace_events.offset, S.grad_disp.offset, G.grad_T.offset, G.workspace_bytes.offset,
            ctypes.sizeof(C.PpeaVslFused), C.PpeaVslFused.workspace.offset, C.PpeaVslFused.workspace_bytes.offset]
    assert got == want


def test_size_queries(lib):
    assert lib.ppea_vsl_sums_floats(12, 4) == 4 * (8 + 4 * 12)
    assert lib.ppea_vsl_workspace_bytes(12, 192, 640, 4) > 0
    det = lib.ppea_vsl_backward_workspace_bytes(12, 192, 640, 4, C.F_DETERMINISTIC)
    nondet = lib.ppea_vsl_backward_workspace_bytes(12, 192, 640, 4, 0)
    assert det - nondet == 4 * 12 * 192 * 640 * 4
    assert lib.ppea_vsl_workspace_bytes(0, 192, 640, 4) == 0


def test_argument_validation_happens_on_the_host(lib):
    """Negative codes are returned before any launch, so they can be exercised without a GPU."""
    assert lib.ppea_vsl_forward(None, None) == -1
    p = C.PpeaVslParams()
    assert lib.ppea_vsl_forward(ctypes.byref(p), None) == -5          # struct_size mismatch
    p.struct_size = ctypes.sizeof(C.PpeaVslParams)
    assert lib.ppea_vsl_forward(ctypes.byref(p), None) == -2          # shape
    p.batch, p.height, p.width, p.num_scales, p.total_scales = 2, 32, 64, 1, 1
    assert lib.ppea_vsl_forward(ctypes.byref(p), None) == -1          # NULL tensors
    p.num_scales = C.MAX_SCALES + 1
    assert lib.ppea_vsl_forward(ctypes.byref(p), None) == -2
    p.num_scales = 1
    p.flags = C.F_MULTI | C.F_GRAD_POSE
    assert lib.ppea_vsl_forward(ctypes.byref(p), None) == -4          # T is detached on the multi path
    assert lib.ppea_vsl_fused_forward(ctypes.byref(p), None, None) == -4      # (the MULTI|GRAD_POSE contradiction is found first)
    assert lib.ppea_vsl_fused_workspace_bytes(None) == 0
    assert lib.ppea_ssim_forward(None, None, None, 3, 8, 8, None) == -1
    assert lib.ppea_ssim_forward(1, 1, 1, 3, 0, 8, None) == -2
    assert lib.ppea_smooth_forward(None, None, None, None, 1, 8, 8, None) == -1


def test_library_is_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", C.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
