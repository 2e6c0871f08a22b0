"""ORACLE-SIDE RECIPE (test / baseline infrastructure, not product code).

Makes the reference's own CPU implementation of the path available where /root/reference does not exist (the
GPU box): copies the `ppeadepth` Python package -- `layers.py`, `trainer.py` and the package modules `trainer.py`
imports at load time -- from the read-only reference tree into `oracle/_ref/`, which is git-ignored (it never
enters the history) but travels with the gpurun snapshot like the built `.so`.  `oracle/ref_import.py` then runs
the UNMODIFIED loss methods from there (`bench.py --impl reference`, `cpu_baseline.kind == "reference"`).

    python oracle/build_ref.py            # no-op when /root/reference is absent
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PPEA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose=True) -> bool:
    src = os.path.join(REF, "ppeadepth")
    if not os.path.isfile(os.path.join(src, "trainer.py")):
        if verbose:
            print("oracle/_ref: reference tree not present at %s -- keeping what is there" % REF)
        return os.path.isfile(os.path.join(DST, "ppeadepth", "trainer.py"))
    dst = os.path.join(DST, "ppeadepth")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.pt", "*.png", "*.jpg"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(dst))
        print("oracle/_ref: %d files of the reference's ppeadepth package copied from %s" % (n, src))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
