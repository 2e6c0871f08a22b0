import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names():
    # (the matching_*.pt / pose_*.pt / decoder_*.pt / pyramid_*.pt fixtures belong to their own test files)
    return sorted(f[:-3] for f in os.listdir(GOLDEN_DIR) if f.endswith(".pt") and not f.startswith(("matching_", "matchdyn_", "pose_", "decoder_", "pyramid_")))


def load_golden(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    fx["inputs"] = {k: (v.float() / 255.0 if v.dtype == torch.uint8 else v) for k, v in fx["inputs"].items()}
    fx["opt"] = SimpleNamespace(**fx["opt"])
    return fx


def to_device(d, device):
    return {k: v.to(device) for k, v in d.items()}
