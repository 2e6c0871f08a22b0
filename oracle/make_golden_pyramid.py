"""Generates tests/golden/pyramid_*.pt by running the REFERENCE's own MonoDataset.preprocess (datasets/mono_dataset.py:96-112: the
LANCZOS resize chain + ToTensor) on a seeded raw frame.  Run in the build container (needs /root/reference, Pillow, torchvision):
    python -m oracle.make_golden_pyramid"""
import os
import sys

import numpy as np
import torch

from . import ref_import

CASES = {"pyramid_kitti_like_93x310_to_48x160": dict(raw=(93, 310), height=48, width=160, scales=4, seed=0),
         "pyramid_ragged_57x83_to_40x72": dict(raw=(57, 83), height=40, width=72, scales=3, seed=1)}


def main():
    from PIL import Image
    ref_import.load_reference()      # (stubs the uninstalled third-party imports of the dataset package: skimage, ...)
    from ppeadepth.datasets.mono_dataset import MonoDataset

    class _DS(MonoDataset):
        def check_depth(self):
            return False

    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    for name, c in CASES.items():
        rng = np.random.default_rng(c["seed"])
        yy, xx = np.mgrid[0:c["raw"][0], 0:c["raw"][1]]
        base = 127 + 90 * np.sin(xx / 7.0)[..., None] * np.cos(yy / 5.0)[..., None] * np.array([1.0, 0.7, -0.8])
        raw = np.clip(base + rng.normal(0, 25, base.shape), 0, 255).astype(np.uint8)      # smooth structure + noise, some clipping
        ds = _DS("", [], c["height"], c["width"], [0], c["scales"], is_train=True)
        inputs = {("color", 0, -1): Image.fromarray(raw)}
        ds.preprocess(inputs, (lambda x: x))
        fx = dict(case=c, raw=torch.from_numpy(raw).permute(2, 0, 1).contiguous())
        for i in range(c["scales"]):
            t = inputs[("color", 0, i)]
            u8 = (t * 255).round().to(torch.uint8)
            assert torch.equal(u8.float() / 255, t)                                     # ToTensor: exactly k / 255
            fx["scale%d" % i] = u8
        torch.save(fx, os.path.join(out, name + ".pt"))
        print(name, {i: tuple(fx["scale%d" % i].shape) for i in range(c["scales"])})


if __name__ == "__main__":
    main()
