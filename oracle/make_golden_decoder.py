"""Generates tests/golden/decoder_*.pt by running the REFERENCE's own DepthDecoderV2 (networks/depth_decoder_v2.py) and
Conv3x3 (layers.py:119-135) with autograd on seeded inputs.  Run in the build container (needs /root/reference):
    python -m oracle.make_golden_decoder"""
import os

import torch

from . import decoder_oracle as D


def main():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    # the whole reference decoder (RepLKNet-31B channel counts) on random encoder features: 16x24 at stride 4 -> 64x96 disparity
    g = torch.Generator().manual_seed(7)
    feats = [torch.randn(1, c, 16 // (2 ** i), 24 // (2 ** i), generator=g) for i, c in enumerate([128, 256, 512, 1024])]
    case = D.run_reference_decoder(feats, seed=11, grad_seed=12)
    torch.save(case, os.path.join(out, "decoder_v2_tail_1x32x64x96.pt"))
    print("decoder_v2_tail", tuple(case["x"].shape), float(case["disp"].mean()))
    # ragged sizes straight through the reference's Conv3x3 + Sigmoid
    for name, (B, C, H, W, seed) in {"decoder_head_ragged_1x5x19x35": (1, 5, 19, 35, 3), "decoder_head_small_2x3x4x3": (2, 3, 4, 3, 4)}.items():
        x, w, b, grad = D.synthetic_head_case(B, C, H, W, seed)
        disp, gx, gw, gb = D.run_reference_conv3x3(x, w, b, grad)
        torch.save(dict(x=x, weight=w, bias=b, grad_disp=grad, disp=disp, grad_x=gx, grad_weight=gw, grad_bias=gb), os.path.join(out, name + ".pt"))
        print(name, float(disp.mean()))


if __name__ == "__main__":
    main()
