"""Decoder tail: the disparity head `sigmoid(Conv3x3(x))` of DepthDecoderV2 (depth_decoder_v2.py:239).

not gpu: the oracle restatement against the fixtures the reference's own DepthDecoderV2 / Conv3x3 produced (and against the
reference live where it is mounted).  gpu: the CUDA head through the C ABI against the fixtures and the oracle -- values,
all three gradients, depth output, module / install_decoder behaviour, fixed-order (bit-reproducible) reductions."""
import os

import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import decoder_oracle as D
from oracle import ref_import

CASES = ["decoder_v2_tail_1x32x64x96", "decoder_head_ragged_1x5x19x35", "decoder_head_small_2x3x4x3"]


def _load(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    fx = _load(name)
    disp = D.disp_head(fx["x"], fx["weight"], fx["bias"])
    assert float((disp - fx["disp"]).abs().max()) <= 2e-7          # same ATen ops on the same machine class: a rounding apart at most
    gx, gw, gb = D.disp_head_grads(fx["x"], fx["weight"], fx["bias"], fx["grad_disp"])
    assert _rel(gx, fx["grad_x"]) <= 1e-6 and _rel(gw, fx["grad_weight"]) <= 1e-5 and _rel(gb, fx["grad_bias"]) <= 1e-5


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_oracle_matches_reference_live():
    x, w, b, grad = D.synthetic_head_case(2, 6, 21, 30, seed=9)
    disp_r, gx_r, gw_r, gb_r = D.run_reference_conv3x3(x, w, b, grad)
    assert float((D.disp_head(x, w, b) - disp_r).abs().max()) <= 2e-7
    gx, gw, gb = D.disp_head_grads(x, w, b, grad)
    assert _rel(gx, gx_r) <= 1e-6 and _rel(gw, gw_r) <= 1e-5 and _rel(gb, gb_r) <= 1e-5


def _cuda_head(x, w, b, grad):
    import ppea_depth_b200 as P
    xd = x.cuda().requires_grad_(True)
    wd = w.cuda().requires_grad_(True)
    bd = b.cuda().requires_grad_(True)
    disp = P.disp_head(xd, wd, bd)
    disp.backward(grad.cuda())
    return disp.detach().cpu(), xd.grad.cpu(), wd.grad.cpu(), bd.grad.cpu()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_head_against_reference_fixture(name):
    fx = _load(name)
    disp, gx, gw, gb = _cuda_head(fx["x"], fx["weight"], fx["bias"], fx["grad_disp"])
    # forward: the sum over C*9 products is associated differently from ATen's -> a few ulp of the pre-activation (|v| <~ 4), sigmoid' <= 1/4
    assert float((disp - fx["disp"]).abs().max()) <= 1e-6
    # gradients against the float64 restatement: at least as close as the reference's own float32 autograd (factor 2), and <= 1e-5 of max
    gx64, gw64, gb64 = D.disp_head_grads(fx["x"], fx["weight"], fx["bias"], fx["grad_disp"], dtype=torch.float64)
    for mine, ref32, ref64, what in ((gx, fx["grad_x"], gx64, "x"), (gw, fx["grad_weight"], gw64, "weight"), (gb, fx["grad_bias"].reshape(1), gb64, "bias")):
        err, ref_err = _rel(mine.reshape(ref64.shape), ref64), _rel(ref32.reshape(ref64.shape), ref64)
        assert err <= max(2 * ref_err, 2e-6), (what, err, ref_err)
        assert _rel(mine.reshape(ref32.shape), ref32) <= 1e-5, what


@pytest.mark.gpu
def test_cuda_head_full_size_depth_and_determinism():
    import ppea_depth_b200 as P
    x, w, b, grad = D.synthetic_head_case(2, 32, 192, 640, seed=5)      # RepLKNet-31B head: 32 channels at the KITTI resolution
    xd, wd, bd = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    disp, depth = P.disp_head_with_depth(xd, wd, bd, 0.1, 100.0)
    ref = D.disp_head(x, w, b)
    assert float((disp.detach().cpu() - ref).abs().max()) <= 1e-6
    _, depth_ref = D.disp_to_depth(disp.detach().cpu(), 0.1, 100.0)     # layers.py:14-23 on the kernel's own disparity: same op order
    assert float(((depth.cpu() - depth_ref) / depth_ref).abs().max()) <= 4e-7
    assert not depth.requires_grad
    disp.backward(grad.cuda())
    g1 = (xd.grad.clone(), wd.grad.clone(), bd.grad.clone())
    gx64, gw64, gb64 = D.disp_head_grads(x, w, b, grad, dtype=torch.float64)
    assert _rel(g1[0].cpu(), gx64) <= 2e-6 and _rel(g1[1].cpu(), gw64) <= 2e-6 and _rel(g1[2].cpu(), gb64) <= 2e-5
    xd.grad = wd.grad = bd.grad = None
    P.disp_head(xd, wd, bd).backward(grad.cuda())
    assert torch.equal(g1[0], xd.grad) and torch.equal(g1[1], wd.grad) and torch.equal(g1[2], bd.grad)     # fixed-order reductions


@pytest.mark.gpu
def test_install_decoder_keeps_state_dict_and_output():
    import torch.nn as nn
    import ppea_depth_b200 as P

    class Conv3x3(nn.Module):          # layers.py:119-135 (the reference's class, restated: the GPU box has no /root/reference)
        def __init__(self, cin, cout):
            super().__init__()
            self.pad = nn.ReflectionPad2d(1)
            self.conv = nn.Conv2d(int(cin), int(cout), 3)

        def forward(self, x):
            return self.conv(self.pad(x))

    class Tail(nn.Module):             # the last statement of DepthDecoderV2.forward (depth_decoder_v2.py:239)
        def __init__(self):
            super().__init__()
            self.disp_convs = nn.ModuleList([Conv3x3(8, 1)])
            self.sigmoid = nn.Sigmoid()

        def forward(self, x):
            return self.sigmoid(self.disp_convs[0](x))

    torch.manual_seed(0)
    dec = Tail().cuda()
    keys = sorted(dec.state_dict().keys())
    x = torch.randn(2, 8, 33, 47, device="cuda")
    before = dec(x)
    P.install_decoder(dec)
    assert isinstance(dec.disp_convs[0], P.FusedDispHead)
    assert sorted(dec.state_dict().keys()) == keys                      # checkpoints of the reference still load
    after = dec(x)
    assert float((after - before).abs().max()) <= 1e-6
    after.sum().backward()
    assert dec.disp_convs[0].conv.weight.grad is not None and dec.disp_convs[0].conv.bias.grad is not None
    with pytest.raises(RuntimeError):
        P.disp_head(x.cpu(), dec.disp_convs[0].conv.weight.cpu(), dec.disp_convs[0].conv.bias.cpu())      # no CPU path
