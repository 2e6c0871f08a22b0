"""-m gpu: the encoder -> loss glue (remainder) against the reference's own statements, run with
PyTorch on the CPU: repdepth.py:615-620, trainer.py:859-869, trainer.py:41-69 (DepthBins), repdepth.py:502-505."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _reference_glue(lowest_cost, confidence, mono_depth):
    H, W = mono_depth.shape[-2:]
    lc = F.interpolate(lowest_cost.unsqueeze(1), [H, W], mode="nearest")[:, 0]                    # repdepth.py:615-617
    cm = F.interpolate(confidence.unsqueeze(1), [H, W], mode="nearest")[:, 0]                     # repdepth.py:618-620
    matching_depth = 1 / lc.unsqueeze(1)                                                          # trainer.py:863
    mask = ((matching_depth - mono_depth) / mono_depth) < 1.0
    mask = mask * (((mono_depth - matching_depth) / matching_depth) < 1.0)
    return lc, cm * mask[:, 0]                                                                    # trainer.py:450-451


class _RefDepthBins:          # trainer.py:41-64 without the torchmetrics base class
    def __init__(self, opt_min_depth):
        self.min_depth, self.max_depth, self.opt_min_depth = torch.tensor(0.1), torch.tensor(10.0), opt_min_depth

    def update(self, mono_depth):
        min_depth = mono_depth.detach().min(-1)[0].min(-1)[0]
        max_depth = mono_depth.detach().max(-1)[0].max(-1)[0]
        min_depth, max_depth = min_depth.mean(), max_depth.mean()
        min_depth = max(self.opt_min_depth, min_depth * 0.9)
        max_depth = max_depth * 1.1
        self.max_depth = self.max_depth * 0.99 + max_depth * 0.01
        self.min_depth = self.min_depth * 0.99 + min_depth * 0.01


@pytest.mark.parametrize("shape", [(2, 48, 160, 192, 640), (3, 12, 20, 50, 70), (1, 7, 9, 33, 47)])
def test_matching_glue_and_depth_bins(shape):
    import ppea_depth_b200 as P
    B, h, w, H, W = shape
    g = torch.Generator().manual_seed(B * 100 + h)
    mono = 0.3 + 20 * torch.rand(B, 1, H, W, generator=g)
    lowest = 1.0 / (0.3 + 20 * torch.rand(B, h, w, generator=g))          # a disparity (1 / depth), as indices_to_disparity returns
    conf = (torch.rand(B, h, w, generator=g) < 0.7).float()
    want_lc, want_cm = _reference_glue(lowest, conf, mono)
    up, cons, extrema = P.matching_glue(lowest.cuda(), conf.cuda(), mono.cuda())
    assert torch.equal(up.cpu(), want_lc)
    assert torch.equal(cons.cpu(), want_cm)
    assert 0.05 < float(want_cm.mean()) < 0.95
    ref = _RefDepthBins(0.1)
    dev = P.DeviceDepthBins(0.1)
    for step in range(3):
        m = mono * (1.0 + 0.1 * step)
        ref.update(m)
        if step == 0:
            dev.update_from(extrema)
        else:
            dev.update(m.cuda())
        lo, hi = dev.compute()
        assert abs(float(lo) - float(ref.min_depth)) <= 1e-6 * float(ref.min_depth) and abs(float(hi) - float(ref.max_depth)) <= 1e-6 * float(ref.max_depth)
    # the clamp at opt_min_depth
    ref2, dev2 = _RefDepthBins(5.0), P.DeviceDepthBins(5.0)
    ref2.update(mono)
    dev2.update(mono.cuda())
    assert abs(float(dev2.compute()[0]) - float(ref2.min_depth)) <= 1e-6 * float(ref2.min_depth)


def test_zero_missing_poses():
    import ppea_depth_b200 as P
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(5, 16, 6, 20, generator=g)
    feats[1] = 0
    feats[4] = 0
    pose = torch.randn(5, 4, 4, generator=g)
    want = pose.clone()
    for b, feat in enumerate(feats):                      # repdepth.py:502-505
        if feat.sum() == 0:
            want[b] *= 0
    got = P.zero_missing_poses(pose.cuda().contiguous(), feats.cuda())
    assert torch.equal(got.cpu(), want)
    with pytest.raises(RuntimeError):
        P.zero_missing_poses(pose, feats)
