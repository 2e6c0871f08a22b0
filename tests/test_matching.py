"""Plane-sweep cost volume.  CPU: the oracle restatement against the fixtures produced by the
reference's own `match_features` (and against the reference live where /root/reference is mounted).  -m gpu: the CUDA
kernel, through the C ABI, against the same fixtures and against the oracle on larger seeded cases."""
import glob
import os

import pytest
import torch

from oracle import matching_oracle as M

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "matching_*.pt")))


def _load(path):
    return torch.load(path, weights_only=False)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_matches_reference_fixture(path):
    fx = _load(path)
    cost, missing = M.match_features(fx["cur"], fx["look"], fx["poses"], fx["K"], fx["invK"], fx["bins"].numpy(), fx["set_missing_to_max"])
    assert torch.equal(missing, fx["missing"])
    assert float((cost - fx["cost"]).abs().max()) <= 1e-6


@pytest.mark.skipif(not os.path.isdir("/root/reference/ppeadepth"), reason="reference not mounted")
def test_oracle_matches_reference_live():
    cur, look, poses, K, invK, bins = M.synthetic_case(B=2, Fr=2, C=12, h=20, w=28, D=10, seed=7, zero_pose_item=0)
    want = M.run_reference_match_features(cur, look, poses, K, invK, bins)
    got = M.match_features(cur, look, poses, K, invK, bins)
    assert torch.equal(got[1], want[1])
    assert float((got[0] - want[0]).abs().max()) <= 1e-6


def _check_kernel(cur, look, poses, K, invK, bins, stm, want_cost, want_missing):
    import ppea_depth_b200 as P
    dev = "cuda"
    cost, missing = P.match_features(cur.to(dev), look.to(dev), poses.to(dev), K.to(dev), invK.to(dev), bins, stm)
    cost, missing = cost.cpu(), missing.cpu()
    # the border masks compare projected coordinates with 2 and size-2: a hypothesis within rounding distance of a
    # threshold may fall on the other side -- those (pixel, bin) entries are excluded, and must be rare
    same = missing == want_missing
    assert float((~same).float().mean()) <= 2e-3, float((~same).float().mean())
    if stm:
        # a flipped entry changes the per-pixel maximum that fills the missing bins: compare pixels without flips
        ok_px = same.all(1, keepdim=True).expand_as(same)
    else:
        ok_px = same
    err = (cost - want_cost).abs()[ok_px]
    assert float(err.max()) <= 2e-6 * max(1.0, float(want_cost.abs().max())), float(err.max())


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_cuda_matches_reference_fixture(path):
    fx = _load(path)
    _check_kernel(fx["cur"], fx["look"], fx["poses"], fx["K"], fx["invK"], fx["bins"].numpy(), fx["set_missing_to_max"],
                  fx["cost"], fx["missing"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(B=2, Fr=1, C=32, h=48, w=160, D=96, seed=11, min_bin=0.3, max_bin=30.0),
                                  dict(B=3, Fr=2, C=7, h=33, w=47, D=13, seed=12, zero_pose_item=1),
                                  dict(B=1, Fr=1, C=64, h=24, w=80, D=128, seed=13, min_bin=0.05, max_bin=80.0)])
@pytest.mark.parametrize("stm", [True, False])
def test_cuda_matches_oracle(case, stm):
    cur, look, poses, K, invK, bins = M.synthetic_case(**case)
    want_cost, want_missing = M.match_features(cur, look, poses, K, invK, bins, stm)
    assert 0.02 < float(want_missing.mean()) < 0.98          # a mixed volume
    _check_kernel(cur, look, poses, K, invK, bins, stm, want_cost, want_missing)


@pytest.mark.gpu
def test_install_matching_rebinds_the_method():
    import types
    import ppea_depth_b200 as P
    cur, look, poses, K, invK, bins = M.synthetic_case(B=1, Fr=1, C=8, h=16, w=24, D=6, seed=3)

    class Enc:
        pass

    P.install_matching(Enc)
    me = Enc()
    me.warp_depths = torch.stack([torch.ones(1, 16, 24) * float(d) for d in bins], 0).cuda()
    me.set_missing_to_max = True
    cost, missing = me.match_features(cur.cuda(), look.cuda(), poses.cuda(), K.cuda(), invK.cuda())
    want = M.match_features(cur, look, poses, K, invK, bins)
    assert torch.equal(missing.cpu(), want[1])
    with pytest.raises(RuntimeError):
        P.match_features(cur, look, poses, K, invK, bins)          # no CPU path


def test_oracle_tail_on_fixture():
    """CPU: the tail restatement is self-consistent on a reference-made volume (confidence <=> no missing bin)."""
    fx = _load([p for p in GOLDEN if p.endswith("_max.pt")][0])
    conf, mins, argmin, masked = M.cost_volume_tail(fx["cost"], fx["missing"])
    assert torch.equal(conf, (fx["missing"].sum(1) == 0).float())
    assert torch.equal(masked, fx["cost"] * conf.unsqueeze(1))
    assert torch.equal(mins, torch.gather(torch.where(fx["cost"] == 0, torch.full_like(fx["cost"], 100.0), fx["cost"]), 1,
                                          argmin.unsqueeze(1)).squeeze(1))


@pytest.mark.gpu
@pytest.mark.parametrize("stm", [True, False])
def test_cuda_tail_matches_oracle(stm):
    import ppea_depth_b200 as P
    cur, look, poses, K, invK, bins = M.synthetic_case(B=2, Fr=1, C=16, h=40, w=72, D=24, seed=21, min_bin=2.0, max_bin=12.0)
    cost, missing = M.match_features(cur, look, poses, K, invK, bins, stm)
    want = M.cost_volume_tail(cost, missing)
    vol = cost.cuda().clone()
    conf, mins, argmin = P.cost_volume_tail(vol, missing.cuda())
    assert torch.equal(conf.cpu(), want[0]) and 0.05 < float(want[0].mean()) < 0.95
    assert torch.equal(mins.cpu(), want[1])
    assert torch.equal(argmin.cpu(), want[2])
    assert torch.equal(vol.cpu(), want[3])

    class Enc:
        num_depth_bins = 24

    P.install_matching(Enc)
    c2 = Enc().compute_confidence_mask((cost * (1 - missing)).cuda())
    assert torch.equal(c2.cpu(), want[0])


@pytest.mark.gpu
def test_matching_no_out_of_bounds_writes():
    """Guard regions around the cost volume / mask / tail outputs stay untouched (ragged size, far bins)."""
    from unittest import mock
    import ppea_depth_b200 as P
    from ppea_depth_b200 import matching as MM
    cur, look, poses, K, invK, bins = M.synthetic_case(B=2, Fr=2, C=5, h=33, w=47, D=13, seed=31, min_bin=0.05, max_bin=60.0)
    G = 1024
    real_empty = torch.empty
    tracked = []

    def guarded_empty(*size, **kw):
        dt = kw.get("dtype", torch.float32)
        n = 1
        for d in size:
            n *= int(d)
        sent = float("nan") if dt == torch.float32 else -7
        buf = torch.full((n + 2 * G,), sent, device=kw["device"], dtype=dt)
        tracked.append((buf, n, dt, sent))
        return buf[G:G + n].view(*size)

    with mock.patch.object(MM.torch, "empty", guarded_empty):
        cost, missing = P.match_features(cur.cuda(), look.cuda(), poses.cuda(), K.cuda(), invK.cuda(), bins, True)
        conf, mins, argmin = P.cost_volume_tail(cost, missing)
        torch.cuda.synchronize()
    assert len(tracked) == 5 and torch.isfinite(cost).all()
    for buf, n, dt, sent in tracked:
        for guard in (buf[:G], buf[G + n:]):
            assert bool(torch.isnan(guard).all()) if dt == torch.float32 else bool((guard == sent).all())
