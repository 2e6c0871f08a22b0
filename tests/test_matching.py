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


@pytest.mark.gpu
@pytest.mark.parametrize("case", [dict(B=2, Fr=2, C=8, h=33, w=47, D=13, seed=41, min_bin=0.05, max_bin=60.0, zero_pose_item=1),
                                  dict(B=2, Fr=1, C=64, h=48, w=160, D=96, seed=42, min_bin=0.3, max_bin=30.0),
                                  dict(B=1, Fr=1, C=12, h=5, w=70, D=3, seed=43)])
def test_channel_quad_path_is_bit_identical_to_planar(case):
    """ppea_match_features_ws (features re-laid as (N, C/4, h, w, 4), 128-bit gathers) against the planar kernel: same
    arithmetic in the same channel order -> identical bits; guard regions around the scratch buffer and both outputs."""
    from unittest import mock
    import ppea_depth_b200 as P
    from ppea_depth_b200 import matching as MM
    cur, look, poses, K, invK, bins = M.synthetic_case(**case)
    g = [t.cuda() for t in (cur, look, poses, K, invK)]
    want_cost, want_missing = P.match_features(*g, bins, True, planar=True)
    G = 4096
    tracked = []

    def guarded_empty(*size, **kw):
        dt = kw.get("dtype", torch.float32)
        n = 1
        for d in size:
            n *= int(d)
        sent = float("nan") if dt == torch.float32 else 0xA5
        buf = torch.full((n + 2 * G,), sent, device=kw["device"], dtype=dt)
        tracked.append((buf, n, dt, sent))
        return buf[G:G + n].view(*size)

    MM._WORKSPACES.clear()
    with mock.patch.object(MM.torch, "empty", guarded_empty):
        cost, missing = P.match_features(*g, bins, True)
        torch.cuda.synchronize()
    MM._WORKSPACES.clear()
    assert len(tracked) == 3          # cost, missing, scratch (first use)
    scratch = [t for t in tracked if t[2] == torch.uint8]
    assert len(scratch) == 1 and scratch[0][1] == P._cabi.lib().ppea_match_workspace_bytes(case["B"], case["Fr"], case["C"], case["h"], case["w"]) > 0
    for buf, n, dt, sent in tracked:
        for guard in (buf[:G], buf[G + n:]):
            assert bool(torch.isnan(guard).all()) if dt == torch.float32 else bool((guard == sent).all())
    assert torch.equal(missing, want_missing) and (case["D"] < 8 or 0.02 < float(missing.mean()) < 0.98)
    assert torch.equal(cost, want_cost)
    # a shape without the fast path (C % 4 != 0) reports no scratch and takes the planar kernel
    assert P._cabi.lib().ppea_match_workspace_bytes(2, 1, 7, 33, 47) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", list(range(6)))
def test_random_shapes_quad_path_fixup_and_tail(seed):
    """Random small shapes around the kernels' partition boundaries (bins not a multiple of four, more bins than the split fix-up
    and tail kernels hold in registers, images narrower than a tile): channel-quad volume == planar volume bit for bit, the
    volume against the oracle, and the tail against the oracle's tail of the SAME volume."""
    import random
    import ppea_depth_b200 as P
    rng = random.Random(7000 + seed)
    case = dict(B=rng.choice([1, 2]), Fr=rng.choice([1, 2]), C=rng.choice([4, 8, 20, 64]), h=rng.choice([3, 7, 16, 33, 40]),
                w=rng.choice([5, 31, 32, 33, 70]), D=rng.choice([1, 5, 31, 33, 96, 128, 130]), seed=90 + seed,
                min_bin=rng.choice([0.05, 0.5]), max_bin=rng.choice([12.0, 60.0]))
    cur, look, poses, K, invK, bins = M.synthetic_case(**case)
    g = [t.cuda() for t in (cur, look, poses, K, invK)]
    for stm in (True, False):
        cost, missing = P.match_features(*g, bins, stm)
        cost_p, missing_p = P.match_features(*g, bins, stm, planar=True)
        assert torch.equal(cost, cost_p) and torch.equal(missing, missing_p), (case, stm)
        want_cost, want_missing = M.match_features(cur, look, poses, K, invK, bins, stm)
        _check_kernel(cur, look, poses, K, invK, bins, stm, want_cost, want_missing)
        want = M.cost_volume_tail(cost.cpu(), missing.cpu())
        vol = cost.clone()
        conf, mins, argmin = P.cost_volume_tail(vol, missing)
        assert torch.equal(conf.cpu(), want[0]) and torch.equal(mins.cpu(), want[1]) and torch.equal(argmin.cpu(), want[2]), (case, stm)
        assert torch.equal(vol.cpu(), want[3]), (case, stm)


# ---------------------------------------------------------------------------------------------------------------
# match_features_dyn (replk_matching_adapter.py:163-258)
# ---------------------------------------------------------------------------------------------------------------
GOLDEN_DYN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "matchdyn_*.pt")))


def _unpack_dyn(fx):
    import numpy as np
    B, _, h, w = fx["cur"].shape
    D = fx["bins"].numel()
    missing = torch.from_numpy(np.unpackbits(fx["missing_bits"].numpy())[:B * D * h * w].reshape(B, D, h, w).astype(np.float32))
    return fx["images_u8"].float() / 255.0, missing


@pytest.mark.parametrize("path", GOLDEN_DYN, ids=[os.path.basename(p)[:-3] for p in GOLDEN_DYN])
def test_oracle_dyn_matches_reference_fixture(path):
    """The restatement against the outputs of the reference's own method (kept layers, whole missing mask, float64 checksums
    of the whole volume)."""
    fx = _load(path)
    img, missing = _unpack_dyn(fx)
    cost, miss = M.match_features_dyn(fx["cur"], fx["look"], fx["poses"], fx["K"], fx["invK"], fx["bins"].numpy(), img, aug_mask=fx["aug"], **fx["opts"])
    assert torch.equal(miss, missing)
    assert float((cost[:, fx["kept_bins"]] - fx["cost"]).abs().max()) <= 1e-6
    assert abs(float(cost.double().sum()) - fx["cost_sum"]) <= 1e-9 * abs(fx["cost_sum"])
    assert abs(float((cost.double() ** 2).sum()) - fx["cost_sq"]) <= 1e-9 * abs(fx["cost_sq"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/ppeadepth"), reason="reference not mounted")
def test_oracle_dyn_matches_reference_live():
    cur, look, poses, K, invK, bins, img, aug = M.synthetic_dyn_case(B=1, C=3, seed=5)
    opts = dict(cv_min=False, set_1=False, pool=True, pool_r=1, pool_th=0.5)
    want = M.run_reference_match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, **opts)
    got = M.match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, **opts)
    assert torch.equal(got[1], want[1])
    assert float((got[0] - want[0]).abs().max()) <= 1e-6


def _check_dyn_kernel(cur, look, poses, K, invK, bins, img, aug, opts, want_cost, want_missing, stm=True):
    import ppea_depth_b200 as P
    d = lambda t: t.to("cuda")
    cost, missing = P.match_features_dyn(d(cur), d(look), d(poses), d(K), d(invK), bins, d(img), aug_mask=d(aug), set_missing_to_max=stm, **opts)
    cost, missing = cost.cpu(), missing.cpu()
    # border masks and the occlusion test (sample > pool_th) are decisions on fp32 coordinates: entries within rounding
    # distance of a threshold may fall the other way; they must be rare, and every other entry must agree
    same = missing == want_missing
    assert float((~same).float().mean()) <= 2e-3, float((~same).float().mean())
    ok_px = same.all(1, keepdim=True).expand_as(same) if stm else same
    err = (cost - want_cost).abs()
    flipped = (err > 2e-6 * max(1.0, float(want_cost.abs().max()))) & ok_px
    assert float(flipped.float().mean()) <= 1e-3, (float(flipped.float().mean()), float(err[ok_px].max()))      # occlusion decisions at the threshold
    return cost, missing


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN_DYN, ids=[os.path.basename(p)[:-3] for p in GOLDEN_DYN])
def test_cuda_dyn_matches_reference_fixture(path):
    fx = _load(path)
    img, missing = _unpack_dyn(fx)
    want_cost, _ = M.match_features_dyn(fx["cur"], fx["look"], fx["poses"], fx["K"], fx["invK"], fx["bins"].numpy(), img, aug_mask=fx["aug"], **fx["opts"])
    cost, _ = _check_dyn_kernel(fx["cur"], fx["look"], fx["poses"], fx["K"], fx["invK"], fx["bins"].numpy(), img, fx["aug"], fx["opts"], want_cost, missing)
    # ... and directly against the layers the reference itself produced
    bad = (cost[:, fx["kept_bins"]] - fx["cost"]).abs() > 2e-6
    assert float(bad.float().mean()) <= 3e-3


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(cv_min=True, set_1=False, pool=True, pool_r=2, pool_th=0.7),
                                  dict(cv_min=False, set_1=False, pool=True, pool_r=1, pool_th=0.3),
                                  dict(cv_min=True, set_1=True, pool=True, pool_r=1, pool_th=0.7),
                                  dict(cv_min=False, set_1=False, pool=False, pool_r=1, pool_th=0.7)])
@pytest.mark.parametrize("stm", [True, False])
def test_cuda_dyn_matches_oracle(opts, stm):
    """Sizes the reference method cannot run at (it hard-codes 48x128x96), two lookup frames, one augmented item."""
    cur, look, poses, K, invK, bins = M.synthetic_case(B=2, Fr=2, C=6, h=30, w=52, D=20, seed=23, min_bin=0.4, max_bin=25.0)
    g = torch.Generator().manual_seed(9)
    img = torch.round((0.2 + 0.6 * torch.rand(2, 3, 120, 208, generator=g)) * 255) / 255
    img[0, :, 20:60, 30:110] = 0.0
    img[1, :, 50:100, 100:190] = 0.0
    aug = torch.tensor([0.0, 0.0]).view(2, 1, 1, 1)
    want_cost, want_missing = M.match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, set_missing_to_max=stm, **opts)
    _check_dyn_kernel(cur, look, poses, K, invK, bins, img, aug, opts, want_cost, want_missing, stm)
    if opts["pool"] or opts["set_1"]:      # the occlusion handling changed something
        plain, _ = M.match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, set_missing_to_max=stm, **dict(opts, set_1=False, pool=False))
        assert float((plain - want_cost).abs().max()) > 1e-3
    aug1 = torch.tensor([1.0, 0.0]).view(2, 1, 1, 1)              # item 0 augmented: its occlusion handling is skipped (:196)
    want_cost, want_missing = M.match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug1, set_missing_to_max=stm, **opts)
    _check_dyn_kernel(cur, look, poses, K, invK, bins, img, aug1, opts, want_cost, want_missing, stm)


@pytest.mark.gpu
def test_install_matching_rebinds_the_dyn_method():
    import types
    import ppea_depth_b200 as P

    class Enc:
        def match_features(self, *a):
            return "reference"

        def match_features_dyn(self, *a, **k):
            return "reference"

        def compute_confidence_mask(self, *a):
            return "reference"

    P.install_matching(Enc)
    cur, look, poses, K, invK, bins, img, aug = M.synthetic_dyn_case(B=1, C=4, seed=3)
    me = Enc()
    me.warp_depths = torch.stack([torch.ones((1, 48, 128)) * float(d) for d in bins], 0).float().cuda()
    me.set_missing_to_max = True
    d = lambda t: t.cuda()
    cost, missing = me.match_features_dyn(d(cur), d(look), d(poses), d(K), d(invK), d(img), cv_min=True, aug_mask=d(aug), set_1=False, pool=True,
                                          pool_r=1, pool_th=0.7)
    want_cost, want_missing = M.match_features_dyn(cur, look, poses, K, invK, bins, img, aug_mask=aug, cv_min=True, set_1=False, pool=True, pool_r=1, pool_th=0.7)
    assert float((missing.cpu() != want_missing).float().mean()) <= 2e-3
