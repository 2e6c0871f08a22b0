"""Pose-network output -> camera transform.  CPU: the oracle restatement against the fixture made by the
reference's own `transformation_from_parameters` (values and autograd gradients) and against the reference live.  -m gpu: the
fused forward / backward kernels through the public function."""
import os

import pytest
import torch

from oracle import pose_oracle as PO

FIX = os.path.join(os.path.dirname(__file__), "golden", "pose_transform.pt")


def _oracle_run(aa, tr, inv, w, dtype=torch.float32):
    a, t = aa.to(dtype).clone().requires_grad_(True), tr.to(dtype).clone().requires_grad_(True)
    M = PO.transformation_from_parameters(a, t, inv)
    (M * w.to(dtype)).sum().backward()
    return M.detach(), a.grad, t.grad


@pytest.mark.parametrize("inv", [False, True])
def test_oracle_matches_reference_fixture(inv):
    fx = torch.load(FIX, weights_only=False)
    ref = fx["inv" if inv else "fwd"]
    M, ga, gt = _oracle_run(fx["axisangle"], fx["translation"], inv, ref["w"])
    assert torch.equal(M, ref["M"])
    assert torch.allclose(ga, ref["g_axisangle"], rtol=1e-6, atol=1e-7) and torch.allclose(gt, ref["g_translation"], rtol=1e-6, atol=1e-7)
    assert torch.isfinite(ga).all()                         # zero rotation: norm() at the origin has a zero gradient


@pytest.mark.skipif(not os.path.isdir("/root/reference/ppeadepth"), reason="reference not mounted")
def test_oracle_matches_reference_live():
    from oracle import ref_import as R
    R.load_reference()
    import ppeadepth.layers as L
    aa, tr = PO.synthetic_poses(B=9, seed=5)
    for inv in (False, True):
        assert torch.equal(PO.transformation_from_parameters(aa, tr, inv), L.transformation_from_parameters(aa, tr, inv))


@pytest.mark.gpu
@pytest.mark.parametrize("inv", [False, True])
def test_cuda_pose_matches_fixture_and_fp64(inv):
    import ppea_depth_b200 as P
    fx = torch.load(FIX, weights_only=False)
    ref = fx["inv" if inv else "fwd"]
    a = fx["axisangle"].cuda().requires_grad_(True)
    t = fx["translation"].cuda().requires_grad_(True)
    M = P.transformation_from_parameters(a, t, inv)
    (M * ref["w"].cuda()).sum().backward()
    assert M.shape == (6, 4, 4)
    assert float((M.detach().cpu() - ref["M"]).abs().max()) <= 1e-6           # sinf / cosf: <= 2 ulp from the host's
    M64, ga64, gt64 = _oracle_run(fx["axisangle"], fx["translation"], inv, ref["w"], torch.float64)
    for got, want32, want64 in ((a.grad, ref["g_axisangle"], ga64), (t.grad, ref["g_translation"], gt64)):
        got = got.cpu()
        assert got.shape == want32.shape and torch.isfinite(got).all()
        scale = float(want64.abs().max())
        # as close to the float64 gradient as the reference's own float32 autograd (factor 2), and within 1e-5 of it
        err = float((got.double() - want64).abs().max())
        assert err <= max(2.0 * float((want32.double() - want64).abs().max()), 1e-5 * scale), (err, scale)


@pytest.mark.gpu
def test_cuda_pose_large_batch_and_loss_chain():
    """dL/dT of the fused loss reaches axis-angle / translation through the fused pose kernels."""
    import ppea_depth_b200 as P
    aa, tr = PO.synthetic_poses(B=200, seed=9)
    for inv in (False, True):
        a, t = aa.cuda().requires_grad_(True), tr.cuda().requires_grad_(True)
        M = P.transformation_from_parameters(a, t, inv)
        w = torch.randn(200, 4, 4, generator=torch.Generator().manual_seed(1))
        (M * w.cuda()).sum().backward()
        M64, ga, gt = _oracle_run(aa, tr, inv, w, torch.float64)
        assert float((M.detach().cpu().double() - M64).abs().max()) <= 2e-6
        assert float((a.grad.cpu().double() - ga).abs().max()) <= 2e-5 * float(ga.abs().max())
        assert float((t.grad.cpu().double() - gt).abs().max()) <= 2e-5 * float(gt.abs().max())


def test_oracle_matching_mask_matches_reference_live():
    """compute_matching_mask restatement vs the reference's own Trainer method (unbound, stand-in self)."""
    if not os.path.isdir("/root/reference/ppeadepth"):
        pytest.skip("reference not mounted")
    from types import SimpleNamespace
    from oracle import ref_import as R
    from oracle import vsl_oracle as O
    T = R.load_reference()
    g = torch.Generator().manual_seed(3)
    mono = 0.5 + 10 * torch.rand(2, 1, 24, 40, generator=g)
    lowest = 1 / (mono[:, 0] * torch.exp(1.2 * torch.randn(2, 24, 40, generator=g)))
    me = SimpleNamespace(device="cpu")
    want = T.Trainer.compute_matching_mask(me, {("mono_depth", 0, 0): mono, "lowest_cost": lowest})
    got = O.compute_matching_mask(mono, lowest)
    assert torch.equal(got, want) and 0.1 < float(want.float().mean()) < 0.9


@pytest.mark.gpu
def test_cuda_matching_mask():
    from oracle import vsl_oracle as O
    from ppea_depth_b200.loss import compute_matching_mask
    g = torch.Generator().manual_seed(4)
    mono = 0.5 + 10 * torch.rand(3, 1, 33, 47, generator=g)
    lowest = 1 / (mono[:, 0] * torch.exp(1.2 * torch.randn(3, 33, 47, generator=g)))
    want = O.compute_matching_mask(mono, lowest)
    got = compute_matching_mask(None, {("mono_depth", 0, 0): mono.cuda(), "lowest_cost": lowest.cuda()})
    assert got.dtype == torch.bool and got.shape == (3, 33, 47)
    assert torch.equal(got.cpu(), want) and 0.1 < float(want.float().mean()) < 0.9
