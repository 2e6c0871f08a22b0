"""-m gpu: the piecewise operators behind the reference's nn.Module / function API
(layers.py SSIM, BackprojectDepth, Project3D, get_smooth_loss, disp_to_depth; F.grid_sample as called
at trainer.py:911-914; Trainer.compute_reprojection_loss / compute_loss_masks) -- forward and backward
through the C ABI against the oracle's restatements (pinned to the reference by tests/test_oracle.py),
and, where /root/reference is mounted, against the reference's own modules."""
import pytest
import torch
import torch.nn.functional as F

import ppea_depth_b200 as P
from oracle import ref_import as R
from oracle import vsl_oracle as O
from ppea_depth_b200 import functional as Fn
from ppea_depth_b200.synth import SynthConfig, make_batch

pytestmark = pytest.mark.gpu

SHAPES = [(2, 24, 40), (1, 33, 47), (3, 16, 64)]


def rel(a, b):
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)


def batch(B, H, W, seed=0):
    cfg = SynthConfig(batch=B, height=H, width=W, num_scales=1, seed=seed)
    return make_batch(cfg)


@pytest.mark.parametrize("shape", SHAPES)
def test_ssim_module(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 3, H, W, generator=g)
    y = (x + 0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    xd, yd = x.double().requires_grad_(True), y.double().requires_grad_(True)
    ref = O.ssim(xd, yd)
    w = torch.rand(B, 3, H, W, generator=g).double()
    (ref * w).sum().backward()
    xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    out = P.SSIM()(xc, yc)
    (out * w.float().cuda()).sum().backward()
    assert out.shape == (B, 3, H, W)
    assert float((out.cpu() - ref.float()).abs().max()) < 5e-5          # fp32 SSIM noise floor (sigma = E[x^2]-mu^2)
    assert float((out.cpu() - O.ssim(x, y)).abs().max()) < 1e-4
    assert rel(xc.grad.cpu().double(), xd.grad) < 2e-3
    assert rel(yc.grad.cpu().double(), yd.grad) < 2e-3
    assert float(P.SSIM()(xc, xc).abs().max()) < 1e-6                    # SSIM(x, x) = 0


@pytest.mark.parametrize("no_ssim", [False, True])
def test_compute_reprojection_loss_and_masks(no_ssim):
    from types import SimpleNamespace
    B, H, W = 2, 24, 40
    g = torch.Generator().manual_seed(2)
    pred = torch.rand(B, 3, H, W, generator=g)
    tgt = (pred + 0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    pd = pred.double().requires_grad_(True)
    ref = O.photometric(pd, tgt.double(), no_ssim)
    w = torch.rand(B, 1, H, W, generator=g).double()
    (ref * w).sum().backward()
    mod = P.ViewSynthesisLoss(SimpleNamespace(no_ssim=no_ssim))
    pc = pred.cuda().requires_grad_(True)
    out = mod.compute_reprojection_loss(pc, tgt.cuda())
    (out * w.float().cuda()).sum().backward()
    assert out.shape == (B, 1, H, W)
    assert float((out.cpu().double() - ref).abs().max()) < 5e-5
    assert rel(pc.grad.cpu().double(), pd.grad) < 2e-3
    m = mod.compute_loss_masks(out, out + 1.0)
    assert float(m.min()) == 1.0 and m.dtype == torch.float32
    assert float(mod.compute_loss_masks(out, out).min()) == 1.0          # ties -> index 0 (torch.argmin first-min rule)
    assert float(mod.compute_loss_masks(out, None).min()) == 1.0


@pytest.mark.parametrize("shape", SHAPES)
def test_backproject_project_grid_sample_chain(shape):
    """BackprojectDepth -> Project3D -> grid_sample exactly as trainer.py:904-914 chains them."""
    B, H, W = shape
    inputs, outputs = batch(B, H, W, seed=3)
    disp = F.interpolate(outputs[("disp", 0)], size=(H, W))
    K, iK, T = inputs[("K", 0)], inputs[("inv_K", 0)], outputs[("cam_T_cam", 0, -1)]
    src = inputs[("color", -1, 0)]

    def chain(dev, dtype):
        d = disp.detach().clone().to(dev, dtype).requires_grad_(True)
        Tt = T.detach().clone().to(dev, dtype).requires_grad_(True)
        if dev == "cpu":
            _, depth = O.disp_to_depth(d, 0.1, 100.0)
            cam = O.backproject(depth, iK.to(dtype), H, W)
            pix = O.project(cam, K.to(dtype), Tt, H, W)
            warped = O.warp(src.to(dtype), pix)
        else:
            _, depth = P.disp_to_depth(d, 0.1, 100.0)
            cam = P.BackprojectDepth(B, H, W)(depth, iK.cuda())
            pix = P.Project3D(B, H, W)(cam, K.cuda(), Tt)
            warped = Fn.grid_sample_border(src.cuda(), pix)
        loss = (warped * torch.linspace(0.5, 1.5, W, device=dev, dtype=dtype)).sum()
        loss.backward()
        return cam.detach().cpu(), pix.detach().cpu(), warped.detach().cpu(), d.grad.cpu(), Tt.grad.cpu()

    cam32, pix32, w32, gd32, gT32 = chain("cpu", torch.float32)
    cam64, pix64, w64, gd64, gT64 = chain("cpu", torch.float64)
    cam, pix, w, gd, gT = chain("cuda", torch.float32)
    assert cam.shape == (B, 4, H * W) and pix.shape == (B, H, W, 2)
    assert rel(cam, cam32) < 2e-6 and torch.equal(cam[:, 3], torch.ones(B, H * W))
    assert float((pix - pix32).abs().max()) < 5e-6                      # normalised grid, op-by-op fp32 order
    assert float((w.double() - w64).abs().max()) < 2e-4
    assert rel(gd.double(), gd64) < max(2e-3, 2 * rel(gd32.double(), gd64))
    assert rel(gT.double(), gT64) < max(2e-3, 2 * rel(gT32.double(), gT64))


def test_project3d_dc_returns_depth():
    B, H, W = 2, 16, 32
    inputs, outputs = batch(B, H, W, seed=4)
    depth = 1.0 + 5.0 * torch.rand(B, 1, H, W)
    cam = O.backproject(depth, inputs[("inv_K", 0)], H, W)
    P4 = torch.matmul(inputs[("K", 0)], outputs[("cam_T_cam", 0, 1)])[:, :3, :]
    z_ref = torch.matmul(P4, cam)[:, 2, :].reshape(B, 1, H, W)
    pix, z = P.Project3D(B, H, W, dc=True)(cam.cuda(), inputs[("K", 0)].cuda(), outputs[("cam_T_cam", 0, 1)].cuda())
    assert z.shape == (B, 1, H, W) and pix.shape == (B, H, W, 2)
    assert rel(z.cpu(), z_ref) < 2e-6


@pytest.mark.parametrize("shape", SHAPES)
def test_get_smooth_loss(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(5)
    disp = torch.rand(B, 1, H, W, generator=g)
    disp[:, :, :2] = 0.5                                                 # exact ties: sign(0) = 0 in the backward
    img = torch.rand(B, 3, H, W, generator=g)
    dd = disp.double().requires_grad_(True)
    ref = O.smoothness(dd, img.double())
    ref.backward()
    dc = disp.cuda().requires_grad_(True)
    out = P.get_smooth_loss(dc, img.cuda())
    (2.0 * out).backward()
    assert out.dim() == 0
    assert abs(float(out) - float(ref)) < 2e-6 * float(ref)
    assert rel(dc.grad.cpu().double(), 2.0 * dd.grad) < 1e-5
    assert float(P.get_smooth_loss(torch.full((1, 1, 8, 8), 0.3).cuda(), torch.rand(1, 3, 8, 8).cuda())) == 0.0


@pytest.mark.skipif(not R.available(), reason="reference tree not mounted")
def test_modules_against_reference_classes():
    """Same constructor / forward signatures and results as the reference's own nn.Modules."""
    R.load_reference()
    import ppeadepth.layers as L
    B, H, W = 2, 24, 40
    inputs, outputs = batch(B, H, W, seed=6)
    depth = 1.0 + 5.0 * torch.rand(B, 1, H, W)
    K, iK, T = inputs[("K", 0)], inputs[("inv_K", 0)], outputs[("cam_T_cam", 0, 1)]
    cam_ref = L.BackprojectDepth(B, H, W)(depth, iK)
    pix_ref = L.Project3D(B, H, W)(cam_ref, K, T)
    x, y = inputs[("color", 0, 0)], inputs[("color", 1, 0)]
    cam = P.BackprojectDepth(B, H, W)(depth.cuda(), iK.cuda())
    pix = P.Project3D(B, H, W)(cam, K.cuda(), T.cuda())
    assert rel(cam.cpu(), cam_ref) < 2e-6
    assert float((pix.cpu() - pix_ref).abs().max()) < 5e-6
    assert float((P.SSIM()(x.cuda(), y.cuda()).cpu() - L.SSIM()(x, y)).abs().max()) < 1e-4
    d = torch.rand(B, 1, H, W)
    assert abs(float(P.get_smooth_loss(d.cuda(), x.cuda())) - float(L.get_smooth_loss(d, x))) < 1e-6
    a, t = 0.1 * torch.randn(B, 1, 3), torch.randn(B, 1, 3)
    for inv in (False, True):
        assert torch.allclose(P.transformation_from_parameters(a, t, inv), L.transformation_from_parameters(a, t, inv), atol=1e-6)
    sd, dp = P.disp_to_depth(d, 0.1, 100.0)
    sd_r, dp_r = L.disp_to_depth(d, 0.1, 100.0)
    assert torch.equal(sd, sd_r) and torch.equal(dp, dp_r)


@pytest.mark.parametrize("n", [256, 3 * 37 * 53, 12 * 3 * 192 * 640 + 5])
def test_images_u8_to_f32_is_totensor(n):
    """uint8 frames expanded on the device == torchvision ToTensor (oracle.images_from_u8), bit for bit; odd sizes and
    unaligned views take the scalar tail / scalar path."""
    from oracle import vsl_oracle as O
    from ppea_depth_b200 import images_to_float
    g = torch.Generator().manual_seed(n)
    u8 = torch.randint(0, 256, (n + 3,), dtype=torch.uint8, generator=g)
    for off in (0, 3):
        view = u8[off:off + n]
        got = images_to_float(view.cuda()) if off == 0 else images_to_float(u8.cuda()[off:off + n])
        assert got.dtype == torch.float32 and torch.equal(got.cpu(), O.images_from_u8(view))
    with pytest.raises(RuntimeError):
        images_to_float(u8)                                   # no CPU path


def test_loss_accepts_uint8_frames():
    """The loss methods fed uint8 colour frames give bit-identical results to the float frames (k/255)."""
    from gpu_helpers import FeedNoise
    from oracle import vsl_oracle as O
    from ppea_depth_b200.loss import ViewSynthesisLoss
    from ppea_depth_b200.synth import SynthConfig, make_batch, make_noise
    cfg = SynthConfig(batch=2, height=64, width=96, num_scales=4, seed=5)
    inputs, outputs = make_batch(cfg)
    noise = make_noise(cfg, 4)
    opt = O.default_opt(sclm=3, height=64, width=96, batch_size=2)

    def run(as_u8):
        ins, outs = O.clone_batch(inputs, outputs, device="cuda")
        if as_u8:
            for k in list(ins):
                if k[0] == "color":
                    ins[k] = torch.round(ins[k] * 255).to(torch.uint8)
        mod = ViewSynthesisLoss(opt)
        with FeedNoise(noise):
            mod.generate_images_pred(ins, outs, False)
            losses, _ = mod.compute_losses(ins, outs, False)
        losses["loss"].backward()
        return float(losses["loss"]), [outs[("disp", s)].grad.cpu() for s in range(4)]

    la, ga = run(False)
    lb, gb = run(True)
    assert la == lb
    # (coarse-scale gradients are float atomics: equal up to summation order)
    assert torch.equal(ga[0], gb[0])
    for a, b in zip(ga[1:], gb[1:]):
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max())
