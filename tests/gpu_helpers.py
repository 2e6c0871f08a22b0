"""Shared helpers of the -m gpu parity tests: run the CUDA path through the reference-facing host
API (ppea_depth_b200.loss -> functional -> C ABI) on a fixture / synthetic batch."""
import torch

from oracle import vsl_oracle as O
from ppea_depth_b200.loss import ViewSynthesisLoss

MARGIN = 1e-4   # see tests/test_emul.py


class FeedNoise:
    """Makes torch.randn inside the host code return the fixture's draws (what the reference's
    CPU generator produced), in order."""

    def __init__(self, draws):
        self.draws, self.i = list(draws), 0

    def __enter__(self):
        self.orig = torch.randn

        def fed(*shape, **kw):
            if kw.get("device") is not None or self.i >= len(self.draws):
                return self.orig(*shape, **kw)
            z = self.draws[self.i]
            self.i += 1
            return z.clone()
        torch.randn = fed
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def run_cuda(inputs, outputs, opt, is_multi, noise, deterministic=False, backward=True, device="cuda", fused=None):
    """Returns (losses, grads, maps) like oracle.run_fwd_bwd, computed by the CUDA path."""
    ins, outs = O.clone_batch(inputs, outputs, device=device)
    mod = ViewSynthesisLoss(opt, deterministic=deterministic, keep_maps=True, fused=fused)
    with FeedNoise(noise if noise is not None else []):
        mod.generate_images_pred(ins, outs, is_multi)
        losses, _ = mod.compute_losses(ins, outs, is_multi)
    if backward:
        losses["loss"].backward()
    torch.cuda.synchronize()
    grads = {k: v.grad.detach().cpu() for k, v in outs.items()
             if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam") and v.grad is not None}
    maps = {}
    for first, n, res in outs[("ppea_maps", bool(is_multi))]:
        for i in range(n):
            sel = res.sel[i].cpu()
            maps[first + i] = dict(depth=res.depth[i].detach().cpu(), r=res.loss_px[i].detach().cpu(),
                                   mask=((sel >> 2) & 1).unsqueeze(1), src_idx=(sel & 3).unsqueeze(1))
    depth_out = {s: outs[("depth", 0, s)] for s in range(opt.sclm + 1)}
    for s in depth_out:
        assert depth_out[s].data_ptr() == outs[("depth", 0, s)].data_ptr()
    return {k: v.detach().cpu() for k, v in losses.items()}, grads, maps


def forced_from(maps):
    return {s: (m["mask"], m["src_idx"]) for s, m in maps.items()}


LAST_REPORT = {}     # which bound decided each assertion of the most recent check_against_oracle (see below)


def knife_edge_pixels(oracle_maps64, inputs, opt, delta=2e-4):
    """Per scale, the (B,1,h_s,w_s) mask of disparity pixels whose gradient is decided by the last bits of fp32:
    a full-resolution pixel whose sampling position in either source lies within `delta` px of an integer coordinate
    (bilinear interpolation has a different slope on each side -- and the border clip switches the gradient off at
    0 and size-1, GridSampler.h clip_coordinates_set_grad), or whose warped value equals the target to 1e-6 in some
    channel (sign(y - x) of the L1 term).  Two correct fp32 implementations that place the sample 1e-5 px apart (the
    reference normalises and un-normalises the coordinate, the kernels do not) legitimately disagree THERE and only
    there: the derivative at such a pixel is O(1) different, everything else is continuous.  The mask is built from
    the float64 oracle and dilated to the coarse pixels the upsample adjoint spreads a full-resolution pixel to."""
    import torch.nn.functional as F
    out = {}
    H, W = opt.height, opt.width
    tgt = inputs[("color", 0, 0)].double()
    for s, m in oracle_maps64.items():
        knife = torch.zeros(tgt.shape[0], 1, H, W, dtype=torch.bool)
        for grid, warped in zip(m["grid"], m["warped"]):
            ix = (grid[..., 0].double() + 1) / 2 * (W - 1)
            iy = (grid[..., 1].double() + 1) / 2 * (H - 1)
            near = ((ix - ix.round()).abs() < delta) | ((iy - iy.round()).abs() < delta)
            knife |= near.unsqueeze(1)
            knife |= ((warped.double() - tgt).abs() < 1e-6).any(1, keepdim=True)
        hs, ws = m["disp_hw"]
        f = max(1, (H + hs - 1) // hs)
        k = knife.float()
        if f > 1:
            k = F.max_pool2d(k, kernel_size=2 * f + 1, stride=1, padding=f)
        out[s] = (F.adaptive_max_pool2d(k, (hs, ws)) > 0, int(knife.sum()))
    return out


def check_against_oracle(inputs, outputs, opt, is_multi, noise, losses, grads, maps, loss_rtol=1e-5, grad_rtol=1e-4,
                         oracle_maps=None, report=None):
    """The contract of BASELINE.json: selection bit-exact outside the fp32 margin, loss within 1e-5
    relative and gradients within 1e-4 relative -- the latter two evaluated at the SAME selection
    (the oracle is re-run with the kernel's own near-tie decisions, `forced`)."""
    # `report` (and gpu_helpers.LAST_REPORT) records per loss / gradient which bound let it pass:
    #   "fp32"  within tolerance of the fp32 reference-order oracle at the kernel's selection (the primary bound),
    #   "fp32-class(n/m)"  disparity gradient: n pixels are further than the tolerance from the float64 gradient -- no
    #           more than twice the m pixels at which the reference's OWN fp32 gradient is (+8), and by no more than twice
    #           its largest deviation.  These are the knife-edge samples of knife_edge_pixels(): the sample position of
    #           a bilinear gather within ~1e-5 px of an integer, where two correct fp32 evaluations pick different slopes,
    #   "fp64"  only within tolerance of the float64 evaluation (the fp32 reference itself is off by more),
    #   "kink"  gradient outside grad_rtol of both, but at least as close to float64 as the fp32 reference is (factor 2).
    report = LAST_REPORT if report is None else report
    report.clear()
    S = opt.sclm + 1
    if oracle_maps is None:
        _, _, oracle_maps = O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise, want_maps=True)
    n_flip = 0
    for s in range(S):
        om, m = oracle_maps[s], maps[s]
        r, ident = om["r"], om["ident"]
        bad_src = m["src_idx"] != om["src_idx"].to(torch.uint8)
        agree = ~bad_src
        assert float((m["r"] - r)[agree].abs().max()) < 1.5e-4, s
        assert abs(float((m["r"] - r)[agree].mean())) < 5e-7, s
        if bad_src.any():
            ps = om.get("per_src")
            if ps is not None:      # a flipped source index needs the two candidates to be near-tied ...
                dark = om["src_idx"] == 2
                if "warped" in om:  # ... or selec_reproj's darkness test (sum_c warped < 0.1, trainer.py:1078-1079) to sit on its threshold
                    for w in om["warped"]:
                        dark = dark | ((w.sum(1, keepdim=True) - 0.1).abs() < MARGIN)
                assert float((ps[:, 0:1] - ps[:, 1:2]).abs()[bad_src & ~dark].max() if (bad_src & ~dark).any() else 0.0) < MARGIN
            assert int(bad_src.sum()) <= max(2, bad_src.numel() // 2000), (s, int(bad_src.sum()))
        if not is_multi:
            bad = m["mask"] != om["mask"].to(torch.uint8)
            if bad.any():
                assert float((r - ident).abs()[bad].max()) < MARGIN, s
            near = int(((r - ident).abs() < MARGIN).sum())
            assert int(bad.sum()) <= max(2, bad.numel() // 2000, near // 2), (s, int(bad.sum()), near)
            n_flip += int(bad.sum())
    forced = forced_from(maps)
    l32, g32, _ = O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise, forced=forced)
    l64, g64, m64 = O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise, dtype=torch.float64, forced=forced, want_maps=True)
    for s in m64:
        m64[s]["disp_hw"] = tuple(outputs[("disp", s)].shape[-2:])
    report["_m64"] = m64
    knife = None
    for k, v in l32.items():
        if k.startswith("smooth_loss"):
            continue
        got = float(losses[k])
        # within loss_rtol of the fp32 reference path or of the fp64 truth (the fp32 reference itself
        # sits up to ~1e-5 from fp64 when SSIM ~ 0, e.g. identity pose)
        e32, e64 = abs(got - float(v)), abs(got - float(l64[k]))
        tol = loss_rtol * abs(float(v)) + 1e-9
        report[k] = "fp32" if e32 <= tol else ("fp64" if e64 <= tol else "FAIL")
        assert min(e32, e64) <= tol, (k, got, float(v), float(l64[k]))
    # Gradients: within grad_rtol (max-norm, relative) of the fp32 reference path at the same
    # selection; where fp32 itself is kink-limited (sign(y-x), clamp and clip boundaries decided by the
    # last bit make the reference's own fp32 gradient deviate from fp64 by more than that), the kernel
    # must be at least as close to the fp64 truth as the reference is (factor 2).
    for k, ref in g64.items():
        assert k in grads, k
        scale = float(ref.abs().max())
        err64 = float((grads[k].double() - ref).abs().max())
        err32 = float((grads[k] - g32[k]).abs().max())
        ref32_err = float((g32[k].double() - ref).abs().max())
        tol = grad_rtol * scale + 1e-9
        report[k] = "fp32" if err32 <= tol else ("fp64" if err64 <= tol else ("kink" if err64 <= 2.0 * ref32_err else "FAIL"))
        if report[k] not in ("fp32", "FAIL") and k[0] == "disp":
            n_bad = int(((grads[k].double() - ref).abs() > tol).sum())
            n_ref = int(((g32[k].double() - ref).abs() > tol).sum())
            if n_bad <= 2 * n_ref + 8 and err64 <= 2.0 * ref32_err:
                report[k] = "fp32-class(%d/%d)" % (n_bad, n_ref)
        if report[k] == "FAIL" and k[0] == "disp" and not getattr(opt, "v1_multiscale", False):
            # "fp32-knife(n)": every one of the n pixels outside the tolerance is a knife-edge sample (or a coarse pixel one
            # spreads to), there are fewer of them than knife-edge samples, and none is off by more than one pixel's own
            # contribution (1e-2 of the largest gradient)
            if knife is None:
                knife = knife_edge_pixels(m64, inputs, opt)
            km, n_knife = knife[k[1]]
            bad = (grads[k] - g32[k]).abs() > tol
            if not bool((bad & ~km).any()) and int(bad.sum()) <= n_knife and err32 <= 1e-2 * scale:
                report[k] = "fp32-knife(%d)" % int(bad.sum())
        assert report[k] != "FAIL", (k, err64, err32, ref32_err, scale)
    if is_multi:
        assert not any(k[0] == "cam_T_cam" for k in grads)
    return n_flip
