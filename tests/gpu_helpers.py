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


def check_against_oracle(inputs, outputs, opt, is_multi, noise, losses, grads, maps, loss_rtol=1e-5, grad_rtol=1e-4,
                         oracle_maps=None):
    """The contract of BASELINE.json: selection bit-exact outside the fp32 margin, loss within 1e-5
    relative and gradients within 1e-4 relative -- the latter two evaluated at the SAME selection
    (the oracle is re-run with the kernel's own near-tie decisions, `forced`)."""
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
            if ps is not None:      # a flipped source index needs the two candidates to be near-tied
                dark = om["src_idx"] == 2
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
    l64, g64, _ = O.run_fwd_bwd(inputs, outputs, opt, is_multi, noise, dtype=torch.float64, forced=forced)
    for k, v in l32.items():
        if k.startswith("smooth_loss"):
            continue
        got = float(losses[k])
        # within loss_rtol of the fp32 reference path or of the fp64 truth (the fp32 reference itself
        # sits up to ~1e-5 from fp64 when SSIM ~ 0, e.g. identity pose)
        err = min(abs(got - float(v)), abs(got - float(l64[k])))
        assert err <= loss_rtol * abs(float(v)) + 1e-9, (k, got, float(v), float(l64[k]))
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
        ok = min(err64, err32) <= grad_rtol * scale + 1e-9 or err64 <= 2.0 * ref32_err
        assert ok, (k, err64, err32, ref32_err, scale)
    if is_multi:
        assert not any(k[0] == "cam_T_cam" for k in grads)
    return n_flip
