"""ORACLE (test infrastructure, not product code).

Runs the reference's OWN `ppeadepth.layers` / `ppeadepth.trainer.Trainer`
loss methods, unmodified, from /root/reference -- only possible in the build
container (the GPU box has no /root/reference).  Used by
`oracle/make_golden.py` to produce the committed fixtures and by
`tests/test_oracle.py` (skipped when the reference is absent) to pin
`oracle/vsl_oracle.py`.

trainer.py imports packages that are not installed here (skimage, matplotlib,
accelerate, torchmetrics, timm); they are registered as empty stub modules
before the import -- the four loss methods never touch them.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import torch

def _reference_root():
    """/root/reference in the build container; on the GPU box the copy that `oracle/build_ref.py` left under
    `oracle/_ref/` (git-ignored, ships with the gpurun snapshot)."""
    for root in (os.environ.get("PPEA_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")):
        if root and os.path.isfile(os.path.join(root, "ppeadepth", "trainer.py")):
            return root
    return os.environ.get("PPEA_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _reference_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ppeadepth", "trainer.py"))


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


def _register_stub(name):
    if name in sys.modules:
        return sys.modules[name]
    mod = _Anything(name)
    mod.__path__ = []
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    sys.modules[name] = mod
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, mod)
    return mod


_TRAINER_MOD = None


def load_reference():
    """Imports ppeadepth.trainer from the reference tree; returns the module."""
    global _TRAINER_MOD
    if _TRAINER_MOD is not None:
        return _TRAINER_MOD
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    nthreads = torch.get_num_threads()
    saved_env = {k: os.environ.get(k) for k in ("MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS", "OMP_NUM_THREADS")}
    for name in ("skimage", "skimage.transform", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
                 "accelerate", "torchmetrics", "timm", "timm.layers"):
        fresh = name not in sys.modules
        try:
            if fresh:
                __import__(name)
        except Exception:
            _register_stub(name)
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, _Anything):
        plt.get_cmap = lambda *a, **k: None                       # trainer.py:72
    tm = sys.modules["torchmetrics"]
    if isinstance(tm, _Anything):
        class Metric(torch.nn.Module):                            # trainer.py:41-46
            def add_state(self, name, default, dist_reduce_fx=None):
                setattr(self, name, default)
        tm.Metric = Metric
    tl = sys.modules["timm.layers"]
    if isinstance(tl, _Anything):
        tl.DropPath = torch.nn.Identity
        tl.trunc_normal_ = lambda t, *a, **k: t
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    argv, sys.argv = sys.argv, ["x"]
    try:
        import ppeadepth.trainer as T
    finally:
        sys.argv = argv
        # trainer.py:8-10 pins the BLAS/OpenMP pools to one thread at import
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        torch.set_num_threads(nthreads)
    _TRAINER_MOD = T
    return T


def make_trainer(opt, device="cpu"):
    """A Trainer instance with only the attributes the loss methods read."""
    T = load_reference()
    from ppeadepth.layers import SSIM, BackprojectDepth, Project3D
    tr = object.__new__(T.Trainer)
    full = types.SimpleNamespace(**vars(opt))
    for k, v in dict(loss_pct=False, debug=False).items():
        if not hasattr(full, k):
            setattr(full, k, v)
    tr.opt = full
    tr.device, tr.step, tr.is_main = torch.device(device), 1, False
    tr.ssim = SSIM().to(device)
    tr.backproject_depth, tr.project_3d = {}, {}
    for s in range(full.sclm + 1):
        h, w = full.height // 2 ** s, full.width // 2 ** s
        tr.backproject_depth[s] = BackprojectDepth(full.batch_size, h, w).to(device)
        tr.project_3d[s] = Project3D(full.batch_size, h, w).to(device)
    return tr


class _NoiseFeed:
    """Replaces torch.randn inside compute_losses with a fixed list of draws
    (trainer.py:1086 draws one (B,1,H,W) tensor per scale)."""

    def __init__(self, draws):
        self.draws, self.i, self.orig = list(draws), 0, None

    def __enter__(self):
        self.orig = torch.randn

        def fed(*shape, **kw):
            z = self.draws[self.i]
            self.i += 1
            shp = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert tuple(z.shape) == shp, (z.shape, shp)
            return z.clone()
        torch.randn = fed
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


def run_reference(inputs, outputs, opt, is_multi=False, noise=None, dtype=torch.float32):
    """generate_images_pred + compute_losses + backward with the reference's code.
    Returns (losses, grads, outputs_dict_after)."""
    from oracle.vsl_oracle import clone_batch
    tr = make_trainer(opt)
    if dtype != torch.float32:
        tr.ssim = tr.ssim.to(dtype)
        for s in tr.backproject_depth:
            tr.backproject_depth[s] = tr.backproject_depth[s].to(dtype)
    ins, outs = clone_batch(inputs, outputs, dtype=dtype)
    tr.generate_images_pred(ins, outs, is_multi)
    if noise is not None:
        with _NoiseFeed([z.to(dtype) for z in noise]):
            losses, _ = tr.compute_losses(ins, outs, is_multi)
    else:
        losses, _ = tr.compute_losses(ins, outs, is_multi)
    losses["loss"].backward()
    grads = {k: v.grad for k, v in outs.items()
             if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam") and getattr(v, "grad", None) is not None}
    return {k: v.detach() for k, v in losses.items()}, grads, (tr, ins, outs)
