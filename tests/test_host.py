"""CPU: host-side logic above the C ABI -- flag mapping, drop-in surface, no-CPU-fallback behaviour,
synthetic batch generator, algorithmic-byte accounting."""
from types import SimpleNamespace

import pytest
import os

import torch

import ppea_depth_b200 as P
from ppea_depth_b200 import _cabi as C
from ppea_depth_b200.functional import VslConfig
from ppea_depth_b200.synth import SynthConfig, algorithmic_bytes, make_batch, make_noise


def test_public_surface_mirrors_the_reference():
    for name in ("BackprojectDepth", "Project3D", "SSIM", "get_smooth_loss", "disp_to_depth", "upsample",
                 "transformation_from_parameters", "generate_images_pred", "compute_reprojection_loss",
                 "compute_loss_masks", "compute_losses", "install"):
        assert hasattr(P, name), name
    import inspect
    assert list(inspect.signature(P.generate_images_pred).parameters) == ["self", "inputs", "outputs", "is_multi"]
    assert list(inspect.signature(P.compute_losses).parameters) == ["self", "inputs", "outputs", "is_multi"]
    assert list(inspect.signature(P.compute_reprojection_loss).parameters) == ["self", "pred", "target"]
    assert list(inspect.signature(P.Project3D.__init__).parameters) == ["self", "batch_size", "height", "width", "dc", "eps"]
    assert list(inspect.signature(P.BackprojectDepth.__init__).parameters) == ["self", "batch_size", "height", "width"]


def test_install_rebinds_trainer_methods():
    class FakeTrainer:
        def compute_losses(self, inputs, outputs, is_multi=False):
            return "reference"
    P.install(FakeTrainer, deterministic=True)
    assert FakeTrainer.compute_losses is P.compute_losses
    assert FakeTrainer.generate_images_pred is P.generate_images_pred
    assert FakeTrainer.ppea_deterministic is True
    m = FakeTrainer.compute_loss_masks(torch.tensor([1.0, 3.0]), torch.tensor([2.0, 2.0]))
    assert m.tolist() == [1.0, 0.0]
    assert FakeTrainer.compute_loss_masks(torch.ones(2), None).tolist() == [1.0, 1.0]
    assert FakeTrainer.compute_loss_masks(torch.tensor([2.0]), torch.tensor([2.0])).tolist() == [1.0]   # first-min tie rule


def test_bit_reproducible_gradients_are_the_default_of_the_host_mirror():
    """install() / ViewSynthesisLoss default to the fixed-point (deterministic) coarse-scale fields; the low-level VslConfig
    stays explicit; bench.py follows the host mirror and `--float-atomics` opts out."""
    import subprocess
    import sys
    from types import SimpleNamespace

    class FakeTrainer:
        pass
    P.install(FakeTrainer)
    assert FakeTrainer.ppea_deterministic is True
    assert P.ViewSynthesisLoss(SimpleNamespace()).ppea_deterministic is True
    assert P.ViewSynthesisLoss(SimpleNamespace(), deterministic=False).ppea_deterministic is False
    assert VslConfig().deterministic is False
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.argv = ['bench.py'] + %r; import bench; a = bench.parse(); print(int(a.deterministic))")
    for argv, want in (([], "1"), (["--deterministic"], "1"), (["--float-atomics"], "0")):
        out = subprocess.run([sys.executable, "-c", code % (argv,)], cwd=root, capture_output=True, text=True)
        assert out.stdout.strip().splitlines()[-1] == want, (argv, out.stdout, out.stderr[-500:])


def test_flags():
    f = VslConfig().flags(grad_pose=True)
    assert f == C.F_AUTOMASK | C.F_SELEC_REPROJ | C.F_MOTION_MASK | C.F_MATCH_AUG | C.F_GRAD_POSE
    f = VslConfig(is_multi=True, selec_reproj=False, no_ssim=True, deterministic=True, motion_mask=False).flags(False)
    assert f == C.F_MULTI | C.F_AUTOMASK | C.F_NO_SSIM | C.F_DETERMINISTIC | C.F_MATCH_AUG


def test_cpu_tensors_raise_instead_of_falling_back():
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.SSIM()(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.get_smooth_loss(torch.rand(1, 1, 8, 8), x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.BackprojectDepth(1, 8, 8)(torch.rand(1, 1, 8, 8), torch.eye(4)[None])
    opt = SimpleNamespace(sclm=0, v1_multiscale=False, height=16, width=32, min_depth=0.1, max_depth=100.0,
                          frame_ids=[0, -1, 1], disable_automasking=False, no_ssim=False, selec_reproj=True,
                          disable_motion_masking=False, no_matching_augmentation=False, batch_size=1,
                          disparity_smoothness=1e-3)
    inputs, outputs = make_batch(SynthConfig(batch=1, height=16, width=32, num_scales=1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.ViewSynthesisLoss(opt).generate_images_pred(inputs, outputs, False)


def test_only_two_source_configuration_is_fused():
    opt = SimpleNamespace(sclm=0, frame_ids=[0, -1], batch_size=1)
    with pytest.raises((NotImplementedError, IndexError)):
        P.ViewSynthesisLoss(opt).generate_images_pred({}, {("disp", 0): torch.zeros(1)}, False)


def test_pose_helpers_match_known_answers():
    aa = torch.zeros(2, 1, 3)
    tr = torch.tensor([[[1.0, 2.0, 3.0]], [[0.0, 0.0, 0.0]]])
    T = P.transformation_from_parameters(aa, tr)
    assert torch.allclose(T[0, :3, 3], torch.tensor([1.0, 2.0, 3.0]))
    assert torch.allclose(T[:, :3, :3], torch.eye(3).expand(2, 3, 3))
    aa = 0.3 * torch.randn(4, 1, 3)
    tr = torch.randn(4, 1, 3)
    prod = P.transformation_from_parameters(aa, tr, invert=True) @ P.transformation_from_parameters(aa, tr)
    assert torch.allclose(prod, torch.eye(4).expand(4, 4, 4), atol=1e-5)        # layers.py:299-307 smoke block
    _, depth = P.disp_to_depth(torch.tensor([0.0, 1.0]), 0.1, 100.0)
    assert torch.allclose(depth, torch.tensor([100.0, 0.1]))
    assert P.upsample(torch.zeros(1, 1, 2, 3)).shape == (1, 1, 4, 6)


def test_synthetic_batch_is_deterministic_and_well_formed():
    cfg = SynthConfig(batch=2, height=32, width=64, num_scales=3, seed=4)
    a_in, a_out = make_batch(cfg)
    b_in, b_out = make_batch(cfg)
    for k in a_in:
        assert torch.equal(a_in[k], b_in[k]), k
    img = a_in[("color", 0, 0)]
    assert img.shape == (2, 3, 32, 64) and float(img.min()) >= 0 and float(img.max()) <= 1
    assert torch.equal(img, torch.round(img * 255) / 255)                       # exact k/255 like ToTensor(uint8)
    assert float((a_in[("color", -1, 0)].sum(1) == 0).float().mean()) > 0.005   # dark pixels for selec_reproj
    assert a_out[("disp", 2)].shape == (2, 1, 8, 16)
    assert torch.allclose(a_in[("K", 1)][0] @ a_in[("inv_K", 1)][0], torch.eye(4), atol=1e-4)
    n = make_noise(cfg, 3)
    assert len(n) == 3 and n[0].shape == (2, 1, 32, 64)


def test_algorithmic_bytes_match_survey_table():
    n = 12 * 192 * 640
    assert abs(algorithmic_bytes(12, 192, 640, 4) / n - 375.9) < 0.05
    assert abs(algorithmic_bytes(12, 192, 640, 4, deterministic=True) / n - 407.9) < 0.05
    assert abs(algorithmic_bytes(12, 192, 640, 4, is_multi=True) / n - 423.9) < 0.05


def test_streaming_step_row_partition_covers_every_column_once():
    """The work split of the streaming step (vsl_common.cuh stream_chunk_rows / stream_pieces, vsl_stream.cu task decode), restated:
    the rows of all (image, strip, scale) columns form one line cut into chunks of L rows; a chunk crossing a column end yields two
    pieces; piece k of a column owns tile slot k, the column's last piece clears the slots above it.  Every (column, slot) is
    written or cleared exactly once, the pieces of a column tile its rows, and no piece index reaches stream_pieces(H, L)."""
    import random
    rnd = random.Random(0)
    for _ in range(400):
        B, strips, S, H, L = rnd.randint(1, 5), rnd.randint(1, 6), rnd.randint(1, 4), rnd.randint(4, 200), rnd.randint(16, 400)
        pmax = (H - 1) // L + 2
        total = B * strips * S * H
        written, cleared = {}, set()
        for j in range(-(-total // L)):
            g, g_end = j * L, min(j * L + L, total)
            while g < g_end:
                col = g // H
                y0 = g - col * H
                y1 = min(H, y0 + (g_end - g))
                g += y1 - y0
                piece = j - (col * H) // L
                assert 0 <= piece < pmax and (col, piece) not in written
                written[(col, piece)] = (y0, y1)
                if y1 == H:
                    cleared.update((col, k) for k in range(piece + 1, pmax))
        for c in range(B * strips * S):
            rows = [written[(c, k)] for k in range(pmax) if (c, k) in written]
            assert all(((c, k) in written) != ((c, k) in cleared) for k in range(pmax))
            assert rows[0][0] == 0 and rows[-1][1] == H and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
