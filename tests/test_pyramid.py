"""Input format, remainder: the dataset's LANCZOS pyramid (datasets/mono_dataset.py:79-85, :96-112) and the
packed RGBx frame format.  Integer work: the bar is bit-exact.

not gpu: the oracle restatement of Pillow's resampler against the fixtures the reference's own MonoDataset.preprocess produced,
against Pillow live, and the library's HOST tap tables against the oracle's.  gpu: the CUDA resize / pyramid / packing through the
C ABI against the fixtures, the oracle and Pillow, at fixture sizes and at the full KITTI sizes."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import pyramid_oracle as O

CASES = ["pyramid_kitti_like_93x310_to_48x160", "pyramid_ragged_57x83_to_40x72"]


def _load(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    fx = _load(name)
    c = fx["case"]
    pyr = O.pyramid(fx["raw"].numpy(), c["height"], c["width"], c["scales"])
    for i in range(c["scales"]):
        assert np.array_equal(pyr[i], fx["scale%d" % i].numpy()), i


@pytest.mark.parametrize("shape", [(375, 1242, 192, 640), (192, 640, 96, 320), (33, 47, 20, 31), (20, 30, 40, 45), (64, 64, 64, 32), (7, 9, 3, 4)])
def test_oracle_matches_pillow_live(shape):
    pytest.importorskip("PIL")
    H, W, h, w = shape
    img = np.random.default_rng(H * W).integers(0, 256, (3, H, W), dtype=np.uint8)
    assert np.array_equal(O.resize_lanczos_u8(img, (h, w)), O.pil_resize(img, (h, w)))


@pytest.mark.parametrize("sizes", [(1242, 640), (640, 320), (375, 192), (47, 31), (30, 45), (64, 64), (5, 2)])
def test_host_tap_tables_equal_the_oracle(sizes):
    from ppea_depth_b200 import _cabi as C
    n_in, n_out = sizes
    ksize = C.lib().ppea_lanczos_ksize(n_in, n_out)
    bounds = np.zeros((n_out, 2), np.int32)
    coeffs = np.zeros((n_out, ksize), np.int32)
    assert C.lib().ppea_lanczos_table(n_in, n_out, bounds.ctypes.data_as(ctypes.c_void_p), coeffs.ctypes.data_as(ctypes.c_void_p)) == ksize
    ob, oc = O.lanczos_table(n_in, n_out)
    assert oc.shape[1] == ksize and np.array_equal(bounds, ob) and np.array_equal(coeffs, oc)
    assert C.lib().ppea_lanczos_ksize(0, 4) < 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_pyramid_against_reference_fixture(name):
    import ppea_depth_b200 as P
    fx = _load(name)
    c = fx["case"]
    raw = fx["raw"].unsqueeze(0).cuda()
    pyr = P.ImagePyramid(c["height"], c["width"], c["scales"])(raw)
    flt = P.ImagePyramid(c["height"], c["width"], c["scales"])(raw, as_float=True)
    for i in range(c["scales"]):
        assert torch.equal(pyr[i][0].cpu(), fx["scale%d" % i]), i
        assert torch.equal(flt[i][0].cpu(), fx["scale%d" % i].float() / 255), i      # ToTensor of the reference's PIL image


@pytest.mark.gpu
def test_cuda_resize_full_size_and_edge_shapes():
    import ppea_depth_b200 as P
    rng = np.random.default_rng(3)
    for (n, H, W, h, w) in [(4, 375, 1242, 192, 640), (36, 192, 640, 96, 320), (2, 20, 30, 40, 45), (1, 64, 64, 64, 32), (1, 64, 64, 16, 64),
                            (1, 7, 9, 3, 4), (1, 12, 12, 12, 12)]:
        img = rng.integers(0, 256, (n, 3, H, W), dtype=np.uint8)
        out = P.resize_lanczos_u8(torch.from_numpy(img).cuda(), (h, w)).cpu().numpy()
        assert np.array_equal(out, O.resize_lanczos_u8(img, (h, w))), (n, H, W, h, w)
        try:
            assert np.array_equal(out[0], O.pil_resize(img[0], (h, w)))             # Pillow itself, where it is installed
        except ImportError:
            pass
    # saturated inputs: the negative Lanczos lobes overshoot and must clip exactly like PIL's clip8
    img = np.zeros((1, 3, 40, 60), np.uint8)
    img[..., ::3, :] = 255
    img[..., :, ::4] = 255
    out = P.resize_lanczos_u8(torch.from_numpy(img).cuda(), (20, 30)).cpu().numpy()
    assert np.array_equal(out, O.resize_lanczos_u8(img, (20, 30)))
    with pytest.raises(RuntimeError):
        P.resize_lanczos_u8(torch.from_numpy(img), (20, 30))                        # no CPU path


@pytest.mark.gpu
def test_pack_rgbx_is_the_gather_format():
    import ppea_depth_b200 as P
    g = torch.Generator().manual_seed(0)
    planar = torch.randint(0, 256, (3, 3, 33, 47), generator=g, dtype=torch.uint8)
    want = planar[:, 0].int() | (planar[:, 1].int() << 8) | (planar[:, 2].int() << 16)
    assert torch.equal(P.pack_rgbx(planar.cuda()).cpu(), want)
    assert torch.equal(P.pack_rgbx(planar.permute(0, 2, 3, 1).contiguous().cuda()).cpu(), want)      # the PIL / dataset layout (H,W,3)
