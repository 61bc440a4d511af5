"""GPU tests of the unified conv kernel (csrc/conv3x3.cu: conv_impl 4 = 64-channel K-blocks, 5 = 32-channel K-blocks)
against the round-1 kernels it replaces: same arithmetic in the same order, so every layer output and every image must
be BIT-IDENTICAL to conv_impl 2 (conv_tc2.cu, validated against the reference goldens) resp. conv_impl 3."""
import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from common import diff_stats, golden, manifest, model_paths, read_convs

pytestmark = pytest.mark.gpu
CASES = manifest()["cases"]


def _engine(tta=False, **opts):
    e = rb.RealSR(0, tta_mode=tta)
    p, b = model_paths("models-DF2K")
    assert e.load(p, b) == 0, e.last_error()
    for k, v in opts.items():
        assert e.set_option(k, v) == 0, e.last_error()
    return e


LAYERS = [(1, 37, 45), (2, 17, 31), (3, 64, 90), (4, 40, 61), (5, 33, 64), (346, 24, 30), (347, 50, 33), (348, 31, 95), (349, 77, 40), (350, 29, 47)]
IMAGES = [(70, 93, 3, 32), (256, 256, 3, 200), (120, 90, 4, 64), (300, 500, 3, 200), (1080, 1920, 3, 400)]


@pytest.mark.parametrize("old,new", [(2, 4), (3, 5)])
def test_layers_bit_identical(models_dir, old, new):
    ref, eng = _engine(conv_impl=old), _engine(conv_impl=new)
    convs = read_convs("models-DF2K")
    for idx, H, W in LAYERS:
        c = convs[idx]
        x = (np.random.default_rng(idx).standard_normal((H, W, c["cin"])) * 0.5).astype(np.float16)
        assert np.array_equal(ref.debug_conv(idx, x, c["cout"]), eng.debug_conv(idx, x, c["cout"])), idx


@pytest.mark.parametrize("old,new,opts", [(2, 4, {}), (2, 4, dict(merge_shapes=1)), (2, 4, dict(residual_fp16=0)), (3, 5, dict(merge_shapes=1))],
                         ids=["k64", "k64-merged", "k64-res32", "k32-merged"])
def test_images_bit_identical(models_dir, old, new, opts):
    base = {k: v for k, v in opts.items() if k != "merge_shapes"}
    ref, eng = _engine(conv_impl=old, **base), _engine(conv_impl=new, **opts)
    rng = np.random.default_rng(7)
    for (h, w, c, tile) in IMAGES:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = np.zeros((4 * h, 4 * w, c), np.uint8)
        b = np.zeros_like(a)
        ref.tilesize = eng.tilesize = tile
        assert ref.process(img, a) == 0 and eng.process(img, b) == 0, eng.last_error()
        assert np.array_equal(a, b), (h, w, c, tile, diff_stats(a, b))


@pytest.mark.parametrize("new", [4, 5])
def test_tta_bit_identical(models_dir, new):
    a, b = _engine(tta=True, conv_impl=new - 2), _engine(tta=True, conv_impl=new, merge_shapes=1)
    rng = np.random.default_rng(17)
    for (h, w, c, tile) in [(45, 61, 3, 32), (70, 50, 4, 32), (200, 330, 3, 200)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        oa, ob = np.zeros((4 * h, 4 * w, c), np.uint8), np.zeros((4 * h, 4 * w, c), np.uint8)
        a.tilesize = b.tilesize = tile
        assert a.process(img, oa) == 0 and b.process(img, ob) == 0, b.last_error()
        assert np.array_equal(oa, ob), (h, w, c, tile, diff_stats(oa, ob))
