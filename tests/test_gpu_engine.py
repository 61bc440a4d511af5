"""GPU tests of the engine's scheduling paths (run with `pytest -m gpu` on a B200): merged multi-shape launches, tile-granular
sharding of one image over several handles, and the C++ command line on the real library.
Numerical parity against the reference is in test_gpu_parity.py; here every check is byte identity between two schedules
of the same arithmetic."""
import os

import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from common import diff_stats, golden, manifest, model_paths

pytestmark = pytest.mark.gpu
CASES = manifest()["cases"]


def _engine(tta=False, **opts):
    e = rb.RealSR(0, tta_mode=tta)
    p, b = model_paths("models-DF2K")
    assert e.load(p, b) == 0, e.last_error()
    for k, v in opts.items():
        assert e.set_option(k, v) == 0, e.last_error()
    return e


def test_merged_shape_launches_are_bit_identical(models_dir):
    """merge_shapes=1 (default) runs every layer as ONE launch over all shape groups of a frame (concatenated activation
    buffers, per-group tensor maps); merge_shapes=0 plans one job per shape group: same per-tile arithmetic, so the images
    must be equal bit for bit.  Also checks the launch count of a 1080p frame: 350 tensor-core convs + per-group Conv_0 /
    upsample launches"""
    ref, new = _engine(merge_shapes=0), _engine()
    rng = np.random.default_rng(13)
    for (h, w, c, tile) in [(70, 93, 3, 32), (256, 256, 3, 200), (120, 90, 4, 64), (1080, 1920, 3, 400), (300, 500, 3, 200)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = np.zeros((4 * h, 4 * w, c), np.uint8)
        b = np.zeros_like(a)
        ref.tilesize = new.tilesize = tile
        assert ref.process(img, a) == 0 and new.process(img, b) == 0, new.last_error()
        assert np.array_equal(a, b), (h, w, c, tile, diff_stats(a, b))
    before = new.kernel_launches()
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    out = np.zeros((4320, 7680, 3), np.uint8)
    new.tilesize = 400
    assert new.process(img, out) == 0
    assert new.kernel_launches() - before == 350 + 3 * 4 and new.last_conv_launches() == 350


def test_merged_shape_launches_tta_bit_identical(models_dir):
    """TTA: both orientations (variants 0..3 as H x W, 4..7 as W x H) of the shape chunks share one launch per layer; result
    slots are addressed through the per-group TTA offset"""
    a, b = _engine(tta=True, merge_shapes=0), _engine(tta=True)
    rng = np.random.default_rng(17)
    for (h, w, c, tile) in [(45, 61, 3, 32), (70, 50, 4, 32), (200, 330, 3, 200)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        oa, ob = np.zeros((4 * h, 4 * w, c), np.uint8), np.zeros((4 * h, 4 * w, c), np.uint8)
        a.tilesize = b.tilesize = tile
        assert a.process(img, oa) == 0 and b.process(img, ob) == 0, b.last_error()
        assert np.array_equal(oa, ob), (h, w, c, tile, diff_stats(oa, ob))


def test_max_batch_does_not_change_the_result(models_dir):
    """tiles per job (max_batch) only changes how many tiles share a launch"""
    a, b = _engine(), _engine(max_batch=1)
    img = np.random.default_rng(5).integers(0, 256, (300, 500, 3), dtype=np.uint8)
    oa, ob = np.zeros((1200, 2000, 3), np.uint8), np.zeros((1200, 2000, 3), np.uint8)
    a.tilesize = b.tilesize = 200
    assert a.process(img, oa) == 0 and b.process(img, ob) == 0, b.last_error()
    assert np.array_equal(oa, ob)


@pytest.mark.parametrize("shape,tile,parts", [((150, 130, 3), 32, 3), ((90, 200, 4), 64, 2), ((1080, 1920, 3), 400, 4)])
def test_tile_sharded_process_is_byte_identical(models_dir, shape, tile, parts):
    """sharded.process_sharded: several handles (here all on GPU 0, one thread each) each run a contiguous tile range of
    ONE image and copy only their regions back; the assembled buffer equals a plain process() call byte for byte"""
    engines = [_engine() for _ in range(parts)]
    for e in engines:
        e.tilesize = tile
    img = np.random.default_rng(11).integers(0, 256, shape, dtype=np.uint8)
    ref = np.zeros((shape[0] * 4, shape[1] * 4, shape[2]), np.uint8)
    out = np.full_like(ref, 0x5A)
    assert engines[0].process(img, ref) == 0
    assert rb.process_sharded(engines, img, out) == 0, [e.last_error() for e in engines]
    assert np.array_equal(out, ref)
    # ranges outside the grid are refused
    n = engines[0].tile_count(shape[1], shape[0])
    assert engines[0].process_tiles(img, out, n - 1, 2) != 0 and "tile range" in engines[0].last_error()


def test_cpp_cli_png_pipeline(tmp_path, models_dir):
    """the C++ command line (host/realsr_cli.cpp, PNG through host/png_codec.h) on the real library: file -> file equals
    the engine's own output for the same input, dir -> dir with two proc threads writes every file"""
    import subprocess
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "realsr-ncnn-vulkan_b200", "realsr-b200")
    img, gold = golden("natural_48x40_df2k_t32")
    src, dst = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    Image.fromarray(img).save(src)
    mdir = os.path.join(root, "models", "models-DF2K")
    r = subprocess.run([exe, "-i", src, "-o", dst, "-m", mdir, "-t", "32", "-v"], capture_output=True, text=True)
    assert r.returncode == 0 and " done" in r.stderr, r.stderr
    e = _engine()
    e.tilesize = 32
    ref = np.zeros_like(gold)
    assert e.process(img, ref) == 0
    assert np.array_equal(np.asarray(Image.open(dst)), ref)
    din, dout = tmp_path / "in", tmp_path / "out"
    din.mkdir(); dout.mkdir()
    for i in range(5):
        Image.fromarray(np.random.default_rng(i).integers(0, 256, (20 + i, 30, 3), dtype=np.uint8)).save(str(din / ("%d.png" % i)))
    r = subprocess.run([exe, "-i", str(din), "-o", str(dout), "-m", mdir, "-j", "2:2:2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sorted(os.listdir(str(dout))) == ["%d.png" % i for i in range(5)]
