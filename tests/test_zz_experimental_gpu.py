"""GPU tests of code paths that are compiled and shipped as OPT-IN options but have not yet had their first run on a
B200 (written after the round's GPU budget was spent).  They only run with RSR_TEST_EXPERIMENTAL=1, so that the
default `pytest -m gpu` stays a statement about validated code; tools/r2_first_call.sh runs them first thing next round."""
import os

import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from common import diff_stats, golden, manifest, model_paths, read_convs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("RSR_TEST_EXPERIMENTAL"), reason="opt-in paths: set RSR_TEST_EXPERIMENTAL=1")]

CASES = manifest()["cases"]


def _engine(tta=False, **opts):
    e = rb.RealSR(0, tta_mode=tta)
    p, b = model_paths("models-DF2K")
    assert e.load(p, b) == 0, e.last_error()
    for k, v in opts.items():
        assert e.set_option(k, v) == 0
    return e


@pytest.mark.parametrize("residual_fp16", [1, 0])
def test_split_epilogue_is_bit_identical(models_dir, residual_fp16):
    """epilogue_split=1 (channel-split epilogue, early accumulator hand-back) performs the same arithmetic in the same
    order as the default epilogue: every layer output and the final image must be bit-identical to the default path"""
    ref = _engine(residual_fp16=residual_fp16)
    new = _engine(residual_fp16=residual_fp16, epilogue_split=1)
    convs = read_convs("models-DF2K")
    for idx, H, W in [(1, 37, 45), (2, 17, 31), (3, 64, 90), (4, 40, 61), (5, 33, 64), (346, 24, 30), (347, 50, 33), (350, 29, 47)]:
        c = convs[idx]
        x = (np.random.default_rng(idx).standard_normal((H, W, c["cin"])) * 0.5).astype(np.float16)
        assert np.array_equal(ref.debug_conv(idx, x, c["cout"]), new.debug_conv(idx, x, c["cout"])), idx
    rng = np.random.default_rng(7)
    for (h, w, c, tile) in [(70, 93, 3, 32), (256, 256, 3, 200), (120, 90, 4, 64), (1080, 1920, 3, 400)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = np.zeros((4 * h, 4 * w, c), np.uint8)
        b = np.zeros_like(a)
        ref.tilesize = new.tilesize = tile
        assert ref.process(img, a) == 0 and new.process(img, b) == 0, new.last_error()
        assert np.array_equal(a, b), (h, w, c, tile, diff_stats(a, b))


def test_serpentine_tile_order_is_bit_identical(models_dir):
    """serpentine=1 only permutes the order in which a conv walks its work tiles (every second conv backwards)"""
    ref, new = _engine(), _engine(serpentine=1)
    rng = np.random.default_rng(9)
    for (h, w, c, tile) in [(70, 93, 3, 32), (256, 256, 3, 200), (120, 90, 4, 64), (1080, 1920, 3, 400)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = np.zeros((4 * h, 4 * w, c), np.uint8)
        b = np.zeros_like(a)
        ref.tilesize = new.tilesize = tile
        assert ref.process(img, a) == 0 and new.process(img, b) == 0, new.last_error()
        assert np.array_equal(a, b), (h, w, c, tile, diff_stats(a, b))


@pytest.mark.parametrize("opts", [dict(merge_shapes=1), dict(merge_shapes=1, epilogue_split=1, serpentine=1)],
                         ids=["merge", "merge+split+serpentine"])
def test_merged_shape_launches_are_bit_identical(models_dir, opts):
    """merge_shapes=1 runs every layer as ONE launch over all shape groups of a frame (concatenated activation buffers,
    per-group tensor maps): same per-tile arithmetic, so the image must equal the default path bit for bit"""
    ref, new = _engine(), _engine(**opts)
    rng = np.random.default_rng(13)
    for (h, w, c, tile) in [(70, 93, 3, 32), (256, 256, 3, 200), (120, 90, 4, 64), (1080, 1920, 3, 400), (300, 500, 3, 200)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        a = np.zeros((4 * h, 4 * w, c), np.uint8)
        b = np.zeros_like(a)
        ref.tilesize = new.tilesize = tile
        assert ref.process(img, a) == 0 and new.process(img, b) == 0, new.last_error()
        assert np.array_equal(a, b), (h, w, c, tile, diff_stats(a, b))
    # the launch count drops to one per layer: 351 tensor-core convs + per-group Conv_0 / upsample launches
    before = new.kernel_launches()
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    out = np.zeros((4320, 7680, 3), np.uint8)
    new.tilesize = 400
    assert new.process(img, out) == 0
    assert new.kernel_launches() - before == 350 + 3 * 4


def test_merged_shape_launches_tta_bit_identical(models_dir):
    """merge_shapes with TTA: both orientations (variants 0..3 as H x W, 4..7 as W x H) of up to two shape chunks share
    one launch per layer; result slots are addressed through the per-group TTA offset"""
    a, b = _engine(tta=True), _engine(tta=True, merge_shapes=1)
    rng = np.random.default_rng(17)
    for (h, w, c, tile) in [(45, 61, 3, 32), (70, 50, 4, 32), (200, 330, 3, 200)]:
        img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        oa, ob = np.zeros((4 * h, 4 * w, c), np.uint8), np.zeros((4 * h, 4 * w, c), np.uint8)
        a.tilesize = b.tilesize = tile
        assert a.process(img, oa) == 0 and b.process(img, ob) == 0, b.last_error()
        assert np.array_equal(oa, ob), (h, w, c, tile, diff_stats(oa, ob))


def test_split_epilogue_tta_and_goldens(models_dir):
    new = _engine(epilogue_split=1)
    for name, natural in (("natural_48x40_df2k_t32", True), ("noise_50x33_df2k_t32", False), ("readme0_crop96x80_df2k_t64", True)):
        case = [c for c in CASES if c["name"] == name][0]
        img, gold = golden(name)
        new.tilesize = case["tilesize"]
        out = np.zeros_like(gold)
        assert new.process(img, out) == 0, new.last_error()
        s = diff_stats(out, gold)
        assert (s["max"] <= 1 and s["frac_diff"] <= 0.05) if natural else (s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 10000)), s
    a, b = _engine(tta=True), _engine(tta=True, epilogue_split=1)
    img = np.random.default_rng(3).integers(0, 256, (45, 61, 3), dtype=np.uint8)
    oa, ob = np.zeros((180, 244, 3), np.uint8), np.zeros((180, 244, 3), np.uint8)
    a.tilesize = b.tilesize = 32
    assert a.process(img, oa) == 0 and b.process(img, ob) == 0
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


@pytest.mark.parametrize("split", [0, 1])
def test_kblock32_kernel_same_contract(models_dir, split):
    """conv_impl=3 (32-channel K-blocks, SWIZZLE_64B boxes) meets the per-layer tolerance against an fp32 torch conv and
    the end-to-end contract against the reference goldens (the accumulation order differs from conv_impl=2, so this is a
    tolerance test, not a bit-identity test)"""
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    e = _engine(conv_impl=3, epilogue_split=split)
    convs = read_convs("models-DF2K")
    for idx, H, W in [(1, 37, 45), (2, 17, 31), (3, 64, 90), (4, 40, 61), (5, 33, 64), (346, 24, 30), (347, 50, 33), (350, 29, 47)]:
        c = convs[idx]
        x = (np.random.default_rng(idx).standard_normal((H, W, c["cin"])) * 0.5).astype(np.float16)
        y = e.debug_conv(idx, x, c["cout"])
        ref = F.conv2d(torch.from_numpy(x.astype(np.float32)).cuda().permute(2, 0, 1)[None], torch.from_numpy(c["w"].astype(np.float32)).cuda(),
                       torch.from_numpy(c["b"].copy()).cuda(), padding=1)
        if c["act"] == 2:
            ref = F.leaky_relu(ref, 0.2)
        ref = ref[0].permute(1, 2, 0).cpu().numpy()
        if c["cout"] != 3:
            ref = ref.astype(np.float16).astype(np.float32)
        err = np.abs(y - ref)
        assert np.all((err <= 2e-3) | (err <= 2e-3 * np.abs(ref))), (idx, float(err.max()))
    for name, natural in (("natural_48x40_df2k_t32", True), ("noise_50x33_df2k_t32", False), ("readme0_crop96x80_df2k_t64", True)):
        case = [c for c in CASES if c["name"] == name][0]
        img, gold = golden(name)
        e.tilesize = case["tilesize"]
        out = np.zeros_like(gold)
        assert e.process(img, out) == 0, e.last_error()
        s = diff_stats(out, gold)
        assert (s["max"] <= 1 and s["frac_diff"] <= 0.05) if natural else (s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 10000)), s


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
