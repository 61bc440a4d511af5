"""GPU parity tests (run with `pytest -m gpu` on a B200).  Everything goes through the C ABI (librealsr_b200.so) via the
Python mirror of the reference's `class RealSR`.

Tolerance contract (vs the reference's own CPU path at the same tile size; stated here because the path is
floating point: fp16 tensor-core operands with fp32 accumulation, fp32 Conv_0, fp32 RRDB skip stream, dense-block residual
carried as fp16 + int8 extension):
  * natural images : every byte within 1 LSB, >= 97 % of bytes exact
  * uniform noise  : max 2 LSB, at most 1 byte in 10 000 off by more than 1 LSB
(measured on B200, tools/parity_report.py: natural 0.6 - 2.4 % of bytes off by one; noise 5 - 14 %, 3 - 8 bytes of 3 M at 2 LSB)
Per-layer: abs-or-relative 2e-3 against an fp32 torch conv on the same fp16 operands (ncnn's own Winograd-eligible
conv tolerance, ncnn/tests/test_convolution.cpp:42-52), result rounded to fp16 as the engine stores it.
"""
import threading

import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from common import Oracle, diff_stats, golden, manifest, model_paths, read_convs

pytestmark = pytest.mark.gpu

_ENG = {}


def engine(model, tta=False):
    key = (model, tta)
    if key not in _ENG:
        e = rb.RealSR(0, tta_mode=tta)
        p, b = model_paths(model)
        assert e.load(p, b) == 0, e.last_error()
        _ENG[key] = e
    return _ENG[key]


def run(model, img, tilesize, tta=False):
    e = engine(model, tta)
    e.tilesize = tilesize
    out = np.zeros((img.shape[0] * 4, img.shape[1] * 4, img.shape[2]), dtype=np.uint8)
    assert e.process(img, out) == 0, e.last_error()
    return out


def check_contract(s, natural):
    if natural:
        assert s["max"] <= 1 and s["frac_diff"] <= 0.03, s
    else:
        assert s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 10000), s


CASES = manifest()["cases"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_vs_reference_cpu_path(case, models_dir):
    """CUDA path vs the committed outputs of the real reference CPU path (oracle/_ref), same tile size"""
    img, gold = golden(case["name"])
    out = run(case["model"], img, case["tilesize"], bool(case["tta"]))
    check_contract(diff_stats(out[..., :3], gold[..., :3]), natural="noise" not in case["name"])
    if img.shape[2] == 4:  # alpha: bicubic x4 in fp32 with the reference's operation order -> expected bit-exact, 1 LSB allowed
        s = diff_stats(out[..., 3], gold[..., 3])
        assert s["max"] <= 1 and s["frac_diff"] <= 0.001, s


@pytest.mark.parametrize("idx,H,W", [(1, 37, 45), (2, 16, 30), (3, 17, 31), (4, 64, 61), (5, 33, 95), (5, 8, 29), (346, 40, 40), (347, 35, 62), (348, 23, 91),
                                     (349, 70, 33), (350, 29, 47), (1, 5, 3), (4, 130, 7)])
def test_conv_layer_vs_torch(idx, H, W, models_dir):
    """one 3x3 conv of the model on the tcgen05 kernel vs torch fp32 conv2d on identical fp16 operands"""
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    c = read_convs("models-DF2K")[idx]
    rng = np.random.default_rng(idx * 1000 + H)
    x = (rng.standard_normal((H, W, c["cin"])) * 0.5).astype(np.float16)
    y = engine("models-DF2K").debug_conv(idx, x, c["cout"])
    xt = torch.from_numpy(x.astype(np.float32)).cuda().permute(2, 0, 1)[None]
    ref = F.conv2d(xt, torch.from_numpy(c["w"].astype(np.float32)).cuda(), torch.from_numpy(c["b"].copy()).cuda(), padding=1)
    if c["act"] == 2:
        ref = F.leaky_relu(ref, 0.2)
    ref = ref[0].permute(1, 2, 0).cpu().numpy()
    if c["cout"] != 3:
        ref = ref.astype(np.float16).astype(np.float32)
    err = np.abs(y - ref)
    assert np.all((err <= 2e-3) | (err <= 2e-3 * np.abs(ref))), float(err.max())


@pytest.mark.parametrize("w,h,tile,tta", [(16, 13, 32, False), (33, 47, 32, False), (61, 32, 32, True), (45, 45, 200, False)])
def test_fresh_inputs_vs_oracle_port(w, h, tile, tta, models_dir):
    """inputs that are not in the golden set, checked against the C restatement run here (edge tiles, ragged sizes)"""
    img = np.random.default_rng(w * 100 + h).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = Oracle("models-DF2K_JPEG").process(img, tile, tta=tta)
    check_contract(diff_stats(run("models-DF2K_JPEG", img, tile, tta), ref), natural=False)


@pytest.mark.parametrize("w,h,tile,tta", [(70, 45, 32, False), (40, 33, 200, True)])
def test_rgba_vs_oracle_port(w, h, tile, tta, models_dir):
    """RGBA incl. several tiles: alpha is up-scaled per un-padded tile region (the reference's GPU-shader semantics,
    realsr_preproc.comp:79-88; its CPU branch is only defined for one tile), colour channels as for RGB"""
    rng = np.random.default_rng(w + h)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    img[..., 3] = np.clip((xx * 5 + yy * 7) % 256, 0, 255).astype(np.uint8)
    img[: h // 3, : w // 3, 3] = 255
    ref = Oracle("models-DF2K").process(img, tile, tta=tta)
    out = run("models-DF2K", img, tile, tta)
    check_contract(diff_stats(out[..., :3], ref[..., :3]), natural=False)
    s = diff_stats(out[..., 3], ref[..., 3])
    assert s["max"] <= 1 and s["frac_diff"] <= 0.001, s


def test_constant_and_extreme_images(models_dir):
    for v in (0, 255, 128):
        img = np.full((40, 52, 3), v, dtype=np.uint8)
        ref = Oracle("models-DF2K").process(img, 32)
        check_contract(diff_stats(run("models-DF2K", img, 32), ref), natural=True)


def test_full_size_properties(models_dir):
    """BASELINE configs[1] size (1920x1080, tile 400), where the CPU oracle would take minutes: size-independent
    properties.  (a) deterministic; (b) device-resident path == host path; (c) a tile's output only depends on its own
    padded input: changing the bottom-right tile leaves the bytes of all non-adjacent tiles unchanged; (d) the top-left
    tile equals a stand-alone run on its padded crop (same tile geometry) -> ties the big run to the small-case parity."""
    import torch
    img = np.random.default_rng(7).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    e = engine("models-DF2K")
    a = run("models-DF2K", img, 400)
    b = run("models-DF2K", img, 400)
    assert np.array_equal(a, b)
    d_in, d_out = torch.from_numpy(img).cuda(), torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
    assert e.process_device(d_in.data_ptr(), 1920, 1080, 3, d_out.data_ptr()) == 0
    assert np.array_equal(d_out.cpu().numpy(), a)
    img2 = img.copy()
    img2[900:, 1700:] = 255 - img2[900:, 1700:]
    c = run("models-DF2K", img2, 400)
    assert np.array_equal(c[:3200, :], a[:3200, :]) and np.array_equal(c[:, :6400], a[:, :6400])
    assert not np.array_equal(c[3600:, 6800:], a[3600:, 6800:])
    crop = np.ascontiguousarray(img[:410, :410])  # tile (0,0) sees rows/cols [0, 410)
    s = run("models-DF2K", crop, 400)
    assert np.array_equal(s[:1600, :1600], a[:1600, :1600])
    assert e.last_padded_pixels() == 420 * 420 + 2 * 420 * 30 + 30 * 30


def test_reentrant_same_handle(models_dir):
    """RealSR::process is const and is called concurrently on one object (src/main.cpp:811-828)"""
    imgs = [np.random.default_rng(50 + i).integers(0, 256, (48, 70, 3), dtype=np.uint8) for i in range(4)]
    serial = [run("models-DF2K", im, 32) for im in imgs]
    e = engine("models-DF2K")
    e.tilesize = 32
    outs = [np.zeros_like(s) for s in serial]
    rcs = [None] * 4

    def work(i):
        rcs[i] = e._lib.rsr_process(e._h, imgs[i].ctypes.data, 70, 48, 3, outs[i].ctypes.data)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert rcs == [0, 0, 0, 0]
    for o, s in zip(outs, serial):
        assert np.array_equal(o, s)


def test_tta_is_dihedral_equivariant(models_dir):
    """TTA averages the 8 dihedral variants, so flipping the input flips the output (up to fp32 summation order)"""
    img = np.random.default_rng(99).integers(0, 256, (36, 36, 3), dtype=np.uint8)
    a = run("models-DF2K", img, 200, tta=True)
    b = run("models-DF2K", np.ascontiguousarray(img[::-1]), 200, tta=True)[::-1]
    assert diff_stats(a, b)["max"] <= 1


def test_errors_are_reported(models_dir):
    e = rb.RealSR(0)
    assert e.process(np.zeros((8, 8, 3), np.uint8), np.zeros((32, 32, 3), np.uint8)) != 0 and "no model" in e.last_error()
    g = engine("models-DF2K")
    assert g.process(np.zeros((8, 8, 2), np.uint8), np.zeros((32, 32, 2), np.uint8)) != 0  # only RGB / RGBA
    assert g.kernel_launches() > 0


def test_cpp_wrapper_cli_concurrent_jobs(tmp_path, models_dir):
    """the C++ `class RealSR` wrapper driven like the reference's proc stage: 3 threads call process() on one object"""
    import os
    import subprocess
    from common import ROOT
    exe = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "realsr_raw_cli")
    if not os.path.exists(exe):
        pytest.skip("realsr_raw_cli not built")
    img, gold = golden("natural_48x40_df2k_t32")
    fi, fo = tmp_path / "i.raw", tmp_path / "o.raw"
    img.tofile(fi)
    r = subprocess.run([exe, os.path.join(models_dir, "models-DF2K"), str(fi), "48", "40", str(fo), "-t", "32", "-j", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(fo, dtype=np.uint8).reshape(gold.shape)
    check_contract(diff_stats(out, gold), natural=True)


def test_cli_directory_mode(tmp_path, models_dir):
    """the reference's command line surface (src/main.cpp) on the B200 engine: dir -> dir, PNG in/out, RGB + RGBA, -t, -j"""
    import os
    from PIL import Image
    import realsr_ncnn_vulkan_b200.cli as cli
    ind, outd = tmp_path / "in", tmp_path / "out"
    ind.mkdir(); outd.mkdir()
    img, gold = golden("natural_48x40_df2k_t32")
    rgba, gold_rgba = golden("rgba_noise_41x30_df2k_t200")
    Image.fromarray(img).save(ind / "a.png")
    Image.fromarray(rgba).save(ind / "b.png")
    rc = cli.main(["-i", str(ind), "-o", str(outd), "-m", os.path.join(models_dir, "models-DF2K"), "-t", "32", "-j", "2:2:2", "-v"])
    assert rc == 0
    a = np.asarray(Image.open(outd / "a.png"))
    check_contract(diff_stats(a, gold), natural=True)
    b = np.asarray(Image.open(outd / "b.png"))
    assert b.shape == gold_rgba.shape  # tile 32 here vs 200 in the golden: only shape / alpha-away-from-tile-borders are comparable
    ref = run("models-DF2K", rgba, 32)
    assert np.array_equal(b, ref)
