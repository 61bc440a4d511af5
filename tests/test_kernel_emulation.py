"""CPU checks of the engine's CUDA-core kernels: csrc/misc_kernels.cu is compiled for the HOST (tests/emu/cuda_host_emu.h: a thread
block = real threads, __syncthreads = a barrier) and run against
  * the straightforward round-2 version of conv_first_kernel that was validated on B200 (tests/emu/conv_first_simple.cu.inc): bit identity
    of all three outputs (fp32 fea, fp16 x0, int8 extension) over borders, ragged tiles and all eight TTA variants;
  * a numpy model of preproc + Conv_0 (reflect-101 crop, x 1/255, zero-padded 3x3 conv, realsr.cpp:735-766 / x4.param:4);
  * numpy nearest x2 for upsample2x_kernel.
No GPU involved: this pins the index arithmetic and the fp32 summation order of those kernels, not their speed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, oracle_lib, read_convs

EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "csrc")


def build(name, kernels_cu, extra=()):
    so = os.path.join(EMU, name)
    srcs = [os.path.join(EMU, "emu_misc.cpp"), os.path.join(EMU, "cuda_host_emu.h"), kernels_cu, os.path.join(CSRC, "kernels.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include", "-I" + CSRC,
                               "-DRSR_MISC_KERNELS_CU=\"%s\"" % kernels_cu, *extra, "-o", so, srcs[0], "-lpthread"])
    lib = ctypes.CDLL(so)
    assert lib.emu_sizeof_tiledesc() == 40
    return lib


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists("/usr/local/cuda/include/cuda_fp16.h"):
        pytest.skip("CUDA headers not installed")
    return build("libemu_misc.so", os.path.join(CSRC, "misc_kernels.cu")), build("libemu_simple.so", os.path.join(EMU, "conv_first_simple.cu.inc"), ("-DRSR_EMU_CONV_FIRST_ONLY",))


def tile_desc(in_x0, in_y0, base_h, base_w, variant):
    return [in_x0, in_y0, 0, 0, 0, 0, variant, base_h, base_w, 0]


def run_first(lib, img, descs, H, W, w, b):
    B = len(descs)
    td = np.array(descs, dtype=np.int32)
    f32 = np.full((B, H, W, 64), np.nan, np.float32)
    d16 = np.zeros((B, H, W, 64), np.uint16)
    ext = np.zeros((B, H, W, 64), np.uint8)
    img = np.ascontiguousarray(img)
    lib.emu_conv_first(img.ctypes.data_as(ctypes.c_void_p), img.shape[1], img.shape[0], img.shape[2], td.ctypes.data_as(ctypes.c_void_p), B, H, W,
                       w.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), f32.ctypes.data_as(ctypes.c_void_p),
                       d16.ctypes.data_as(ctypes.c_void_p), ext.ctypes.data_as(ctypes.c_void_p))
    return f32, d16, ext


def reflect101(v, n):
    v = np.abs(v)
    v = (n - 1) - np.abs(v - (n - 1))
    return np.clip(v, 0, n - 1)


def numpy_first(img, d, w, b):
    """float64 model of preproc + TTA variant + Conv_0 for one tile descriptor"""
    in_x0, in_y0, variant, bh, bw = d[0], d[1], d[6], d[7], d[8]
    ii, jj = np.meshgrid(np.arange(bh), np.arange(bw), indexing="ij")
    base = img[reflect101(in_y0 + ii, img.shape[0]), reflect101(in_x0 + jj, img.shape[1]), :3].astype(np.float32) * np.float32(1 / 255.0)
    v = {0: base, 1: base[::-1], 2: base[:, ::-1], 3: base[::-1, ::-1], 4: base.transpose(1, 0, 2), 5: base[:, ::-1].transpose(1, 0, 2),
         6: base[::-1].transpose(1, 0, 2), 7: base[::-1, ::-1].transpose(1, 0, 2)}[variant]  # variant_to_base of misc_kernels.cu, inverted
    H, W = v.shape[:2]
    p = np.zeros((H + 2, W + 2, 3), np.float64)
    p[1:-1, 1:-1] = v
    out = np.tile(b.astype(np.float64), (H, W, 1))
    for c in range(3):
        for ky in range(3):
            for kx in range(3):
                out += p[ky:ky + H, kx:kx + W, c, None] * w[:, c, ky, kx].astype(np.float64)
    return out


CASES = [  # (image h, w, channels, padded tile h, w, list of (in_x0, in_y0)) -- tiles hang over every image border (reflect-101)
    (33, 50, 3, 40, 37, [(-10, -10), (20, 3), (23, -4)]),
    (20, 28, 4, 40, 48, [(-10, -10)]),          # RGBA source, tile larger than the image on both sides
    (11, 12, 3, 31, 32, [(-10, -10)]),          # smallest well-defined shapes
    (70, 90, 3, 64, 96, [(5, 3), (-3, 10)]),    # exact multiples of the 32 x 8 block
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("tta", [False, True])
def test_conv_first_matches_the_validated_simple_kernel_and_numpy(libs, case, tta, models_dir):
    new, simple = libs
    ih, iw, ic, bh, bw, origins = CASES[case]
    img = np.random.default_rng(100 + case).integers(0, 256, (ih, iw, ic), dtype=np.uint8)
    c0 = read_convs("models-DF2K_JPEG")[0]
    w, b = np.ascontiguousarray(c0["w"], np.float32), np.ascontiguousarray(c0["b"], np.float32)
    assert w.shape == (64, 3, 3, 3)
    for variants in ([0, 1, 2, 3], [4, 5, 6, 7]) if tta else ([0],):  # variants 4..7 run transposed: a group has ONE padded shape
        descs = [tile_desc(x0, y0, bh, bw, v) for (x0, y0) in (origins[:1] if tta else origins) for v in variants]  # (host threads are slow)
        H, W = (bw, bh) if variants[0] >= 4 else (bh, bw)
        got = run_first(new, img, descs, H, W, w, b)
        ref = run_first(simple, img, descs, H, W, w, b)
        assert not np.isnan(got[0]).any()
        for g, r, what in zip(got, ref, ("fp32 fea", "fp16 x0", "int8 extension")):
            assert np.array_equal(g.view(np.uint8), r.view(np.uint8)), what
        for k, d in enumerate(descs):
            model = numpy_first(img, d, w, b)
            assert np.abs(got[0][k] - model).max() < 2e-6
        # carrier: fp16 + extension/256 ulp reproduces the fp32 value to 1/256 of an fp16 ulp (kernels.h res_ext_encode)
        h16 = got[1].view(np.float16).astype(np.float32)
        dec = (h16.view(np.int32) + got[2].view(np.int8).astype(np.int32) * 32).view(np.float32)
        big = np.abs(got[0]) > 2.0 ** -14
        assert (np.abs(dec - got[0])[big] <= np.abs(got[0])[big] * 2.0 ** -18).all()


def test_upsample2x_is_nearest(libs):
    new, _ = libs
    B, H, W = 2, 5, 7
    x = np.random.default_rng(3).integers(0, 65536, (B, H, W, 64), dtype=np.uint16)
    out = np.zeros((B, 2 * H, 2 * W, 64), np.uint16)
    new.emu_upsample2x(x.ctypes.data_as(ctypes.c_void_p), 64, out.ctypes.data_as(ctypes.c_void_p), B, H, W)
    assert np.array_equal(out, x.repeat(2, axis=1).repeat(2, axis=2))  # ncnn Interp nearest: in = out >> 1 (interp.cpp:598-618)


@pytest.mark.parametrize("w,h,tile", [(41, 30, 200), (36, 28, 16), (9, 7, 200), (40, 38, 32)])
def test_alpha_bicubic_matches_the_oracle_per_tile(libs, w, h, tile):
    """alpha of RGBA images: bicubic x4 of the UN-padded tile region, truncated (realsr.cpp:431-442; ncnn interp.cpp:178-249,:645-668)"""
    new, _ = libs
    rng = np.random.default_rng(w * 100 + h)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[: h // 2, : w // 2, 3] = 255  # a flat region: must stay exactly 255
    out = np.zeros((4 * h, 4 * w, 4), np.uint8)
    ntiles = ((w + tile - 1) // tile) * ((h + tile - 1) // tile)
    new.emu_alpha_bicubic4x(img.ctypes.data_as(ctypes.c_void_p), w, h, tile, 0, ntiles, out.ctypes.data_as(ctypes.c_void_p))
    lib = oracle_lib()
    lib.rsr_oracle_bicubic4x.restype = None
    lib.rsr_oracle_bicubic4x.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    want = np.zeros((4 * h, 4 * w), np.uint8)
    for y0 in range(0, h, tile):
        for x0 in range(0, w, tile):
            src = np.ascontiguousarray(img[y0:y0 + tile, x0:x0 + tile, 3].astype(np.float32))
            th, tw = src.shape
            if th < 4 or tw < 4:  # ncnn's border folding needs >= 4 source samples: the reference is undefined below that
                pytest.skip("tile narrower than the cubic support")
            dst = np.zeros((4 * th, 4 * tw), np.float32)
            lib.rsr_oracle_bicubic4x(src.ctypes.data_as(ctypes.c_void_p), tw, th, dst.ctypes.data_as(ctypes.c_void_p))
            want[4 * y0:4 * (y0 + th), 4 * x0:4 * (x0 + tw)] = np.clip(dst.astype(np.int32), 0, 255)  # (int) truncation, then clamp
    assert np.array_equal(out[:, :, 3], want)
    assert (out[: 2 * h - 8, : 2 * w - 8, 3] == 255).all() if tile >= max(w, h) else True
    assert not out[:, :, :3].any()  # the kernel touches the alpha byte only

    # a handle that owns only a tile range (rsr_process_tiles) writes only its tiles' pixels
    if ntiles > 1:
        part = np.zeros_like(out)
        new.emu_alpha_bicubic4x(img.ctypes.data_as(ctypes.c_void_p), w, h, tile, 1, ntiles - 1, part.ctypes.data_as(ctypes.c_void_p))
        t0w, t0h = min(tile, w), min(tile, h)
        assert not part[: 4 * t0h, : 4 * t0w].any()
        part[: 4 * t0h, : 4 * t0w] = out[: 4 * t0h, : 4 * t0w]
        assert np.array_equal(part, out)


def test_tta_gather_inverts_the_eight_variants(libs):
    """tta_gather_kernel: every variant plane holds the same base picture in its own orientation -> the average is the picture itself
    (realsr.cpp:691-724), cropped by the scaled prepadding and post-processed (x255 + 0.5, truncate, clamp)"""
    new, _ = libs
    bh, bw, ps = 9, 13, 2 * 4  # padded base tile 9 x 13 (LR), prepadding 2 -> 8 at 4x
    oh, ow = 4 * bh, 4 * bw
    rng = np.random.default_rng(11)
    base = rng.random((oh, ow, 4), dtype=np.float32) * 1.2 - 0.1  # some values outside [0,1]: clamp
    planes = [base, base[::-1], base[:, ::-1], base[::-1, ::-1], base.transpose(1, 0, 2), base[:, ::-1].transpose(1, 0, 2),
              base[::-1].transpose(1, 0, 2), base[::-1, ::-1].transpose(1, 0, 2)]
    slot = oh * ow  # float4 per slot
    buf = np.zeros((8, slot, 4), np.float32)
    for v, p in enumerate(planes):
        buf[v] = np.ascontiguousarray(p).reshape(-1, 4)
    crop_h, crop_w = oh - 2 * ps, ow - 2 * ps
    td = np.array([[0, 0, 5, 3, crop_w, crop_h, v, bh, bw, v] for v in range(8)], dtype=np.int32)
    img_w, img_h = crop_w + 9, crop_h + 6
    out = np.zeros((img_h, img_w, 3), np.uint8)
    new.emu_tta_gather(buf.ctypes.data_as(ctypes.c_void_p), td.ctypes.data_as(ctypes.c_void_p), 1, ctypes.c_longlong(slot), out.ctypes.data_as(ctypes.c_void_p),
                       img_w, img_h, 3, ps, crop_w, crop_h)
    c = base[ps:ps + crop_h, ps:ps + crop_w, :3]
    avg = (c + c + c + c + c + c + c + c) / np.float32(8)  # the kernel's summation order on identical values
    want = np.clip((avg * np.float32(255) + np.float32(0.5)).astype(np.int32), 0, 255).astype(np.uint8)
    assert np.array_equal(out[3:3 + crop_h, 5:5 + crop_w], want)
    out[3:3 + crop_h, 5:5 + crop_w] = 0
    assert not out.any()
