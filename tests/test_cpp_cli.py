"""CPU tests of the C++ host side: the PNG codec (host/png_codec.h) and the reference-shaped C++ command line
(host/realsr_cli.cpp: flag surface and messages of /root/reference/src/main.cpp:422-879, file planning :606-659, the
load -> proc -> save thread pipeline :179-416).  The CLI is linked against tests/stub/stub_realsr.c, a nearest-x4
stand-in for the C ABI, so the whole pipeline runs without a GPU; the real library is exercised by the gpu tests."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "host")


@pytest.fixture(scope="module")
def bins(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("cppcli"))
    subprocess.run(["gcc", "-shared", "-fPIC", "-O2", "-o", os.path.join(d, "librealsr_b200.so"), os.path.join(ROOT, "tests", "stub", "stub_realsr.c")], check=True)
    subprocess.run(["g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", os.path.join(d, "realsr-b200"), os.path.join(HOST, "realsr_cli.cpp"),
                    "-L" + d, "-lrealsr_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"], check=True)
    subprocess.run(["g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", os.path.join(d, "png_tool"), os.path.join(HOST, "png_tool.cpp"), "-lz", "-lpthread"], check=True)
    return d


def decode(bins, path, tmp):
    raw = os.path.join(tmp, "o.raw")
    r = subprocess.run([os.path.join(bins, "png_tool"), "decode", path, raw], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w, h, c = map(int, r.stdout.split())
    return np.fromfile(raw, dtype=np.uint8).reshape(h, w, c)


def expected(im):
    """what the reference's load stage hands to the engine (main.cpp:243-263): RGB, or RGBA when the file carries alpha"""
    alpha = im.mode in ("RGBA", "LA", "PA") or "transparency" in im.info
    return np.asarray(im.convert("RGBA" if alpha else "RGB"))


def test_png_decode_matches_pillow(bins, tmp_path):
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    cases = {
        "rgb": Image.fromarray(rgb),
        "rgba": Image.fromarray(np.dstack([rgb, rng.integers(0, 256, (37, 53), dtype=np.uint8)])),
        "gray": Image.fromarray(rgb[..., 0]),
        "la": Image.fromarray(np.dstack([rgb[..., 0], rgb[..., 1]]), "LA"),
        "pal": Image.fromarray(rgb).quantize(200),
        "pal16": Image.fromarray(rgb).quantize(16),
        "bilevel": Image.fromarray(rgb[..., 0] > 127),
    }
    for name, im in cases.items():
        p = str(tmp_path / (name + ".png"))
        im.save(p, optimize=(name == "pal16"))  # optimize packs small palettes to 4 bits per pixel
        got = decode(bins, p, str(tmp_path))
        assert np.array_equal(got, expected(Image.open(p))), name
    # palette with transparency, 16-bit gray, 16-bit RGB
    pal = Image.fromarray(rgb).quantize(64)
    p = str(tmp_path / "paltr.png")
    pal.save(p, transparency=3)
    got = decode(bins, p, str(tmp_path))
    assert got.shape[2] == 4 and np.array_equal(got, expected(Image.open(p)))
    g16 = rng.integers(0, 65536, (20, 31), dtype=np.uint16)
    p = str(tmp_path / "g16.png")
    Image.fromarray(g16).save(p)
    got = decode(bins, p, str(tmp_path))
    assert np.array_equal(got, np.repeat((g16 >> 8).astype(np.uint8)[..., None], 3, 2))


def _write_png(path, w, h, depth, ctype, raw, interlace=0):
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace)) +
                chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def test_png_decode_adam7_and_all_filters(bins, tmp_path):
    """hand-built files: Adam7 interlace (Pillow cannot write it) and every scanline filter type"""
    rng = np.random.default_rng(6)
    img = rng.integers(0, 256, (13, 19, 3), dtype=np.uint8)
    h, w, _ = img.shape
    raw = b""
    for (x0, y0, dx, dy) in [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]:
        sub = img[y0::dy, x0::dx]
        if sub.size:
            for row in sub:
                raw += b"\x00" + row.tobytes()
    p = str(tmp_path / "adam7.png")
    _write_png(p, w, h, 8, 2, raw, interlace=1)
    assert np.array_equal(decode(bins, p, str(tmp_path)), img)
    assert np.array_equal(np.asarray(Image.open(p)), img)  # and Pillow agrees the file is well formed
    # filters 0..4, one per row, computed here from the specification
    def paeth(a, b, c):
        pp = a + b - c
        pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
        return a if pa <= pb and pa <= pc else (b if pb <= pc else c)
    raw = b""
    prev = np.zeros(w * 3, dtype=np.int32)
    for y in range(h):
        cur = img[y].reshape(-1).astype(np.int32)
        ft = y % 5
        out = bytearray([ft])
        for i in range(w * 3):
            a = cur[i - 3] if i >= 3 else 0
            b = prev[i]
            c = prev[i - 3] if i >= 3 else 0
            pred = [0, a, b, (a + b) // 2, paeth(a, b, c)][ft]
            out.append((cur[i] - pred) & 255)
        raw += bytes(out)
        prev = cur
    p = str(tmp_path / "filters.png")
    _write_png(p, w, h, 8, 2, raw)
    assert np.array_equal(decode(bins, p, str(tmp_path)), img)
    # corrupt CRC is refused
    data = bytearray(open(p, "rb").read())
    data[-20] ^= 0x55
    bad = str(tmp_path / "bad.png")
    open(bad, "wb").write(bytes(data))
    r = subprocess.run([os.path.join(bins, "png_tool"), "decode", bad, str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert r.returncode != 0


@pytest.mark.parametrize("c", [3, 4])
def test_png_encode_roundtrip(bins, tmp_path, c):
    rng = np.random.default_rng(7 + c)
    yy, xx = np.mgrid[0:61, 0:47]
    smooth = np.stack([(xx * 5 + yy * 2) % 256, (xx + yy * 3) % 256, (xx * yy) % 256, (xx * 7) % 256][:c], -1).astype(np.uint8)
    for name, img in (("noise", rng.integers(0, 256, (61, 47, c), dtype=np.uint8)), ("smooth", smooth)):
        raw, png = str(tmp_path / (name + ".raw")), str(tmp_path / (name + ".png"))
        img.tofile(raw)
        subprocess.run([os.path.join(bins, "png_tool"), "encode", raw, "47", "61", str(c), png], check=True)
        back = np.asarray(Image.open(png))
        assert back.shape == img.shape and np.array_equal(back, img), name
        assert np.array_equal(decode(bins, png, str(tmp_path)), img)
        # parallel row bands (independent raw deflate streams spliced into one zlib stream) and the adaptive-filter mode
        for lvl, th in (("1", "3"), ("6", "1"), ("6", "4")):
            subprocess.run([os.path.join(bins, "png_tool"), "encode", raw, "47", "61", str(c), png, lvl, th], check=True)
            assert np.array_equal(np.asarray(Image.open(png)), img), (name, lvl, th)
    assert os.path.getsize(str(tmp_path / "smooth.png")) < smooth.size // 3  # filters + deflate do compress gradients


def test_png_fast_writer_blocks_runs_and_stored_fallback(bins, tmp_path):
    """the level-1 writer's own deflate (png_codec.h fastz: run-length matches + one dynamic Huffman code per 128-KB block, stored blocks
    for incompressible data, bands spliced into one zlib stream): inflates bit-exactly with zlib (Pillow) and with the own reader over
    degenerate sizes, long runs across block and band borders, noise, sparse and mixed content; and it is not larger than zlib level 1"""
    rng = np.random.default_rng(77)

    def rt(arr, name, threads):
        h, w, c = arr.shape
        raw, png = str(tmp_path / "rt.raw"), str(tmp_path / "rt.png")
        np.ascontiguousarray(arr).tofile(raw)
        subprocess.run([os.path.join(bins, "png_tool"), "encode", raw, str(w), str(h), str(c), png, "1", str(threads)], check=True)
        got = np.asarray(Image.open(png))
        assert got.shape == arr.shape and np.array_equal(got, arr), name
        assert np.array_equal(decode(bins, png, str(tmp_path)), arr), name
        return os.path.getsize(png)

    for c in (3, 4):
        for (h, w) in ((1, 1), (2, 1), (3, 5), (64, 300), (33, 5000)):
            yy, xx = np.mgrid[0:h, 0:w]
            sparse = np.zeros((h, w, c), np.uint8)
            m = rng.random((h, w)) < 0.01
            sparse[m] = rng.integers(0, 256, (int(m.sum()), c), dtype=np.uint8)
            cases = {"noise": rng.integers(0, 256, (h, w, c), dtype=np.uint8), "flat": np.full((h, w, c), 200, np.uint8), "sparse": sparse,
                     "runs": np.repeat(rng.integers(0, 256, (h, (w + 40) // 41, c), dtype=np.uint8), 41, axis=1)[:, :w],
                     "smooth": np.stack([(xx * 3 + yy) % 256, (xx + yy * 2) // 3 % 256, (xx * yy) % 251, (xx + yy) % 256][:c], -1).astype(np.uint8)}
            for name, a in cases.items():
                for th in (1, 3):
                    size = rt(a, "%s %dx%dx%d t%d" % (name, w, h, c, th), th)
                if name == "flat":
                    assert size < 200 + a.size // 200
                if name == "noise" and a.size > 10000:
                    assert size < a.size + a.size // 100 + 200  # stored blocks: no expansion beyond the headers
    mixed = np.zeros((600, 1000, 3), np.uint8)  # several 128-KB blocks per band, runs that cross block and band borders
    mixed[100:500, 200:800] = rng.integers(0, 256, (400, 600, 3), dtype=np.uint8)
    s1, s4 = rt(mixed, "mixed", 1), rt(mixed, "mixed", 4)
    assert abs(s1 - s4) < 2000
    photo = np.asarray(Image.fromarray(rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)).resize((640, 480), Image.BICUBIC))  # smooth, like an SR result
    import io
    buf = io.BytesIO()
    Image.fromarray(photo).save(buf, "PNG", compress_level=1)
    assert rt(photo, "photo", 2) < len(buf.getvalue())


def test_jpeg_decode_matches_libjpeg(bins, tmp_path):
    """JPEG reader (sequential and progressive) vs Pillow (libjpeg-turbo): IDCT / upsampling / colour rounding differ by a
    few LSB at most"""
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:75, 0:101]
    img = np.stack([np.sin(xx / 9.) * 80 + 128 + yy / 2, np.cos(yy / 7.) * 70 + 120 + xx / 3, ((xx * yy) % 200) + 20.], -1)
    img = np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
    cases = [("q95_444", img, dict(quality=95, subsampling=0)), ("q90_420", img, dict(quality=90, subsampling=2)),
             ("q85_422", img, dict(quality=85, subsampling=1)), ("q75", img, dict(quality=75)), ("gray", img[..., 0], dict(quality=90)),
             ("rst", img, dict(quality=90, subsampling=2, restart_marker_blocks=3)), ("odd", img[:33, :17], dict(quality=92, subsampling=2))]
    for name, arr, kw in cases:
        p = str(tmp_path / (name + ".jpg"))
        try:
            Image.fromarray(arr).save(p, **kw)
        except TypeError:  # older Pillow without restart_marker_blocks
            continue
        got = decode(bins, p, str(tmp_path)).astype(int)
        ref = np.asarray(Image.open(p).convert("RGB")).astype(int)
        d = np.abs(got - ref)
        assert got.shape == ref.shape and d.max() <= 4 and d.mean() < 0.35 and (d > 2).mean() < 0.002, (name, d.max(), d.mean())
    # progressive (spectral selection + successive approximation), with and without restart intervals / subsampling
    big = np.asarray(Image.fromarray(img).resize((517, 389), Image.BICUBIC))
    for name, arr, kw in [("p420", img, dict(quality=90, subsampling=2, progressive=True)), ("p444", img, dict(quality=97, subsampling=0, progressive=True)),
                          ("pgray", img[..., 0], dict(quality=80, progressive=True)), ("pbig", big, dict(quality=85, progressive=True)),
                          ("prst", big, dict(quality=85, progressive=True, restart_marker_blocks=5)),
                          ("p422", big, dict(quality=70, subsampling=1, progressive=True)), ("opt", big, dict(quality=88, optimize=True))]:
        p = str(tmp_path / (name + ".jpg"))
        Image.fromarray(arr).save(p, **kw)
        got = decode(bins, p, str(tmp_path)).astype(int)
        ref = np.asarray(Image.open(p).convert("RGB")).astype(int)
        d = np.abs(got - ref)
        assert got.shape == ref.shape and d.max() <= 4 and d.mean() < 0.35 and (d > 2).mean() < 0.002, (name, d.max(), d.mean())
    # a truncated file is an error, not a crash
    data = open(p, "rb").read()
    bad = str(tmp_path / "trunc.jpg")
    open(bad, "wb").write(data[:200])
    r = subprocess.run([os.path.join(bins, "png_tool"), "decode", bad, str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "decode failed" in r.stderr


def test_jpeg_encode_quality_100(bins, tmp_path):
    """baseline 4:4:4 writer at the reference's quality (stbi_write_jpg(..., 100), main.cpp:397): libjpeg reads it back
    within rounding"""
    rng = np.random.default_rng(22)
    for name, img in (("noise", rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)),
                      ("flat", np.full((9, 9, 3), 200, np.uint8)), ("ramp", np.stack([np.tile(np.arange(64, dtype=np.uint8) * 4, (40, 1))] * 3, -1))):
        raw, jpg = str(tmp_path / (name + ".raw")), str(tmp_path / (name + ".jpg"))
        img.tofile(raw)
        h, w, _ = img.shape
        subprocess.run([os.path.join(bins, "png_tool"), "encode", raw, str(w), str(h), "3", jpg], check=True)
        back = np.asarray(Image.open(jpg).convert("RGB")).astype(int)
        d = np.abs(back - img.astype(int))
        assert back.shape == img.shape and d.max() <= 4 and d.mean() < 0.8, (name, d.max(), d.mean())


def test_webp_lossless_reader_and_writer(bins, tmp_path):
    """VP8L: files written by libwebp (Pillow) at several effort levels -- palette / pixel packing, colour cache, LZ77,
    all transforms -- decode bit-exactly; files written here are decoded bit-exactly by libwebp"""
    rng = np.random.default_rng(31)
    yy, xx = np.mgrid[0:75, 0:101]
    img = np.stack([np.sin(xx / 9.) * 80 + 128 + yy / 2, np.cos(yy / 7.) * 70 + 120 + xx / 3, ((xx * yy) % 200) + 20.], -1)
    img = np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
    big = np.asarray(Image.fromarray(img).resize((517, 389), Image.BICUBIC))
    flat = np.zeros((60, 80, 3), np.uint8)
    flat[10:40, 20:60] = (200, 30, 40)
    cases = {"smooth": img, "big": big, "flat": flat, "pal12": np.asarray(Image.fromarray(big).quantize(12).convert("RGB")),
             "bilevel": (rng.integers(0, 2, (40, 50, 1)) * 255).astype(np.uint8).repeat(3, 2),
             "rgba": np.dstack([img, (xx * 2 % 256).astype(np.uint8)]), "noise": rng.integers(0, 256, (64, 64, 3), dtype=np.uint8),
             "one": img[:1, :1], "row": img[:1], "col": img[:, :1], "const": np.full((33, 17, 3), 77, np.uint8)}
    for name, arr in cases.items():
        arr = np.ascontiguousarray(arr)
        mode = "RGBA" if arr.shape[2] == 4 else "RGB"
        for method in (0, 3, 6):
            p = str(tmp_path / ("%s_%d.webp" % (name, method)))
            Image.fromarray(arr).save(p, lossless=True, method=method)
            assert np.array_equal(decode(bins, p, str(tmp_path)), np.asarray(Image.open(p).convert(mode))), (name, method)
        raw, out = str(tmp_path / "e.raw"), str(tmp_path / (name + "_own.webp"))
        arr.tofile(raw)
        subprocess.run([os.path.join(bins, "png_tool"), "encode", raw, str(arr.shape[1]), str(arr.shape[0]), str(arr.shape[2]), out], check=True)
        assert np.array_equal(np.asarray(Image.open(out).convert(mode)), arr), name
        assert np.array_equal(decode(bins, out, str(tmp_path)), arr), name


def libwebp_encoder():
    """libwebp's advanced encoder API (WebPConfig / WebPPicture / WebPEncode) of the libwebp that ships with Pillow, through ctypes: Pillow's
    own save() exposes only quality / method, the bitstream features a decoder has to get right need more knobs"""
    import ctypes
    import glob
    import PIL
    libs = os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs")
    so = glob.glob(os.path.join(libs, "libwebp-*.so*"))
    if not so:
        return None
    for dep in glob.glob(os.path.join(libs, "libsharpyuv-*.so*")):
        ctypes.CDLL(dep, mode=ctypes.RTLD_GLOBAL)
    L = ctypes.CDLL(so[0])
    if (L.WebPGetEncoderVersion() >> 16) != 1:
        return None  # the struct layouts below are those of libwebp 1.x
    fields = ["lossless", "quality", "method", "image_hint", "target_size", "target_PSNR", "segments", "sns_strength", "filter_strength",
              "filter_sharpness", "filter_type", "autofilter", "alpha_compression", "alpha_filtering", "alpha_quality", "pass_", "show_compressed",
              "preprocessing", "partitions", "partition_limit", "emulate_jpeg_size", "thread_level", "low_memory", "near_lossless", "exact",
              "use_delta_palette", "use_sharp_yuv", "qmin", "qmax"]

    class Config(ctypes.Structure):
        _fields_ = [(n, ctypes.c_float if n in ("quality", "target_PSNR") else ctypes.c_int) for n in fields]

    class MemWriter(ctypes.Structure):
        _fields_ = [("mem", ctypes.c_void_p), ("size", ctypes.c_size_t), ("max_size", ctypes.c_size_t), ("pad", ctypes.c_uint32 * 1)]

    L.WebPConfigInitInternal.argtypes = [ctypes.POINTER(Config), ctypes.c_int, ctypes.c_float, ctypes.c_int]
    abi = 0x020f  # WEBP_ENCODER_ABI_VERSION: only the major byte is compared

    def encode(arr, quality=75, **kw):
        cfg = Config()
        assert L.WebPConfigInitInternal(ctypes.byref(cfg), 0, ctypes.c_float(quality), abi)
        for k, v in kw.items():
            setattr(cfg, k, v)
        assert L.WebPValidateConfig(ctypes.byref(cfg)), kw
        pic = (ctypes.c_uint8 * 512)()  # WebPPicture: width / height are ints 2 and 3, writer and custom_ptr sit at bytes 96 and 104
        assert L.WebPPictureInitInternal(pic, abi)
        h, w, c = arr.shape
        ints = ctypes.cast(pic, ctypes.POINTER(ctypes.c_int))
        ints[2], ints[3] = w, h
        arr = np.ascontiguousarray(arr)
        assert (L.WebPPictureImportRGBA if c == 4 else L.WebPPictureImportRGB)(pic, arr.ctypes.data_as(ctypes.c_void_p), w * c)
        mw = MemWriter()
        L.WebPMemoryWriterInit(ctypes.byref(mw))
        ctypes.cast(ctypes.addressof(pic) + 96, ctypes.POINTER(ctypes.c_void_p))[0] = ctypes.cast(L.WebPMemoryWrite, ctypes.c_void_p).value
        ctypes.cast(ctypes.addressof(pic) + 104, ctypes.POINTER(ctypes.c_void_p))[0] = ctypes.addressof(mw)
        assert L.WebPEncode(ctypes.byref(cfg), pic), "WebPEncode failed"
        data = ctypes.string_at(mw.mem, mw.size)
        L.WebPPictureFree(pic)
        return data

    return encode


def test_webp_lossy_reader_matches_libwebp(bins, tmp_path):
    """VP8 key frames (host/vp8_codec.h) + ALPH planes: bit-exact RGB(A) against libwebp's WebPDecodeRGB(A) (through Pillow) -- intra
    prediction incl. the frame-edge rules, token contexts, segments, 1..8 token partitions, simple and normal loop filter at several
    strengths / sharpness, the 9-3-3-1 chroma upsampler and the fixed-point YUV->RGB; alpha raw / VP8L-coded with every prediction filter"""
    rng = np.random.default_rng(41)
    yy, xx = np.mgrid[0:90, 0:121]
    smooth = np.stack([np.sin(xx / 9.) * 80 + 128 + yy / 2, np.cos(yy / 7.) * 70 + 120 + xx / 3, ((xx * yy) % 200) + 20.], -1)
    smooth = np.clip(smooth + rng.normal(0, 4, smooth.shape), 0, 255).astype(np.uint8)
    noise = rng.integers(0, 256, (70, 83, 3), dtype=np.uint8)
    n = [0]

    def check(data_or_arr, name, **save_kw):
        p = str(tmp_path / (name + ".webp"))
        if isinstance(data_or_arr, bytes):
            open(p, "wb").write(data_or_arr)
        else:
            Image.fromarray(data_or_arr).save(p, "WEBP", **save_kw)
        im = Image.open(p)
        want = np.asarray(im.convert("RGBA" if im.mode == "RGBA" else "RGB"))
        got = decode(bins, p, str(tmp_path))
        assert got.shape == want.shape and np.array_equal(got, want), name
        n[0] += 1

    for src, sname in ((smooth, "smooth"), (noise, "noise"), (smooth[:5, :3], "tiny"), (smooth[:16, :16], "onemb"), (smooth[:17, :33], "ragged")):
        for q in (0, 30, 75, 100):
            for m in (0, 4, 6):
                check(src, "%s_q%d_m%d" % (sname, q, m), quality=q, method=m)
    rgba = np.dstack([smooth, (xx * 2 % 256).astype(np.uint8)])
    rgba[:40, :50, 3] = 255
    for aq in (100, 50):
        check(rgba, "rgba_aq%d" % aq, quality=60, alpha_quality=aq)
    enc = libwebp_encoder()
    if enc is not None:
        for src, sname in ((smooth, "smooth"), (noise, "noise")):
            for ft in (0, 1):
                for sharp, fs in ((0, 0), (0, 60), (3, 20), (7, 100)):
                    check(enc(src, quality=40, filter_type=ft, filter_sharpness=sharp, filter_strength=fs, method=4), "%s_ft%d_sh%d_fs%d" % (sname, ft, sharp, fs))
            for parts in (1, 2, 3):
                check(enc(src, quality=60, partitions=parts, method=3), "%s_parts%d" % (sname, parts))
            for seg, sns in ((1, 0), (2, 50), (4, 100)):
                check(enc(src, quality=30, segments=seg, sns_strength=sns, method=5), "%s_seg%d" % (sname, seg))
        for ac in (0, 1):
            for af in (0, 1, 2):
                check(enc(rgba, quality=50, alpha_compression=ac, alpha_filtering=af, alpha_quality=100 - 40 * af, method=4), "alpha_c%d_f%d" % (ac, af))
    assert n[0] >= 62
    # corrupt input is an error, not a crash
    bad = bytearray(open(str(tmp_path / "smooth_q75_m4.webp"), "rb").read())
    bad[23] ^= 0xff  # the VP8 start code
    badp = str(tmp_path / "bad.webp")
    open(badp, "wb").write(bytes(bad))
    r = subprocess.run([os.path.join(bins, "png_tool"), "decode", badp, str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "VP8" in r.stderr


def cli(bins, *args):
    return subprocess.run([os.path.join(bins, "realsr-b200")] + list(args), capture_output=True, text=True)


def test_cli_argument_validation(bins, tmp_path):
    """same checks and messages as the reference (main.cpp:510-604, :661-733)"""
    img = str(tmp_path / "a.png")
    Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(img)
    out = str(tmp_path / "o.png")
    assert cli(bins).returncode != 0 and "Usage:" in cli(bins).stderr
    assert "invalid scale argument" in cli(bins, "-i", img, "-o", out, "-s", "2").stderr
    assert "invalid tilesize argument" in cli(bins, "-i", img, "-o", out, "-t", "16").stderr
    assert "invalid tilesize argument" in cli(bins, "-i", img, "-o", out, "-t", "64,64", "-g", "0").stderr
    assert "invalid thread count argument" in cli(bins, "-i", img, "-o", out, "-j", "0:2:2").stderr
    assert "invalid jobs_proc thread count argument" in cli(bins, "-i", img, "-o", out, "-j", "1:2,2,2:2", "-g", "0,1").stderr
    assert "invalid argument" in cli(bins, "-i", img, "-o", out, "-j", "4").stderr
    assert "invalid outputpath extension type" in cli(bins, "-i", img, "-o", str(tmp_path / "o.bmp")).stderr
    assert "invalid format argument" in cli(bins, "-i", str(tmp_path), "-o", str(tmp_path), "-f", "tiff").stderr
    assert "must be either file or directory at the same time" in cli(bins, "-i", img, "-o", str(tmp_path)).stderr
    assert "unknown model dir type" in cli(bins, "-i", img, "-o", out, "-m", "models-other").stderr
    assert "no CPU mode" in cli(bins, "-i", img, "-o", out, "-g", "-1", "-m", "models-DF2K").stderr
    r = cli(bins, "-i", img, "-o", out, "-m", "missing/models-DF2K")
    assert r.returncode != 0 and "load failed" in r.stderr


def test_cli_pipeline_directory_mode(bins, tmp_path):
    """dir -> dir through load/proc/save threads on two (stub) GPUs: sorted names, the collision rename, -v lines,
    undecodable files reported and skipped, RGBA kept, every output = the engine's result for its input"""
    din, dout = tmp_path / "in", tmp_path / "out"
    din.mkdir(); dout.mkdir()
    rng = np.random.default_rng(8)
    imgs = {}
    for i in range(11):
        c = 4 if i % 4 == 3 else 3
        im = rng.integers(0, 256, (9 + i, 14 - i % 3, c), dtype=np.uint8)
        imgs["img%02d.png" % i] = im
        Image.fromarray(im).save(str(din / ("img%02d.png" % i)))
    (din / "img03.txt").write_text("not an image")  # same stem as img03.png -> collision rule, and undecodable
    r = cli(bins, "-i", str(din), "-o", str(dout), "-m", "models-DF2K", "-g", "0,1", "-j", "2:3:2", "-t", "64,64", "-v")
    assert r.returncode == 0, r.stderr
    assert "both img03.txt and img03.png output img03.png ! img03.txt will output img03.txt.png" in r.stderr
    assert "decode image %s failed" % str(din / "img03.txt") in r.stderr
    assert r.stderr.count(" done\n") == 11
    assert sorted(os.listdir(str(dout))) == sorted(imgs)
    for name, im in imgs.items():
        got = np.asarray(Image.open(str(dout / name)))
        assert np.array_equal(got, np.repeat(np.repeat(im, 4, 0), 4, 1)), name
    # jpg in, jpg out (-f jpg), and the reference's rule for alpha inputs: <name>.jpg.png (main.cpp:279-288)
    dj, djo = tmp_path / "jin", tmp_path / "jout"
    dj.mkdir(); djo.mkdir()
    smooth = np.stack([np.tile(np.arange(40, dtype=np.uint8) * 6, (30, 1))] * 3, -1)
    Image.fromarray(smooth).save(str(dj / "a.jpg"), quality=95, subsampling=0)
    Image.fromarray(imgs["img03.png"]).save(str(dj / "b.png"))  # RGBA
    r = cli(bins, "-i", str(dj), "-o", str(djo), "-m", "models-DF2K", "-f", "jpg")
    assert r.returncode == 0, r.stderr
    assert "has alpha channel" in r.stderr and sorted(os.listdir(str(djo))) == ["a.jpg", "b.jpg.png"]
    up = np.repeat(np.repeat(np.asarray(Image.open(str(dj / "a.jpg")).convert("RGB")), 4, 0), 4, 1).astype(int)
    assert np.abs(np.asarray(Image.open(str(djo / "a.jpg")).convert("RGB")).astype(int) - up).max() <= 6
    assert np.array_equal(np.asarray(Image.open(str(djo / "b.jpg.png"))), np.repeat(np.repeat(imgs["img03.png"], 4, 0), 4, 1))
    # webp in (lossless), webp out
    wj = str(tmp_path / "w.webp")
    Image.fromarray(imgs["img05.png"]).save(wj, lossless=True)
    r = cli(bins, "-i", wj, "-o", str(tmp_path / "w4.webp"), "-m", "models-DF2K")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "w4.webp")).convert("RGB")), np.repeat(np.repeat(imgs["img05.png"], 4, 0), 4, 1))
    # lossy webp in (VP8 key frame, host/vp8_codec.h), png out: the pixels the engine sees are libwebp's
    wl = str(tmp_path / "lossy.webp")
    Image.fromarray(smooth).save(wl, quality=70)
    r = cli(bins, "-i", wl, "-o", str(tmp_path / "lossy4.png"), "-m", "models-DF2K")
    assert r.returncode == 0, r.stderr
    assert np.array_equal(np.asarray(Image.open(str(tmp_path / "lossy4.png"))), np.repeat(np.repeat(np.asarray(Image.open(wl).convert("RGB")), 4, 0), 4, 1))
    # file -> file
    one = str(tmp_path / "one.png")
    r = cli(bins, "-i", str(din / "img05.png"), "-o", one, "-m", "x/models-DF2K_JPEG")
    assert r.returncode == 0 and np.array_equal(np.asarray(Image.open(one)), np.repeat(np.repeat(imgs["img05.png"], 4, 0), 4, 1))
