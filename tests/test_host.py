"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/realsr_b200.h declares, the
model loader accepts the shipped files / rejects broken ones, weight packing, and loud failure without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from common import ROOT, model_paths, read_convs


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "realsr_b200.h")).read()
    return sorted(set(re.findall(r"RSR_API\s+[\w\s\*]+?\b(rsr_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = rb.load_library()
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "missing export " + s


def check_model(param, binf):
    lib = rb.load_library()
    lib.rsr_check_model.restype = ctypes.c_int
    nr, nc = ctypes.c_int(0), ctypes.c_int(0)
    err = ctypes.create_string_buffer(512)
    r = lib.rsr_check_model(param.encode(), binf.encode(), ctypes.byref(nr), ctypes.byref(nc), err, 512)
    return r, nr.value, nc.value, err.value.decode()


@pytest.mark.parametrize("model", ["models-DF2K", "models-DF2K_JPEG"])
def test_shipped_models_parse(model, models_dir):
    """351 convs = 1 + 23*15 + 5; .bin consumed exactly (33 424 520 B)"""
    p, b = model_paths(model)
    r, nr, nc, err = check_model(p, b)
    assert (r, nr, nc) == (0, 23, 351), err
    assert os.path.getsize(b) == 33424520


def test_loader_rejects_bad_files(models_dir, tmp_path):
    p, b = model_paths("models-DF2K")
    r, _, _, err = check_model(str(tmp_path / "nope.param"), b)
    assert r != 0 and "cannot open" in err
    # truncated weights
    tb = tmp_path / "trunc.bin"
    tb.write_bytes(open(b, "rb").read()[:1000000])
    r, _, _, err = check_model(p, str(tb))
    assert r != 0 and "truncated" in err
    # a graph that is not RRDBNet: drop one Eltwise line
    lines = open(p).read().split("\n")
    idx = [i for i, l in enumerate(lines) if l.startswith("Eltwise")][5]
    bad = tmp_path / "bad.param"
    bad.write_text("\n".join(lines[:idx] + lines[idx + 1:]).replace("\n999 ", "\n998 ", 1))
    r, _, _, err = check_model(str(bad), b)
    assert r != 0 and "RRDBNet" in err
    # wrong magic
    bad.write_text("123\n1 1\n")
    r, _, _, err = check_model(str(bad), b)
    assert r != 0 and "magic" in err


def test_weight_rows_for_cta_pair_kernel(models_dir):
    """row-major [rows][64] weight matrix of the cta_group::2 kernel: per (plane, ky) the three kx tap groups stacked along N"""
    lib = rb.load_library()
    lib.rsr_debug_pack_rows.restype = ctypes.c_longlong
    p, b = model_paths("models-DF2K_JPEG")
    convs = read_convs("models-DF2K_JPEG")
    for idx in (1, 4, 5, 348, 350):
        c = convs[idx]
        grp = 16 if c["cout"] <= 16 else (32 if c["cout"] <= 32 else 64)
        per_ky = 64 if grp == 16 else 3 * grp
        nchunk = (c["cin"] + 63) // 64
        rows = lib.rsr_debug_pack_rows(p.encode(), b.encode(), idx, None, ctypes.c_longlong(0))
        assert rows == nchunk * 3 * per_ky
        buf = np.zeros(rows * 64, dtype=np.uint16)
        assert lib.rsr_debug_pack_rows(p.encode(), b.encode(), idx, buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(buf.size)) == rows
        ref = np.zeros((nchunk, 3, per_ky, 64), dtype=np.float16)
        for ch in range(nchunk):
            k = min(64, c["cin"] - ch * 64)
            for ky in range(3):
                for kx in range(3):
                    ref[ch, ky, kx * grp:kx * grp + c["cout"], :k] = c["w"][:, ch * 64:ch * 64 + k, ky, kx]
        assert np.array_equal(buf.view(np.float16).reshape(ref.shape).view(np.uint16), ref.view(np.uint16))


def test_fails_loudly_without_gpu_or_with_cpu_id(models_dir):
    """no CPU fallback: gpuid -1 (the reference's ncnn CPU mode) and a missing device are errors, not silent paths"""
    p, b = model_paths("models-DF2K")
    e = rb.RealSR(-1)
    assert e.load(p, b) == -4 and "CPU" in e.last_error()
    out = np.zeros((128, 128, 3), dtype=np.uint8)
    assert e.process(np.zeros((32, 32, 3), dtype=np.uint8), out) != 0
    e = rb.RealSR(63)
    assert e.load(p, b) != 0 and "device" in e.last_error()


def test_set_validation():
    e = rb.RealSR(0)
    e.scale = 2
    with pytest.raises(rb.RealSRError):
        e._push_params()
    e.scale, e.tilesize = 4, 16
    with pytest.raises(rb.RealSRError):
        e._push_params()


def test_cpp_wrapper_compiles_and_fails_loudly(tmp_path, models_dir):
    """class RealSR (host/realsr.h, the reference's signatures) over the C ABI: builds with a plain g++ and reports the
    missing GPU instead of falling back to anything"""
    import subprocess
    exe = tmp_path / "cli"
    host = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "host")
    libdir = os.path.join(ROOT, "realsr-ncnn-vulkan_b200")
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++11", "-Wall", "-Werror", "-o", str(exe), os.path.join(host, "realsr_raw_cli.cpp"),
                           "-L" + libdir, "-lrealsr_b200", "-Wl,-rpath," + libdir, "-lpthread"])
    raw = tmp_path / "in.raw"
    np.zeros((8, 8, 3), np.uint8).tofile(raw)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([str(exe), os.path.join(models_dir, "models-DF2K"), str(raw), "8", "8", str(tmp_path / "o.raw")], capture_output=True, text=True)
    if not has_gpu:
        assert r.returncode == 4 and "invalid gpu device" in r.stderr


def test_options_and_tile_api_validate_without_gpu():
    """rsr_set_option rejects unknown keys / out-of-range values; rsr_tile_count follows the reference tile grid;
    rsr_process_tiles on a handle without a model is an error (no GPU needed for any of it)"""
    e = rb.RealSR(0)
    assert e.set_option("no_such_option", 1) != 0 and "unknown option" in e.last_error()
    for key in ("merge_shapes", "serpentine", "max_batch", "time_convs"):
        assert e.set_option(key, 1) == 0, key
    e.tilesize = 400
    assert e.tile_count(1920, 1080) == 15
    e.tilesize = 200
    assert e.tile_count(3840, 2160) == 220 and e.tile_count(1, 1) == 1
    img, out = np.zeros((40, 40, 3), np.uint8), np.zeros((160, 160, 3), np.uint8)
    assert e.process_tiles(img, out, 0, 1) != 0 and "no model" in e.last_error()


def test_build_variant_compiles():
    """-DRSR_EARLY_HANDBACK (csrc/conv3x3.cu, csrc/build.sh RSR_BUILD_VARIANTS): an A/B variant for hardware runs that is never the
    default product -- it must at least keep compiling for sm_100a without spills"""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not installed")
    src = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "csrc", "conv3x3.cu")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DRSR_EARLY_HANDBACK", "-Xptxas", "-v", "-c", src,
                        "-o", os.devnull], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    spills = [int(x) for x in re.findall(r"(\d+) bytes spill stores", r.stderr)]
    assert spills and max(spills) <= 64, r.stderr[-2000:]
