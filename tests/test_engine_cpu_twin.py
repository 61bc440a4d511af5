"""Whole images through the REAL engine host code (csrc/engine.cu: planner, workspace, plane rotation, tensor maps, launch arguments,
TTA slots, tile crops, alpha) on the CPU, against the reference goldens.  The CUDA runtime is tests/fakecuda/fake_cuda.cpp; the kernel
launchers are tests/fakecuda/fake_kernels.cpp (the CUDA-core kernels run from their real source under the host emulation) and
fake_conv.cpp (the tensor-core conv replaced by a CPU model of its CONTRACT: fp16 operands through the engine's tensor maps, fp32 accumulation, the
epilogues and rounding points of DESIGN.md section 2).  This validates the engine's data flow and the precision design end to end through
the C ABI without a GPU; the tensor-core kernel itself is validated only on hardware (-m gpu)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from common import diff_stats, golden, manifest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "csrc")
FAKE = os.path.join(ROOT, "tests", "fakecuda")


@pytest.fixture(scope="module")
def twin(tmp_path_factory, models_dir):
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h") or not shutil.which("g++"):
        pytest.skip("CUDA headers or g++ not installed")
    d = str(tmp_path_factory.mktemp("twin"))
    inc = ["-I/usr/local/cuda/include", "-I" + CSRC]
    shim = ["-include", os.path.join(FAKE, "shim.h")]
    jobs = [
        ["g++", "-O2", "-std=c++17", "-w", *inc, *shim, "-c", "-x", "c++", os.path.join(CSRC, "engine.cu"), "-o", d + "/engine.o"],
        ["g++", "-O2", "-std=c++17", "-w", "-c", os.path.join(CSRC, "model.cpp"), "-o", d + "/model.o"],
        ["g++", "-O2", "-std=c++17", "-w", "-DFAKE_CUDA_FUNCTIONAL", *inc, *shim, "-c", os.path.join(FAKE, "fake_cuda.cpp"), "-o", d + "/fake_cuda.o"],
        ["g++", "-O2", "-std=c++17", "-w", "-ffp-contract=off", *inc, "-c", os.path.join(FAKE, "fake_kernels.cpp"), "-o", d + "/fake_kernels.o"],
        ["g++", "-O3", "-march=native", "-std=c++17", "-w", *inc, *shim, "-c", os.path.join(FAKE, "fake_conv.cpp"), "-o", d + "/fake_conv.o"],
        ["g++", "-O2", "-std=c++17", "-w", "-c", os.path.join(FAKE, "engine_host_test.cpp"), "-o", d + "/test.o"],
    ]
    procs = [subprocess.Popen(j) for j in jobs]
    assert all(p.wait() == 0 for p in procs)
    subprocess.run(["g++", "-o", d + "/engine_twin"] + [d + "/" + o for o in ("engine.o", "model.o", "fake_cuda.o", "fake_kernels.o", "fake_conv.o", "test.o")] + ["-lpthread"], check=True)

    cache = {}

    def run(name, env=None):
        key = (name, tuple(sorted((env or {}).items())))
        if key not in cache:
            cache[key] = run_uncached(name, env)
        return cache[key]

    def run_uncached(name, env=None):
        case = [c for c in manifest()["cases"] if c["name"] == name][0]
        img, gold = golden(name)
        h, w, c = img.shape
        img.tofile(d + "/in.raw")
        e = dict(os.environ)
        e.update(env or {})
        r = subprocess.run([d + "/engine_twin", models_dir, "image", case["model"], d + "/in.raw", str(w), str(h), str(c), str(case["tilesize"]), str(case["tta"]),
                            d + "/out.raw"], capture_output=True, text=True, env=e, timeout=900)
        assert r.returncode == 0 and "image ok" in r.stdout, (r.stdout[-300:], r.stderr[-2000:])
        return np.fromfile(d + "/out.raw", np.uint8).reshape(gold.shape), gold

    return run


@pytest.mark.parametrize("name,natural", [("natural_48x40_df2k_t32", True),        # four tiles in four padded shapes: one merged job
                                          ("noise_12x11_df2k_tta_t32", False),      # 8-way TTA: transposed shape groups, fp32 result planes, gather
                                          ("rgba_natural_36x28_jpeg_t200", True)])  # RGBA: the alpha plane through the bicubic kernel
def test_goldens_through_the_engine_on_the_cpu(twin, name, natural):
    out, gold = twin(name)
    s = diff_stats(out, gold)
    if natural:
        assert s["max"] <= 1 and s["frac_diff"] <= 0.03, s  # >= 97 % of bytes exact on natural images
    else:
        assert s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 10000), s


def test_packing_options_and_tile_ranges_do_not_change_the_image(twin):
    """one job per shape group instead of merged launches, one tile per job, and the image cut into tile ranges (rsr_process_tiles, the
    sharded path): the same per-tile arithmetic, so the bytes are identical"""
    ref, _ = twin("natural_48x40_df2k_t32")  # (cached from the golden test)
    env = {"RSR_TWIN_OPTS": "merge_shapes=0,serpentine=0,max_batch=1", "RSR_TWIN_SHARDS": "3"}  # all at once: the host cores are slow at convolving
    out, _ = twin("natural_48x40_df2k_t32", env)
    assert np.array_equal(out, ref)
