"""The engine's HOST code under AddressSanitizer / UndefinedBehaviorSanitizer / ThreadSanitizer, without a GPU: csrc/engine.cu and
csrc/model.cpp are compiled as plain C++ and linked against tests/fakecuda/fake_cuda.cpp, a CPU stand-in for the CUDA runtime whose
allocations are tracked and whose kernel launchers and tensor-map encoder CHECK every range the real kernels would touch (and the TMA
encoding rules) instead of computing anything.
  sweep   : one thread; image sizes from 1x1 to 640x360, tile sizes 32..400, RGB / RGBA, TTA, merge_shapes / max_batch / serpentine,
            tile ranges (rsr_process_tiles), argument errors, model reload and rsr_destroy returning every allocation.
  threads : six proc threads on ONE handle calling rsr_set + rsr_process like RealSR::process() const does (reference main.cpp:811-828),
            failing calls, rsr_last_error and the statistics getters concurrently.
What this cannot see: anything the kernels themselves do (compute-sanitizer on hardware would; it took the GPU box down, profiles/README.md)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "csrc")
FAKE = os.path.join(ROOT, "tests", "fakecuda")


def build(tmp, name, flags):
    exe = os.path.join(tmp, name)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-w", *flags, "-I/usr/local/cuda/include", "-I" + CSRC, "-include", os.path.join(FAKE, "shim.h"), "-o", exe,
           "-x", "c++", os.path.join(CSRC, "engine.cu"), "-x", "c++", os.path.join(CSRC, "model.cpp"), os.path.join(FAKE, "fake_cuda.cpp"),
           os.path.join(FAKE, "engine_host_test.cpp"), "-lpthread"]
    subprocess.run(cmd, check=True)
    return exe


@pytest.fixture(scope="module")
def have_headers():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h") or not shutil.which("g++"):
        pytest.skip("CUDA headers or g++ not installed")


def test_engine_host_code_asan_ubsan_sweep(have_headers, tmp_path, models_dir):
    exe = build(str(tmp_path), "engine_host_asan", ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"])
    r = subprocess.run([exe, models_dir, "sweep"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sweep ok" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
    launches, maps = (int(x) for x in __import__("re").findall(r"(\d+) launches, (\d+) tensor maps", r.stdout)[0])
    assert launches > 50000 and maps > 20000


def test_engine_host_code_tsan_shared_handle(have_headers, tmp_path, models_dir):
    exe = build(str(tmp_path), "engine_host_tsan", ["-fsanitize=thread"])
    r = subprocess.run([exe, models_dir, "threads"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "threads ok" in r.stdout and "ThreadSanitizer" not in r.stderr, (r.stdout[-500:], r.stderr[-4000:])
