#!/usr/bin/env python3
"""Differential fuzz on the CPU: random small images (both dimensions >= 11: below that the reference's reflect border is undefined), tile
sizes, RGB / RGBA, TTA, both models through the REAL engine host code over the fake CUDA runtime (tests/fakecuda: emulated CUDA-core
kernels + contract model of the conv kernel) against the oracle port.   tools/twin_fuzz.py [seed] [cases]"""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import Oracle, diff_stats
CSRC, FAKE, B = "realsr-ncnn-vulkan_b200/csrc", "tests/fakecuda", "/tmp/twin"
os.makedirs(B, exist_ok=True)
inc = ["-I/usr/local/cuda/include", "-I" + CSRC]
shim = ["-include", FAKE + "/shim.h"]
for cmd in (["g++", "-O2", "-std=c++17", "-w", *inc, *shim, "-c", "-x", "c++", CSRC + "/engine.cu", "-o", B + "/engine.o"],
            ["g++", "-O2", "-std=c++17", "-w", "-c", CSRC + "/model.cpp", "-o", B + "/model.o"],
            ["g++", "-O2", "-std=c++17", "-w", "-DFAKE_CUDA_FUNCTIONAL", *inc, *shim, "-c", FAKE + "/fake_cuda.cpp", "-o", B + "/fake_cuda.o"],
            ["g++", "-O2", "-std=c++17", "-w", "-ffp-contract=off", *inc, "-c", FAKE + "/fake_kernels.cpp", "-o", B + "/fake_kernels.o"],
            ["g++", "-O3", "-march=native", "-std=c++17", "-w", *inc, *shim, "-c", FAKE + "/fake_conv.cpp", "-o", B + "/fake_conv.o"],
            ["g++", "-O2", "-std=c++17", "-w", "-c", FAKE + "/engine_host_test.cpp", "-o", B + "/test.o"],
            ["g++", "-o", B + "/engine_twin"] + [B + "/" + o for o in ("engine.o", "model.o", "fake_cuda.o", "fake_kernels.o", "fake_conv.o", "test.o")] + ["-lpthread"]):
    subprocess.check_call(cmd)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
orc = {m: Oracle(m) for m in ("models-DF2K", "models-DF2K_JPEG")}
bad = 0
for it in range(N):
    w, h = int(rng.integers(11, 90)), int(rng.integers(11, 70))
    c = int(rng.choice([3, 4])); tta = int(rng.random() < 0.3); tile = int(rng.choice([32, 33, 40, 64, 200]))
    if tta and w * h > 900: w, h = max(11, w // 3), max(11, h // 3)
    model = str(rng.choice(list(orc)))
    kind = rng.choice(["noise", "smooth"])
    if kind == "noise": img = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([(xx * 3 + yy * 2) % 256, (xx + yy * 5) % 256, (xx * yy + 7) % 256, (xx * 9 + yy) % 256][:c], -1).astype(np.uint8)
    img = np.ascontiguousarray(img); img.tofile("/tmp/twin/in.raw")
    t0 = time.time()
    r = subprocess.run(["/tmp/twin/engine_twin", "models", "image", model, "/tmp/twin/in.raw", str(w), str(h), str(c), str(tile), str(tta), "/tmp/twin/out.raw"], capture_output=True, text=True)
    if r.returncode:
        print("FAILED", (w, h, c, tile, tta, model), r.stderr[-500:]); bad += 1; continue
    out = np.fromfile("/tmp/twin/out.raw", np.uint8).reshape(4 * h, 4 * w, c)
    ref = orc[model].process(img, tile, tta=bool(tta))
    s = diff_stats(out, ref)
    ok = s["max"] <= 2 and s["n_gt1"] <= max(2, s["n"] // 5000)
    print("ok " if ok else "BAD", (w, h, c, tile, tta, model, kind), "%.1fs" % (time.time() - t0), s)
    bad += not ok
print("bad", bad)
