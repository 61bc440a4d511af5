#!/usr/bin/env python3
"""CPU benchmark of the C++ command line's PNG writer on an SR-like 3840x2160 image (a quarter of a BASELINE configs[2] result):
prints out-MP/s for 1 and 4 row-band threads, the file size, and Pillow / zlib level 1 for comparison."""
import io, os, subprocess, sys, time
import numpy as np
from PIL import Image
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(1)
lr = rng.integers(0, 256, (540 + 8, 960 + 8, 3)).astype(np.float32)
c = np.cumsum(np.cumsum(np.pad(lr, ((1, 0), (1, 0), (0, 0))), 0, dtype=np.float64), 1)
k = 7
b = (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)
b = b[:540, :960]
b = (b - b.min()) / (b.max() - b.min()) * 255
img = np.asarray(Image.fromarray(np.clip(np.rint(b), 0, 255).astype(np.uint8)).resize((3840, 2160), Image.BICUBIC))
img.tofile("/tmp/enc.raw")
exe = "/tmp/bench_png"
subprocess.check_call(["g++", "-O2", "-std=c++11", "-I" + os.path.join(ROOT, "realsr-ncnn-vulkan_b200", "host"), "-o", exe, os.path.join(ROOT, "tools", "bench_png.cpp"), "-lz", "-lpthread"])
for th in ("1", "4"):
    print(subprocess.run([exe, th], capture_output=True, text=True).stdout.strip().splitlines()[-1], "| file %.2f MB" % (os.path.getsize("/dev/shm/enc_b.png") / 1e6))
t0 = time.time(); buf = io.BytesIO(); Image.fromarray(img).save(buf, "PNG", compress_level=1); dt = time.time() - t0
print("Pillow (libpng + zlib level 1): %.3f s  %.1f MP/s | file %.2f MB" % (dt, 3840 * 2160 / 1e6 / dt, len(buf.getvalue()) / 1e6))
