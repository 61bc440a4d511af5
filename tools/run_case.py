#!/usr/bin/env python3
"""Runs the engine on one synthetic image (for profiling under ncu): run_case.py W H TILE [iters] [tta]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb

w, h, tile = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 1
tta = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
e = rb.RealSR(0, tta_mode=tta)
assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0, e.last_error()
e.tilesize = tile
if os.environ.get("RSR_MAX_BATCH"):
    e.set_option("max_batch", int(os.environ["RSR_MAX_BATCH"]))
if os.environ.get("RSR_DEBUG"):
    e.set_option("debug_mask", int(os.environ["RSR_DEBUG"]))
if os.environ.get("RSR_TIME_CONVS"):
    e.set_option("time_convs", 1)
img = np.random.default_rng(1000).integers(0, 256, (h, w, 3), dtype=np.uint8)
out = np.zeros((h * 4, w * 4, 3), dtype=np.uint8)
for i in range(iters):
    t0 = time.time()
    r = e.process(img, out)
    if r:
        print("process failed", r, e.last_error(), flush=True)
        break
    print("rc", r, "wall %.2f ms device %.2f ms" % ((time.time() - t0) * 1e3, e.last_ms()), "launches", e.kernel_launches(),
          "conv ms %.2f" % e.last_ms(1), "padded px", e.last_padded_pixels(), "-> %.1f out-MP/s, conv %.0f TFLOP/s" %
          (w * h * 16 / 1e6 / (e.last_ms() * 1e-3), e.last_padded_pixels() * 35.853696e6 / max(e.last_ms(1), 1e-9) / 1e9), flush=True)
