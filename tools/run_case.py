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
img = np.random.default_rng(1000).integers(0, 256, (h, w, 3), dtype=np.uint8)
out = np.zeros((h * 4, w * 4, 3), dtype=np.uint8)
for i in range(iters):
    t0 = time.time()
    r = e.process(img, out)
    print("rc", r, "wall %.2f ms device %.2f ms" % ((time.time() - t0) * 1e3, e.last_ms()), "launches", e.kernel_launches(), flush=True)
