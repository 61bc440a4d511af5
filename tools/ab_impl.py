#!/usr/bin/env python3
"""Interleaved A/B of the two conv implementations on BASELINE config 1 (device-resident, CUDA-event ms per call)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
eng = {}
for impl in (1, 2):
    e = rb.RealSR(0)
    assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0
    e.tilesize = 400
    e.set_option("conv_impl", impl)
    eng[impl] = e
img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
ms = {1: [], 2: []}
for rnd in range(6):
    for impl in (1, 2):
        for _ in range(3):
            assert eng[impl].process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr()) == 0
            if rnd > 0:
                ms[impl].append(eng[impl].last_ms())
for impl in (1, 2):
    a = np.array(ms[impl])
    print("impl %d: median %.2f ms  min %.2f  max %.2f  -> %.1f out-MP/s" % (impl, np.median(a), a.min(), a.max(), 33.1776 / np.median(a) * 1e3))
