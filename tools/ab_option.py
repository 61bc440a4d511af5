#!/usr/bin/env python3
"""Interleaved A/B of one engine option on BASELINE configs[1] (device-resident, CUDA-event ms per call):
   ab_option.py <option> <value_a> <value_b> [tile]      e.g.  ab_option.py epilogue_split 0 1"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
opt, vals = sys.argv[1], (int(sys.argv[2]), int(sys.argv[3]))
tile = int(sys.argv[4]) if len(sys.argv) > 4 else 400
eng = []
for v in vals:
    e = rb.RealSR(0)
    assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0
    e.tilesize = tile
    assert e.set_option(opt, v) == 0, e.last_error()
    eng.append(e)
img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
outs = [torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda") for _ in vals]
ms = [[], []]
for rnd in range(6):
    for i in (0, 1):
        for _ in range(3):
            assert eng[i].process_device(img.data_ptr(), 1920, 1080, 3, outs[i].data_ptr()) == 0, eng[i].last_error()
            if rnd > 0:
                ms[i].append(eng[i].last_ms())
for i in (0, 1):
    a = np.array(ms[i])
    print("%s=%d: median %.2f ms  min %.2f  max %.2f  -> %.1f out-MP/s" % (opt, vals[i], np.median(a), a.min(), a.max(), 33.1776 / np.median(a) * 1e3))
print("outputs identical:", bool(torch.equal(outs[0], outs[1])))
