#!/usr/bin/env python3
"""Tiles-per-launch sweep (`max_batch` option) on 1920x1080: device ms per call and output identity across batch sizes."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
e = rb.RealSR(0)
assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0
for kv in os.environ.get("RSR_OPTS", "").split(","):  # e.g. RSR_OPTS=serpentine=0,merge_shapes=0
    if "=" in kv:
        k, v = kv.split("=")
        assert e.set_option(k, int(v)) == 0, e.last_error()
img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
for tile, mbs in ((400, (0, 1, 2, 3, 4, 0)), (200, (0, 2, 3, 4, 6, 8, 16)), (256, (0, 2, 3, 4, 6))):
    e.tilesize = tile
    ref = None
    for mb in mbs:
        e.set_option("max_batch", mb)
        ms = []
        for i in range(6):
            assert e.process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr()) == 0
            if i >= 2:
                ms.append(e.last_ms())
        o = out.clone()
        same = True if ref is None else bool(torch.equal(o, ref))
        if ref is None:
            ref = o
        print("tile %d max_batch %2d: median %.2f ms (min %.2f) -> %.1f out-MP/s  identical=%s" %
              (tile, mb, np.median(ms), min(ms), 33.1776 / np.median(ms) * 1e3, same), flush=True)
