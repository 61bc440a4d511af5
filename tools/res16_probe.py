#!/usr/bin/env python3
"""Precision and speed of the residual_fp16 option vs the default fp32 residual stream (goldens + cfg2 timing)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import realsr_ncnn_vulkan_b200 as rb
from common import diff_stats, golden, manifest, model_paths
for mode in (0, 1):
    print("residual_fp16 =", mode)
    engs = {}
    for case in manifest()["cases"]:
        key = (case["model"], case["tta"])
        if key not in engs:
            e = rb.RealSR(0, tta_mode=bool(case["tta"])); p, b = model_paths(case["model"]); assert e.load(p, b) == 0
            e.set_option("residual_fp16", mode); engs[key] = e
        e = engs[key]; e.tilesize = case["tilesize"]
        img, gold = golden(case["name"]); out = np.zeros_like(gold)
        assert e.process(img, out) == 0
        s = diff_stats(out[..., :3], gold[..., :3])
        print("  %-32s max %d differ %.3f%% >1: %d" % (case["name"], s["max"], 100 * s["frac_diff"], s["n_gt1"]))
    e = engs[("models-DF2K", 0)]; e.tilesize = 400
    img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
    out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(8):
        assert e.process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr()) == 0
        if i >= 3: ms.append(e.last_ms())
    print("  cfg2 median %.2f ms" % np.median(ms))
