#!/usr/bin/env python3
"""Measures every single-GPU BASELINE.json config (not only the headline one bench.py reports):
   cfg1 256x256 (T=200, T=256, DF2K_JPEG), cfg2 1920x1080 T=400, cfg4 3840x2160 T=200, cfg5 1280x720 TTA T=200.
Device-resident input/output (rsr_process_device), CUDA-event time of the call, median of 5 after 2 warm-ups.
Writes gpurun_out/r2_configs.json (copy it to profiles/)."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb

CFG = [("cfg1 256x256 T=200 DF2K_JPEG", "models-DF2K_JPEG", 256, 256, 200, False),
       ("cfg1 256x256 T=256 DF2K_JPEG", "models-DF2K_JPEG", 256, 256, 256, False),
       ("cfg2 1920x1080 T=400 DF2K", "models-DF2K", 1920, 1080, 400, False),
       ("cfg4 3840x2160 T=200 DF2K", "models-DF2K", 3840, 2160, 200, False),
       ("cfg5 1280x720 T=200 TTA DF2K", "models-DF2K", 1280, 720, 200, True)]
out = []
for name, model, w, h, tile, tta in CFG:
    e = rb.RealSR(0, tta_mode=tta)
    d = os.path.join(ROOT, "models", model)
    assert e.load(os.path.join(d, "x4.param"), os.path.join(d, "x4.bin")) == 0, e.last_error()
    e.tilesize = tile
    img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (h, w, 3), dtype=np.uint8)).cuda()
    res = torch.empty((h * 4, w * 4, 3), dtype=torch.uint8, device="cuda")
    ms = []
    for i in range(7):
        assert e.process_device(img.data_ptr(), w, h, 3, res.data_ptr()) == 0, e.last_error()
        if i >= 2:
            ms.append(e.last_ms())
    e.set_option("time_convs", 1)
    e.process_device(img.data_ptr(), w, h, 3, res.data_ptr())
    cms, pp = e.last_ms(1), e.last_padded_pixels()
    m = float(np.median(ms))
    rec = dict(config=name, ms=m, out_mp_s=w * h * 16 / 1e6 / (m * 1e-3), padded_px=pp, conv_ms=cms,
               conv_tflops=pp * 35.853696e6 / (cms * 1e-3) / 1e12, executed_tflop=pp * 35.853696e6 / 1e12)
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del e
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_configs.json"), "w"), indent=1)
