#!/usr/bin/env python3
"""Per-tile pipeline timestamps of one conv launch (RSR_DBG_TS=1): ts_probe.py <conv_index> [H W]"""
import os, sys
import numpy as np
os.environ["RSR_DBG_TS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
idx = int(sys.argv[1]); H = int(sys.argv[2]) if len(sys.argv) > 2 else 1200; W = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
e = rb.RealSR(0)
assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0
cin = {1: 64, 2: 96, 3: 128, 4: 160, 5: 192}.get(idx, 64); cout = 64 if idx in (5,) or idx > 340 else 32
x = np.zeros((H, W, cin), dtype=np.float16); x[::7, ::5] = 0.25
e.debug_conv(idx, x, cout)
