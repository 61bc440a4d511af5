#!/usr/bin/env python3
"""Interleaved A/B/... of engine option sets on one box (device-resident, CUDA-event ms per call):
   ab.py [--tile T] [--size WxH] [--tta] "" "merge_shapes=0" "serpentine=0,max_batch=8" ...
Prints the median / min ms of every set, the conv-kernel share (time_convs run) and whether the outputs are identical."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb
args = sys.argv[1:]
tile, W, H, tta, rounds = 400, 1920, 1080, False, 5
while args and args[0].startswith("--"):
    k = args.pop(0)
    if k == "--tile": tile = int(args.pop(0))
    elif k == "--size": W, H = map(int, args.pop(0).split("x"))
    elif k == "--tta": tta = True
    elif k == "--rounds": rounds = int(args.pop(0))
sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv) for s in args]
eng = []
for o in sets:
    e = rb.RealSR(0, tta_mode=tta)
    assert e.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin")) == 0
    e.tilesize = tile
    for k, v in o.items():
        assert e.set_option(k, v) == 0, e.last_error()
    eng.append(e)
img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda()
outs = [torch.empty((4 * H, 4 * W, 3), dtype=torch.uint8, device="cuda") for _ in sets]
ms = [[] for _ in sets]
for rnd in range(rounds + 1):
    for i, e in enumerate(eng):
        for _ in range(3):
            assert e.process_device(img.data_ptr(), W, H, 3, outs[i].data_ptr()) == 0, e.last_error()
            if rnd > 0:
                ms[i].append(e.last_ms())
conv = []
for i, e in enumerate(eng):  # one extra call with per-conv events (adds launch gaps: only the conv sum is used)
    e.set_option("time_convs", 1)
    assert e.process_device(img.data_ptr(), W, H, 3, outs[i].data_ptr()) == 0
    conv.append((e.last_ms(1), e.last_padded_pixels()))
    e.set_option("time_convs", 0)
mp = 16.0 * W * H / 1e6
for i, o in enumerate(sets):
    a = np.array(ms[i])
    cms, px = conv[i]
    print("%-40s median %.2f ms  min %.2f  max %.2f -> %.1f out-MP/s | conv %.2f ms = %.0f TFLOP/s (timed separately) | same as set 0: %s"
          % (args[i], np.median(a), a.min(), a.max(), mp / np.median(a) * 1e3, cms, 35.853696e6 * px / (cms * 1e-3) / 1e12,
             bool(torch.equal(outs[0], outs[i]))))
