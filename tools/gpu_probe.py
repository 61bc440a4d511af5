#!/usr/bin/env python3
"""Bring-up probe (run under gpurun): per-layer check of the tcgen05 conv against torch, end-to-end vs goldens, timing."""
import json, os, sys, time
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realsr_ncnn_vulkan_b200 as rb

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def read_convs(model):
    """independent numpy reader of the .bin (format: modelbin.cpp:88-296) for the torch reference"""
    mdir = os.path.join(ROOT, "models", model)
    convs = []
    for line in open(os.path.join(mdir, "x4.param")):
        p = line.split()
        if p and p[0] == "Convolution":
            kv = dict(t.split("=", 1) for t in p if "=" in t)
            convs.append(dict(cout=int(kv["0"]), wsize=int(kv["6"]), act=int(kv.get("9", 0))))
    buf = open(os.path.join(mdir, "x4.bin"), "rb").read()
    off = 0
    for c in convs:
        tag = int.from_bytes(buf[off:off + 4], "little"); off += 4
        assert tag == 0x01306B47
        n = c["wsize"]
        w = np.frombuffer(buf, dtype=np.float16, count=n, offset=off); off += (n * 2 + 3) & ~3
        c["cin"] = n // (c["cout"] * 9)
        c["w"] = w.reshape(c["cout"], c["cin"], 3, 3)
        c["b"] = np.frombuffer(buf, dtype=np.float32, count=c["cout"], offset=off); off += 4 * c["cout"]
    assert off == len(buf)
    return convs


def main():
    out = {}
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    eng = rb.RealSR(0)
    r = eng.load(os.path.join(ROOT, "models/models-DF2K/x4.param"), os.path.join(ROOT, "models/models-DF2K/x4.bin"))
    print("load", r, eng.last_error())
    assert r == 0
    convs = read_convs("models-DF2K")
    rng = np.random.default_rng(0)
    nconv = len(convs)
    for mode in (0,):
        for idx in (1, 2, 3, 4, 5, nconv - 5, nconv - 1):
            c = convs[idx]
            H, W = 37, 45
            x = (rng.standard_normal((H, W, c["cin"])) * 0.5).astype(np.float16)
            try:
                y = eng.debug_conv(idx, x, c["cout"])
            except Exception as e:
                print("mode", mode, "conv", idx, "FAILED", e)
                continue
            xt = torch.from_numpy(x.astype(np.float32)).to(dev).permute(2, 0, 1)[None]
            wt = torch.from_numpy(c["w"].astype(np.float32)).to(dev)
            bt = torch.from_numpy(c["b"].copy()).to(dev)
            ref = F.conv2d(xt, wt, bt, padding=1)
            if c["act"] == 2:
                ref = F.leaky_relu(ref, 0.2)
            ref = ref[0].permute(1, 2, 0).cpu().numpy()
            if c["cout"] != 3:
                ref = ref.astype(np.float16).astype(np.float32)
            err = np.abs(y - ref)
            print("conv %3d cin %3d cout %2d: max err %.5f  mean %.6f  (ref absmax %.3f)" %
                  (idx, c["cin"], c["cout"], err.max(), err.mean(), np.abs(ref).max()), flush=True)
            out["conv_m%d_%d" % (mode, idx)] = float(err.max())
    # pick the descriptor mode that works
    good = [m for m in (0,) if all(out.get("conv_m%d_%d" % (m, i), 1e9) < 0.02 for i in (1, 2, 3, 4, 5))]
    print("per-layer checks ok:", bool(good))
    out["good_modes"] = good
    if good:
        man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
        engines = {}
        for case in man["cases"]:
            key = (case["model"], case["tta"])
            if key not in engines:
                e = rb.RealSR(0, tta_mode=bool(case["tta"]))
                assert e.load(os.path.join(ROOT, "models", case["model"], "x4.param"), os.path.join(ROOT, "models", case["model"], "x4.bin")) == 0
                engines[key] = e
            e = engines[key]
            e.tilesize = case["tilesize"]
            img = np.load(os.path.join(ROOT, "tests/golden", case["name"] + ".in.npy"))
            gold = np.load(os.path.join(ROOT, "tests/golden", case["name"] + ".out.npy"))
            res = np.zeros_like(gold)
            r = e.process(img, res)
            d = np.abs(res.astype(int) - gold.astype(int))
            print("%-32s rc %d  max %d  differ %.3f%%  >1: %d   (%.2f ms, %d padded px)" %
                  (case["name"], r, d.max(), 100.0 * (d > 0).mean(), int((d > 1).sum()), e.last_ms(), e.last_padded_pixels()), flush=True)
            out["gold_" + case["name"]] = [int(d.max()), float((d > 0).mean())]
        # timing: cfg2
        e = engines[("models-DF2K", 0)]
        e.tilesize = 400
        img = np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        res = np.zeros((4320, 7680, 3), dtype=np.uint8)
        e.set_option("time_convs", 1)
        for it in range(3):
            t0 = time.time()
            r = e.process(img, res)
            t1 = time.time()
            ms, cms, pp = e.last_ms(), e.last_ms(1), e.last_padded_pixels()
            print("cfg2 iter %d rc %d: wall %.1f ms, device %.1f ms, conv kernels %.1f ms, %.1f out-MP/s, conv TFLOP/s %.1f" %
                  (it, r, (t1 - t0) * 1e3, ms, cms, 33.1776 / ms * 1e3, pp * 35.853696e6 / (cms * 1e-3) / 1e12), flush=True)
        out["cfg2_ms"] = ms
        out["cfg2_conv_ms"] = cms
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out/probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
