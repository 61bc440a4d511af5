#!/usr/bin/env python3
"""Parity of the engine against every committed golden (outputs of the real reference CPU path), per option set:
   parity_report.py "" "merge_shapes=0" ...
Prints max |diff|, % of bytes that differ and bytes beyond 1 LSB per case; writes gpurun_out/parity_report.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import realsr_ncnn_vulkan_b200 as rb
from common import diff_stats, golden, manifest, model_paths
sets = sys.argv[1:] or [""]
cases = manifest()["cases"]
report = {}
for s in sets:
    opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv)
    eng = {}
    rows = []
    for c in cases:
        key = (c["model"], bool(c["tta"]))
        if key not in eng:
            e = rb.RealSR(0, tta_mode=bool(c["tta"]))
            assert e.load(*model_paths(c["model"])) == 0, e.last_error()
            for k, v in opts.items():
                assert e.set_option(k, v) == 0, e.last_error()
            eng[key] = e
        e = eng[key]
        img, gold = golden(c["name"])
        e.tilesize = c["tilesize"]
        out = np.zeros_like(gold)
        assert e.process(img, out) == 0, e.last_error()
        st = diff_stats(out[..., :3], gold[..., :3])
        rows.append(dict(name=c["name"], **st))
        print("%-46s %-44s max %d  differ %.3f %%  >1LSB %d / %d" % (s or "defaults", c["name"], st["max"], 100 * st["frac_diff"], st["n_gt1"], st["n"]))
    report[s or "defaults"] = rows
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
