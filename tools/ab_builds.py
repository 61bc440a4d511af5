#!/usr/bin/env python3
"""Same-box A/B of several BUILDS of librealsr_b200.so (one subprocess per build: RSR_LIB selects the library), BASELINE configs[1],
device-resident, CUDA-event ms per call.  Usage: ab_builds.py [--rounds R] name=path[:opt=v,opt=v] ...
Builds of earlier commits: git worktree add /tmp/wt_<c> <c>; bash /tmp/wt_<c>/realsr-ncnn-vulkan_b200/csrc/build.sh; copy the .so."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
import numpy as np, torch
sys.path.insert(0, %r)
import realsr_ncnn_vulkan_b200 as rb
opts = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in sys.argv[1].split(",") if kv)
e = rb.RealSR(0)
assert e.load(os.path.join(%r, "models/models-DF2K/x4.param"), os.path.join(%r, "models/models-DF2K/x4.bin")) == 0
e.tilesize = 400
for k, v in opts.items():
    assert e.set_option(k, v) == 0, e.last_error()
img = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)).cuda()
out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
ms = []
for i in range(int(sys.argv[2]) + 3):
    assert e.process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr()) == 0, e.last_error()
    if i >= 3: ms.append(e.last_ms())
e.set_option("time_convs", 1)
e.process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr())
print(json.dumps({"ms": ms, "conv_ms": e.last_ms(1), "sum": int(out[::7, ::5].to(torch.int64).sum())}))
''' % (ROOT, ROOT, ROOT)
args = sys.argv[1:]
rounds, iters = 2, 8
if args and args[0] == "--rounds":
    rounds = int(args[1]); args = args[2:]
builds = []
for a in args:
    name, rest = a.split("=", 1)
    path, _, opts = rest.partition(":")
    builds.append((name, path, opts))
res = {b[0]: {"ms": [], "conv": []} for b in builds}
for r in range(rounds):
    for name, path, opts in builds:
        env = dict(os.environ, RSR_LIB=os.path.abspath(path))
        p = subprocess.run([sys.executable, "-c", CHILD, opts, str(iters)], capture_output=True, text=True, env=env, timeout=300)
        if p.returncode:
            print(name, "FAILED", p.stderr[-300:]); continue
        d = json.loads(p.stdout.strip().splitlines()[-1])
        res[name]["ms"] += d["ms"]; res[name]["conv"].append(d["conv_ms"]); res[name]["sum"] = d["sum"]
for name, path, opts in builds:
    m = np.array(res[name]["ms"])
    if len(m):
        print("%-34s median %.2f ms  min %.2f  max %.2f | conv (events) %s | checksum %s" % (name + (" [" + opts + "]" if opts else ""), np.median(m), m.min(), m.max(),
              ["%.1f" % c for c in res[name]["conv"]], res[name].get("sum")))
