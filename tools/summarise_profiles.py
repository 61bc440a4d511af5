#!/usr/bin/env python3
"""Turns the raw captures of tools/final_run.sh (gpurun_out/) into the committed summaries under profiles/.
Needs `ncu` (reads the .ncu-rep); run in the build container after the gpurun call came back."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r1"

# ---- launch list ----
lines = [l for l in open(os.path.join(G, R + "_launches.csv")) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void unnamed>::", "").replace("unnamed>::", "").strip()
    v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    tot[k] += v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    cnt[k] += 1
T = sum(tot.values())
summ = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -s <3rd call> -c <one call> --csv python bench.py ... (tools/final_run.sh)",
        "note": "one process() call of the bench after warm-up calls; ncu times are cold-cache and serialised (programmatic-dependent-launch overlap is lost "
                "under the profiler, and the GPU is not power-capped between replays): compare SHARES with the live CUDA-event numbers, not absolutes",
        "total_ms": T, "launches": sum(cnt.values()),
        "kernels": [{"kernel": k, "launches": cnt[k], "total_ms": v, "share": v / T, "avg_us": v / cnt[k] * 1e3}
                    for k, v in sorted(tot.items(), key=lambda x: -x[1])]}
json.dump(summ, open(os.path.join(P, R + "_launches_summary.json"), "w"), indent=1)
shutil.copy(os.path.join(G, R + "_launches.csv"), os.path.join(P, R + "_launches.csv"))

# ---- full capture ----
rep = os.path.join(G, R + "_conv_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier"]
NAMES = {64: "conv1 64->32", 96: "conv2 96->32", 128: "conv3 128->32", 160: "conv4 160->32", 192: "conv5 192->64 + residual"}
PX = 8 * 420 * 420
kernels = []
for d in data:
    rec = {}
    for k in KEEP:
        if k in ix:
            rec[k] = (d[ix[k]] + " " + units[ix[k]]).strip()
    t_us = float(d[ix["gpu__time_duration.sum"]].replace(",", ""))
    rd, wr = float(d[ix["dram__bytes_read.sum"]].replace(",", "")), float(d[ix["dram__bytes_write.sum"]].replace(",", ""))
    rec["derived_dram_TBps"] = (rd + wr) / t_us  # Mbyte / us
    rec["derived_dram_bytes_per_px"] = (rd + wr) * 1e6 / PX
    kernels.append(rec)
# identify the conv by its DRAM read volume (planes read) and kernel template
order = []
for r in kernels:
    rd = float(r["dram__bytes_read.sum"].split()[0])
    planes = round(rd / 180.7)
    r["planes_read"] = planes
json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2_kernel -s 4247 -c 10 python bench.py --steps 1 --warmup 1 --no-cpu-baseline",
           "note": "10 consecutive tensor-core conv launches of the 8 x 420x420 batch (3024 pair tiles on 74 persistent CTA pairs): "
                   "conv3, conv4, conv5, conv1, conv2, conv3, conv4, conv5, conv1, conv2 of consecutive dense blocks; "
                   "derived_* = (dram read + write) / duration and / 1 411 200 px",
           "kernels": kernels}, open(os.path.join(P, R + "_conv_full_summary.json"), "w"), indent=1)
shutil.copy(rep, os.path.join(P, R + "_conv_full.ncu-rep"))

# ---- roofline.traffic for bench.py: DRAM bytes per launch, averaged over the dense-block convs (1380 of the 1400 conv launches of a step)
by_planes_mode = {}
for r in kernels:
    key = (r["planes_read"], "<64, 5>" in r["Kernel Name"], round(float(r["gpu__time_duration.sum"].split()[0]) / 20))
    by_planes_mode.setdefault((r["Kernel Name"], r["planes_read"]), []).append(r)
per_px = {}
for r in kernels:
    rd, wr = float(r["dram__bytes_read.sum"].split()[0]), float(r["dram__bytes_write.sum"].split()[0])
    name = ("conv5" if "<64, 5>" in r["Kernel Name"] else {1: "conv1", 3: "conv4"}.get(r["planes_read"]))
    if name is None:  # two planes: conv2 (96 ch) is the faster one
        name = "conv2or3"
    per_px.setdefault(name, []).append((rd + wr) * 1e6 / PX)
avg = {k: sum(v) / len(v) for k, v in per_px.items()}
rdb_bytes_per_px = avg["conv1"] + 2 * avg["conv2or3"] + avg["conv4"] + avg["conv5"]
padded_px, nrdb = 2302800, 69
traffic = {"source": "profiles/%s_conv_full.ncu-rep (ncu --set full, 10 consecutive conv launches of the 8 x 420x420 batch)" % R,
           "dram_bytes_per_px": avg, "dram_bytes_per_px_per_dense_block": rdb_bytes_per_px,
           "dram_bytes_per_launch": rdb_bytes_per_px * padded_px * nrdb / (nrdb * 5 * 4),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum of the five convs of a dense block, per network-input pixel, scaled to the "
                   "2 302 800 padded px of the 1080p step and divided by its 69 x 5 x 4 dense-block launches (4 shape groups). The 20 other conv "
                   "launches of a step (RRDB-closing extras, trunk, 4x tail) are not in the capture. Reads equal the planes a conv must fetch "
                   "(180.6 MB per 64-channel fp16 plane of this batch): halo rows are re-read from L2, not from DRAM."}
json.dump(traffic, open(os.path.join(P, R + "_conv_traffic.json"), "w"), indent=1)
for f in (R + "_configs.json",):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
shutil.copy(os.path.join(G, "bench_%s_n1.json" % R), os.path.join(P, R + "_bench_n1.json"))
for f in ("ts_1.log", "ts_4.log", "ts_5.log", "knockout.log"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, R + "_" + f.replace(".log", ".txt")))
print(json.dumps(traffic, indent=1))
print("step total %.2f ms" % T, [(k["kernel"], round(k["share"], 3)) for k in summ["kernels"][:4]])
