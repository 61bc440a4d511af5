#!/usr/bin/env python3
"""Turns the raw ncu captures of a gpurun call (gpurun_out/<R>_launches.csv, <R>_conv_full.ncu-rep) into the committed summaries
under profiles/.  Needs `ncu` (reads the .ncu-rep); run in the build container after the gpurun call came back.
   summarise_profiles.py <R>        e.g. r2"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
PX = 2302800  # network-input pixels of the 1080p / tile-400 step: every conv launch (merged shapes) covers all of them

# ---- launch list ----
lines = [l for l in open(os.path.join(G, R + "_launches.csv")) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void unnamed>::", "").replace("unnamed>::", "").replace("rsr::<", "").strip()
    v, u = float(row["Metric Value"].replace(",", "")), row["Metric Unit"]
    tot[k] += v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    cnt[k] += 1
T = sum(tot.values())
summ = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -s <3rd call> -c <one call> --csv python bench.py --profile --steps 1 --warmup 2",
        "note": "one process() call of the bench (BASELINE configs[1]) after two warm-up calls; ncu times are cold-cache and serialised (programmatic-"
                "dependent-launch overlap is lost under the profiler, and the GPU is not power-capped between replays): compare SHARES with the live "
                "CUDA-event numbers, not absolutes.  Template arguments: conv3x3_kernel<Cout_pad, ConvMode>",
        "total_ms": T, "launches": sum(cnt.values()),
        "kernels": [{"kernel": k, "launches": cnt[k], "total_ms": v, "share": v / T, "avg_us": v / cnt[k] * 1e3}
                    for k, v in sorted(tot.items(), key=lambda x: -x[1])]}
json.dump(summ, open(os.path.join(P, R + "_launches_summary.json"), "w"), indent=1)
with open(os.path.join(P, R + "_launches.csv"), "w") as f:
    f.writelines(lines)

# ---- full capture ----
rep = os.path.join(G, R + "_conv_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]
FLOP = {64: 2 * 9 * 64 * 32, 96: 2 * 9 * 96 * 32, 128: 2 * 9 * 128 * 32, 160: 2 * 9 * 160 * 32, 192: 2 * 9 * 192 * 64}
def num(s):
    return float(s.replace(",", ""))
def to_bytes(v, u):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_us(v, u):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
kernels = []
seq = ["conv1 64->32", "conv2 96->32", "conv3 128->32", "conv4 160->32", "conv5 192->64 (+0.2x, +x0)"] * 2
cin = [64, 96, 128, 160, 192] * 2
for n, d in enumerate(data):
    rec = {"layer": seq[n] if n < len(seq) else "?"}
    for k in KEEP:
        if k in ix:
            rec[k] = (d[ix[k]] + " " + units[ix[k]]).strip()
    t_us = to_us(d[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
    rd = to_bytes(d[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
    wr = to_bytes(d[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
    rec["derived_dram_TBps"] = (rd + wr) / t_us / 1e6
    rec["derived_dram_bytes_per_px"] = (rd + wr) / PX
    rec["derived_dram_bytes"] = rd + wr
    if n < len(cin):
        rec["derived_useful_TFLOPs"] = FLOP[cin[n]] * PX / t_us / 1e6
    kernels.append(rec)
json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s <2 calls + 35> -c 10 python bench.py --profile --steps 1 --warmup 2",
           "note": "10 consecutive conv launches of the third process() call: dense blocks 7 and 8 (block 8 closes an RRDB: its conv5 also reads and "
                   "writes the fp32 RRDB skip stream); every launch covers all 15 padded tiles of the frame (2 302 800 px, merged shape groups)",
           "kernels": kernels}, open(os.path.join(P, R + "_conv_full_summary.json"), "w"), indent=1)
tb = [k["derived_dram_bytes"] for k in kernels]
traffic = {"source": "profiles/%s_conv_full_summary.json (ncu --set full capture of 10 consecutive conv launches = two dense blocks)" % R,
           "dram_bytes_per_launch": sum(tb) / len(tb),
           "dram_bytes_per_px_per_dense_block": sum(tb) / 2 / PX,
           "note": "dram__bytes_read.sum + dram__bytes_write.sum, average over the 10 captured launches (a step has 345 dense-block launches of this "
                   "mix + 5 tail launches)"}
json.dump(traffic, open(os.path.join(P, R + "_conv_traffic.json"), "w"), indent=1)
print(json.dumps(traffic, indent=1))
print("step total %.2f ms" % T, [(k["kernel"], round(k["share"], 3)) for k in summ["kernels"][:5]])
for k in kernels:
    print(k["layer"], k["gpu__time_duration.sum"], "tensor", k.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
          "dram %.2f TB/s" % k["derived_dram_TBps"], "%.0f B/px" % k["derived_dram_bytes_per_px"], "%.0f TF/s" % k.get("derived_useful_TFLOPs", 0))
