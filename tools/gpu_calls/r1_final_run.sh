#!/bin/bash
# Round-end validation + profile capture on the GPU box (one gpurun call): tools/final_run.sh
# Outputs under gpurun_out/: bench_r1_n1.json, r1_launches.csv, r1_conv_full.ncu-rep, ts_<idx>.log, knockout.log, cfg.log
# Environment: RSR_BENCH_OPTS="--opts merge_shapes=1,..." to measure a non-default configuration; then also set
#   RSR_LAUNCHES_PER_STEP (all kernels of one process() call: 1412 by default, 362 with merge_shapes=1) and
#   RSR_TC2_PER_STEP (tensor-core conv launches of one call: 1404 by default, 351 merged) so that ncu skips to the 4th call.
R=${RSR_ROUND:-r1}
L=${RSR_LAUNCHES_PER_STEP:-1412}
K=${RSR_TC2_PER_STEP:-1404}
t0=$(date +%s); mark() { echo "[final_run] $1 @ $(( $(date +%s) - t0 )) s"; }
python -m pytest tests -m gpu -x -q --durations=5 -k "${RSR_FINAL_K:-constant or full_size or errors or single_cta or cpp_wrapper or cli_directory}" 2>&1 | tail -10; mark tests
python __graft_entry__.py smoke 2>&1 | tail -1; mark smoke
python bench.py --steps 10 --warmup 3 $RSR_BENCH_OPTS 2>&1 | tail -1 > gpurun_out/bench_${R}_n1.json; cut -c1-260 gpurun_out/bench_${R}_n1.json; mark bench
# under ncu every launch costs ~0.1 s: profile 3 process() calls only (bench.py --profile: no e2e leg, no event pass) and keep the 3rd
ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --profile --steps 1 --warmup 2 $RSR_BENCH_OPTS > gpurun_out/${R}_bench_under_ncu.log 2>&1; mark launch-list
# 10 consecutive tensor-core conv launches of the 8 x 420x420 batch (largest shape runs first): RDB 7 conv1..5 and RDB 8 (closes RRDB 2) conv1..5
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s $((2 * K + 35)) -c 10 -o gpurun_out/${R}_conv_full -f \
    python bench.py --profile --steps 1 --warmup 2 $RSR_BENCH_OPTS > gpurun_out/${R}_full.log 2>&1; tail -c 120 gpurun_out/${R}_full.log; mark ncu-full
for i in 1 4; do timeout 60 python tools/ts_probe.py $i > gpurun_out/ts_$i.log 2>&1; done
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts_5.log 2>&1; mark ts
bash tools/run_knockout.sh > gpurun_out/knockout.log 2>&1; mark knockout
timeout 120 python tools/bench_configs.py > gpurun_out/cfg.log 2>&1; mark configs
