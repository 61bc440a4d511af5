#!/bin/bash
t0=$(date +%s); mark() { echo "[call10] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; mark gpu-tests
timeout 300 python tools/parity_report.py "" 2>&1 | grep -E "natural|readme|cfg" ; mark parity
# separate processes: cudaLimitPersistingL2CacheSize is device-wide state
for o in "" "l2_persist_mb=24" "l2_persist_mb=48" "" "serpentine=0"; do timeout 120 python tools/ab.py --rounds 4 "$o" 2>&1 | tail -1; done; mark ab
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts10_5.log 2>&1; sed -n 30,34p gpurun_out/ts10_5.log
RSR_DBG_RDB=2 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts10_5e.log 2>&1; sed -n 30,34p gpurun_out/ts10_5e.log; mark ts
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_call10.json 2> gpurun_out/bench_call10.err; python -c "
import json;d=json.load(open('gpurun_out/bench_call10.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_pageable','clocks','parity')});print(d['roofline'])"; tail -3 gpurun_out/bench_call10.err; mark bench
