#!/bin/bash
# 6th gpurun call of round 2: early hand-back for the Cout=64 residual convs, serpentine, fp16+int8 residual carrier
t0=$(date +%s); mark() { echo "[call6] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; mark gpu-tests
timeout 300 python tools/parity_report.py "" "residual_ext=0" "residual_fp16=0" 2>&1 | grep -E "natural|readme|cfg" ; mark parity
timeout 300 python tools/ab.py "" "serpentine=0" "residual_ext=0" "serpentine=0,residual_ext=0" 2>&1 | tail -5; mark ab
timeout 200 python tools/ab.py --tile 200 "" "serpentine=0" 2>&1 | tail -3; mark ab-t200
for i in 1 2; do timeout 60 python tools/ts_probe.py $i > gpurun_out/ts6_$i.log 2>&1; echo "ts conv $i rc=$?"; sed -n 30,33p gpurun_out/ts6_$i.log; done
timeout 60 python tools/ts_probe.py 347 > gpurun_out/ts6_347.log 2>&1; sed -n 30,36p gpurun_out/ts6_347.log
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts6_5.log 2>&1; sed -n 30,36p gpurun_out/ts6_5.log; mark ts
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_call6.json 2> gpurun_out/bench_call6.err; python -c "
import json;d=json.load(open('gpurun_out/bench_call6.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_pageable','clocks','parity')});print(d['roofline'])"; tail -3 gpurun_out/bench_call6.err; mark bench
