#!/bin/bash
# 5th gpurun call of round 2: single conv kernel (N=192 Cout=64 path, per-K-block planes), full GPU test suite, bench
t0=$(date +%s); mark() { echo "[call5] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv_layer_vs_torch" 2>&1 | tail -6; mark layer-tests
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; mark gpu-tests
timeout 200 python tools/ab.py "" "merge_shapes=0" "residual_fp16=0" 2>&1 | tail -4; mark ab
timeout 200 python tools/ab.py --tile 200 "" "merge_shapes=0" 2>&1 | tail -3; mark ab-t200
for i in 1 2 4; do timeout 60 python tools/ts_probe.py $i > gpurun_out/ts5_$i.log 2>&1; echo "ts conv $i rc=$?"; sed -n 30,34p gpurun_out/ts5_$i.log; done
timeout 60 python tools/ts_probe.py 347 > gpurun_out/ts5_347.log 2>&1; sed -n 30,36p gpurun_out/ts5_347.log
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts5_5.log 2>&1; sed -n 30,36p gpurun_out/ts5_5.log; mark ts
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_call5.json 2> gpurun_out/bench_call5.err; python -c "
import json;d=json.load(open('gpurun_out/bench_call5.json'));print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_pageable','clocks','parity')});print(d['roofline'])"; tail -3 gpurun_out/bench_call5.err; mark bench
timeout 300 python tools/parity_report.py "" 2>&1 | tail -16; mark parity
