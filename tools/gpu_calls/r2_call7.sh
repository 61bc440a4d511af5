#!/bin/bash
t0=$(date +%s); mark() { echo "[call7] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4; mark parity-tests
timeout 300 python tools/parity_report.py "" 2>&1 | grep -E "natural|readme|cfg" ; mark parity
timeout 300 python tools/ab.py "" "residual_ext=0" "residual_fp16=0" "serpentine=0" 2>&1 | tail -5; mark ab
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts7_5.log 2>&1; sed -n 30,36p gpurun_out/ts7_5.log; mark ts
timeout 200 python tools/bench_configs.py 2>&1 | tail -6; mark configs
