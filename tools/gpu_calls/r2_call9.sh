#!/bin/bash
t0=$(date +%s); mark() { echo "[call9] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4; mark gpu-tests
timeout 300 python tools/ab.py "" "residual_ext=0" "l2_persist_mb=48" "l2_persist_mb=80" "residual_fp16=0" 2>&1 | tail -6; mark ab
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts9_5.log 2>&1; sed -n 30,36p gpurun_out/ts9_5.log
RSR_DBG_RDB=2 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts9_5e.log 2>&1; sed -n 30,34p gpurun_out/ts9_5e.log; mark ts
timeout 120 python tools/bench_dir.py --gpus 1 --images 6 --e2e-images 6 2>&1 | tail -1 | cut -c1-1200; mark bench_dir-smoke
