#!/bin/bash
# 4th gpurun call of round 2: tcgen05.shift probe, parity report, full GPU test suite on the new defaults, ncu launch list + full capture
t0=$(date +%s); mark() { echo "[call4] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 60 tools/ubench3 > gpurun_out/r2_ubench3.txt 2>&1; head -60 gpurun_out/r2_ubench3.txt; mark ubench3
timeout 300 python tools/parity_report.py "" "residual_fp16=0" 2>&1 | tail -40; mark parity
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; mark gpu-tests
L=$(python - <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import realsr_ncnn_vulkan_b200 as rb
e = rb.RealSR(0); e.load("models/models-DF2K/x4.param", "models/models-DF2K/x4.bin"); e.tilesize = 400
img = torch.zeros((1080, 1920, 3), dtype=torch.uint8, device="cuda"); out = torch.empty((4320, 7680, 3), dtype=torch.uint8, device="cuda")
l0 = e.kernel_launches(); e.process_device(img.data_ptr(), 1920, 1080, 3, out.data_ptr()); print(e.kernel_launches() - l0)
PY
)
echo "launches per step: $L"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --profile --steps 1 --warmup 2 > gpurun_out/r2_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r2_bench_under_ncu.log; mark launch-list
# RDB 7 conv1..5 and RDB 8 conv1..5 (10 consecutive conv launches) of the third call
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s $((2 * 350 + 35)) -c 10 -o gpurun_out/r2_conv_full -f \
    python bench.py --profile --steps 1 --warmup 2 > gpurun_out/r2_full.log 2>&1; tail -c 200 gpurun_out/r2_full.log; mark ncu-full
