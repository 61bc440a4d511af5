#!/bin/bash
# ncu evidence of the current build on BASELINE configs[1] (one gpurun call): launch list of one step + --set full capture of 10 consecutive
# conv launches (dense blocks 7 and 8: block 8 closes an RRDB) + --set full of the tail convs.  Outputs gpurun_out/<R>_*; then run
# tools/summarise_profiles.py <R> in the build container.
R=${RSR_ROUND:-r2}
t0=$(date +%s); mark() { echo "[profile] $1 @ $(( $(date +%s) - t0 )) s"; }
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
ncu --metrics gpu__time_duration.sum --clock-control none -s $((2 * L)) -c $L --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --profile --steps 1 --warmup 2 > gpurun_out/${R}_bench_under_ncu.log 2>&1; tail -1 gpurun_out/${R}_bench_under_ncu.log; mark launch-list
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s $((2 * 350 + 35)) -c 10 -o gpurun_out/${R}_conv_full -f \
    python bench.py --profile --steps 1 --warmup 2 > gpurun_out/${R}_full.log 2>&1; tail -c 150 gpurun_out/${R}_full.log; mark ncu-full
