#!/bin/bash
# accumulator hand-back before the tap shuffles: tests, same-box A/B against the previous build, job-size A/B, sanitizers, all configs
t0=$(date +%s); mark() { echo "[call14] $1 @ $(( $(date +%s) - t0 )) s"; }
L=realsr-ncnn-vulkan_b200/librealsr_b200.so
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; mark gpu-tests
timeout 600 python tools/ab_builds.py --rounds 3 "head=$L" "tma_rdb=ablibs/tma_rdb.so" 2>&1 | tail -4; mark ab-builds
timeout 300 python tools/ab.py "" "max_batch=8" "max_batch=4" "serpentine=0" 2>&1 | tail -5; mark ab-opts
for i in 1 2; do timeout 60 python tools/ts_probe.py $i > gpurun_out/ts14_$i.log 2>&1; sed -n 30,33p gpurun_out/ts14_$i.log; done; mark ts
bash tools/run_sanitizers.sh; mark sanitizers
timeout 600 python tools/bench_configs.py > gpurun_out/configs14.log 2>&1; tail -8 gpurun_out/configs14.log | cut -c1-400; mark configs
