#!/bin/bash
# conv5 with the TMA-native residual epilogue: layer tests, full GPU suite, same-box A/B against the previous build, parity, ts probes
t0=$(date +%s); mark() { echo "[call13] $1 @ $(( $(date +%s) - t0 )) s"; }
L=realsr-ncnn-vulkan_b200/librealsr_b200.so
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv_layer_vs_torch" 2>&1 | tail -6; mark layer-tests
timeout 300 python tools/parity_report.py "" 2>&1 | tail -16; mark parity
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; mark gpu-tests
timeout 600 python tools/ab_builds.py --rounds 3 "head=$L" "prev=ablibs/prev_head.so" 2>&1 | tail -6; mark ab-builds
RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts13_5.log 2>&1; sed -n 30,36p gpurun_out/ts13_5.log
RSR_DBG_RDB=2 timeout 60 python tools/ts_probe.py 5 > gpurun_out/ts13_5e.log 2>&1; sed -n 30,36p gpurun_out/ts13_5e.log; mark ts
