#!/bin/bash
# second gpurun call of round 2: epilogue micro-benchmarks + first run of the unified conv kernel (csrc/conv3x3.cu)
t0=$(date +%s); mark() { echo "[call2] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 120 tools/ubench2 > gpurun_out/r2_ubench2.txt 2>&1; cat gpurun_out/r2_ubench2.txt; mark ubench2
timeout 120 tools/ubench > gpurun_out/r2_ubench.txt 2>&1; head -6 gpurun_out/r2_ubench.txt; mark ubench
for k in "layers_bit_identical and 2-4" "layers_bit_identical and 3-5" "images_bit_identical and k64 and not merged and not res32" k64-merged k64-res32 k32-merged "tta_bit_identical and 4" "tta_bit_identical and 5"; do
  timeout 200 python -m pytest tests/test_conv_unified_gpu.py -m gpu -q -x -k "$k" 2>&1 | tail -5; echo "[$k] rc=$?"
done; mark tests
timeout 200 python tools/ab.py "conv_impl=2" "conv_impl=4" "conv_impl=3" "conv_impl=5" 2>&1 | tail -5; mark ab-plain
timeout 200 python tools/ab.py "conv_impl=2,merge_shapes=1" "conv_impl=4,merge_shapes=1" "conv_impl=5,merge_shapes=1" "conv_impl=5,merge_shapes=1,residual_fp16=0" 2>&1 | tail -5; mark ab-merged
timeout 200 python tools/ab.py --tile 200 "conv_impl=2" "conv_impl=4,merge_shapes=1" "conv_impl=5,merge_shapes=1" 2>&1 | tail -4; mark ab-t200
for impl in 4 5; do for i in 1 2 4 5 347; do RSR_IMPL=$impl RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py $i > gpurun_out/ts_i${impl}_$i.log 2>&1; echo "ts impl $impl conv $i rc=$?"; sed -n 30,34p gpurun_out/ts_i${impl}_$i.log; done; done; mark ts
