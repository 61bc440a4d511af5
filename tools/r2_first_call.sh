#!/bin/bash
# First gpurun call of the next round: first run of the opt-in paths written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# Every step is wrapped in `timeout` (a wrong mbarrier count in the split epilogue would hang the kernel).
t0=$(date +%s); mark() { echo "[r2_first_call] $1 @ $(( $(date +%s) - t0 )) s"; }
for i in 1 4 5; do RSR_EPI_SPLIT=1 RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py $i > gpurun_out/ts_split_$i.log 2>&1; echo "ts_probe $i rc=$?"; tail -4 gpurun_out/ts_split_$i.log; done; mark ts
# one pytest process per test function: a hang (wrong barrier count) or a sticky CUDA error in one variant must not hide the others
for k in split_epilogue_is_bit_identical serpentine merged_shape_launches_are merged_shape_launches_tta split_epilogue_tta kblock32 tile_sharded cpp_cli; do
  RSR_TEST_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q -k "$k" 2>&1 | tail -4; echo "[$k] rc=$?"
done; mark experimental-tests
timeout 120 python tools/ab_option.py epilogue_split 0 1 2>&1 | tail -3; mark ab
timeout 120 python tools/ab_option.py epilogue_split 0 1 200 2>&1 | tail -3; mark ab-t200
timeout 120 python tools/ab_option.py serpentine 0 1 2>&1 | tail -3; mark ab-serpentine
timeout 120 python tools/ab_option.py merge_shapes 0 1 2>&1 | tail -3; mark ab-merge
timeout 120 python tools/ab_option.py merge_shapes 0 1 200 2>&1 | tail -3; mark ab-merge-t200
timeout 120 python tools/ab_option.py conv_impl 2 3 2>&1 | tail -3; mark ab-kblock32
for i in 2 4 5; do RSR_IMPL=3 RSR_DBG_RDB=1 timeout 60 python tools/ts_probe.py $i > gpurun_out/ts_k32_$i.log 2>&1; tail -3 gpurun_out/ts_k32_$i.log; done; mark ts-k32
RSR_OPTS=serpentine=1,epilogue_split=1 timeout 200 python tools/batch_sweep.py 2>&1 | tail -20; mark batch-sweep-serpentine
timeout 120 tools/ubench > gpurun_out/r2_ubench.txt 2>&1; head -8 gpurun_out/r2_ubench.txt; mark ubench
