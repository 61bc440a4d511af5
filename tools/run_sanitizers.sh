#!/bin/bash
# compute-sanitizer over the hot path: memcheck and racecheck of a whole RealSR::process call on BASELINE configs[0]
# (256x256, tile 200: four padded shapes -> merged launches, edge tiles, every epilogue mode), plus a TTA call.
#   gpurun -- 'bash tools/run_sanitizers.sh'      -> gpurun_out/r2_sanitizer_*.txt (copy to profiles/)
# WARNING (round 2, gpurun calls 59-61): this script took the GPU box down three times in a row, ~60 s into the first memcheck run
# (box lost, no output came back -- most likely the host ran out of memory instrumenting the 227-KB-smem cluster kernel).  It therefore
# refuses to run unless RSR_SANITIZERS_ANYWAY=1; on a shared box start with ONE tool on the smallest case and watch `free -g`.
if [ "${RSR_SANITIZERS_ANYWAY:-0}" != "1" ]; then echo "run_sanitizers.sh: disabled (see the warning in this script)"; exit 0; fi
CS=/usr/local/cuda/bin/compute-sanitizer
t0=$(date +%s); mark() { echo "[sanitizers] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 600 $CS --tool memcheck --print-limit 20 python tools/run_case.py 256 256 200 1 > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck.txt; mark memcheck
timeout 600 $CS --tool memcheck --print-limit 20 python tools/run_case.py 61 45 32 1 1 > gpurun_out/r2_sanitizer_memcheck_tta.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck_tta.txt; mark memcheck-tta
timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python tools/run_case.py 96 80 64 1 > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.txt; mark racecheck
timeout 600 $CS --tool synccheck --print-limit 20 python tools/run_case.py 96 80 64 1 > gpurun_out/r2_sanitizer_synccheck.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_synccheck.txt; mark synccheck
