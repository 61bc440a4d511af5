#!/bin/bash
# compute-sanitizer over the hot path: memcheck and racecheck of a whole RealSR::process call on BASELINE configs[0]
# (256x256, tile 200: four padded shapes -> merged launches, edge tiles, every epilogue mode), plus a TTA call.
#   gpurun -- 'bash tools/run_sanitizers.sh'      -> gpurun_out/r2_sanitizer_*.txt (copy to profiles/)
CS=/usr/local/cuda/bin/compute-sanitizer
t0=$(date +%s); mark() { echo "[sanitizers] $1 @ $(( $(date +%s) - t0 )) s"; }
timeout 600 $CS --tool memcheck --print-limit 20 python tools/run_case.py 256 256 200 1 > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/r2_sanitizer_memcheck.txt; mark memcheck
timeout 600 $CS --tool memcheck --print-limit 20 python tools/run_case.py 61 45 32 1 1 > gpurun_out/r2_sanitizer_memcheck_tta.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck_tta.txt; mark memcheck-tta
timeout 900 $CS --tool racecheck --racecheck-report analysis --print-limit 20 python tools/run_case.py 96 80 64 1 > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.txt; mark racecheck
timeout 600 $CS --tool synccheck --print-limit 20 python tools/run_case.py 96 80 64 1 > gpurun_out/r2_sanitizer_synccheck.txt 2>&1; tail -3 gpurun_out/r2_sanitizer_synccheck.txt; mark synccheck
