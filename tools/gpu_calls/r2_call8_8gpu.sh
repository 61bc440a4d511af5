#!/bin/bash
# BASELINE configs[2] on the 8-GPU box: gpurun --gpus 8 -- 'bash tools/gpu_calls/r2_call8_8gpu.sh'
t0=$(date +%s); mark() { echo "[call8] $1 @ $(( $(date +%s) - t0 )) s"; }
nvidia-smi -L | head -8; df -h /dev/shm /tmp | tail -2; nproc
N=${RSR_GPUS:-8}
timeout 900 python tools/bench_dir.py --gpus $N --images ${RSR_IMAGES:-100} --content natural --e2e-images ${RSR_E2E_IMAGES:-48} > gpurun_out/r2_cfg3_${N}gpu.log 2>&1; tail -2 gpurun_out/r2_cfg3_${N}gpu.log | cut -c1-1500; mark bench_dir
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; cut -c1-300 gpurun_out/r2_bench_${N}gpu.json; tail -2 gpurun_out/r2_bench_${N}gpu.err; mark torchrun-bench
