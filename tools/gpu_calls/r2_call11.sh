#!/bin/bash
t0=$(date +%s); mark() { echo "[call11] $1 @ $(( $(date +%s) - t0 )) s"; }
L=realsr-ncnn-vulkan_b200/librealsr_b200.so
timeout 900 python tools/ab_builds.py --rounds 2 "head=$L" "r32ext=ablibs/gpurun_libs_eb12348.so" "r32noext=ablibs/gpurun_libs_eb12348.so:residual_ext=0" "n192late=ablibs/gpurun_libs_6c8f3d2.so" "v2n128=ablibs/gpurun_libs_ab49cf3.so:conv_impl=5,merge_shapes=1" "v2n128k64=ablibs/gpurun_libs_ab49cf3.so:conv_impl=4,merge_shapes=1" 2>&1 | tail -8; mark ab-builds
