#!/bin/bash
# Builds librealsr_b200.so in-tree for sm_100a (the .so is git-ignored but travels to the GPU box with gpurun).
set -e
cd "$(dirname "$0")"
OUT=../librealsr_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-Wall -ccbin /usr/bin/g++"
mkdir -p build
pids=()
for f in conv3x3.cu misc_kernels.cu engine.cu; do
  $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f -o build/${f%.cu}.o &
  pids+=($!)
done
/usr/bin/g++ -O2 -std=c++17 -fPIC -fvisibility=hidden -Wall -c model.cpp -o build/model.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
# link to a temporary name and rename: a gpurun snapshot taken during a rebuild never sees a half-written library
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT.tmp build/conv3x3.o build/misc_kernels.o build/engine.o build/model.o -ccbin /usr/bin/g++
mv -f $OUT.tmp $OUT
echo "built $(realpath $OUT)"
# host-side C++ mirror of the reference interface (class RealSR over the C ABI) + a minimal caller written like the
# reference's proc stage
/usr/bin/g++ -O2 -std=c++11 -Wall -o ../realsr_raw_cli ../host/realsr_raw_cli.cpp -L.. -lrealsr_b200 -Wl,-rpath,'$ORIGIN' -lpthread
echo "built $(realpath ../realsr_raw_cli)"
# the reference-shaped C++ command line (PNG through zlib)
/usr/bin/g++ -O2 -std=c++11 -Wall -o ../realsr-b200 ../host/realsr_cli.cpp -L.. -lrealsr_b200 -Wl,-rpath,'$ORIGIN' -lz -lpthread
echo "built $(realpath ../realsr-b200)"
