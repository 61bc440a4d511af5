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
# the reference-shaped C++ command line (own PNG / JPEG / WebP codecs; zlib for inflate and checksums)
/usr/bin/g++ -O2 -std=c++11 -Wall -o ../realsr-b200 ../host/realsr_cli.cpp -L.. -lrealsr_b200 -Wl,-rpath,'$ORIGIN' -lz -lpthread
echo "built $(realpath ../realsr-b200)"
# optional build variants for A/B runs on hardware (tools/ab_builds.py picks a library through RSR_LIB); never the default product
if [ -n "$RSR_BUILD_VARIANTS" ]; then
  $NVCC $FLAGS -DRSR_EARLY_HANDBACK -c conv3x3.cu -o build/conv3x3_early.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../librealsr_b200_early.so.tmp build/conv3x3_early.o build/misc_kernels.o build/engine.o build/model.o -ccbin /usr/bin/g++
  mv -f ../librealsr_b200_early.so.tmp ../librealsr_b200_early.so
  echo "built $(realpath ../librealsr_b200_early.so) (-DRSR_EARLY_HANDBACK: accumulator hand-back before the tap shuffles; not validated on hardware)"
fi
