// cuda_host_emu.h -- test infrastructure: runs the engine's CUDA-core kernels (csrc/misc_kernels.cu, UNMODIFIED source) on the host,
// so that their index arithmetic and fp32 math can be checked against the oracle without a GPU (`-m "not gpu"` tests).
// A thread block is emulated by blockDim real threads that execute the kernel body concurrently; __syncthreads() is a pthread
// barrier, __shared__ variables are function-local statics (blocks run one after the other).  Only what those kernels use is
// provided; tensor-core / TMA code (conv3x3.cu) cannot be emulated this way.  Compile with g++ -ffp-contract=off (an `a * b + c`
// in the source must stay two roundings on the host like the explicit __fmul_rn/__fmaf_rn chain on the device).
#pragma once
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <thread>
#include <vector>

// device intrinsics used by kernels.h / misc_kernels.cu, declared before those headers are parsed
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
using std::max;
using std::min;

#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <typename T> static inline T __ldg(const T* p) { return *p; }

namespace cuda_emu {
inline pthread_barrier_t& barrier() { static pthread_barrier_t b; return b; }
}
static thread_local uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::barrier()); }

#undef __shared__
#define __shared__ static
#undef __global__
#define __global__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __forceinline__
#define __forceinline__ inline
#undef __restrict__
#define __restrict__

namespace cuda_emu {
// kernel<<<grid, block>>>(args...): block.x*y*z host threads walk the grid block by block
template <typename F> void launch(dim3 grid, dim3 block, F body)
{
    const unsigned nt = block.x * block.y * block.z;
    gridDim = grid;
    blockDim = block;
    pthread_barrier_init(&barrier(), nullptr, nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=] {
            threadIdx.x = t % block.x; threadIdx.y = (t / block.x) % block.y; threadIdx.z = t / (block.x * block.y);
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                        body();
                        pthread_barrier_wait(&barrier());  // the next block re-uses the "shared memory"
                    }
        });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&barrier());
}
}  // namespace cuda_emu
