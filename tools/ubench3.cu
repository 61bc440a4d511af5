// tools/ubench3.cu -- semantics and cost of tcgen05.shift (UTCSHIFT.DOWN) on B200: can the MMA pipe align the three horizontal tap
// groups of the Cout=32 conv's accumulators (shift the kx=0 group down by 2 rows, the kx=1 group by 1) so that the epilogue adds
// columns of the SAME lane instead of fetching its neighbours' values with 64 warp shuffles per row?  Not part of the product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench3 tools/ubench3.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../realsr-ncnn-vulkan_b200/csrc/ptx.cuh"
using namespace rsr;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tshift1(uint32_t taddr) { asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory"); }

// 128 threads = 4 warps = the 4 lane quarters.  out[lane][col] for 32 columns after `nshift` shifts issued at column `col0`, lane `lane0`
__global__ void __launch_bounds__(128, 1) shift_semantics(int nshift, int col0, int lane0, uint32_t* out, long long* cyc, int iters)
{
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_ptr, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int half = 0; half < 2; half++) {
        for (int j = 0; j < 16; j++) v[j] = (uint32_t)((warp * 32 + lane) * 256 + half * 16 + j);
        tmem_st16(taddr + half * 16, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t phase = 0;
    if (warp == 0) {
        const bool leader = elect_one();
        for (int s = 0; s < nshift; s++)
            if (leader) tshift1(tm + ((uint32_t)lane0 << 16) + col0);
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
        // throughput: iters x 4 shifts (a 32-column group), one commit at the end
        long long t0 = clock64();
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (leader) tshift1(tm + 32 + c * 8);
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        if (leader) cyc[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t r[32];
    tmem_ld32(taddr, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; j++) out[(warp * 32 + lane) * 32 + j] = r[j];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

int main()
{
    uint32_t* d; long long* dc;
    CK(cudaMalloc(&d, 128 * 32 * 4)); CK(cudaMalloc(&dc, 8));
    static uint32_t h[128 * 32];
    struct { int n, col0, lane0; } tests[] = {{1, 0, 0}, {2, 0, 0}, {1, 4, 0}, {1, 0, 32}, {1, 8, 64}};
    for (auto& t : tests) {
        shift_semantics<<<1, 128>>>(t.n, t.col0, t.lane0, d, dc, 1000);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
        long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        printf("== %d shift(s) at col %d lane %d (entry = source_lane.col; unshifted entry would be lane.col); 4000 shifts took %lld clk = %.1f clk/shift\n",
               t.n, t.col0, t.lane0, c, c / 4000.0);
        const int lanes[] = {0, 1, 2, 3, 30, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 126, 127};
        for (int l : lanes) {
            printf(" lane %3d:", l);
            for (int j : {0, 3, 4, 7, 8, 11, 12, 15, 16, 24})
                printf("  c%-2d=%3u.%-2u", j, h[l * 32 + j] >> 8, h[l * 32 + j] & 255);
            printf("\n");
        }
    }
    return 0;
}
