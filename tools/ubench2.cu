// tools/ubench2.cu -- B200 micro-benchmarks for the conv kernel's EPILOGUE side (not part of the product):
//   1. tcgen05.ld (LDTM) throughput per SM with 4 / 8 / 16 reading warps, alone and under a running cta_group::2 MMA stream
//      (does draining accumulators slow the MMAs down, and how many bytes per clock does TMEM give the register file?)
//   2. warp shuffle and packed-fp32 (add.f32x2) issue rates with 8 / 16 warps (the horizontal tap combination)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../realsr-ncnn-vulkan_b200/csrc/ptx.cuh"

using namespace rsr;

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e = (x);                                                         \
        if (e != cudaSuccess) {                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                 \
        }                                                                            \
    } while (0)

// warps: 0 = MMA issuer (leader CTA), 1..NE = readers.  do_mma / do_ld select what runs; ld_mode: 0 = .x32 loads, 1 = .x16 loads,
// 2 = .x32 loads + 2 shuffles + 2 adds per value (the Cout=32 tap combination), 3 = shuffles only (no TMEM)
template <int NE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((1 + NE) * 32, 1)
epi_bench(int iters, int do_mma, int do_ld, int ld_mode, long long* out_mma, long long* out_ld, float* sink)
{
    constexpr int N = 96;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smem + 65536;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster_sync();
    if (warp == 0) tmem_alloc2(&tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync();
    const uint32_t tm = tmem_ptr;
    if (warp == 0) {
        if (rank == 0 && do_mma) {
            const uint32_t idesc = umma_idesc_f16(256, N);
            const uint64_t da0 = umma_desc_sw128(smem_u32(smA), 1024, 0), db0 = umma_desc_sw128(smem_u32(smB), 1024, 0);
            const bool leader = elect_one();
            long long t0 = clock64();
            uint32_t phase = 0;
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int tap = 0; tap < 9; tap++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++)
                        if (leader) umma2_f16(tm + ((tap + ks) & 1) * 128, da0 + ks * 2 + (tap % 3) * 8, db0 + ks * 2, idesc, 1);
                if ((it & 7) == 7) {
                    if (leader) umma2_commit(&bar, 1);
                    mbar_wait(&bar, phase);
                    phase ^= 1;
                }
            }
            if (leader) umma2_commit(&bar, 1);
            mbar_wait(&bar, phase);
            long long t1 = clock64();
            if (leader) out_mma[blockIdx.x / 2] = t1 - t0;
        }
    } else if (do_ld) {
        const int q4 = warp & 3;
        const uint32_t taddr = tm + ((uint32_t)(q4 * 32) << 16) + 256;  // the half of TMEM the MMAs do not write
        float acc = 0.f;
        // the MMA stream (36 MMAs x ~48 clk per iteration) is ~1.7 k clk per iteration; readers run a comparable amount of loads
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int rep = 0; rep < 4; rep++) {
                if (ld_mode == 1) {
                    uint32_t v[16], u[16], w[16];
                    tmem_ld16(taddr + 0, v);
                    tmem_ld16(taddr + 32, u);
                    tmem_ld16(taddr + 64, w);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j++) acc += __uint_as_float(v[j] ^ u[j] ^ w[j]);
                } else if (ld_mode == 3) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = acc + j;
#pragma unroll
                    for (int j = 0; j < 32; j++) f[j] = f[j] + __shfl_down_sync(0xffffffffu, f[(j + 1) & 31], 1) + __shfl_down_sync(0xffffffffu, f[(j + 2) & 31], 2);
#pragma unroll
                    for (int j = 0; j < 32; j++) acc += f[j];
                } else {
                    uint32_t v[32], u[32], w[32];
                    tmem_ld32(taddr + 0, v);
                    tmem_ld32(taddr + 32, u);
                    tmem_ld32(taddr + 64, w);
                    tmem_ld_wait();
                    if (ld_mode == 2) {
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            acc += __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                   __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) acc += __uint_as_float(v[j] ^ u[j] ^ w[j]);
                    }
                }
            }
        }
        long long t1 = clock64();
        if (lane == 0) out_ld[blockIdx.x * NE + warp - 1] = t1 - t0;
        if (acc == 123.456f) sink[0] = acc;
    }
    tc_fence_before();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tm, 512);
}

template <int NE>
void run(int do_mma, int do_ld, int ld_mode, const char* what)
{
    long long *d_mma, *d_ld;
    float* sink;
    CK(cudaMalloc(&d_mma, 148 * 8));
    CK(cudaMalloc(&d_ld, 148 * NE * 8));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(d_mma, 0, 148 * 8));
    CK(cudaMemset(d_ld, 0, 148 * NE * 8));
    const int smem = 1024 + 65536 + 32768;
    CK(cudaFuncSetAttribute(epi_bench<NE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 1000;
    epi_bench<NE><<<148, (1 + NE) * 32, smem>>>(20, do_mma, do_ld, ld_mode, d_mma, d_ld, sink);
    CK(cudaDeviceSynchronize());
    epi_bench<NE><<<148, (1 + NE) * 32, smem>>>(iters, do_mma, do_ld, ld_mode, d_mma, d_ld, sink);
    CK(cudaDeviceSynchronize());
    long long hm[74], hl[148 * NE];
    CK(cudaMemcpy(hm, d_mma, sizeof hm, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hl, d_ld, sizeof(long long) * 148 * NE, cudaMemcpyDeviceToHost));
    double am = 0, al = 0;
    for (int i = 0; i < 74; i++) am += (double)hm[i];
    am /= 74;
    for (int i = 0; i < 148 * NE; i++) al += (double)hl[i];
    al /= 148 * NE;
    // per SM and iteration: NE warps x 4 reps x 3 loads (x32: 4 KB each, x16: 2 KB each)
    const double bytes_per_it = (double)NE * 4 * 3 * (ld_mode == 1 ? 2048 : 4096);
    printf("%-44s NE=%2d:", what, NE);
    if (do_mma) printf("  MMA N=96: %.1f clk/MMA", am / (iters * 36.0));
    if (do_ld && ld_mode != 3) printf("  LDTM: %.1f clk per warp-iter-set, %.1f B/clk/SM", al / (iters * 4.0), bytes_per_it * iters / al);
    if (do_ld && ld_mode == 3) printf("  SHFL: %.2f warp-shuffles/clk/SM", (double)NE * 4 * 64 * iters / al);
    printf("\n");
    cudaFree(d_mma); cudaFree(d_ld); cudaFree(sink);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<8>(1, 0, 0, "MMA alone");
    run<4>(0, 1, 0, "LDTM.x32 alone");
    run<8>(0, 1, 0, "LDTM.x32 alone");
    run<16>(0, 1, 0, "LDTM.x32 alone");
    run<8>(0, 1, 1, "LDTM.x16 alone");
    run<16>(0, 1, 1, "LDTM.x16 alone");
    run<8>(1, 1, 0, "MMA + LDTM.x32");
    run<16>(1, 1, 0, "MMA + LDTM.x32");
    run<8>(0, 1, 2, "LDTM.x32 + 2 shfl + 2 add per value");
    run<16>(0, 1, 2, "LDTM.x32 + 2 shfl + 2 add per value");
    run<8>(1, 1, 2, "MMA + LDTM.x32 + shfl/add");
    run<16>(1, 1, 2, "MMA + LDTM.x32 + shfl/add");
    run<8>(0, 1, 3, "shuffles only");
    run<16>(0, 1, 3, "shuffles only");
    return 0;
}
