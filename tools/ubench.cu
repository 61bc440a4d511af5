// tools/ubench.cu -- B200 micro-benchmarks that size the conv kernel's design (not part of the product).
//   1. tcgen05.mma kind::f16 throughput vs N with both operands in shared memory (SS), M=128, K=16 per instruction:
//      is a small-N MMA limited by the tensor pipe (N/2 clk) or by the smem operand feed?
//   2. same with A in TMEM (TS).
//   3. L2 -> smem and HBM -> smem streaming bandwidth with cp.async.bulk, all SMs, deep pipeline.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu
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

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

#define UMMA_VARIANT(name, opc)                                                                                        \
    __device__ __forceinline__ void name(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) \
    {                                                                                                                  \
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" opc " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),   \
                     "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)                                                    \
                     : "memory");                                                                                      \
    }
UMMA_VARIANT(umma_a_fill, "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill")
UMMA_VARIANT(umma_a_use, "tcgen05.mma.cta_group::1.kind::f16.collector::a::use")
UMMA_VARIANT(umma_a_last, "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse")
UMMA_VARIANT(umma_ws_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill")
UMMA_VARIANT(umma_ws_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use")
UMMA_VARIANT(umma_ws_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse")
UMMA_VARIANT(umma_ws_discard, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::discard")
UMMA_VARIANT(umma_ws1_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::fill")
UMMA_VARIANT(umma_ws1_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::use")
UMMA_VARIANT(umma_ws1_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::lastuse")
UMMA_VARIANT(umma_ws2_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b2::fill")
UMMA_VARIANT(umma_ws2_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b2::use")
UMMA_VARIANT(umma_ws2_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b2::lastuse")
UMMA_VARIANT(umma_ws3_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b3::fill")
UMMA_VARIANT(umma_ws3_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b3::use")
UMMA_VARIANT(umma_ws3_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b3::lastuse")
// mode 6: like mode 4 but successive groups rotate through collector buffers b0..b3
// mode 7: like mode 4 with groups of 8 M-tiles sharing B (fill, use x6, lastuse)

// mode 3: groups of 4 MMAs share A (A_KEEP then A_REUSE), B differs; mode 4: .ws, groups of 4 share B (fill, use x3), A differs;
// mode 5: .ws, every MMA discards (no reuse)
// mode 0: SS aligned A; mode 1: SS with row-shifted A starts (like the conv taps); mode 2: TS (A in TMEM)
template <int N>
__global__ void __launch_bounds__(128, 1) mma_bench(int iters, int mode, int nacc, long long* out_cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;                       // 64 KB: 512 rows x 128 B
    uint8_t* smB = smem + 65536;               // 256 rows x 128 B = 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_ptr, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_ptr;
    if (threadIdx.x < 32) {  // warp-uniform loop; one elected lane issues
        constexpr uint32_t idesc = umma_idesc_f16(128, N);
        const uint32_t a0 = smem_u32(smA), b0 = smem_u32(smB);
        const uint64_t da0 = umma_desc_sw128(a0, 1024, 0), db0 = umma_desc_sw128(b0, 1024, 0);
        const bool leader = elect_one();
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const uint64_t da = da0 + (uint64_t)(mode == 1 ? (((tap / 3) * 32 + tap % 3) * 128) >> 4 : 0);
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint32_t d = tm + (nacc == 2 ? ((tap + ks) & 1) * N : 0);
                    if (leader) {
                        if (mode == 2) umma_f16_ts(d, tm + 256 + ks * 8, db0 + ks * 2, idesc, 1);
                        else if (mode == 3) {
                            // same A (tap), B advanced per ks (pretend different weights): fill, use, use, lastuse
                            if (ks == 0) umma_a_fill(d, da, db0 + ks * 2, idesc, 1);
                            else if (ks == 3) umma_a_last(d, da, db0 + ks * 2, idesc, 1);
                            else umma_a_use(d, da, db0 + ks * 2, idesc, 1);
                        } else if (mode == 4) {
                            // same B (tap), A differs per ks (pretend 4 M-tiles, 16 KB apart)
                            const uint64_t dam = da0 + (uint64_t)(ks * 16384 >> 4);
                            const uint32_t dd = tm + ks * N % 512;
                            if (ks == 0) umma_ws_fill(dd, dam, db0 + (tap & 3) * 2, idesc, 1);
                            else if (ks == 3) umma_ws_last(dd, dam, db0 + (tap & 3) * 2, idesc, 1);
                            else umma_ws_use(dd, dam, db0 + (tap & 3) * 2, idesc, 1);
                        } else if (mode == 6) {
                            const uint64_t dam = da0 + (uint64_t)(ks * 16384 >> 4);
                            const uint32_t dd = tm + ks * N % 512;
                            const uint64_t dbb = db0 + (tap & 3) * 2;
                            switch (tap & 3) {
                            case 0: if (ks == 0) umma_ws_fill(dd, dam, dbb, idesc, 1); else if (ks == 3) umma_ws_last(dd, dam, dbb, idesc, 1); else umma_ws_use(dd, dam, dbb, idesc, 1); break;
                            case 1: if (ks == 0) umma_ws1_fill(dd, dam, dbb, idesc, 1); else if (ks == 3) umma_ws1_last(dd, dam, dbb, idesc, 1); else umma_ws1_use(dd, dam, dbb, idesc, 1); break;
                            case 2: if (ks == 0) umma_ws2_fill(dd, dam, dbb, idesc, 1); else if (ks == 3) umma_ws2_last(dd, dam, dbb, idesc, 1); else umma_ws2_use(dd, dam, dbb, idesc, 1); break;
                            default: if (ks == 0) umma_ws3_fill(dd, dam, dbb, idesc, 1); else if (ks == 3) umma_ws3_last(dd, dam, dbb, idesc, 1); else umma_ws3_use(dd, dam, dbb, idesc, 1); break;
                            }
                        } else if (mode == 7) {
                            // 8 MMAs per B: (tap even, ks 0..3) then (tap odd, ks 0..3) share the B of the even tap
                            const int g = (tap & 1) * 4 + ks;
                            const uint64_t dam = da0 + (uint64_t)((g & 3) * 16384 >> 4) + (g >> 2) * 8;
                            const uint32_t dd = tm + (g * N) % 512;
                            const uint64_t dbb = db0 + ((tap >> 1) & 3) * 2;
                            if (tap == 8) umma_ws_discard(dd, dam, dbb, idesc, 1);
                            else if (g == 0) umma_ws_fill(dd, dam, dbb, idesc, 1);
                            else if (g == 7) umma_ws_last(dd, dam, dbb, idesc, 1);
                            else umma_ws_use(dd, dam, dbb, idesc, 1);
                        } else if (mode == 5) umma_ws_discard(d, da + ks * 2, db0 + ks * 2, idesc, 1);
                        else umma_f16(d, da + ks * 2, db0 + ks * 2, idesc, 1);
                    }
                }
            }
            if ((it & 7) == 7) {  // bound the number of MMAs in flight
                if (leader) umma_commit(&bar);
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
        if (leader) umma_commit(&bar);
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        if (leader) out_cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

// ---- cta_group::2: one instruction drives the tensor cores of both SMs of a CTA pair (M = 256) -----------------------------
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_bench(int iters, long long* out_cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smem + 65536;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u ^ (i * 2654435761u & 0x03ff03ffu);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t tm = tmem_ptr;
    if (rank == 0 && threadIdx.x < 32) {
        const uint32_t idesc = umma_idesc_f16(256, N);
        const uint64_t da0 = umma_desc_sw128(smem_u32(smA), 1024, 0), db0 = umma_desc_sw128(smem_u32(smB), 1024, 0);
        const bool leader = elect_one();
        long long t0 = clock64();
        uint32_t phase = 0;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    if (leader) umma2_f16(tm + ((tap + ks) & 1) * N, da0 + ks * 2 + (tap % 3) * 8, db0 + ks * 2, idesc, 1);
                }
            }
            if ((it & 7) == 7) {
                if (leader)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                                 "h"((uint16_t)1)
                                 : "memory");
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
        if (leader)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                         "h"((uint16_t)1)
                         : "memory");
        mbar_wait(&bar, phase);
        long long t1 = clock64();
        if (leader) out_cycles[blockIdx.x / 2] = t1 - t0;
    }
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int N>
void run_mma2()
{
    long long* d;
    CK(cudaMalloc(&d, 148 * 8));
    const int smem = 1024 + 65536 + 32768;
    CK(cudaFuncSetAttribute(mma2_bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000;
    mma2_bench<N><<<148, 128, smem>>>(50, d);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma2_bench<N><<<148, 128, smem>>>(iters, d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[74];
    CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 74; i++) avg += (double)h[i];
    avg /= 74;
    const double nmma = (double)iters * 36;
    const double flops = nmma * 2.0 * 256 * N * 16 * 74;
    printf("cta_group::2 M=256 N=%3d: %.2f clk/MMA (per-SM floor %d)  -> %.1f TFLOP/s chip (%.3f ms)\n", N, avg / nmma, N / 2, flops / (ms * 1e-3) / 1e12, ms);
    cudaFree(d);
}

template <int N>
void run_mma(int mode, int nacc)
{
    long long* d;
    CK(cudaMalloc(&d, 148 * 8));
    const int smem = 1024 + 65536 + 32768;
    CK(cudaFuncSetAttribute(mma_bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 2000;
    mma_bench<N><<<148, 128, smem>>>(50, mode, nacc, d);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mma_bench<N><<<148, 128, smem>>>(iters, mode, nacc, d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double nmma = (double)iters * 36;
    const double flops = nmma * 2.0 * 128 * N * 16 * 148;
    printf("N=%3d mode=%d nacc=%d: %.2f clk/MMA (tensor floor %d)  -> %.1f TFLOP/s chip (%.3f ms)\n", N, mode, nacc, avg / nmma, N / 2,
           flops / (ms * 1e-3) / 1e12, ms);
    cudaFree(d);
}

// ---- streaming bandwidth: every CTA keeps `depth` bulk copies of `chunk` bytes in flight -------------------------------
__global__ void __launch_bounds__(128, 1) stream_bench(const uint8_t* src, size_t span_bytes, int chunk, int depth, int nchunks, long long* out_cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t nslots = span_bytes / chunk;
        size_t slot = ((size_t)blockIdx.x * 7919) % nslots;
        long long t0 = clock64();
        for (int i = 0; i < nchunks + depth; i++) {
            const int s = i % depth;
            if (i >= depth) mbar_wait(&bars[s], ((i / depth) - 1) & 1);
            if (i < nchunks) {
                mbar_expect_tx(&bars[s], chunk);
                bulk_load(smem + (size_t)s * chunk, src + slot * chunk, chunk, &bars[s]);
                slot += gridDim.x;
                if (slot >= nslots) slot -= nslots;
            }
        }
        out_cycles[blockIdx.x] = clock64() - t0;
    }
}

void run_stream(const char* name, size_t span, int chunk, int depth, int ctas_per_sm)
{
    uint8_t* buf;
    CK(cudaMalloc(&buf, span));
    CK(cudaMemset(buf, 1, span));
    long long* d;
    const int grid = 148 * ctas_per_sm;
    CK(cudaMalloc(&d, grid * 8));
    const int smem = 1024 + chunk * depth;
    CK(cudaFuncSetAttribute(stream_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int nchunks = (int)((size_t)(64 << 20) / chunk);  // 64 MB per CTA
    stream_bench<<<grid, 128, smem>>>(buf, span, chunk, depth, 256, d);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    stream_bench<<<grid, 128, smem>>>(buf, span, chunk, depth, nchunks, d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 4];
    CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += (double)h[i];
    avg /= grid;
    const double bytes = (double)nchunks * chunk * grid;
    printf("%-4s span %5zu MB chunk %3d KB depth %2d x%d CTA/SM: %.2f TB/s, %.1f B/clk/SM (%.3f ms)\n", name, span >> 20, chunk >> 10, depth, ctas_per_sm,
           bytes / (ms * 1e-3) / 1e12, (double)nchunks * chunk * ctas_per_sm / avg, ms);
    cudaFree(buf);
    cudaFree(d);
}

// ---- TMA tensor (tiled, SWIZZLE_128B) box streaming: the conv kernel's activation load pattern in isolation -------------
#include <cuda.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void __launch_bounds__(128, 1) box_bench(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int depth, int W, int H, int B,
                                                    int nrounds, long long* out_cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bars[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; i++) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tiles_x = (W + 29) / 30, tiles_y = (H + box_rows - 3) / (box_rows - 2);
        const int per_img = tiles_x * tiles_y, ntiles = per_img * B;
        long long t0 = clock64();
        int issued = 0, done = 0;
        for (int r = 0; r < nrounds; r++)
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                if (issued - done >= depth) { mbar_wait(&bars[done % depth], (done / depth) & 1); done++; }
                const int b = t / per_img, rem = t - b * per_img, ty = rem / tiles_x, tx = rem - ty * tiles_x;
                const int s = issued % depth;
                mbar_expect_tx(&bars[s], box_bytes);
                tma_load_4d(smem + (size_t)s * (box_bytes + 1024), &map, &bars[s], 0, tx * 30 - 1, ty * (box_rows - 2) - 1, b);
                issued++;
            }
        while (done < issued) { mbar_wait(&bars[done % depth], (done / depth) & 1); done++; }
        out_cycles[blockIdx.x] = clock64() - t0;
    }
}
void run_box(int box_rows, int depth)
{
    const int W = 420, H = 420, B = 8;
    void* buf;
    const size_t bytes = (size_t)B * H * W * 128;
    CK(cudaMalloc(&buf, bytes));
    CK(cudaMemset(buf, 1, bytes));
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    CUtensorMap map;
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, 32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (((EncodeTiledFn)fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); exit(1); }
    const int box_bytes = box_rows * 32 * 128;
    const int smem = 1024 + depth * (box_bytes + 1024);
    CK(cudaFuncSetAttribute(box_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* d;
    CK(cudaMalloc(&d, 148 * 8));
    const int nrounds = 6;
    box_bench<<<148, 128, smem>>>(map, box_bytes, box_rows, depth, W, H, B, 1, d);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    box_bench<<<148, 128, smem>>>(map, box_bytes, box_rows, depth, W, H, B, nrounds, d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const int tiles = ((W + 29) / 30) * ((H + box_rows - 3) / (box_rows - 2)) * B;
    const double moved = (double)tiles * box_bytes * nrounds, useful = (double)bytes * nrounds;
    printf("TMA box {64ch,32px,%2d rows} depth %d: %.2f TB/s moved into smem, %.2f TB/s of distinct tensor bytes (%.3f ms)\n", box_rows, depth,
           moved / (ms * 1e-3) / 1e12, useful / (ms * 1e-3) / 1e12, ms);
    cudaFree(buf);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    // CTA-pair MMA issue rates, incl. the shapes the conv kernel actually uses (N=96) and a candidate (N=192)
    run_mma2<64>();
    run_mma2<96>();
    run_mma2<128>();
    run_mma2<192>();
    run_mma2<256>();
    run_box(10, 2);
    run_box(10, 3);
    run_box(10, 4);
    run_box(10, 5);
    run_box(18, 2);
    run_box(18, 3);
    run_box(34, 1);
    run_stream("HBM", (size_t)4096 << 20, 40 << 10, 4, 1);

    return 0;
}
