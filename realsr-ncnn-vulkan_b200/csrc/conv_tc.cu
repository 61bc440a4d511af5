// conv_tc.cu -- 3x3 / stride 1 / pad 1 convolution as an im2col-free implicit GEMM on Blackwell tensor cores.
//
// Replaces ncnn's Convolution layer for the 350 non-first convs of the RRDB graph (arithmetic spec:
// /root/reference/src/ncnn/src/layer/convolution.cpp:133-204, activation fused_activation.h:31-36) together with
// the layers that ncnn runs around it: Concat/Split (channel-slice addressing), Eltwise 0.2*a+b
// (layer/eltwise.cpp:116-131), BinaryOp ADD (layer/binaryop.cpp:27-40) and the reference's postproc
// (src/realsr.cpp:791-833 / realsr_postproc.comp:47-89) -- all fused into the epilogue.
//
// GEMM view: M = output pixels, N = Cout (16/32/64), K = 9 taps x Cin.  fp16 operands, fp32 accumulation in TMEM.
//
// Tiling.  A CTA owns BH x 30 output pixels of one tile of the batch.  For every 64-channel chunk of the input it
// TMA-loads ONE haloed box {64 ch, P=32 px, BH+2 rows} (out-of-bounds = zeros = the conv's zero padding) into
// shared memory as (BH+2)*32 rows of 128 B (SWIZZLE_128B, K-major).  Because the box is stored with a fixed pitch
// of P pixels, the A operand of tap (ky,kx) for 128 consecutive "flattened" output positions q = r*P + c is just
// the 128 consecutive smem rows starting at row q + ky*P + kx: the 9 taps re-use the same bytes through 9
// shifted UMMA descriptors -- no im2col, no per-tap reload.  Positions with c >= 30 compute garbage that is never
// stored (6.25 % of the MMA work).  Weights are pre-packed on the host as the exact swizzled smem image of every
// (chunk, tap) slab [N][64] and streamed with plain cp.async.bulk through a small ring.
//
// Warp roles (224 threads): warp 0 = TMA producer for A, warp 1 = tcgen05.mma issuer (+ TMEM alloc),
// warp 2 = bulk-copy producer for B, warps 3..6 = epilogue (TMEM -> registers -> global).
#include "kernels.h"
#include "ptx.cuh"

namespace rsr {

namespace {

constexpr int P = CONV_P;        // smem pitch (pixels) of the haloed box
constexpr int BW = P - 2;        // valid output columns per CTA
constexpr int BH = CONV_BH;      // output rows per CTA
constexpr int MT = BH * P / 128; // 128-row MMA tiles per CTA
constexpr int NA = 2;            // A stages
constexpr int A_BYTES = (BH + 2) * P * 128;
constexpr int A_STAGE = A_BYTES + 1024;  // taps of the last (garbage) positions read up to 2 rows past the box
constexpr int NTHREADS = 224;

template <int N> struct Cfg {
    static constexpr int B_SLAB = N * 128;
    static constexpr int NB = (N == 64) ? 3 : 6;
    static constexpr int TMEM_COLS = (MT * N < 32) ? 32 : MT * N;
    static constexpr int SMEM = 1024 /*align*/ + NA * A_STAGE + NB * B_SLAB + 256 /*barriers*/;
};

struct Bars {
    uint64_t a_full[NA], a_empty[NA];
    uint64_t b_full[6], b_empty[6];
    uint64_t acc_full;
    uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int N>
__global__ void __launch_bounds__(NTHREADS, 2)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a)
{
    using C = Cfg<N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smA = smem;
    uint8_t* smB = smem + NA * A_STAGE;
    Bars* bars = reinterpret_cast<Bars*>(smB + C::NB * C::B_SLAB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH, b = blockIdx.z;
    const int nchunk = (a.cin + 63) >> 6;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA; i++) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < C::NB; i++) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
        mbar_init(&bars->acc_full, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmap);
    }
    if (warp == 1) tmem_alloc(&bars->tmem_ptr, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_ptr;

    if (warp == 0) {
        const bool leader = elect_one();
        for (int c = 0; c < nchunk; c++) {
            const int s = c % NA;
            mbar_wait(&bars->a_empty[s], ((c / NA) & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&bars->a_full[s], A_BYTES);
                tma_load_4d(smA + s * A_STAGE, &tmap, &bars->a_full[s], c * 64, x0 - 1, y0 - 1, b);
            }
        }
    } else if (warp == 2) {
        const bool leader = elect_one();
        const uint8_t* w = reinterpret_cast<const uint8_t*>(a.wpk);
        const int nit = nchunk * 9;
        for (int it = 0; it < nit; it++) {
            const int s = it % C::NB;
            mbar_wait(&bars->b_empty[s], ((it / C::NB) & 1) ^ 1);
            if (leader) {
                mbar_expect_tx(&bars->b_full[s], C::B_SLAB);
                bulk_load(smB + s * C::B_SLAB, w + (size_t)it * C::B_SLAB, C::B_SLAB, &bars->b_full[s]);
            }
        }
    } else if (warp == 1) {
        // The whole warp runs the (warp-uniform) loop and one elected lane issues: descriptors then live in uniform
        // registers and every tcgen05.mma is a single predicated UTCHMMA.  (Issuing from inside `if (lane == 0)`
        // made the compiler wrap each MMA in a divergence loop: ~120 clk per MMA, measured.)
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t desc_hi = umma_desc_sw128(0, 1024, 0);
        int it = 0;
        for (int c = 0; c < nchunk; c++) {
            const int sA = c % NA;
            mbar_wait(&bars->a_full[sA], (c / NA) & 1);
            tc_fence_after();
            const int rem = a.cin - c * 64;
            const int ksteps = rem >= 64 ? 4 : (rem >> 4);
            const uint32_t a_base = smem_u32(smA + sA * A_STAGE);
            for (int tap = 0; tap < 9; tap++, it++) {
                const int sB = it % C::NB;
                mbar_wait(&bars->b_full[sB], (it / C::NB) & 1);
                tc_fence_after();
                const int ky = tap / 3, kx = tap - ky * 3;
                const uint64_t db0 = desc_hi | (uint64_t)((smem_u32(smB + sB * C::B_SLAB) & 0x3FFFFu) >> 4);
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    const uint64_t da0 = desc_hi | (uint64_t)(((a_base + (uint32_t)(ky * P + kx + mt * 128) * 128u) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        if (ks < ksteps && leader)
                            umma_f16(tmem_base + mt * N, da0 + ks * 2, db0 + ks * 2, idesc, (c | tap | ks) ? 1u : 0u);
                    }
                }
                if (leader) umma_commit(&bars->b_empty[sB]);
            }
            if (leader) umma_commit(&bars->a_empty[sA]);
        }
        if (leader) umma_commit(&bars->acc_full);
    } else {
        // ---------------- epilogue: warps 3..6, TMEM lane quarter = warp % 4 ----------------
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
        const int q4 = warp & 3;
#pragma unroll 1
        for (int mt = 0; mt < MT; mt++) {
            const int q = mt * 128 + q4 * 32 + lane;
            const int r = q / P, cp = q - r * P;
            const int y = y0 + r, x = x0 + cp;
            const bool valid = (cp < BW) && (x < a.W) && (y < a.H);
            const size_t pix = ((size_t)b * a.H + y) * a.W + x;
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + mt * N;
            if constexpr (N == 16) {
                uint32_t v[16];
                tmem_ld16(taddr, v);
                tmem_ld_wait();
                if (valid) {
                    const TileDesc td = a.tiles[b];
                    float o[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) o[j] = __uint_as_float(v[j]) + __ldg(a.bias + j);
                    if (a.mode == CONV_LAST_U8) {
                        const int i = y - a.prepad_s, jx = x - a.prepad_s;
                        if (i >= 0 && i < td.crop_h && jx >= 0 && jx < td.crop_w) {
                            uint8_t* dst = a.img_out + ((size_t)(td.out_y0 + i) * a.img_w + td.out_x0 + jx) * 3;
#pragma unroll
                            for (int j = 0; j < 3; j++) {
                                int iv = (int)(o[j] * 255.f + 0.5f);
                                dst[j] = (uint8_t)min(max(iv, 0), 255);
                            }
                        }
                    } else {
                        float4* dst = reinterpret_cast<float4*>(a.tta_out) + (size_t)td.tta_slot * ((size_t)a.H * a.W) + (size_t)y * a.W + x;
                        *dst = make_float4(o[0], o[1], o[2], 0.f);
                    }
                }
            } else {
#pragma unroll 1
                for (int nb = 0; nb < N / 32; nb++) {
                    uint32_t v[32];
                    tmem_ld32(taddr + nb * 32, v);
                    tmem_ld_wait();
                    if (valid) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) f[j] = __uint_as_float(v[j]) + __ldg(a.bias + nb * 32 + j);
                        if (a.mode == CONV_STORE16) {
                            if (a.act == 2) {
#pragma unroll
                                for (int j = 0; j < 32; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * a.slope;
                            }
                        } else if (a.mode == CONV_RDB_OUT) {
                            float4* tp = reinterpret_cast<float4*>(a.t32 + pix * 64 + nb * 32);
                            float4* rp = reinterpret_cast<float4*>(a.r32 + pix * 64 + nb * 32);
#pragma unroll
                            for (int j4 = 0; j4 < 8; j4++) {
                                const float4 t = tp[j4];
                                float4 o;
                                o.x = f[j4 * 4 + 0] * 0.2f + t.x;
                                o.y = f[j4 * 4 + 1] * 0.2f + t.y;
                                o.z = f[j4 * 4 + 2] * 0.2f + t.z;
                                o.w = f[j4 * 4 + 3] * 0.2f + t.w;
                                if (a.rrdb_end) {
                                    const float4 rr = rp[j4];
                                    o.x = o.x * 0.2f + rr.x;
                                    o.y = o.y * 0.2f + rr.y;
                                    o.z = o.z * 0.2f + rr.z;
                                    o.w = o.w * 0.2f + rr.w;
                                    rp[j4] = o;
                                }
                                tp[j4] = o;
                                f[j4 * 4 + 0] = o.x; f[j4 * 4 + 1] = o.y; f[j4 * 4 + 2] = o.z; f[j4 * 4 + 3] = o.w;
                            }
                        } else {  // CONV_TRUNK
                            const float4* fp = reinterpret_cast<const float4*>(a.f32 + pix * 64 + nb * 32);
#pragma unroll
                            for (int j4 = 0; j4 < 8; j4++) {
                                const float4 t = __ldg(fp + j4);
                                f[j4 * 4 + 0] += t.x; f[j4 * 4 + 1] += t.y; f[j4 * 4 + 2] += t.z; f[j4 * 4 + 3] += t.w;
                            }
                        }
                        uint4* dst = reinterpret_cast<uint4*>(a.out16 + pix * a.out_cs + a.out_co + nb * 32);
#pragma unroll
                        for (int j8 = 0; j8 < 4; j8++) {
                            uint4 o;
                            o.x = pack_h2(f[j8 * 8 + 0], f[j8 * 8 + 1]);
                            o.y = pack_h2(f[j8 * 8 + 2], f[j8 * 8 + 3]);
                            o.z = pack_h2(f[j8 * 8 + 4], f[j8 * 8 + 5]);
                            o.w = pack_h2(f[j8 * 8 + 6], f[j8 * 8 + 7]);
                            dst[j8] = o;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int N>
cudaError_t launch_n(const CUtensorMap* in_map, const ConvArgs& a, cudaStream_t stream)
{
    using C = Cfg<N>;
    static bool attr_done[64] = {};  // per device: function attributes live in the device's context
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (!attr_done[dev & 63]) {
        e = cudaFuncSetAttribute(conv3x3_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    dim3 grid((a.W + BW - 1) / BW, (a.H + BH - 1) / BH, a.B);
    conv3x3_tc_kernel<N><<<grid, NTHREADS, C::SMEM, stream>>>(*in_map, a);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv3x3_tc(const CUtensorMap* in_map, const ConvArgs& a, int cout_pad, cudaStream_t stream)
{
    switch (cout_pad) {
    case 16: return launch_n<16>(in_map, a, stream);
    case 32: return launch_n<32>(in_map, a, stream);
    case 64: return launch_n<64>(in_map, a, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace rsr
