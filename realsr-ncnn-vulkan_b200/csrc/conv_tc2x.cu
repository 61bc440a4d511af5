// conv_tc2x.cu -- the CTA-pair conv kernel of conv_tc2.cu plus the variants written after round 1's GPU budget was spent.
//
// conv_tc2.cu is the kernel every round-1 measurement and parity run used; it stays byte-for-byte what was validated and
// remains the default.  This file is the same kernel with four additions, each behind an engine option and each
// compiled out / inactive unless that option is set (engine.cu routes a launch here only then):
//   * SPLIT  (epilogue_split=1): all 8 epilogue warps work on every tile, channel halves per warp, accumulator set handed
//     back right after the TMEM read -- see epilogue_split() below;
//   * KB=32  (conv_impl=3): 32-channel K-blocks, SWIZZLE_64B boxes, up to 8 A stages;
//   * MULTI  (merge_shapes=1): one launch per layer over up to four same-shape tile groups (ConvArgs::grp);
//   * reverse (serpentine=1): every second conv walks its work tiles backwards.
// Status: compiles for sm_100a; first run on a B200 pending (tests/test_zz_experimental_gpu.py, tools/r2_first_call.sh).
// Once validated this file replaces conv_tc2.cu.  Read conv_tc2.cu / conv_tc.cu for the algorithm itself.
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace rsr {

namespace {

constexpr int P = CONV_P;
constexpr int BW = P - 2;
constexpr int BH = 8;                         // output rows per CTA
constexpr int BAND = 2 * BH;                  // output rows per CTA pair
constexpr int MT = BH * P / 128;              // 2 M-tiles per CTA
// K-block KB = input channels per A stage / per resident weight block: 64 (one SWIZZLE_128B box per 64-channel plane) or
// 32 (option conv_impl=3: SWIZZLE_64B half-plane boxes -- the 96- and 160-channel convs stop fetching a half plane of
// zero-weighted channels, and twice as many, half as large stages pipeline the 192-channel conv more finely).
constexpr int NA_MAX = 8;                     // A stages (run-time count: as many as fit next to the resident weights; KB=64: <= 4)
__host__ __device__ constexpr int na_max(int kb) { return kb == 64 ? 4 : 8; }
__host__ __device__ constexpr int a_bytes(int kb) { return (BH + 2) * P * kb * 2; }   // 40 KB / 20 KB
__host__ __device__ constexpr int a_slack(int kb) { return kb == 64 ? 1024 : 512; }    // zeroed rows behind the box (kx taps of the last positions)
__host__ __device__ constexpr int a_stage(int kb) { return a_bytes(kb) + a_slack(kb); }
constexpr int NEPI = 8;
constexpr int NTHREADS = (3 + NEPI) * 32;
constexpr int STAGE_BYTES = 32 * 20 * 4;
constexpr int DCOLS = 128;                    // TMEM columns per M-tile
constexpr int ACC_COLS = MT * DCOLS;          // 256 per accumulator set
constexpr int TMEM_COLS = 2 * ACC_COLS;
constexpr int SMEM_MAX = 227 * 1024 - 1024;  // dynamic part; the kernel also has ~300 B of static shared memory
constexpr int SMEM_FIXED = 1024 /*align*/ + NEPI * STAGE_BYTES + 256 /*barriers*/;

// this CTA's half of the weights of one (plane, ky): Cout=64 -> pair half (64 rows) + single half (32 rows);
// Cout<=32 -> triple half (64 rows); Cout<=16 -> 32 rows
__host__ __device__ constexpr int w_bytes_per_ky(int nreal, int kb) { return (nreal == 64 ? 12288 : (nreal == 32 ? 6144 : 4096)) * kb / 64; }

struct Bars {
    uint64_t a_full[NA_MAX], a_empty[NA_MAX];
    uint64_t w_full;
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}


// ---------------------------------------------------------------------------------------------------------------------
// Channel-split epilogue (option `epilogue_split`, template SPLIT): all 8 epilogue warps work on EVERY tile.  Warps w and
// w+4 share a TMEM lane quarter (= one pixel row of each M-tile) and split the output channels in halves, each warp pulls
// its share of BOTH M-tiles into registers, combines the horizontal taps, and hands the accumulator set back to the MMA
// issuer right there -- before any bias / activation / residual math or store.  Why: per-tile timestamps
// (profiles/r1_ts_1.txt) show the set-per-group scheme serialising MMA completion (~2 k clk) -> epilogue of that set
// (~3 k clk) per accumulator set, so the Cout=32 convs run at 2.6 k clk per tile where MMAs / A loads need 1.4 k.  With
// the early hand-back only the TMEM read (~0.3 k clk) stays on that chain.
// Where work tile `tt` lives: image b of its group, band by, column tile tx, the group's image size, pixel offset and maps.
struct TileGeo {
    int b, by, tx, H, W, desc0, g;
    size_t pxoff, aux;
};
template <bool MULTI>
__device__ __forceinline__ TileGeo decode_tile(const ConvArgs& a, const int tt, const int tiles_per_img, const int tiles_x)
{
    TileGeo r;
    if constexpr (MULTI) {
        int g = 0;
#pragma unroll
        for (int k = 1; k < 4; k++)
            if (k < a.ngroups && tt >= a.grp[k].tile_start) g = k;
        // explicit selects: a dynamically indexed kernel parameter would be copied to local memory
        const int H = g == 0 ? a.grp[0].H : g == 1 ? a.grp[1].H : g == 2 ? a.grp[2].H : a.grp[3].H;
        const int W = g == 0 ? a.grp[0].W : g == 1 ? a.grp[1].W : g == 2 ? a.grp[2].W : a.grp[3].W;
        const int tx_n = g == 0 ? a.grp[0].tiles_x : g == 1 ? a.grp[1].tiles_x : g == 2 ? a.grp[2].tiles_x : a.grp[3].tiles_x;
        const int tpi = g == 0 ? a.grp[0].tiles_per_img : g == 1 ? a.grp[1].tiles_per_img : g == 2 ? a.grp[2].tiles_per_img : a.grp[3].tiles_per_img;
        const int t0 = g == 0 ? a.grp[0].tile_start : g == 1 ? a.grp[1].tile_start : g == 2 ? a.grp[2].tile_start : a.grp[3].tile_start;
        r.desc0 = g == 0 ? a.grp[0].desc0 : g == 1 ? a.grp[1].desc0 : g == 2 ? a.grp[2].desc0 : a.grp[3].desc0;
        r.pxoff = (size_t)(g == 0 ? a.grp[0].px_off : g == 1 ? a.grp[1].px_off : g == 2 ? a.grp[2].px_off : a.grp[3].px_off);
        r.aux = (size_t)(g == 0 ? a.grp[0].aux_off : g == 1 ? a.grp[1].aux_off : g == 2 ? a.grp[2].aux_off : a.grp[3].aux_off);
        const int lt = tt - t0;
        r.b = lt / tpi;
        const int rem = lt - r.b * tpi;
        r.by = rem / tx_n;
        r.tx = rem - r.by * tx_n;
        r.H = H; r.W = W; r.g = g;
    } else {
        r.b = tt / tiles_per_img;
        const int rem = tt - r.b * tiles_per_img;
        r.by = rem / tiles_x;
        r.tx = rem - r.by * tiles_x;
        r.H = a.H; r.W = a.W; r.desc0 = 0; r.g = 0; r.pxoff = 0; r.aux = 0;
    }
    return r;
}

template <int NREAL, int MODE, bool MULTI>
__device__ __forceinline__ void epilogue_split(const ConvArgs& a, Bars* bars, const uint32_t tmem_base, uint8_t* smStage, const float* s_bias,
                                               const uint32_t rank, const int cluster_id, const int nclusters, const int ntiles,
                                               const int tiles_per_img, const int tiles_x, const int warp, const int lane)
{
    constexpr bool RDB16 = MODE == CONV_RDB_OUT16;
    constexpr bool RDB = MODE == CONV_RDB_OUT || RDB16;
    constexpr int NF = NREAL == 64 ? 32 : 16;  // accumulator values per M-tile and thread after the tap combination
    const int q4 = warp & 3;
    const int ew = warp - 3;
    const int chh = ew >> 2;  // channel half owned by this warp (Cout=3: the M-tile it owns)
    uint32_t* stg = reinterpret_cast<uint32_t*>(smStage + ew * STAGE_BYTES);  // [32 rows][20 words]
    const int tr = lane >> 2, tc = lane & 3;
    const bool dbg_noepi = (a.debug_mask & 8) != 0;
    const uint32_t acc_empty_leader0 = mapa_u32(smem_u32(&bars->acc_empty[0]), 0);
    const uint32_t acc_empty_leader1 = mapa_u32(smem_u32(&bars->acc_empty[1]), 0);
    int ti = 0;
    for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
        const int tt = a.reverse ? ntiles - 1 - t : t;  // serpentine layer order, see ConvArgs::reverse
        const TileGeo geo = decode_tile<MULTI>(a, tt, tiles_per_img, tiles_x);
        const int b = geo.b, by = geo.by, tx = geo.tx;
        const int x0 = tx * BW;
        const int buf = ti & 1;
        int yy[2];
        size_t rowpix[2];
        bool vrow[2];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            yy[i] = by * BAND + (int)rank * BH + i * 4 + q4;
            vrow[i] = yy[i] < geo.H && !dbg_noepi;
            rowpix[i] = geo.pxoff + ((size_t)b * geo.H + yy[i]) * geo.W + x0;
        }
        bool vcol[4];
#pragma unroll
        for (int r = 0; r < 4; r++) vcol[r] = (tr + 8 * r < BW) && (x0 + tr + 8 * r < geo.W);

        // ---- residual operands requested before the accumulators are waited for ----
        uint2 x0raw[2][2][4];  // RDB16: [i][hf][r], channels chh*32 + hf*16 + tc*4 .. +3 of pixel tr + 8r
        float4 res[2][4];      // fp32 modes: one M-tile ahead
        auto load_res = [&](int i, float4 (&dst)[2][4]) {
            if constexpr (MODE == CONV_RDB_OUT || MODE == CONV_TRUNK) {
                const float* base = (MODE == CONV_RDB_OUT ? a.t32 : a.f32) + rowpix[i] * 64 + chh * 32 + tc * 4;
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (vcol[r] && vrow[i]) dst[hf][r] = *reinterpret_cast<const float4*>(base + (size_t)(tr + 8 * r) * 64 + hf * 16);
            }
        };
        if constexpr (RDB16) {
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (vcol[r] && vrow[i])
                            x0raw[i][hf][r] = *reinterpret_cast<const uint2*>(a.res16 + (rowpix[i] + tr + 8 * r) * 64 + chh * 32 + hf * 16 + tc * 4);
        } else {
            load_res(0, res);
        }
        if constexpr (RDB) {
            if (a.rrdb_end) {  // RRDB-level skip operands of THIS tile: pull this warp's half of each 256-B pixel into L2
#pragma unroll
                for (int i = 0; i < 2; i++)
                    if (vrow[i]) {
                        const int npx = min(BW, geo.W - x0);
                        const char* p1 = reinterpret_cast<const char*>(a.r32 + rowpix[i] * 64 + chh * 32);
                        for (int l = lane; l < npx; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + (size_t)l * 256));
                    }
            }
        }
        const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && rank == 0 && ew == 0 && lane == 0;
        if (ts_on) a.dbg_ts[ti * 8 + 4] = clock64();
        mbar_wait(&bars->acc_full[buf], (ti >> 1) & 1);
        tc_fence_after();
        if (ts_on) a.dbg_ts[ti * 8 + 5] = clock64();

        // ---- accumulators of both M-tiles -> registers (taps combined), then give the set back ----
        float f[2][NF];
        float o3[3] = {0.f, 0.f, 0.f};
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS;
        if constexpr (NREAL == 16) {
            // Cout = 3: 16 + 16 + 16 columns per M-tile; the warp of group chh takes M-tile chh, the other one only signs off
            uint32_t v[32], u[16];
            tmem_ld32(taddr0 + chh * DCOLS, v);
            tmem_ld16(taddr0 + chh * DCOLS + 32, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 3; j++)
                o3[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[16 + j]), 1) +
                        __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 2);
        } else if constexpr (NREAL == 32) {
            // D[c] + right neighbour's D[32 + c] + second right neighbour's D[64 + c], c = chh*16 .. +15
            uint32_t v[2][16], u[2][16], w[2][16];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                tmem_ld16(taddr0 + i * DCOLS + chh * 16, v[i]);
                tmem_ld16(taddr0 + i * DCOLS + 32 + chh * 16, u[i]);
                tmem_ld16(taddr0 + i * DCOLS + 64 + chh * 16, w[i]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 16; j++)
                    f[i][j] = __uint_as_float(v[i][j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[i][j]), 1) +
                              __shfl_down_sync(0xffffffffu, __uint_as_float(w[i][j]), 2);
        } else {
            // D[c] + right neighbour's D[64 + c], c = chh*32 .. +31
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint32_t v[32], u[32];
                tmem_ld32(taddr0 + i * DCOLS + chh * 32, v);
                tmem_ld32(taddr0 + i * DCOLS + 64 + chh * 32, u);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) f[i][j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1);
            }
        }
        // Plain (CTA-scope release) remote arrive: only the TMEM reads have to be ordered (tcgen05 fence); a cluster-scope
        // release would also wait for this thread's global stores to drain.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(buf ? acc_empty_leader1 : acc_empty_leader0);
        if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();

        // ---- math + stores from registers ----
        if constexpr (NREAL == 16) {
            const int yyi = chh ? yy[1] : yy[0];       // (selects, not a dynamic index: keeps yy / vrow in registers)
            const bool vri = chh ? vrow[1] : vrow[0];
            const int x = x0 + lane;
            if (lane < BW && x < geo.W && vri) {
                const TileDesc td = a.tiles[geo.desc0 + b];
                float o[3];
#pragma unroll
                for (int j = 0; j < 3; j++) o[j] = o3[j] + s_bias[j];
                if constexpr (MODE == CONV_LAST_U8) {
                    const int ii = yyi - a.prepad_s, jx = x - a.prepad_s;
                    if (ii >= 0 && ii < td.crop_h && jx >= 0 && jx < td.crop_w) {
                        uint8_t* dst = a.img_out + ((size_t)(td.out_y0 + ii) * a.img_w + td.out_x0 + jx) * a.img_c;
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            int iv = (int)(o[j] * 255.f + 0.5f);
                            dst[j] = (uint8_t)min(max(iv, 0), 255);
                        }
                    }
                } else {
                    float4* dst = reinterpret_cast<float4*>(a.tta_out) + geo.aux + (size_t)td.tta_slot * ((size_t)geo.H * geo.W) + (size_t)yyi * geo.W + x;
                    *dst = make_float4(o[0], o[1], o[2], 0.f);
                }
            }
        } else if constexpr (MODE == CONV_STORE16 && NREAL == 32) {
            // 16 channels = 32 B per pixel: lanes 0..15 carry channels chh*16 .. +7 of 16 pixels, lanes 16..31 the other 8
            const int p = lane & 15, tc2 = lane >> 4;
#pragma unroll
            for (int i = 0; i < 2; i++) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float x = f[i][j] + s_bias[chh * 16 + j];
                    if (a.act == 2) x = x > 0.f ? x : x * a.slope;
                    f[i][j] = x;
                }
#pragma unroll
                for (int k = 0; k < 2; k++)
                    *reinterpret_cast<uint4*>(stg + lane * 20 + 4 * k) =
                        make_uint4(pack_h2(f[i][8 * k + 0], f[i][8 * k + 1]), pack_h2(f[i][8 * k + 2], f[i][8 * k + 3]),
                                   pack_h2(f[i][8 * k + 4], f[i][8 * k + 5]), pack_h2(f[i][8 * k + 6], f[i][8 * k + 7]));
                __syncwarp();
                __half* orow = a.out16 + rowpix[i] * a.out_cs + a.out_co + chh * 16 + tc2 * 8;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const int px = p + 16 * r;
                    const uint4 o = *reinterpret_cast<const uint4*>(stg + px * 20 + 4 * tc2);
                    if (px < BW && x0 + px < geo.W && vrow[i]) *reinterpret_cast<uint4*>(orow + (size_t)px * a.out_cs) = o;
                }
                __syncwarp();
            }
        } else if constexpr (MODE == CONV_STORE16) {
            // Cout = 64: this warp's 32 channels = 64 B per pixel, 4 lanes per pixel
#pragma unroll
            for (int i = 0; i < 2; i++) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    float x = f[i][j] + s_bias[chh * 32 + j];
                    if (a.act == 2) x = x > 0.f ? x : x * a.slope;
                    f[i][j] = x;
                }
#pragma unroll
                for (int k = 0; k < 4; k++)
                    *reinterpret_cast<uint4*>(stg + lane * 20 + 4 * k) =
                        make_uint4(pack_h2(f[i][8 * k + 0], f[i][8 * k + 1]), pack_h2(f[i][8 * k + 2], f[i][8 * k + 3]),
                                   pack_h2(f[i][8 * k + 4], f[i][8 * k + 5]), pack_h2(f[i][8 * k + 6], f[i][8 * k + 7]));
                __syncwarp();
                __half* orow = a.out16 + rowpix[i] * a.out_cs + a.out_co + chh * 32 + tc * 8;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const uint4 o = *reinterpret_cast<const uint4*>(stg + (tr + 8 * r) * 20 + 4 * tc);
                    if (vcol[r] && vrow[i]) *reinterpret_cast<uint4*>(orow + (size_t)(tr + 8 * r) * a.out_cs) = o;
                }
                __syncwarp();
            }
        } else {
            // CONV_RDB_OUT / CONV_RDB_OUT16 / CONV_TRUNK (Cout = 64): fp32 math against the residual stream, transposed domain
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float4 nxt[2][4];
                if constexpr (!RDB16) {
                    if (i == 0) load_res(1, nxt);
                }
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        *reinterpret_cast<float4*>(stg + lane * 20 + 4 * k) =
                            make_float4(f[i][16 * hf + 4 * k], f[i][16 * hf + 4 * k + 1], f[i][16 * hf + 4 * k + 2], f[i][16 * hf + 4 * k + 3]);
                    __syncwarp();
                    const int ch0 = chh * 32 + hf * 16 + tc * 4;
                    const float4 bb = *reinterpret_cast<const float4*>(&s_bias[ch0]);
                    float4 acc4[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) acc4[r] = *reinterpret_cast<const float4*>(stg + (tr + 8 * r) * 20 + 4 * tc);
                    __syncwarp();
                    if constexpr (RDB16) {
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&x0raw[i][hf][r].x));
                            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&x0raw[i][hf][r].y));
                            res[hf][r] = make_float4(lo.x, lo.y, hi.x, hi.y);
                        }
                    }
                    float4 rr[4];
                    if (RDB && a.rrdb_end) {
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (vcol[r] && vrow[i]) rr[r] = *reinterpret_cast<const float4*>(a.r32 + (rowpix[i] + tr + 8 * r) * 64 + ch0);
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float4 o;
                        if constexpr (RDB) {
                            o.x = (acc4[r].x + bb.x) * 0.2f + res[hf][r].x;
                            o.y = (acc4[r].y + bb.y) * 0.2f + res[hf][r].y;
                            o.z = (acc4[r].z + bb.z) * 0.2f + res[hf][r].z;
                            o.w = (acc4[r].w + bb.w) * 0.2f + res[hf][r].w;
                            if (a.rrdb_end) {
                                o.x = o.x * 0.2f + rr[r].x; o.y = o.y * 0.2f + rr[r].y;
                                o.z = o.z * 0.2f + rr[r].z; o.w = o.w * 0.2f + rr[r].w;
                            }
                        } else {
                            o.x = acc4[r].x + bb.x + res[hf][r].x; o.y = acc4[r].y + bb.y + res[hf][r].y;
                            o.z = acc4[r].z + bb.z + res[hf][r].z; o.w = acc4[r].w + bb.w + res[hf][r].w;
                        }
                        if (vcol[r] && vrow[i]) {
                            const size_t px = rowpix[i] + tr + 8 * r;
                            if constexpr (RDB) {
                                if (a.rrdb_end) *reinterpret_cast<float4*>(a.r32 + px * 64 + ch0) = o;
                                if constexpr (!RDB16) *reinterpret_cast<float4*>(a.t32 + px * 64 + ch0) = o;
                            }
                            *reinterpret_cast<uint2*>(a.out16 + px * a.out_cs + a.out_co + ch0) = make_uint2(pack_h2(o.x, o.y), pack_h2(o.z, o.w));
                        }
                    }
                }
                if constexpr (!RDB16) {
                    if (i == 0) {
#pragma unroll
                        for (int h2 = 0; h2 < 2; h2++)
#pragma unroll
                            for (int r = 0; r < 4; r++) res[h2][r] = nxt[h2][r];
                    }
                }
            }
        }
        if (ts_on) a.dbg_ts[ti * 8 + 7] = clock64();
    }
}

template <int NREAL, int MODE, bool SPLIT, int KB, bool MULTI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
conv3x3_tc2x_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs a, const int NA)
{
    constexpr int A_BYTES = a_bytes(KB), A_STAGE = a_stage(KB);
    constexpr int ROWB = KB * 2;                     // bytes per smem row of an operand tile (one pixel / one weight row)
    constexpr int GOFF = 64 * ROWB;                  // Cout=64: offset of the single-tap group behind the 64 pair rows
    constexpr int NGROUPS = NREAL == 64 ? 2 : 1;           // MMA groups per kernel row (pair + single, or one triple)
    constexpr int ROWS_PER_KY = NREAL == 64 ? 192 : (NREAL == 32 ? 96 : 64);  // weight rows per (plane, ky), see model.cpp
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nchunk = (a.cin + KB - 1) / KB;       // K-blocks
    constexpr int W_KY = w_bytes_per_ky(NREAL, KB);
    uint8_t* smA = smem;
    uint8_t* smW = smem + NA * A_STAGE;                  // resident weights: [plane][ky] x W_KY bytes
    uint8_t* smStage = smW + nchunk * 3 * W_KY;
    Bars* bars = reinterpret_cast<Bars*>(smStage + NEPI * STAGE_BYTES);
    __shared__ __align__(16) float s_bias[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int tiles_x = (a.W + BW - 1) / BW, bands_y = (a.H + BAND - 1) / BAND;
    const int tiles_per_img = tiles_x * bands_y;
    const int ntiles = MULTI ? a.ntiles_multi : tiles_per_img * a.B;  // pair tiles (MULTI: tiles_x / tiles_per_img come per group)
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA_MAX; i++) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], 1); }
        mbar_init(&bars->w_full, 1);
        for (int i = 0; i < 2; i++) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], SPLIT ? 2 * NEPI : NEPI); }  // acc_empty: 4 (SPLIT: 8) warps x 2 CTAs
        mbar_fence_init();
        for (int i = 0; i < 3; i++) tma_prefetch_desc(&maps.m[i]);
        if constexpr (MULTI) {
            for (int g = 1; g < a.ngroups; g++)
                for (int i = 0; i < 3; i++) tma_prefetch_desc(&maps.mg[g - 1][i]);
        }
        tma_prefetch_desc(&maps.w);
        tma_prefetch_desc(&maps.w32);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + NREAL) s_bias[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
    // rows behind each A box: read by the kx=2 taps of the last two (never stored) positions; keep them finite
    {
        constexpr int SL16 = a_slack(KB) / 16;  // uint4 per stage: 64 (KB=64, NA <= 4) or 32 (KB=32, NA <= 8): <= 256 threads
        if (threadIdx.x >= 96 && threadIdx.x < 96 + NA * SL16) {
            const int i = threadIdx.x - 96;
            reinterpret_cast<uint4*>(smA + (i / SL16) * A_STAGE + A_BYTES)[i % SL16] = make_uint4(0, 0, 0, 0);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cluster_sync();  // barriers of both CTAs are initialised before anybody signals across
    if (warp == 1) tmem_alloc2(&bars->tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_ptr;
    // Resident weights: every CTA fetches ITS half of all weight blocks once (before the grid dependency is resolved --
    // weights are constants, so this overlaps the previous conv's tail).  All bytes are counted on the leader's w_full.
    if (warp == 2) {
        if (elect_one() && !(a.debug_mask & 4)) {
            const uint32_t wbar = mapa_u32(smem_u32(&bars->w_full), 0);
            if (is_leader) mbar_expect_tx(&bars->w_full, 2u * (uint32_t)(nchunk * 3 * W_KY));
            for (int c = 0; c < nchunk; c++)
                for (int ky = 0; ky < 3; ky++) {
                    uint8_t* dst = smW + (c * 3 + ky) * W_KY;
                    // weight rows are stored per (64-channel plane, ky); a 32-channel K-block is the left or right half of a row
                    const int plane = KB == 64 ? c : (c >> 1), col = KB == 64 ? 0 : (c & 1) * 32;
                    const int row0 = (plane * 3 + ky) * ROWS_PER_KY;
                    if (NREAL == 16) tma2_load_2d(dst, &maps.w32, wbar, col, row0 + (int)rank * 32);
                    else tma2_load_2d(dst, &maps.w, wbar, col, row0 + (int)rank * (NREAL == 32 ? 48 : 64));  // maps.w box: 48 rows for Cout<=32
                    if (NREAL == 64) tma2_load_2d(dst + GOFF, &maps.w32, wbar, col, row0 + 128 + (int)rank * 32);
                }
        } else if (is_leader && (a.debug_mask & 4) && lane == 0) {
            mbar_arrive(&bars->w_full);
        }
    }
    // Programmatic dependent launch: let the next conv's CTAs be scheduled onto SMs as soon as this grid's CTAs retire
    // (they run their prologue and then block here), and wait until the previous grid has completed and flushed its
    // outputs before touching any activation.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ---------------- A producer (both CTAs; bytes are counted on the leader's barrier) ----------------
        const bool leader = elect_one();
        int ca = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters) {
            const int tt = a.reverse ? ntiles - 1 - t : t;  // serpentine layer order, see ConvArgs::reverse
            const TileGeo geo = decode_tile<MULTI>(a, tt, tiles_per_img, tiles_x);
            const int b = geo.b, by = geo.by, tx = geo.tx;
            const CUtensorMap* gmaps = (MULTI && geo.g > 0) ? maps.mg[geo.g - 1] : maps.m;
            for (int c = 0; c < nchunk; c++, ca++) {
                const int s = ca % NA;
                mbar_wait(&bars->a_empty[s], ((ca / NA) & 1) ^ 1);
                if (leader) {
                    if (a.debug_mask & 2) {
                        if (is_leader) mbar_arrive(&bars->a_full[s]);
                    } else {
                        if (is_leader) mbar_expect_tx(&bars->a_full[s], 2 * A_BYTES);
                        tma2_load_4d(smA + s * A_STAGE, &gmaps[KB == 64 ? c : (c >> 1)], mapa_u32(smem_u32(&bars->a_full[s]), 0),
                                     KB == 64 ? 0 : (c & 1) * 32, tx * BW - 1,
                                     by * BAND + (int)rank * BH - 1, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (is_leader) {
            const bool leader = elect_one();
            constexpr uint32_t idesc128 = umma_idesc_f16(256, NREAL == 32 ? 96 : 128);  // Cout<=32: three 32-wide tap groups
            constexpr uint32_t idesc64 = umma_idesc_f16(256, 64);
            const uint64_t desc_hi = KB == 64 ? umma_desc_sw128(0, 1024, 0) : umma_desc_sw64(0, 512, 0);
            int ca = 0, ti = 0;
            mbar_wait(&bars->w_full, 0);
            tc_fence_after();
            const uint32_t w_base = smem_u32(smW);
            for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
                const int buf = ti & 1;
                const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && leader;
                if (ts_on) a.dbg_ts[ti * 8 + 0] = clock64();
                mbar_wait(&bars->acc_empty[buf], ((ti >> 1) & 1) ^ 1);
                tc_fence_after();
                if (ts_on) a.dbg_ts[ti * 8 + 1] = clock64();
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                for (int c = 0; c < nchunk; c++, ca++) {
                    const int sA = ca % NA;
                    mbar_wait(&bars->a_full[sA], (ca / NA) & 1);
                    tc_fence_after();
                    if (ts_on && c == 0) a.dbg_ts[ti * 8 + 2] = clock64();
                    const int rem = a.cin - c * KB;
                    const int ksteps = rem >= KB ? KB / 16 : (rem >> 4);
                    const uint32_t a_base = smem_u32(smA + sA * A_STAGE);
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int g = 0; g < NGROUPS; g++) {
                            const uint32_t idesc = (NREAL == 16 || g == 1) ? idesc64 : idesc128;
                            const uint64_t db0 = desc_hi | (uint64_t)(((w_base + (uint32_t)((c * 3 + ky) * W_KY + g * GOFF)) & 0x3FFFFu) >> 4);
                            const uint64_t da0 = desc_hi | (uint64_t)(((a_base + (uint32_t)(ky * P + g * 2) * (uint32_t)ROWB) & 0x3FFFFu) >> 4);
#pragma unroll
                            for (int ks = 0; ks < KB / 16; ks++) {
                                if (ks < ksteps && leader && !(a.debug_mask & 1)) {
                                    const uint32_t accum = (c | ky | g | ks) ? 1u : 0u;
                                    umma2_f16(acc + 0 * DCOLS, da0 + ks * 2, db0 + ks * 2, idesc, accum);
                                    umma2_f16(acc + 1 * DCOLS, da0 + ks * 2 + (128 * ROWB) / 16, db0 + ks * 2, idesc, accum);  // second M-tile: 128 rows further
                                }
                            }
                        }
                    if (leader) umma2_commit(&bars->a_empty[sA], 3);
                }
                if (leader) umma2_commit(&bars->acc_full[buf], 3);
                if (ts_on) a.dbg_ts[ti * 8 + 3] = clock64();
            }
        }
    } else if (warp == 2) {
        // weight warp: its work (the one-time resident-weight fetch) happened in the prologue
    } else if constexpr (SPLIT) {
        epilogue_split<NREAL, MODE, MULTI>(a, bars, tmem_base, smStage, s_bias, rank, cluster_id, nclusters, ntiles, tiles_per_img, tiles_x, warp, lane);
    } else {
        // ---------------- epilogue: warps 3..10 in two groups of four ----------------
        // Group g = (warp-3)/4 owns accumulator set g, i.e. every second work tile; inside a group warp q4 = warp % 4
        // (its TMEM lane quarter) handles row q4 of both M-tiles.  Two tiles are therefore in the epilogue stage at any
        // time and each group has two tile periods to drain its set (timestamps showed a single 8-warp epilogue taking
        // 3-5 k clk per tile and stalling the 1-4 k clk MMA stream of the Cout=32 convs).
        const int q4 = warp & 3;
        const int ew = warp - 3;
        const int group = ew >> 2;
        uint32_t* stg = reinterpret_cast<uint32_t*>(smStage + ew * STAGE_BYTES);  // [32 rows][20 words]
        const int tr = lane >> 2, tc = lane & 3;
        const bool dbg_noepi = (a.debug_mask & 8) != 0;
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(&bars->acc_empty[group]), 0);
        int ti = 0;
        for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
            if ((ti & 1) != group) continue;
            const int tt = a.reverse ? ntiles - 1 - t : t;  // serpentine layer order, see ConvArgs::reverse
            const TileGeo geo = decode_tile<MULTI>(a, tt, tiles_per_img, tiles_x);
            const int b = geo.b, by = geo.by, tx = geo.tx;
            const int x0 = tx * BW;
            const int buf = group;
            int yy[2];
            size_t rowpix[2];
            bool vrow[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                yy[i] = by * BAND + (int)rank * BH + i * 4 + q4;
                vrow[i] = yy[i] < geo.H && !dbg_noepi;
                rowpix[i] = geo.pxoff + ((size_t)b * geo.H + yy[i]) * geo.W + x0;
            }
            bool vcol[4];
#pragma unroll
            for (int r = 0; r < 4; r++) vcol[r] = (tr + 8 * r < BW) && (x0 + tr + 8 * r < geo.W);
            // residual operands.  fp32 modes (CONV_RDB_OUT with T32, CONV_TRUNK with F32): one (M-tile, 32-channel block) is
            // fetched one block ahead of use.  CONV_RDB_OUT16: ALL fp16 x0 values this thread needs for the tile (4 blocks x
            // 8 x 8 B = 64 registers) are requested here, before waiting for the MMAs -- per-tile timestamps showed the
            // one-block-ahead scheme paying one memory latency (~3.7 k clk) per block, 15 k clk per tile against 7.5 k clk of
            // MMAs.
            constexpr bool RDB16 = MODE == CONV_RDB_OUT16;
            constexpr bool RDB = MODE == CONV_RDB_OUT || RDB16;
            float4 res[2][4];
            uint2 x0raw[2][2][2][4];  // [i][nb][hf][r], RDB16 only
            auto load_res = [&](int i, int nb, float4 (&dst)[2][4]) {
                if constexpr (MODE == CONV_RDB_OUT || MODE == CONV_TRUNK) {
                    const float* base = (MODE == CONV_RDB_OUT ? a.t32 : a.f32) + rowpix[i] * 64 + nb * 32 + tc * 4;
#pragma unroll
                    for (int hf = 0; hf < 2; hf++)
#pragma unroll
                        for (int r = 0; r < 4; r++)
                            if (vcol[r] && vrow[i]) dst[hf][r] = *reinterpret_cast<const float4*>(base + (size_t)(tr + 8 * r) * 64 + hf * 16);
                }
            };
            if constexpr (RDB16) {
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++)
#pragma unroll
                        for (int hf = 0; hf < 2; hf++)
#pragma unroll
                            for (int r = 0; r < 4; r++)
                                if (vcol[r] && vrow[i])
                                    x0raw[i][nb][hf][r] = *reinterpret_cast<const uint2*>(a.res16 + (rowpix[i] + tr + 8 * r) * 64 + nb * 32 + hf * 16 + tc * 4);
            } else {
                load_res(0, 0, res);
            }
            if constexpr (RDB) {
                if (a.rrdb_end) {  // the RRDB-level skip operands of THIS tile: pull the rows into L2 while the MMAs run
#pragma unroll
                    for (int i = 0; i < 2; i++)
                        if (vrow[i]) {
                            const int npx = min(BW, geo.W - x0);
                            const char* p1 = reinterpret_cast<const char*>(a.r32 + rowpix[i] * 64);
                            for (int l = lane; l < npx * 2; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + (size_t)l * 128));
                        }
                }
            }
            // (An L2 prefetch of the next tile's residual rows was tried and measured 5 % slower end to end: the Cout=64
            // convs are HBM-bandwidth-bound in situ -- 1.44 GB per launch at 4.5 TB/s -- not latency-bound.)
            const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && is_leader && (ew & 3) == 0 && lane == 0;
            if (ts_on) a.dbg_ts[ti * 8 + 4] = clock64();
            mbar_wait(&bars->acc_full[buf], (ti >> 1) & 1);
            tc_fence_after();
            if (ts_on) a.dbg_ts[ti * 8 + 5] = clock64();
            if constexpr (MODE == CONV_STORE16 && NREAL == 32) {
                if (a.epi_split == 2) {
                    // Early hand-back (option epilogue_split=2): both M-tiles' accumulators go to registers (taps combined, same
                    // operation order as below), the set returns to the MMA issuer, and only then bias/activation/pack/store
                    // run -- the chain "MMA done -> epilogue of that set -> arrive" (profiles/r1_ts_1.txt: ~3.0 k clk per
                    // tile) shrinks to the TMEM read.  The two 4-warp groups still alternate tiles, so one group's stores
                    // overlap the other group's TMEM reads.
                    float f[2][32];
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS + i * DCOLS;
                        uint32_t v[32], u[32], w[32];
                        tmem_ld32(taddr, v);
                        tmem_ld32(taddr + 32, u);
                        tmem_ld32(taddr + 64, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            f[i][j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                      __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2) + s_bias[j];
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(acc_empty_leader);
                    if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        if (a.act == 2) {
#pragma unroll
                            for (int j = 0; j < 32; j++) f[i][j] = f[i][j] > 0.f ? f[i][j] : f[i][j] * a.slope;
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            *reinterpret_cast<uint4*>(stg + lane * 20 + 4 * k) =
                                make_uint4(pack_h2(f[i][8 * k + 0], f[i][8 * k + 1]), pack_h2(f[i][8 * k + 2], f[i][8 * k + 3]),
                                           pack_h2(f[i][8 * k + 4], f[i][8 * k + 5]), pack_h2(f[i][8 * k + 6], f[i][8 * k + 7]));
                        __syncwarp();
                        __half* orow = a.out16 + rowpix[i] * a.out_cs + a.out_co + tc * 8;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint4 o = *reinterpret_cast<const uint4*>(stg + (tr + 8 * r) * 20 + 4 * tc);
                            if (vcol[r] && vrow[i]) *reinterpret_cast<uint4*>(orow + (size_t)(tr + 8 * r) * a.out_cs) = o;
                        }
                        __syncwarp();
                    }
                    if (ts_on) a.dbg_ts[ti * 8 + 7] = clock64();
                    continue;
                }
            }
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS + i * DCOLS;
                if constexpr (NREAL == 16) {
                    uint32_t v[32], u[16];
                    tmem_ld32(taddr, v);
                    tmem_ld16(taddr + 32, u);
                    tmem_ld_wait();
                    float o[3];
#pragma unroll
                    for (int j = 0; j < 3; j++)
                        o[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[16 + j]), 1) +
                               __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 2) + s_bias[j];
                    const int x = x0 + lane;
                    if (lane < BW && x < geo.W && vrow[i]) {
                        const TileDesc td = a.tiles[geo.desc0 + b];
                        if constexpr (MODE == CONV_LAST_U8) {
                            const int ii = yy[i] - a.prepad_s, jx = x - a.prepad_s;
                            if (ii >= 0 && ii < td.crop_h && jx >= 0 && jx < td.crop_w) {
                                uint8_t* dst = a.img_out + ((size_t)(td.out_y0 + ii) * a.img_w + td.out_x0 + jx) * a.img_c;
#pragma unroll
                                for (int j = 0; j < 3; j++) {
                                    int iv = (int)(o[j] * 255.f + 0.5f);
                                    dst[j] = (uint8_t)min(max(iv, 0), 255);
                                }
                            }
                        } else {
                            float4* dst = reinterpret_cast<float4*>(a.tta_out) + geo.aux + (size_t)td.tta_slot * ((size_t)geo.H * geo.W) + (size_t)yy[i] * geo.W + x;
                            *dst = make_float4(o[0], o[1], o[2], 0.f);
                        }
                    }
                } else if constexpr (MODE == CONV_STORE16) {
#pragma unroll
                    for (int nb = 0; nb < NREAL / 32; nb++) {
                        float f[32];
                        if constexpr (NREAL == 32) {  // D[0:32] + right neighbour's D[32:64] + second right neighbour's D[64:96]
                            uint32_t v[32], u[32], w[32];
                            tmem_ld32(taddr, v);
                            tmem_ld32(taddr + 32, u);
                            tmem_ld32(taddr + 64, w);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                f[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                       __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2) + s_bias[j];
                        } else {  // D[nb*32..] + right neighbour's D[64 + nb*32..]
                            uint32_t v[32], u[32];
                            tmem_ld32(taddr + nb * 32, v);
                            tmem_ld32(taddr + 64 + nb * 32, u);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                f[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) + s_bias[nb * 32 + j];
                        }
                        if (a.act == 2) {
#pragma unroll
                            for (int j = 0; j < 32; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * a.slope;
                        }
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            *reinterpret_cast<uint4*>(stg + lane * 20 + 4 * k) =
                                make_uint4(pack_h2(f[8 * k + 0], f[8 * k + 1]), pack_h2(f[8 * k + 2], f[8 * k + 3]),
                                           pack_h2(f[8 * k + 4], f[8 * k + 5]), pack_h2(f[8 * k + 6], f[8 * k + 7]));
                        __syncwarp();
                        __half* orow = a.out16 + rowpix[i] * a.out_cs + a.out_co + nb * 32 + tc * 8;
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const uint4 o = *reinterpret_cast<const uint4*>(stg + (tr + 8 * r) * 20 + 4 * tc);
                            if (vcol[r] && vrow[i]) *reinterpret_cast<uint4*>(orow + (size_t)(tr + 8 * r) * a.out_cs) = o;
                        }
                        __syncwarp();
                    }
                } else {
                    // CONV_RDB_OUT / CONV_TRUNK (Cout = 64): fp32 math against the residual stream in the transposed domain
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        uint32_t v[32], u[32];
                        tmem_ld32(taddr + nb * 32, v);
                        tmem_ld32(taddr + 64 + nb * 32, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            v[j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1));
                        float4 nxt[2][4];
                        if constexpr (!RDB16) {  // prefetch the residual operands of the next (M-tile, 32-channel block)
                            const int ni = nb == 1 ? i + 1 : i, nnb = nb ^ 1;
                            if (ni < 2) load_res(ni, nnb, nxt);
                        }
#pragma unroll
                        for (int hf = 0; hf < 2; hf++) {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                *reinterpret_cast<uint4*>(stg + lane * 20 + 4 * k) = make_uint4(v[16 * hf + 4 * k], v[16 * hf + 4 * k + 1], v[16 * hf + 4 * k + 2], v[16 * hf + 4 * k + 3]);
                            __syncwarp();
                            const int ch0 = nb * 32 + hf * 16 + tc * 4;
                            const float4 bb = *reinterpret_cast<const float4*>(&s_bias[ch0]);
                            float4 acc4[4];
#pragma unroll
                            for (int r = 0; r < 4; r++) acc4[r] = *reinterpret_cast<const float4*>(stg + (tr + 8 * r) * 20 + 4 * tc);
                            __syncwarp();
                            if constexpr (RDB16) {
#pragma unroll
                                for (int r = 0; r < 4; r++) {
                                    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&x0raw[i][nb][hf][r].x));
                                    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&x0raw[i][nb][hf][r].y));
                                    res[hf][r] = make_float4(lo.x, lo.y, hi.x, hi.y);
                                }
                            }
                            float4 rr[4];
                            if (RDB && a.rrdb_end) {
#pragma unroll
                                for (int r = 0; r < 4; r++)
                                    if (vcol[r] && vrow[i]) rr[r] = *reinterpret_cast<const float4*>(a.r32 + (rowpix[i] + tr + 8 * r) * 64 + ch0);
                            }
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                float4 o;
                                if constexpr (RDB) {
                                    o.x = (acc4[r].x + bb.x) * 0.2f + res[hf][r].x;
                                    o.y = (acc4[r].y + bb.y) * 0.2f + res[hf][r].y;
                                    o.z = (acc4[r].z + bb.z) * 0.2f + res[hf][r].z;
                                    o.w = (acc4[r].w + bb.w) * 0.2f + res[hf][r].w;
                                    if (a.rrdb_end) {
                                        o.x = o.x * 0.2f + rr[r].x; o.y = o.y * 0.2f + rr[r].y;
                                        o.z = o.z * 0.2f + rr[r].z; o.w = o.w * 0.2f + rr[r].w;
                                    }
                                } else {
                                    o.x = acc4[r].x + bb.x + res[hf][r].x; o.y = acc4[r].y + bb.y + res[hf][r].y;
                                    o.z = acc4[r].z + bb.z + res[hf][r].z; o.w = acc4[r].w + bb.w + res[hf][r].w;
                                }
                                if (vcol[r] && vrow[i]) {
                                    const size_t px = rowpix[i] + tr + 8 * r;
                                    if constexpr (RDB) {
                                        if (a.rrdb_end) *reinterpret_cast<float4*>(a.r32 + px * 64 + ch0) = o;
                                        if constexpr (!RDB16) *reinterpret_cast<float4*>(a.t32 + px * 64 + ch0) = o;
                                    }
                                    *reinterpret_cast<uint2*>(a.out16 + px * a.out_cs + a.out_co + ch0) = make_uint2(pack_h2(o.x, o.y), pack_h2(o.z, o.w));
                                }
                            }
                        }
                        if constexpr (!RDB16) {
#pragma unroll
                            for (int h2 = 0; h2 < 2; h2++)
#pragma unroll
                                for (int r = 0; r < 4; r++) res[h2][r] = nxt[h2][r];
                        }
                    }
                }
            }
            // accumulator set drained by this warp: tell the leader's MMA issuer.  Plain (CTA-scope release) remote arrive:
            // only the TMEM reads have to be ordered (tcgen05 fence above); a cluster-scope release would also wait for
            // this thread's global stores to drain.
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc_empty_leader);
            if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
        }
    }

    tc_fence_before();
    cluster_sync();  // nobody frees TMEM / exits while the peer may still signal or be read
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <int NREAL, int MODE, bool SPLIT, int KB, bool MULTI>
cudaError_t launch_nmskm(const ConvMaps* in_map, const ConvArgs& a, cudaStream_t stream)
{
    static bool attr_done[64] = {};
    static int num_sms[64] = {};
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    dev &= 63;
    if (!attr_done[dev]) {
        e = cudaFuncSetAttribute(conv3x3_tc2x_kernel<NREAL, MODE, SPLIT, KB, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
        if (e != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    const long long ntiles = MULTI ? (long long)a.ntiles_multi : (long long)((a.W + BW - 1) / BW) * ((a.H + BAND - 1) / BAND) * a.B;
    if (ntiles <= 0) return cudaErrorInvalidValue;
    long long nclusters = num_sms[dev] / 2;
    if (ntiles < nclusters) nclusters = ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(NTHREADS);
    // resident weights first, then as many 41 KB A stages as fit (2..4)
    const int nchunk = (a.cin + KB - 1) / KB;
    const int wbytes = nchunk * 3 * w_bytes_per_ky(NREAL, KB);
    int na = (SMEM_MAX - SMEM_FIXED - wbytes) / a_stage(KB);
    if (na > na_max(KB)) na = na_max(KB);
    if (na < 2) return cudaErrorInvalidValue;
    cfg.dynamicSmemBytes = SMEM_FIXED + wbytes + na * a_stage(KB);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const int use_pdl = getenv("RSR_NOPDL") ? 0 : 1;  // bring-up switch
    cfg.numAttrs = use_pdl;
    if ((e = cudaLaunchKernelEx(&cfg, conv3x3_tc2x_kernel<NREAL, MODE, SPLIT, KB, MULTI>, *in_map, a, na)) != cudaSuccess) return e;
    return cudaGetLastError();
}

template <int NREAL, int MODE>
cudaError_t launch_nm(const ConvMaps* in_map, const ConvArgs& a, cudaStream_t stream)
{
    if (a.ngroups > 0) {  // merged multi-shape launch
        if (a.ngroups > 4) return cudaErrorInvalidValue;
        if (a.kblock == 32)
            return a.epi_split == 1 ? launch_nmskm<NREAL, MODE, true, 32, true>(in_map, a, stream) : launch_nmskm<NREAL, MODE, false, 32, true>(in_map, a, stream);
        return a.epi_split == 1 ? launch_nmskm<NREAL, MODE, true, 64, true>(in_map, a, stream) : launch_nmskm<NREAL, MODE, false, 64, true>(in_map, a, stream);
    }
    if (a.kblock == 32)
        return a.epi_split == 1 ? launch_nmskm<NREAL, MODE, true, 32, false>(in_map, a, stream) : launch_nmskm<NREAL, MODE, false, 32, false>(in_map, a, stream);
    return a.epi_split == 1 ? launch_nmskm<NREAL, MODE, true, 64, false>(in_map, a, stream) : launch_nmskm<NREAL, MODE, false, 64, false>(in_map, a, stream);
}

}  // namespace

cudaError_t launch_conv3x3_tc2x(const ConvMaps* in_map, const ConvArgs& a, int cout_pad, cudaStream_t stream)
{
    if (cout_pad == 32 && a.mode == CONV_STORE16) return launch_nm<32, CONV_STORE16>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_STORE16) return launch_nm<64, CONV_STORE16>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_RDB_OUT) return launch_nm<64, CONV_RDB_OUT>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_RDB_OUT16) return launch_nm<64, CONV_RDB_OUT16>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_TRUNK) return launch_nm<64, CONV_TRUNK>(in_map, a, stream);
    if (cout_pad == 16 && a.mode == CONV_LAST_U8) return launch_nm<16, CONV_LAST_U8>(in_map, a, stream);
    if (cout_pad == 16 && a.mode == CONV_LAST_F32) return launch_nm<16, CONV_LAST_F32>(in_map, a, stream);
    return cudaErrorInvalidValue;
}

}  // namespace rsr
