// conv3x3.cu -- the 3x3 stride-1 pad-1 convolution of the RRDB net (reference arithmetic: ncnn layer/convolution.cpp:133-204,
// activation fused_activation.h:31-36) as an implicit GEMM on B200 CTA PAIRS: tcgen05.mma cta_group::2, TMA-staged
// NHWC boxes, accumulators in TMEM, everything that follows the conv in the graph (bias, LeakyReLU, Concat slice, Eltwise /
// BinaryOp residuals, post-processing) fused into the epilogue.  This is the only tensor-core kernel of the engine.
//
// GEMM view: M = pixels, N = Cout x 3 stacked taps (see below), K = 9 * Cin; fp16 x fp16 -> fp32.
//   * Work tile = 30 columns x 16 rows (Cout <= 32) or 8 rows (Cout = 64) of output pixels, owned by a cluster of 2 CTAs; CTA r
//     computes its half of the rows as M-tiles of 128 "positions" (4 rows x pitch 32): two M-tiles (Cout <= 32) or one (Cout = 64,
//     whose N = 192 accumulators need 192 TMEM columns per M-tile: two accumulator sets = 384 of the 512 columns).  One
//     cta_group::2 MMA (M = 256) covers one M-tile in each CTA.
//   * A operand: per 32-channel K-block one haloed TMA box {32 ch, 32 px, 10 or 6 rows} (out-of-bounds = zero fill = the conv's
//     padding); K-block k of a conv's input has its own tensor map (ConvMaps::m[k]: the two halves of the 64-channel x0 plane, then
//     one 32-channel plane per growth tensor x1..x4), so a conv reads exactly the channels it needs.  With a fixed pitch of 32 px
//     the A rows of tap (ky,kx) for 128 consecutive positions are the 128 consecutive smem rows starting at q + 32*ky + kx:
//     row-shifted UMMA descriptors instead of im2col (SWIZZLE_64B is a function of the absolute smem address, so any 64-B aligned
//     start works).
//   * Horizontal tap stacking keeps the MMAs wide (an MMA re-reads its 4 KB A tile from smem, so narrow N is feed-bound,
//     tools/ubench2.cu): the three kx taps of a kernel row are stacked along N -- Cout=32: one N=96 MMA per k-step, Cout=64: one
//     N=192 MMA, Cout=3 (padded 16): N=64.  Column group g of position q then holds tap (ky,g)'s contribution to the output at
//     q-g; the 32 lanes of an epilogue warp are 32 consecutive positions of one tile row, so the epilogue fetches it with
//     __shfl_down.
//   * B operand: each CTA keeps ITS half of the conv's weights resident in smem for the whole launch (B is split N/2 + N/2
//     between the two CTAs of a pair), fetched once by TMA before the grid dependency is resolved.
//   * Warp roles (12 warps): warp 0 A producer (TMA), warp 1 MMA issuer (leader CTA), warp 2 weight fetch, warps 4..11
//     epilogue.  setmaxnreg moves registers from the control warpgroup to the epilogue warpgroups.
//   * Epilogues, all with EARLY accumulator hand-back (pull the columns into registers, combine the taps, return the set to the
//     MMA issuer, only then do the math and the stores; per-tile timestamps of the round-1 scheme showed the chain "MMAs done ->
//     whole epilogue -> hand-back" bounding the Cout=32 convs at 2.6 k clk per tile where their MMAs need 1.2 k):
//       CONV_STORE16 (bias + LeakyReLU: 4 of the 5 convs of a dense block, the tail): one (M-tile, row) unit per warp; the row is
//         packed to fp16 into a swizzled staging buffer and leaves through ONE TMA store.
//       CONV_RDB (conv5 of a dense block): the residual stream (fp16 + int8 extension, kernels.h) ALSO moves by TMA: the operand
//         row boxes arrive in the warp's staging buffer while the MMAs run, every lane works on its own pixel (the TMEM-native
//         layout), the result row and its extension leave through two TMA stores.  No per-thread global access.
//       CONV_TRUNK / last conv (one launch per frame each): two 4-warp groups alternate tiles; fp32 operands go through a per-warp
//         smem transpose so that global accesses are 64-B runs.
//   * Tile geometry is walked incrementally (add-with-carry per tile; integer divisions only when a merged launch enters a
//     new shape group); every second conv of a job walks its tiles backwards (ConvArgs::reverse).
//   * The MMA issuer's loop carries no divisions, no tile geometry and no 64-bit descriptor arithmetic: stage index and
//     barrier parities are counters, descriptors are {lo32 = base + immediate, hi32 = constant}.
//   * Merged multi-shape launches: one launch per layer covers up to four same-shape tile groups (ConvArgs::grp), so a frame
//     pays the per-launch fill/drain once per layer instead of once per layer and shape.
// Synchronisation: "full" barriers live in the leader CTA and collect the TMA bytes of both CTAs (cp.async.bulk.tensor
// .cta_group::2 signalling a shared::cluster mbarrier); "stage free" and "accumulators ready" are signalled in both CTAs at
// once by tcgen05.commit ... multicast::cluster; the peer's epilogue warps hand an accumulator set back with a remote
// (CTA-scope release) mbarrier.arrive.  Programmatic dependent launch overlaps the prologue with the previous conv's tail.
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace rsr {

namespace {

constexpr int P = CONV_P;
constexpr int BW = P - 2;
constexpr int KB = 32;                        // K-block: input channels per A stage / per resident weight block (SWIZZLE_64B boxes)
// Work-tile height per CTA: 8 rows (two M-tiles) for the N = 96 / 64 convs, 4 rows (one M-tile) for the Cout = 64 convs, whose
// single N = 192 MMA per k-step needs 192 TMEM columns per M-tile (two accumulator sets = 384 of the 512 columns).
__host__ __device__ constexpr int bh_of(int nreal) { return conv_bh(nreal); }
__host__ __device__ constexpr int mma_n(int nreal) { return nreal == 64 ? 192 : (nreal == 32 ? 96 : 64); }   // three stacked kx taps
__host__ __device__ constexpr int dcols_of(int nreal) { return nreal == 64 ? 192 : 128; }                    // TMEM columns per M-tile
constexpr int NA_MAX = 8;                     // A stages (run-time count: as many as fit next to the resident weights)
__host__ __device__ constexpr int a_bytes(int bh) { return (bh + 2) * P * KB * 2; }   // 20 KB (8 rows) / 12 KB (4 rows)
constexpr int A_SLACK = 512;                  // zeroed rows behind the box (kx taps of the last, never stored, positions)
__host__ __device__ constexpr int a_stage(int bh) { return a_bytes(bh) + A_SLACK; }
constexpr int NEPI = 8;                       // epilogue warps
// Build variant -DRSR_EARLY_HANDBACK (librealsr_b200_early.so, NOT the default and never run on hardware so far): the epilogue returns
// the accumulator set right after its last tcgen05.ld, BEFORE the tap-combining shuffles (~570 clk of the SM's shuffle unit per tile that
// sit on the chain  MMAs done -> epilogue reads -> hand-back -> next MMAs of this set; matters for the convs with few K-blocks).  A/B it
// with tools/ab_builds.py once a GPU is at hand.  Without the macro the preprocessed source -- and the SASS -- are those of the
// validated kernel.
#ifdef RSR_EARLY_HANDBACK
#define RSR_HAND_BACK(bar_addr)                               \
    do {                                                      \
        tc_fence_before();                                    \
        __syncwarp();                                         \
        if (lane == 0) mbar_arrive_remote(bar_addr);          \
        if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();          \
    } while (0)
#endif
constexpr int EPI_WARP0 = 4;                  // first epilogue warp: warps 0..3 are the control warpgroup
constexpr int NTHREADS = (EPI_WARP0 + NEPI) * 32;
// epilogue staging per warp: CONV_STORE16 stages one 32-pixel row of its output as a dense, swizzled TMA-store source
// (32 x Cout fp16); CONV_RDB keeps one row of the residual stream there (32 px x 64 fp16 + 32 px x 64 int8: TMA-loaded input, overwritten
// in place by the output row); CONV_TRUNK transposes 32 px x 32 ch fp32 through a padded tile (row stride 36 words); the last conv stores
// straight from registers
__host__ __device__ constexpr int stage_bytes(int nreal, int mode)
{
    return mode == CONV_STORE16 ? 32 * nreal * 2 : (mode == CONV_RDB ? 32 * 128 + 32 * 64 : (nreal == 64 ? 32 * 36 * 4 : 1024));
}
constexpr int TMEM_COLS = 512;
constexpr int SMEM_MAX = 227 * 1024 - 1024;   // dynamic part; the kernel also has ~300 B of static shared memory
__host__ __device__ constexpr int smem_fixed(int nreal, int mode) { return 1024 /*align*/ + NEPI * stage_bytes(nreal, mode) + 256 /*barriers*/; }
constexpr int REGS_CTRL = 64, REGS_EPI = 216; // setmaxnreg: 128 x 64 + 256 x 216 <= 64 K registers

// this CTA's half (N/2 rows of 32 channels) of the weights of one (K-block, ky): Cout=64 -> 96 rows, Cout<=32 -> 48 rows (padded
// to 3072 B so that every block stays 512-B aligned for SWIZZLE_64B), Cout<=16 -> 32 rows
__host__ __device__ constexpr int w_bytes_per_ky(int nreal) { return nreal == 64 ? 6144 : (nreal == 32 ? 3072 : 2048); }

struct Bars {
    uint64_t a_full[NA_MAX], a_empty[NA_MAX];
    uint64_t w_full;
    uint64_t acc_full[2], acc_empty[2];
    uint64_t epi_ld[NEPI];  // CONV_RDB: one per epilogue warp, completion of the TMA loads of its residual rows
    uint32_t tmem_ptr;
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// tcgen05.mma cta_group::2 with the descriptors given as {lo32, hi32} halves and an issue predicate inside the asm block: the
// surrounding loop has no lane-dependent control flow at all, only the instruction itself is predicated on the elected lane
__device__ __forceinline__ void umma2_issue(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate,
                                            uint32_t issue)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(issue)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_if(uint64_t* bar, uint16_t mask, uint32_t issue)
{
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(smem_u32(bar)), "h"(mask), "r"(issue)
        : "memory");
}

// Incremental walk over the work tiles of one warp role: logical indices u0, u0 + stride, ... < ntiles, physical tile t = u, or
// t = ntiles-1-u when the launch walks backwards (ConvArgs::reverse: the engine alternates the direction from one conv to the next,
// so that a conv starts on the pixels the previous conv touched last -- the part of its inputs that is still in the 126 MB L2; the
// 32-output-channel convs run at 5.3-6.2 TB/s of DRAM traffic, i.e. they are HBM-bound).  A tile index decodes to (group g, image b
// of the group, band by, column tile tx); the per-tile path is three add-with-carry steps, the integer divisions run only when the
// walk enters a new shape group (merged launches: at most four times per launch).
struct Walk {
    int u, t, gstart, gend, g;
    int b, by, tx;
    int tiles_x, bands_y, H, W, desc0;
    int sb, sby, stx;  // the stride decomposed in this group's (image, band, column) grid
    size_t pxoff, aux;
};
__device__ __forceinline__ void walk_seek(Walk& w, const ConvArgs& a, const int u, const int stride, const int ntiles, const int tiles_x0, const int bands_y0)
{
    int t0 = 0, tpi;
    const int t = a.reverse ? ntiles - 1 - u : u;
    w.u = u;
    w.t = t;
    if (a.ngroups > 0) {
        int g = 0;
#pragma unroll
        for (int k = 1; k < 4; k++)
            if (k < a.ngroups && t >= a.grp[k].tile_start) g = k;
        // explicit selects: a dynamically indexed kernel parameter would be copied to local memory
        w.H = g == 0 ? a.grp[0].H : g == 1 ? a.grp[1].H : g == 2 ? a.grp[2].H : a.grp[3].H;
        w.W = g == 0 ? a.grp[0].W : g == 1 ? a.grp[1].W : g == 2 ? a.grp[2].W : a.grp[3].W;
        w.tiles_x = g == 0 ? a.grp[0].tiles_x : g == 1 ? a.grp[1].tiles_x : g == 2 ? a.grp[2].tiles_x : a.grp[3].tiles_x;
        tpi = g == 0 ? a.grp[0].tiles_per_img : g == 1 ? a.grp[1].tiles_per_img : g == 2 ? a.grp[2].tiles_per_img : a.grp[3].tiles_per_img;
        t0 = g == 0 ? a.grp[0].tile_start : g == 1 ? a.grp[1].tile_start : g == 2 ? a.grp[2].tile_start : a.grp[3].tile_start;
        w.desc0 = g == 0 ? a.grp[0].desc0 : g == 1 ? a.grp[1].desc0 : g == 2 ? a.grp[2].desc0 : a.grp[3].desc0;
        w.pxoff = (size_t)(g == 0 ? a.grp[0].px_off : g == 1 ? a.grp[1].px_off : g == 2 ? a.grp[2].px_off : a.grp[3].px_off);
        w.aux = (size_t)(g == 0 ? a.grp[0].aux_off : g == 1 ? a.grp[1].aux_off : g == 2 ? a.grp[2].aux_off : a.grp[3].aux_off);
        w.gend = g + 1 >= a.ngroups ? a.ntiles_multi : (g == 0 ? a.grp[1].tile_start : g == 1 ? a.grp[2].tile_start : a.grp[3].tile_start);
        w.bands_y = tpi / w.tiles_x;
        w.g = g;
    } else {
        w.H = a.H; w.W = a.W; w.tiles_x = tiles_x0; w.bands_y = bands_y0; tpi = tiles_x0 * bands_y0;
        w.desc0 = 0; w.g = 0; w.pxoff = 0; w.aux = 0; w.gend = 0x7fffffff;
    }
    w.gstart = t0;
    const int lt = t - t0;
    w.b = lt / tpi;
    int rem = lt - w.b * tpi;
    w.by = rem / w.tiles_x;
    w.tx = rem - w.by * w.tiles_x;
    w.sb = stride / tpi;
    rem = stride - w.sb * tpi;
    w.sby = rem / w.tiles_x;
    w.stx = rem - w.sby * w.tiles_x;
}
__device__ __forceinline__ void walk_start(Walk& w, const ConvArgs& a, const int u0, const int stride, const int ntiles, const int tiles_x0, const int bands_y0)
{
    if (u0 < ntiles) walk_seek(w, a, u0, stride, ntiles, tiles_x0, bands_y0);
    else w.u = ntiles;
}
__device__ __forceinline__ void walk_step(Walk& w, const ConvArgs& a, const int stride, const int ntiles, const int tiles_x0, const int bands_y0)
{
    const int u = w.u + stride;
    w.u = u;
    if (u >= ntiles) return;
    if (!a.reverse) {
        const int t = w.t + stride;
        if (t >= w.gend) { walk_seek(w, a, u, stride, ntiles, tiles_x0, bands_y0); return; }
        w.t = t;
        w.tx += w.stx;
        int c = w.tx >= w.tiles_x ? 1 : 0;
        w.tx -= c ? w.tiles_x : 0;
        w.by += w.sby + c;
        c = w.by >= w.bands_y ? 1 : 0;
        w.by -= c ? w.bands_y : 0;
        w.b += w.sb + c;
    } else {
        const int t = w.t - stride;
        if (t < w.gstart) { walk_seek(w, a, u, stride, ntiles, tiles_x0, bands_y0); return; }
        w.t = t;
        w.tx -= w.stx;
        int c = w.tx < 0 ? 1 : 0;
        w.tx += c ? w.tiles_x : 0;
        w.by -= w.sby + c;
        c = w.by < 0 ? 1 : 0;
        w.by += c ? w.bands_y : 0;
        w.b -= w.sb + c;
    }
}

template <int NREAL, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
conv3x3_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs a, const int NA)
{
    constexpr int BH = bh_of(NREAL);                 // output rows per CTA
    constexpr int BAND = 2 * BH;                     // output rows per CTA pair
    constexpr int MT = BH / 4;                       // M-tiles (128 positions = 4 rows x pitch 32) per CTA and work tile
    constexpr int DCOLS = dcols_of(NREAL);           // TMEM columns per M-tile
    constexpr int ACC_COLS = MT * DCOLS;             // per accumulator set (two sets)
    constexpr int A_BYTES = a_bytes(BH), A_STAGE = a_stage(BH);
    constexpr int ROWB = KB * 2;                     // bytes per smem row of an operand tile (one pixel / one weight row)
    constexpr int ROWS_PER_KY = NREAL == 64 ? 192 : (NREAL == 32 ? 96 : 64);  // weight rows per (plane, ky), see model.cpp
    constexpr int KSTEPS = KB / 16;
    // early accumulator hand-back (see header): the modes whose epilogue needs nothing but the accumulators
    constexpr bool EARLY = MODE == CONV_STORE16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nchunk = a.cin / KB;                   // K-blocks (Cin % 32 == 0 for every tensor-core conv of the net)
    constexpr int W_KY = w_bytes_per_ky(NREAL);
    constexpr int STAGE_BYTES = stage_bytes(NREAL, MODE);
    uint8_t* smStage = smem;                         // epilogue staging first: 1024-B aligned (SWIZZLE_128B TMA-store source)
    uint8_t* smA = smem + NEPI * STAGE_BYTES;
    uint8_t* smW = smA + NA * A_STAGE;               // resident weights: [K-block][ky] x W_KY bytes
    Bars* bars = reinterpret_cast<Bars*>(smW + nchunk * 3 * W_KY);
    __shared__ __align__(16) float s_bias[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool is_leader = rank == 0;
    const int tiles_x = (a.W + BW - 1) / BW, bands_y = (a.H + BAND - 1) / BAND;
    const int tiles_per_img = tiles_x * bands_y;
    const int ntiles = a.ngroups > 0 ? a.ntiles_multi : tiles_per_img * a.B;  // pair tiles (merged launch: per-group grids)
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NA_MAX; i++) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], 1); }
        mbar_init(&bars->w_full, 1);
        // hand-back arrivals per accumulator set: EARLY with two M-tiles: 8 warps x 2 CTAs per tile; otherwise the 4 warps of
        // one group x 2 CTAs
        for (int i = 0; i < 2; i++) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], (EARLY && MT == 2) ? 2 * NEPI : NEPI); }
        for (int i = 0; i < NEPI; i++) mbar_init(&bars->epi_ld[i], 1);
        mbar_fence_init();
        for (int i = 0; i < nchunk; i++) tma_prefetch_desc(&maps.m[i]);
        for (int g = 1; g < a.ngroups; g++)
            for (int i = 0; i < nchunk; i++) tma_prefetch_desc(&maps.mg[g - 1][i]);
        tma_prefetch_desc(&maps.w);
        if (EARLY || MODE == CONV_RDB)
            for (int g = 0; g < (a.ngroups > 0 ? a.ngroups : 1); g++) tma_prefetch_desc(&maps.o[g]);
        if (MODE == CONV_RDB)
            for (int g = 0; g < (a.ngroups > 0 ? a.ngroups : 1); g++) {
                tma_prefetch_desc(&maps.rx[g]); tma_prefetch_desc(&maps.re[g]); tma_prefetch_desc(&maps.reo[g]);
                if (a.rrdb_end) { tma_prefetch_desc(&maps.rrx[g]); tma_prefetch_desc(&maps.rre[g]); }
            }
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + NREAL) s_bias[threadIdx.x - 64] = a.bias[threadIdx.x - 64];
    // rows behind each A box: read by the kx=2 taps of the last two (never stored) positions; keep them finite
    {
        constexpr int SL16 = A_SLACK / 16;  // 32 uint4 per stage, NA <= 8: <= 256 threads
        if (threadIdx.x >= 128 && threadIdx.x < 128 + NA * SL16) {
            const int i = threadIdx.x - 128;
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

    if (warp < EPI_WARP0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
        // Resident weights: every CTA fetches ITS half of all weight blocks once (before the grid dependency is resolved --
        // weights are constants, so this overlaps the previous conv's tail).  All bytes are counted on the leader's w_full.
        if (warp == 2) {
            if (elect_one() && !(a.debug_mask & 4)) {
                const uint32_t wbar = mapa_u32(smem_u32(&bars->w_full), 0);
                if (is_leader) mbar_expect_tx(&bars->w_full, 2u * (uint32_t)(nchunk * 3 * W_KY));
                for (int c = 0; c < nchunk; c++)
                    for (int ky = 0; ky < 3; ky++) {
                        // weight rows are stored per (64-channel plane, ky) as [W(ky,0) | W(ky,1) | W(ky,2)]; a 32-channel K-block
                        // is the left or right half of a row; CTA r takes rows r*N/2 .. (r+1)*N/2-1 (maps.w box: N/2 rows)
                        const int row0 = ((c >> 1) * 3 + ky) * ROWS_PER_KY;
                        tma2_load_2d(smW + (c * 3 + ky) * W_KY, &maps.w, wbar, (c & 1) * 32, row0 + (int)rank * (mma_n(NREAL) / 2));
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
            uint32_t s = 0, ph = 1;  // stage counter and the parity to wait for on its "empty" barrier (first pass: free)
            Walk geo;
            walk_start(geo, a, cluster_id, nclusters, ntiles, tiles_x, bands_y);
            for (; geo.u < ntiles; walk_step(geo, a, nclusters, ntiles, tiles_x, bands_y)) {
                const CUtensorMap* gmaps = geo.g > 0 ? maps.mg[geo.g - 1] : maps.m;
                const int cx = geo.tx * BW - 1, cy = geo.by * BAND + (int)rank * BH - 1;
                for (int c = 0; c < nchunk; c++) {
                    mbar_wait(&bars->a_empty[s], ph);
                    if (leader) {
                        if (a.debug_mask & 2) {
                            if (is_leader) mbar_arrive(&bars->a_full[s]);
                        } else {
                            if (is_leader) mbar_expect_tx(&bars->a_full[s], 2 * A_BYTES);
                            tma2_load_4d(smA + s * A_STAGE, &gmaps[c], mapa_u32(smem_u32(&bars->a_full[s]), 0), c < 2 ? c * 32 : 0, cx, cy, geo.b);
                        }
                    }
                    if (++s == (uint32_t)NA) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1 && is_leader) {
            // ---------------- MMA issuer (leader CTA only): counters and immediates only ----------------
            const uint32_t issue = (elect_one() && !(a.debug_mask & 1)) ? 1u : 0u;
            const uint32_t commit = elect_one() ? 1u : 0u;
            constexpr uint32_t idesc = umma_idesc_f16(256, mma_n(NREAL));
            const uint32_t desc_hi = (uint32_t)(umma_desc_sw64(0, 512, 0) >> 32);
            const uint32_t a_lo0 = (smem_u32(smA) & 0x3FFFFu) >> 4, w_lo0 = (smem_u32(smW) & 0x3FFFFu) >> 4;
            uint32_t s = 0, ph = 0;       // A stage counter, parity of its "full" barrier
            uint32_t buf = 0, phe = 3;    // accumulator set, per-set parity bits to wait for on acc_empty (first use: free)
            int ti = 0;
            mbar_wait(&bars->w_full, 0);
            tc_fence_after();
            for (int t = cluster_id; t < ntiles; t += nclusters, ti++) {
                const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && commit;
                if (ts_on) a.dbg_ts[ti * 8 + 0] = clock64();
                mbar_wait(&bars->acc_empty[buf], (phe >> buf) & 1);
                phe ^= 1u << buf;
                tc_fence_after();
                if (ts_on) a.dbg_ts[ti * 8 + 1] = clock64();
                const uint32_t acc = tmem_base + buf * ACC_COLS;
                uint32_t w_lo = w_lo0;
                for (int c = 0; c < nchunk; c++) {
                    mbar_wait(&bars->a_full[s], ph);
                    tc_fence_after();
                    if (ts_on && c == 0) a.dbg_ts[ti * 8 + 2] = clock64();
                    const uint32_t a_lo = a_lo0 + s * (uint32_t)(A_STAGE >> 4);
#pragma unroll
                    for (int ky = 0; ky < 3; ky++)
#pragma unroll
                        for (int ks = 0; ks < KSTEPS; ks++) {
                            const uint32_t accum = (c | ky | ks) ? 1u : 0u;
                            const uint32_t ao = (uint32_t)((ky * P * ROWB) >> 4) + ks * 2;
                            const uint32_t bo = (uint32_t)((ky * W_KY) >> 4) + ks * 2;
#pragma unroll
                            for (int mt = 0; mt < MT; mt++)  // the M-tiles of a work tile share the B operand; M-tile mt starts 128 rows further
                                umma2_issue(acc + mt * DCOLS, a_lo + ao + mt * ((128 * ROWB) >> 4), w_lo + bo, desc_hi, idesc, accum, issue);
                        }
                    umma2_commit_if(&bars->a_empty[s], 3, commit);
                    w_lo += (uint32_t)((3 * W_KY) >> 4);
                    if (++s == (uint32_t)NA) { s = 0; ph ^= 1; }
                }
                umma2_commit_if(&bars->acc_full[buf], 3, commit);
                if (ts_on) a.dbg_ts[ti * 8 + 3] = clock64();
                buf ^= 1;
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int q4 = warp & 3;  // TMEM lane quarter of this warp = row q4 of an M-tile
        const int ew = warp - EPI_WARP0;
        const bool dbg_noepi = (a.debug_mask & 8) != 0;
      if constexpr (EARLY) {
        // ---------------- epilogue, CONV_STORE16: one (M-tile, row) unit per warp, early hand-back, TMA store ----------------
        // All 8 warps work on EVERY tile: warp (mi, q4) pulls row q4 of M-tile mi -- 32 positions x 96 / 128 accumulator columns --
        // into registers (taps combined, + bias), returns its share of the accumulator set (16 arrivals per tile), and then
        // applies the activation, packs to fp16 and writes the 32-pixel row into a swizzled staging buffer that leaves through ONE
        // TMA store (box {Cout, 30 px, 1 row}: the two garbage columns and everything outside the image are clipped by the
        // hardware).  Per-tile timestamps of the previous scheme (a group of 4 warps per tile: both M-tiles, smem transpose,
        // predicated 16-B stores) showed 1.4 k clk until hand-back + 1.2 k clk of stores + 1.1 k clk of index arithmetic per tile.
        // (Cout = 64: a work tile has one M-tile per CTA, i.e. four units; the two 4-warp groups then alternate work tiles.)
        const int grp = ew >> 2;
        const int mi = MT == 2 ? grp : 0;
        constexpr int TSTEP = MT == 2 ? 1 : 2;      // this warp works on every TSTEP-th tile of the cluster
        uint8_t* stg = smStage + ew * STAGE_BYTES;  // [32 px][NREAL fp16], 16-B chunks swizzled like the store map (64B / 128B)
        const uint32_t acc_empty0 = mapa_u32(smem_u32(&bars->acc_empty[0]), 0), acc_empty1 = mapa_u32(smem_u32(&bars->acc_empty[1]), 0);
        constexpr int ROWB_O = NREAL * 2;
        const int sw = NREAL == 32 ? ((lane >> 1) & 3) : (lane & 7);
        uint32_t phf = 0;  // parity bits to wait for on acc_full[0], acc_full[1]
        int ti = MT == 2 ? 0 : grp;
        Walk geo;
        walk_start(geo, a, cluster_id + ti * nclusters, TSTEP * nclusters, ntiles, tiles_x, bands_y);
        for (; geo.u < ntiles; walk_step(geo, a, TSTEP * nclusters, ntiles, tiles_x, bands_y), ti += TSTEP) {
            const int buf = ti & 1;
            const int y = geo.by * BAND + (int)rank * BH + mi * 4 + q4;
            const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && is_leader && ew == 0 && lane == 0;
            if (ts_on) a.dbg_ts[ti * 8 + 4] = clock64();
            mbar_wait(&bars->acc_full[buf], (phf >> buf) & 1);
            phf ^= 1u << buf;
            tc_fence_after();
            if (ts_on) a.dbg_ts[ti * 8 + 5] = clock64();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS + mi * DCOLS;
            float f[NREAL];
            if constexpr (NREAL == 32) {  // D[0:32] + right neighbour's D[32:64] + second right neighbour's D[64:96]
                uint32_t v[32], u[32], w[32];
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + 32, u);
                tmem_ld32(taddr + 64, w);
                tmem_ld_wait();
#ifdef RSR_EARLY_HANDBACK
                RSR_HAND_BACK(buf ? acc_empty1 : acc_empty0);
#endif
#pragma unroll
                for (int j = 0; j < 32; j++)
                    f[j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                           __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2) + s_bias[j];
            } else {  // D[c] + right neighbour's D[64 + c] + second right neighbour's D[128 + c], 32 channels at a time
#pragma unroll
                for (int nb = 0; nb < 2; nb++) {
                    uint32_t v[32], u[32], w[32];
                    tmem_ld32(taddr + nb * 32, v);
                    tmem_ld32(taddr + 64 + nb * 32, u);
                    tmem_ld32(taddr + 128 + nb * 32, w);
                    tmem_ld_wait();
#ifdef RSR_EARLY_HANDBACK
                    if (nb == 1) RSR_HAND_BACK(buf ? acc_empty1 : acc_empty0);
#endif
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        f[nb * 32 + j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                         __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2) + s_bias[nb * 32 + j];
                }
            }
            // Plain (CTA-scope release) remote arrive: only the TMEM reads have to be ordered (tcgen05 fence); a cluster-scope
            // release would also wait for this thread's global stores to drain.
#ifndef RSR_EARLY_HANDBACK
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(buf ? acc_empty1 : acc_empty0);
            if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
#endif
            if (a.act == 2) {
#pragma unroll
                for (int j = 0; j < NREAL; j++) f[j] = f[j] > 0.f ? f[j] : f[j] * a.slope;
            }
            if (lane == 0) bulk_wait_read<0>();  // the previous tile's TMA store has finished reading this staging buffer
            __syncwarp();
            uint8_t* row = stg + lane * ROWB_O;
#pragma unroll
            for (int j = 0; j < NREAL / 8; j++)
                *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) =
                    make_uint4(pack_h2(f[8 * j + 0], f[8 * j + 1]), pack_h2(f[8 * j + 2], f[8 * j + 3]), pack_h2(f[8 * j + 4], f[8 * j + 5]),
                               pack_h2(f[8 * j + 6], f[8 * j + 7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !dbg_noepi && y < geo.H) {
                tma_store_4d(&maps.o[geo.g], stg, 0, geo.tx * BW, y, geo.b);
                bulk_commit();
            }
            if (ts_on) a.dbg_ts[ti * 8 + 7] = clock64();
        }
        if (lane == 0) bulk_wait<0>();  // all of this warp's stores have been performed before the CTA retires
      } else if constexpr (MODE == CONV_RDB) {
        // ---------------- epilogue, CONV_RDB (conv5 of a dense block): residual stream through TMA, no transposes ----------------
        // Two 4-warp groups alternate work tiles (one per accumulator set); warp q4 owns tile row q4: lane = pixel, registers =
        // the 64 channels (the TMEM-native layout).  The row of the residual operand x0 (fp16 + int8 extension, kernels.h) arrives by
        // TMA in a swizzled staging buffer while the MMAs run, every lane reads ITS pixel's 128 + 64 bytes, computes
        // t = 0.2*(acc + b) + x0 [, t = 0.2*t + R with R fetched the same way], writes fp16(t) and its extension back over its own
        // row, and the row leaves through two TMA stores: no per-thread global access, no predicates (the hardware clips), no smem
        // transpose.  (The previous transposed-domain version cost ~4 k clk per tile row next to 3.5 k clk of MMAs.)
        const int group = ew >> 2;
        uint8_t* xb = smStage + ew * STAGE_BYTES;   // [32 px][64 fp16], SWIZZLE_128B: input x0 row, then output row
        uint8_t* eb = xb + 32 * 128;                 // [32 px][64 int8], SWIZZLE_64B: input extension row, then output extension row
        uint64_t* ldbar = &bars->epi_ld[ew];
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(&bars->acc_empty[group]), 0);
        const int buf = group;
        const int swx = lane & 7, swe = (lane >> 1) & 3;
        constexpr uint32_t ROW_BYTES = BW * 128 + BW * 64;  // TMA bytes of one fp16 row box + one int8 row box (out-of-image parts are zero-filled but counted)
        uint32_t phf = 0, phl = 0;  // parities: this group's acc_full, this warp's load barrier
        int ti = group;
        Walk geo;
        walk_start(geo, a, cluster_id + group * nclusters, 2 * nclusters, ntiles, tiles_x, bands_y);
        for (; geo.u < ntiles; walk_step(geo, a, 2 * nclusters, ntiles, tiles_x, bands_y), ti += 2) {
            const int y = geo.by * BAND + (int)rank * BH + q4, xc = geo.tx * BW;
            if (lane == 0) {  // the previous tile's stores have finished reading the staging rows: fetch this tile's x0 row
                bulk_wait_read<0>();
                mbar_expect_tx(ldbar, ROW_BYTES);
                tma_load_4d(xb, &maps.rx[geo.g], ldbar, 0, xc, y, geo.b);
                tma_load_4d(eb, &maps.re[geo.g], ldbar, 0, xc, y, geo.b);
            }
            const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && is_leader && (ew & 3) == 0 && lane == 0;
            if (ts_on) a.dbg_ts[ti * 8 + 4] = clock64();
            mbar_wait(&bars->acc_full[buf], phf);
            phf ^= 1;
            tc_fence_after();
            if (ts_on) a.dbg_ts[ti * 8 + 5] = clock64();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS;
            float t[64];
#pragma unroll
            for (int nb = 0; nb < 2; nb++) {
                uint32_t v[32], u[32], w[32];  // D[c] + right neighbour's D[64 + c] + second right neighbour's D[128 + c]
                tmem_ld32(taddr + nb * 32, v);
                tmem_ld32(taddr + 64 + nb * 32, u);
                tmem_ld32(taddr + 128 + nb * 32, w);
                tmem_ld_wait();
#ifdef RSR_EARLY_HANDBACK
                if (nb == 1) RSR_HAND_BACK(acc_empty_leader);
#endif
#pragma unroll
                for (int j = 0; j < 32; j++)
                    t[nb * 32 + j] = __uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                     __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2);
            }
#ifndef RSR_EARLY_HANDBACK
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(acc_empty_leader);  // early hand-back (CTA-scope release: only the TMEM reads have to be ordered)
            if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
#endif
            // one row of (fp16, int8) operands -> t = t * scale + operand, 8 channels per 16-B / 8-B chunk
            auto add_row = [&](const bool first) {
                mbar_wait(ldbar, phl);
                phl ^= 1;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint4 hw = *reinterpret_cast<const uint4*>(xb + lane * 128 + ((j ^ swx) << 4));
                    const uint2 ew2 = *reinterpret_cast<const uint2*>(eb + lane * 64 + (((j >> 1) ^ swe) << 4) + (j & 1) * 8);
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&hw.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&hw.y));
                    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&hw.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&hw.w));
                    const float r0 = res_ext_decode(f0.x, ew2.x, 0), r1 = res_ext_decode(f0.y, ew2.x, 1), r2 = res_ext_decode(f1.x, ew2.x, 2), r3 = res_ext_decode(f1.y, ew2.x, 3);
                    const float r4 = res_ext_decode(f2.x, ew2.y, 0), r5 = res_ext_decode(f2.y, ew2.y, 1), r6 = res_ext_decode(f3.x, ew2.y, 2), r7 = res_ext_decode(f3.y, ew2.y, 3);
                    float* tt = &t[8 * j];
                    if (first) {  // t = 0.2 * (acc + bias) + x0   (Eltwise of the dense block, x4.param:18)
                        const float4 b0 = *reinterpret_cast<const float4*>(&s_bias[8 * j]), b1 = *reinterpret_cast<const float4*>(&s_bias[8 * j + 4]);
                        tt[0] = (tt[0] + b0.x) * 0.2f + r0; tt[1] = (tt[1] + b0.y) * 0.2f + r1; tt[2] = (tt[2] + b0.z) * 0.2f + r2; tt[3] = (tt[3] + b0.w) * 0.2f + r3;
                        tt[4] = (tt[4] + b1.x) * 0.2f + r4; tt[5] = (tt[5] + b1.y) * 0.2f + r5; tt[6] = (tt[6] + b1.z) * 0.2f + r6; tt[7] = (tt[7] + b1.w) * 0.2f + r7;
                    } else {      // t = 0.2 * t + R                 (Eltwise at the end of an RRDB, x4.param:47)
                        tt[0] = tt[0] * 0.2f + r0; tt[1] = tt[1] * 0.2f + r1; tt[2] = tt[2] * 0.2f + r2; tt[3] = tt[3] * 0.2f + r3;
                        tt[4] = tt[4] * 0.2f + r4; tt[5] = tt[5] * 0.2f + r5; tt[6] = tt[6] * 0.2f + r6; tt[7] = tt[7] * 0.2f + r7;
                    }
                }
            };
            add_row(true);
            if (a.rrdb_end) {
                // the RRDB input R comes through the same staging rows: every lane is done reading x0 (generic proxy) before the TMA
                // (async proxy) overwrites them
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_expect_tx(ldbar, ROW_BYTES);
                    tma_load_4d(xb, &maps.rrx[geo.g], ldbar, 0, xc, y, geo.b);
                    tma_load_4d(eb, &maps.rre[geo.g], ldbar, 0, xc, y, geo.b);
                }
                add_row(false);
            }
            // next block's x0 = fp16(t) + int8 extension, written over this lane's own staging row
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float* tt = &t[8 * j];
                const uint32_t p0 = pack_h2(tt[0], tt[1]), p1 = pack_h2(tt[2], tt[3]), p2 = pack_h2(tt[4], tt[5]), p3 = pack_h2(tt[6], tt[7]);
                const float2 g0 = __half22float2(*reinterpret_cast<const __half2*>(&p0)), g1 = __half22float2(*reinterpret_cast<const __half2*>(&p1));
                const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&p2)), g3 = __half22float2(*reinterpret_cast<const __half2*>(&p3));
                *reinterpret_cast<uint4*>(xb + lane * 128 + ((j ^ swx) << 4)) = make_uint4(p0, p1, p2, p3);
                *reinterpret_cast<uint2*>(eb + lane * 64 + (((j >> 1) ^ swe) << 4) + (j & 1) * 8) =
                    make_uint2(res_ext_pack4(res_ext_encode(tt[0], g0.x), res_ext_encode(tt[1], g0.y), res_ext_encode(tt[2], g1.x), res_ext_encode(tt[3], g1.y)),
                               res_ext_pack4(res_ext_encode(tt[4], g2.x), res_ext_encode(tt[5], g2.y), res_ext_encode(tt[6], g3.x), res_ext_encode(tt[7], g3.y)));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !dbg_noepi && y < geo.H) {
                tma_store_4d(&maps.o[geo.g], xb, 0, xc, y, geo.b);
                tma_store_4d(&maps.reo[geo.g], eb, 0, xc, y, geo.b);
                bulk_commit();
            }
            if (ts_on) a.dbg_ts[ti * 8 + 7] = clock64();
        }
        if (lane == 0) bulk_wait<0>();  // all of this warp's stores have been performed before the CTA retires
      } else {
        // ---------------- epilogue, CONV_TRUNK and the last conv: warps 4..11 in two groups of four ----------------
        // Group g = (warp-4)/4 owns accumulator set g, i.e. every second work tile; inside a group warp q4 handles row q4 of the
        // tile's M-tile(s).  (One launch each per frame; the last conv is HBM-bound.)
        const int group = ew >> 2;
        uint32_t* stg = reinterpret_cast<uint32_t*>(smStage + ew * STAGE_BYTES);  // [32 px][36 words]
        const int tr = lane >> 2, tc = lane & 3;
        const uint32_t acc_empty_leader = mapa_u32(smem_u32(&bars->acc_empty[group]), 0);
        const int buf = group;
        uint32_t phf = 0;  // parity to wait for on this group's acc_full
        int ti = group;
        Walk geo;
        walk_start(geo, a, cluster_id + group * nclusters, 2 * nclusters, ntiles, tiles_x, bands_y);
        for (; geo.u < ntiles; walk_step(geo, a, 2 * nclusters, ntiles, tiles_x, bands_y), ti += 2) {
            const int b = geo.b, by = geo.by, tx = geo.tx;
            const int x0 = tx * BW;
            int yy[MT];
            size_t rowpix[MT];
            bool vrow[MT];
#pragma unroll
            for (int i = 0; i < MT; i++) {
                yy[i] = by * BAND + (int)rank * BH + i * 4 + q4;
                vrow[i] = yy[i] < geo.H && !dbg_noepi;
                rowpix[i] = geo.pxoff + ((size_t)b * geo.H + yy[i]) * geo.W + x0;
            }
            bool vcol[4];
#pragma unroll
            for (int r = 0; r < 4; r++) vcol[r] = (tr + 8 * r < BW) && (x0 + tr + 8 * r < geo.W);
            // CONV_TRUNK (acc + b + F32): the fp32 math runs in a TRANSPOSED domain: a 32-channel block of the tile row goes through a
            // [32 px][32 ch] smem tile, after which thread (tr, tc) holds 8 consecutive channels (tc*8..) of the four pixels tr, tr+8,
            // tr+16, tr+24 -- every global access is then 32 B of fp32 / 16 B of fp16 per lane and a full 128-B / 64-B run per pixel.
            // The fp32 operands are fetched one 32-channel block ahead of use.
            float4 res[4][2];             // [r][8 channels] fp32 residual operands of the current block
            auto load_res = [&](int i, int nb, float4 (&dst)[4][2]) {
                if constexpr (MODE == CONV_TRUNK) {
                    const float* base = a.f32 + rowpix[i] * 64 + nb * 32 + tc * 8;
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (vcol[r] && vrow[i]) {
                            dst[r][0] = *reinterpret_cast<const float4*>(base + (size_t)(tr + 8 * r) * 64);
                            dst[r][1] = *reinterpret_cast<const float4*>(base + (size_t)(tr + 8 * r) * 64 + 4);
                        }
                }
            };
            if constexpr (NREAL == 64) load_res(0, 0, res);
            const bool ts_on = a.dbg_ts != nullptr && cluster_id == 0 && ti < 64 && is_leader && (ew & 3) == 0 && lane == 0;
            if (ts_on) a.dbg_ts[ti * 8 + 4] = clock64();
            mbar_wait(&bars->acc_full[buf], phf);
            phf ^= 1;
            tc_fence_after();
            if (ts_on) a.dbg_ts[ti * 8 + 5] = clock64();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * ACC_COLS;
            // Cout = 64 (trunk): the accumulators go to registers first (taps combined) and the set returns to the MMA issuer before
            // any residual math or store
            uint32_t accv[NREAL == 64 ? MT : 1][NREAL == 64 ? 64 : 1];
            if constexpr (NREAL == 64) {
#pragma unroll
                for (int i = 0; i < MT; i++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        uint32_t v[32], u[32], w[32];  // D[c] + right neighbour's D[64 + c] + second right neighbour's D[128 + c]
                        tmem_ld32(taddr0 + i * DCOLS + nb * 32, v);
                        tmem_ld32(taddr0 + i * DCOLS + 64 + nb * 32, u);
                        tmem_ld32(taddr0 + i * DCOLS + 128 + nb * 32, w);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            accv[i][nb * 32 + j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(u[j]), 1) +
                                                                   __shfl_down_sync(0xffffffffu, __uint_as_float(w[j]), 2));
                    }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(acc_empty_leader);  // CTA-scope release: only the TMEM reads have to be ordered
                if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
            }

            {
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    const uint32_t taddr = taddr0 + i * DCOLS;
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
                    } else {
                        // CONV_TRUNK (Cout = 64): fp32 add of `fea` in the transposed domain
#pragma unroll
                        for (int nb = 0; nb < 2; nb++) {
                            const uint32_t* v = &accv[i][nb * 32];
                            float4 nxt[4][2];
                            {  // prefetch the operands of the next (M-tile, 32-channel block)
                                const int ni = nb == 1 ? i + 1 : i, nnb = nb ^ 1;
                                if (ni < MT) load_res(ni, nnb, nxt);
                            }
#pragma unroll
                            for (int k = 0; k < 8; k++)  // this lane's pixel, 32 channels -> row `lane` of the tile (stride 36 words: conflict-free)
                                *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * k) = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                            __syncwarp();
                            const int ch0 = nb * 32 + tc * 8;
                            const float4 bb0 = *reinterpret_cast<const float4*>(&s_bias[ch0]), bb1 = *reinterpret_cast<const float4*>(&s_bias[ch0 + 4]);
                            float4 acc8[4][2];
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                acc8[r][0] = *reinterpret_cast<const float4*>(stg + (tr + 8 * r) * 36 + 8 * tc);
                                acc8[r][1] = *reinterpret_cast<const float4*>(stg + (tr + 8 * r) * 36 + 8 * tc + 4);
                            }
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                float4 o[2];
#pragma unroll
                                for (int q = 0; q < 2; q++) {
                                    const float4 ac = acc8[r][q], bb = q ? bb1 : bb0, rs = res[r][q];
                                    o[q].x = ac.x + bb.x + rs.x; o[q].y = ac.y + bb.y + rs.y;
                                    o[q].z = ac.z + bb.z + rs.z; o[q].w = ac.w + bb.w + rs.w;
                                }
                                if (vcol[r] && vrow[i]) {
                                    const size_t px = rowpix[i] + tr + 8 * r;
                                    *reinterpret_cast<uint4*>(a.out16 + px * a.out_cs + a.out_co + ch0) =
                                        make_uint4(pack_h2(o[0].x, o[0].y), pack_h2(o[0].z, o[0].w), pack_h2(o[1].x, o[1].y), pack_h2(o[1].z, o[1].w));
                                }
                            }
#pragma unroll
                            for (int r = 0; r < 4; r++) { res[r][0] = nxt[r][0]; res[r][1] = nxt[r][1]; }
                        }
                    }
                }
                if constexpr (NREAL != 64) {  // accumulator set drained by this warp: tell the leader's MMA issuer (CTA-scope release)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(acc_empty_leader);
                    if (ts_on) a.dbg_ts[ti * 8 + 6] = clock64();
                } else if (ts_on) {
                    a.dbg_ts[ti * 8 + 7] = clock64();
                }
            }
        }
      }
    }

    tc_fence_before();
    cluster_sync();  // nobody frees TMEM / exits while the peer may still signal or be read
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <int NREAL, int MODE>
cudaError_t launch_nm(const ConvMaps* in_map, const ConvArgs& a, cudaStream_t stream)
{
    static bool attr_done[64] = {};
    static int num_sms[64] = {};
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    dev &= 63;
    if (!attr_done[dev]) {
        e = cudaFuncSetAttribute(conv3x3_kernel<NREAL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
        if (e != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    constexpr int BAND = 2 * bh_of(NREAL);
    if (a.ngroups < 0 || a.ngroups > 4 || (a.cin % KB) != 0) return cudaErrorInvalidValue;
    const long long ntiles = a.ngroups > 0 ? (long long)a.ntiles_multi : (long long)((a.W + BW - 1) / BW) * ((a.H + BAND - 1) / BAND) * a.B;
    if (ntiles <= 0) return cudaErrorInvalidValue;
    long long nclusters = num_sms[dev] / 2;
    if (ntiles < nclusters) nclusters = ntiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * nclusters));
    cfg.blockDim = dim3(NTHREADS);
    // resident weights first, then as many A stages as fit (20.5 KB each for 8-row tiles, 12.5 KB for 4-row tiles; at most 8)
    const int wbytes = (a.cin / KB) * 3 * w_bytes_per_ky(NREAL);
    int na = (SMEM_MAX - smem_fixed(NREAL, MODE) - wbytes) / a_stage(bh_of(NREAL));
    if (na > NA_MAX) na = NA_MAX;
    if (na < 2) return cudaErrorInvalidValue;
    cfg.dynamicSmemBytes = smem_fixed(NREAL, MODE) + wbytes + na * a_stage(bh_of(NREAL));
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const int use_pdl = getenv("RSR_NOPDL") ? 0 : 1;  // bring-up switch
    cfg.numAttrs = use_pdl;
    if ((e = cudaLaunchKernelEx(&cfg, conv3x3_kernel<NREAL, MODE>, *in_map, a, na)) != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_conv3x3(const ConvMaps* in_map, const ConvArgs& a, int cout_pad, cudaStream_t stream)
{
    if (cout_pad == 32 && a.mode == CONV_STORE16) return launch_nm<32, CONV_STORE16>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_STORE16) return launch_nm<64, CONV_STORE16>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_RDB) return launch_nm<64, CONV_RDB>(in_map, a, stream);
    if (cout_pad == 64 && a.mode == CONV_TRUNK) return launch_nm<64, CONV_TRUNK>(in_map, a, stream);
    if (cout_pad == 16 && a.mode == CONV_LAST_U8) return launch_nm<16, CONV_LAST_U8>(in_map, a, stream);
    if (cout_pad == 16 && a.mode == CONV_LAST_F32) return launch_nm<16, CONV_LAST_F32>(in_map, a, stream);
    return cudaErrorInvalidValue;
}

}  // namespace rsr
