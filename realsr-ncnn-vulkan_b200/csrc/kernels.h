// kernels.h -- device-side interface of the RealSR B200 engine (all launches are asynchronous on `stream`).
//
// Data layout in HBM (per batch of same-shape padded tiles, B tiles of H x W network-input pixels):
//   D16[2]  fp16 [3][B][H][W][64] dense-block buffer as three 64-channel planes: plane 0 = RDB input x0, plane 1 =
//                                 [x1,x2], plane 2 = [x3,x4].  Concat (ncnn layer/concat.cpp:140-175) is never
//                                 materialised: conv k of an RDB reads planes 0..ceil((64+32(k-1))/64)-1 and writes its
//                                 32 outputs into half a plane.  Planar (not 192-interleaved) so that one row of a TMA
//                                 box is 32 px x 128 B = 4 KB contiguous in DRAM.
//   T32     fp32 [B][H][W][64]   residual stream (RDB input, exact fp32), updated in place by conv5's epilogue
//   R32     fp32 [B][H][W][64]   RRDB input, updated in place at every third RDB
//   F32     fp32 [B][H][W][64]   Conv_0 output `fea` (added back after the trunk conv, x4.param:994-995)
//   U/V     fp16 [B][sH][sW][64] activations of the 2x / 4x tail
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsr {

// Geometry of one network input ("padded tile", reference src/realsr.cpp:545-553,:757-766) inside the source image
// and of its cropped result inside the output image.
struct TileDesc {
    int in_x0, in_y0;    // image coords of padded-tile pixel (0,0) of TTA variant 0 (may be < 0: reflect-101)
    int out_x0, out_y0;  // output-image coords of the cropped result
    int crop_w, crop_h;  // size of the cropped result (tile_w_nopad*4, tile_h_nopad*4)
    int variant;         // TTA dihedral variant 0..7 (0 when TTA is off), src/realsr.cpp:618-664
    int base_h, base_w;  // padded-tile size of variant 0 (variants 4..7 run transposed)
    int tta_slot;        // index of the fp32 result plane this tile writes in TTA mode
};

enum ConvMode {
    CONV_STORE16 = 0,   // out = fp16(act(acc + b)) through ConvMaps::o (TMA store)
    CONV_RDB = 1,       // t = 0.2*(acc+b) + x0 ; [rrdb_end: t = 0.2*t + R] ; next x0 = fp16(t) + int8 extension (see res_ext_encode)
    CONV_TRUNK = 2,     // out16 = fp16(acc + b + F32)
    CONV_LAST_U8 = 3,   // v = acc+b -> crop, v*255+0.5, (int) trunc, clamp -> u8 HWC output image
    CONV_LAST_F32 = 4,  // v = acc+b -> fp32 [slot][4H][4W][4] (TTA; averaged by tta_gather)
};


// One same-shape group of padded tiles inside a merged multi-shape launch (ConvArgs::ngroups > 0): the groups' activations
// are concatenated pixel-major in every buffer, group g starting px_off pixels into it.
struct ConvGroup {
    int H, W, B;                 // this conv's input == output size of the group, tiles in the group
    int tiles_x, tiles_per_img;  // work-tile grid of one image of the group (pair tiles of 30 x 2*conv_bh(cout_pad) px)
    int tile_start;              // first work tile of the group in the launch-wide numbering
    int desc0;                   // index of the group's first TileDesc in ConvArgs::tiles
    long long px_off;            // pixel offset of the group inside the activation buffers (already scaled for 2x / 4x layers)
    long long aux_off;           // CONV_LAST_F32: float4 offset of the group's first TTA slot inside ConvArgs::tta_out
};

struct ConvArgs {
    const float* bias;  // [N]
    int cin;            // real input channels (multiple of 16)
    int H, W, B;        // spatial size of this conv's input == output, batch
    int mode, act;      // ConvMode; act: 0 none, 2 leaky relu
    float slope;
    __half* out16;
    int out_cs, out_co;  // channels per pixel of the output buffer, channel offset (residual modes; CONV_STORE16 goes through ConvMaps::o)
    const float* f32;     // CONV_TRUNK: Conv_0's output `fea` (BinaryOp ADD, x4.param:994-995)
    // CONV_RDB: the dense-block residual stream.  A residual operand travels as the fp16 plane that is also the next block's MMA
    // operand plus an int8 extension plane [px][64] that counts 1/256 of the fp16 value's ulp (res_ext_encode below): 8 more
    // mantissa bits for 64 B/px (an fp32 copy is 256 B/px; without the extension the README-crop golden has 3.2 % of its bytes off
    // by one, outside the 97 % contract).  res16/ext_in = this block's input x0; rr16/rr_ext = the input of the enclosing RRDB (read
    // only when rrdb_end: t = 0.2*t + R, Eltwise x4.param:47); out16/ext_out = the next block's x0.  The kernel reaches all of them
    // through the row maps ConvMaps::rx / re / rrx / rre / o / reo (TMA); the pointers below name the same planes for the host side.
    const __half* res16;
    const uint8_t* ext_in;
    const __half* rr16;
    const uint8_t* rr_ext;
    uint8_t* ext_out;
    int rrdb_end;
    // CONV_LAST_*
    const TileDesc* tiles;
    uint8_t* img_out;
    int img_w, img_h;    // output image size (pixels)
    int img_c;           // bytes per pixel of the output image (3 or 4)
    int prepad_s;        // prepadding*scale
    float* tta_out;
    long long* dbg_ts;   // optional (bring-up): per-tile clock64 stamps of cluster 0, 8 slots per tile, see conv3x3.cu
    // Merged multi-shape launch (engine option "merge_shapes", default on): up to 4 same-shape tile groups in ONE launch per layer,
    // so that a frame pays the per-launch pipeline fill/drain once per layer instead of once per layer and shape.
    // ngroups == 0: plain launch described by H, W, B above.
    int ngroups;
    int ntiles_multi;     // total work tiles of the merged launch
    ConvGroup grp[4];
    int reverse;          // walk the work tiles from last to first (the engine alternates it from conv to conv: L2 reuse, see conv3x3.cu Walk)
    int debug_mask;       // debug mask for timing experiments (0 in production): 1 no MMA, 2 no A loads, 4 no B loads, 8 no epilogue I/O
};

// Tensor maps of one conv launch.  Activations: one map per 32-channel K-block of the conv's input, box {32 ch, 32 px,
// conv_bh(cout_pad) + 2 rows, 1}, SWIZZLE_64B.  K-blocks 0 and 1 are the two halves of a 64-channel plane (the kernel passes channel
// coordinate 32k), K-blocks >= 2 are 32-channel planes of their own (the dense-block growth channels x1..x4, see engine.cu JobMaps).
constexpr int CONV_MAX_KBLOCKS = 6;  // Cin <= 192
struct alignas(64) ConvMaps {
    CUtensorMap m[CONV_MAX_KBLOCKS];
    CUtensorMap w;         // 2-D map over the conv's weight rows [rows][64 ch], box {32 ch, N/2 rows} (N = 192 / 96 / 64), SWIZZLE_64B
    CUtensorMap mg[3][CONV_MAX_KBLOCKS];  // merged multi-shape launch: the K-block maps of groups 1..3 (group 0 uses m[])
    CUtensorMap o[4];      // CONV_STORE16 / CONV_RDB: the OUTPUT plane of group g, box {cout_pad ch, 30 px, 1 row}, SWIZZLE_64B (Cout=32:
                           // a 32-channel plane) / SWIZZLE_128B (Cout=64: a 64-channel plane): the epilogue leaves through TMA stores
    // CONV_RDB: row boxes {64, 30 px, 1 row} of the residual stream of group g (ConvArgs::res16 ...): rx / re = this block's input x0
    // (fp16, SWIZZLE_128B) and its int8 extension (SWIZZLE_64B), rrx / rre = the enclosing RRDB's input, reo = the output extension
    CUtensorMap rx[4], re[4], rrx[4], rre[4], reo[4];
};
// conv3x3.cu: conv3x3 s1 p1 (+ fused epilogue, see ConvMode) as implicit GEMM on tcgen05 CTA pairs.  cout_pad in {16, 32, 64}.
cudaError_t launch_conv3x3(const ConvMaps* in_maps, const ConvArgs& a, int cout_pad, cudaStream_t stream);

// Residual carrier (ConvArgs::res16 / ext_in ...): a value v travels as its round-to-nearest fp16 h (the MMA operand) plus an int8 q that
// counts 1/256 of h's ulp: q = floor((bits(v) - bits(float(h))) / 32) on the fp32 bit patterns (13 more mantissa bits in fp32 than in
// fp16, 8 of them kept; bit-pattern arithmetic is monotonic across binades and sign-symmetric), in [-128, 127] (the one value that can
// exceed it, the exact round-to-even tie +4096, saturates).  Decoding adds q * 32 fp32-ulps back: |error| <= 1/256 of an fp16 ulp.
// Below the fp16 normal range q may wrap; the absolute error there stays under 2^-25 either way.
__device__ __forceinline__ int res_ext_encode(float v, float hf) { return min((__float_as_int(v) - __float_as_int(hf)) >> 5, 127); }
// q = byte `byte` (0..3) of the packed word e, sign-extended
__device__ __forceinline__ float res_ext_decode(float hf, unsigned e, int byte)
{
    int q;
    asm("bfe.s32 %0, %1, %2, 8;" : "=r"(q) : "r"(e), "r"(byte * 8));
    return __int_as_float(__float_as_int(hf) + q * 32);
}
// the low bytes of four encoded values -> one word (byte i = value i)
__device__ __forceinline__ unsigned res_ext_pack4(int a, int b, int c, int d)
{
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}
constexpr int CONV_P = 32;      // position pitch of a work tile (32 px: 30 output columns + the two halo columns)
constexpr int CONV_BW = 30;     // output columns per work tile
// output rows per CTA of a pair and work tile: 4 for the Cout = 64 convs (one N = 192 MMA per k-step: 192 TMEM columns per M-tile),
// 8 otherwise; a work tile of the pair is 2 x conv_bh rows, the activation boxes have conv_bh + 2 rows
__host__ __device__ constexpr int conv_bh(int cout_pad) { return cout_pad == 64 ? 4 : 8; }

// Conv_0 (3->64, fp32 CUDA cores) fused with preproc (crop, reflect-101, x 1/255, TTA variant), conv_first.cu
struct FirstArgs {
    const uint8_t* img;  // u8 HWC RGB source image on the device
    int img_w, img_h;
    int img_c;           // bytes per pixel of the source image: 3 (RGB) or 4 (RGBA; alpha is ignored here)
    const TileDesc* tiles;
    const float* w;      // [64][3][3][3] fp32 (OIHW as in the .bin), bias [64]
    const float* bias;
    int H, W, B;
    __half* d16;         // x0 of the first dense block (64 ch per pixel)
    uint8_t* ext;        // its int8 extension [px][64] (see ConvArgs::ext_in)
    float* f32;          // fp32 copy of Conv_0's output for the trunk add
};
cudaError_t launch_conv_first(const FirstArgs& a, cudaStream_t stream);

// nearest x2 (ncnn layer/interp.cpp:598-618): NHWC fp16, reads channels 0..63 of a source with in_cs channels per
// pixel, writes a dense 64-ch tensor
cudaError_t launch_upsample2x(const __half* in, int in_cs, __half* out, int B, int H, int W, cudaStream_t stream);

// TTA gather: average the 8 inverse-transformed fp32 results, x255 + 0.5, truncate, clamp (src/realsr.cpp:691-724)
struct TtaGatherArgs {
    const float* tta_out;  // [slots][4*H0][4*W0][4] per variant, see conv_tc.cu CONV_LAST_F32
    const TileDesc* tiles; // the 8 descs of every tile, variant-major within a tile group
    int ngroups;           // number of source tiles (each has 8 consecutive slots)
    size_t slot_stride;    // floats per slot
    uint8_t* img_out;
    int img_w, img_h;
    int img_c;             // bytes per pixel of the output image (3 or 4)
    int prepad_s;
};
cudaError_t launch_tta_gather(const TtaGatherArgs& a, int max_crop_w, int max_crop_h, cudaStream_t stream);

// Alpha channel of RGBA images: bicubic x4 (A = -0.75; ncnn layer/interp.cpp:178-249,:645-668, used by
// src/realsr.cpp:129-140,:431-442) of the UN-PADDED tile region (realsr_preproc.comp:79-88), truncated to u8 like the
// CPU branch's to_pixels (no +0.5), written into byte 3 of the RGBA output pixels.  One launch per image.
struct AlphaArgs {
    const uint8_t* img;    // RGBA source
    uint8_t* img_out;      // RGBA result
    int img_w, img_h;      // source size
    int tilesize;
    int tile_first, tile_count;  // only output pixels of tiles [tile_first, tile_first + tile_count) (row-major tile index) are written
};
cudaError_t launch_alpha_bicubic4x(const AlphaArgs& a, cudaStream_t stream);

}  // namespace rsr
