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
    CONV_STORE16 = 0,   // out16 = fp16(act(acc + b))
    CONV_RDB_OUT = 1,   // t = 0.2*(acc+b) + T32 ; [rrdb_end: t = 0.2*t + R32 ; R32 = t] ; T32 = t ; out16 = fp16(t)
    CONV_TRUNK = 2,     // out16 = fp16(acc + b + F32)
    CONV_LAST_U8 = 3,   // v = acc+b -> crop, v*255+0.5, (int) trunc, clamp -> u8 HWC output image
    CONV_LAST_F32 = 4,  // v = acc+b -> fp32 [slot][4H][4W][4] (TTA; averaged by tta_gather)
    CONV_RDB_OUT16 = 5, // like CONV_RDB_OUT, but the RDB input is read from the fp16 plane `res16` and T32 is not written
};

// One same-shape group of padded tiles inside a merged multi-shape launch (ConvArgs::ngroups > 0): the groups' activations
// are concatenated pixel-major in every buffer, group g starting px_off pixels into it.
struct ConvGroup {
    int H, W, B;                 // this conv's input == output size of the group, tiles in the group
    int tiles_x, tiles_per_img;  // work-tile grid of one image of the group (CTA-pair kernel: 30 x 16 px pair tiles)
    int tile_start;              // first work tile of the group in the launch-wide numbering
    int desc0;                   // index of the group's first TileDesc in ConvArgs::tiles
    long long px_off;            // pixel offset of the group inside the activation buffers (already scaled for 2x / 4x layers)
    long long aux_off;           // CONV_LAST_F32: float4 offset of the group's first TTA slot inside ConvArgs::tta_out
};

struct ConvArgs {
    const __half* wpk;  // packed weights [chunk][tap][N][64] fp16, pre-swizzled smem image (see model.cpp)
    const float* bias;  // [N]
    int cin;            // real input channels (multiple of 16)
    int H, W, B;        // spatial size of this conv's input == output, batch
    int mode, act;      // ConvMode; act: 0 none, 2 leaky relu
    float slope;
    __half* out16;
    int out_cs, out_co;  // channels per pixel of the output buffer, channel offset
    float* t32;
    float* r32;
    const float* f32;
    const __half* res16;  // CONV_RDB_OUT, optional: take the RDB input from this fp16 plane (64 ch/px) instead of T32 and do
                          // not write T32 ("residual_fp16" engine option; the RRDB-level skip R32 stays fp32)
    int rrdb_end;
    // CONV_LAST_*
    const TileDesc* tiles;
    uint8_t* img_out;
    int img_w, img_h;    // output image size (pixels)
    int img_c;           // bytes per pixel of the output image (3 or 4)
    int prepad_s;        // prepadding*scale
    float* tta_out;
    long long* dbg_ts;   // optional (bring-up): per-tile clock64 stamps of cluster 0, 8 slots per tile, see conv_tc2.cu
    // CTA-pair kernel, merged multi-shape launch ("merge_shapes" option): up to 4 same-shape groups in ONE launch per layer, so
    // that a frame pays the per-launch pipeline fill/drain (7.7 us for the Cout=32 convs, 15 us for the Cout=64 conv, measured)
    // once per layer instead of once per layer and shape.  ngroups == 0: plain launch described by H, W, B above.
    int ngroups;
    int ntiles_multi;     // total work tiles of the merged launch
    ConvGroup grp[4];
    int reverse;          // CTA-pair kernel: walk the work tiles from last to first.  The engine alternates it from one conv to the next
                          // ("serpentine" option) so that a conv starts on the pixels the previous conv touched last -- the part of its
                          // inputs (and of the previous conv's outputs) that is still in the 126 MB L2
    int kblock;           // CTA-pair kernel: input channels per A stage, 64 (default) or 32 (tensor maps must have 32-channel SWIZZLE_64B boxes)
    int epi_split;        // CTA-pair kernel: 1 = channel-split epilogue with early accumulator hand-back (conv_tc2.cu: epilogue_split)
    int debug_mask;       // debug mask for timing experiments (0 in production): 1 no MMA, 2 no A loads, 4 no B loads, 8 no epilogue I/O
};

// one tensor map per 64-channel input plane (chunk) of a conv
struct alignas(64) ConvMaps {
    CUtensorMap m[3];
    CUtensorMap w;    // 2-D map over the conv's weight rows [rows][64], box 64 rows (CTA-pair kernel only)
    CUtensorMap w32;  // same tensor, box 32 rows
    CUtensorMap mg[3][3];  // merged multi-shape launch: the planes of groups 1..3 (group 0 uses m[])
};
// conv3x3 s1 p1 as implicit GEMM on tcgen05 (conv_tc.cu). cout_pad in {16, 32, 64}.
cudaError_t launch_conv3x3_tc(const ConvMaps* in_maps, const ConvArgs& a, int cout_pad, cudaStream_t stream);
// same contract, CTA-pair (cta_group::2) implementation (conv_tc2.cu); needs in_maps->w and 10-row activation boxes
cudaError_t launch_conv3x3_tc2(const ConvMaps* in_maps, const ConvArgs& a, int cout_pad, cudaStream_t stream);
// conv_tc2x.cu: the same kernel with the opt-in variants (ConvArgs::epi_split / kblock / ngroups / reverse); conv_tc2.cu ignores those
cudaError_t launch_conv3x3_tc2x(const ConvMaps* in_maps, const ConvArgs& a, int cout_pad, cudaStream_t stream);
// conv3x3.cu: the unified CTA-pair kernel (early accumulator hand-back, lean MMA issuer, merged launches, K-block 64 / 32)
cudaError_t launch_conv3x3(const ConvMaps* in_maps, const ConvArgs& a, int cout_pad, cudaStream_t stream);
constexpr int CONV2_BH = 8;
constexpr int CONV2_BW = 30;    // output columns per work tile of the CTA-pair kernel (CONV_P - 2)
constexpr int CONV2_BAND = 16;  // output rows per work tile of the CTA-pair kernel (2 x CONV2_BH)
// box dims the kernel expects from the tensor map: {64 ch, 32 px, CONV_BH+2 rows, 1}
constexpr int CONV_BH = 16;
constexpr int CONV_P = 32;

// Conv_0 (3->64, fp32 CUDA cores) fused with preproc (crop, reflect-101, x 1/255, TTA variant), conv_first.cu
struct FirstArgs {
    const uint8_t* img;  // u8 HWC RGB source image on the device
    int img_w, img_h;
    int img_c;           // bytes per pixel of the source image: 3 (RGB) or 4 (RGBA; alpha is ignored here)
    const TileDesc* tiles;
    const float* w;      // [64][3][3][3] fp32 (OIHW as in the .bin), bias [64]
    const float* bias;
    int H, W, B;
    __half* d16;         // plane 0 of D16[0] (64 ch per pixel)
    float* t32;
    float* r32;
    float* f32;
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
