// Host emulation of the CUDA-core kernels of the engine (see cuda_host_emu.h): the kernel SOURCE is included as is.
// Exports plain C entry points for the ctypes tests in tests/test_kernel_emulation.py.
#include "cuda_host_emu.h"
#define RSR_HOST_EMULATION 1
#include RSR_MISC_KERNELS_CU

using namespace rsr;

extern "C" {

// conv_first_kernel over one group of B same-shape padded tiles (descriptors td[B]); outputs f32 [B][H][W][64], d16, ext
int emu_conv_first(const uint8_t* img, int img_w, int img_h, int img_c, const TileDesc* td, int B, int H, int W, const float* w,
                   const float* bias, float* f32, uint16_t* d16, uint8_t* ext)
{
    FirstArgs a{};
    a.img = img; a.img_w = img_w; a.img_h = img_h; a.img_c = img_c; a.tiles = td; a.w = w; a.bias = bias; a.H = H; a.W = W; a.B = B;
    a.f32 = f32; a.d16 = reinterpret_cast<__half*>(d16); a.ext = ext;
    dim3 grid, block;
    conv_first_launch_dims(a, grid, block);
    cuda_emu::launch(grid, block, [=] { conv_first_kernel(a); });
    return (int)(grid.x * grid.y);
}

int emu_upsample2x(const uint16_t* in, int in_cs, uint16_t* out, int B, int H, int W)
{
    cuda_emu::launch(dim3(3), dim3(256), [=] {
        upsample2x_kernel(reinterpret_cast<const uint4*>(in), in_cs / 8, reinterpret_cast<uint4*>(out), B, H, W);
    });
    return 0;
}

#ifndef RSR_EMU_CONV_FIRST_ONLY  // the conv_first_simple fixture has only that kernel
// alpha_bicubic4x_kernel over a whole RGBA image (all tiles): writes byte 3 of every output pixel
int emu_alpha_bicubic4x(const uint8_t* img, int img_w, int img_h, int tilesize, int tile_first, int tile_count, uint8_t* out)
{
    AlphaArgs a{};
    a.img = img; a.img_out = out; a.img_w = img_w; a.img_h = img_h; a.tilesize = tilesize; a.tile_first = tile_first; a.tile_count = tile_count;
    cuda_emu::launch(dim3((img_w * 4 + 31) / 32, (img_h * 4 + 7) / 8), dim3(32, 8), [=] { alpha_bicubic4x_kernel(a); });
    return 0;
}

// tta_gather_kernel: ngroups tiles x 8 variant planes [slot][4H][4W][4] fp32 -> averaged, post-processed u8 crop in the output image
int emu_tta_gather(const float* tta_out, const TileDesc* td, int ngroups, long long slot_stride, uint8_t* img_out, int img_w, int img_h, int img_c,
                   int prepad_s, int max_crop_w, int max_crop_h)
{
    TtaGatherArgs a{};
    a.tta_out = tta_out; a.tiles = td; a.ngroups = ngroups; a.slot_stride = (size_t)slot_stride; a.img_out = img_out; a.img_w = img_w; a.img_h = img_h;
    a.img_c = img_c; a.prepad_s = prepad_s;
    cuda_emu::launch(dim3((max_crop_w + 31) / 32, (max_crop_h + 7) / 8, ngroups), dim3(32, 8), [=] { tta_gather_kernel(a); });
    return 0;
}

#endif

int emu_sizeof_tiledesc() { return (int)sizeof(TileDesc); }
}
