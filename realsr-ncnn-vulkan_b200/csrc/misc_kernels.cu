// misc_kernels.cu -- the bandwidth-bound kernels around the conv stack.
//   conv_first : reference preproc (src/realsr.cpp:735-766 / realsr_preproc.comp:47-95: crop, reflect-101 at the image
//                border, x (1/255.f)), the TTA scatter (src/realsr.cpp:618-664) and Conv_0 (3->64, x4.param:4) in ONE
//                fp32 CUDA-core kernel.  Conv_0 must see the exact u8/255 inputs in fp32 (rounding the
//                input image to fp16 costs 4-5 LSB), so it never goes through the fp16 tensor-core path.
//   upsample2x : ncnn Interp nearest (layer/interp.cpp:598-618) on NHWC fp16.
//   tta_gather : inverse dihedral gather + average + postproc (src/realsr.cpp:691-724, :816-833).
#include "kernels.h"

namespace rsr {

namespace {

__device__ __forceinline__ int reflect101(int v, int n)
{
    v = abs(v);
    v = (n - 1) - abs(v - (n - 1));
    return min(max(v, 0), n - 1);  // only differs from the reference for images narrower than the prepadding
}

// variant-space coords (a = row, b = col) -> base-tile coords (i, j); inverse of src/realsr.cpp:640-660
__device__ __forceinline__ void variant_to_base(int variant, int a, int b, int ph, int pw, int& i, int& j)
{
    switch (variant) {
    default: i = a; j = b; break;
    case 1: i = ph - 1 - a; j = b; break;
    case 2: i = a; j = pw - 1 - b; break;
    case 3: i = ph - 1 - a; j = pw - 1 - b; break;
    case 4: i = b; j = a; break;
    case 5: i = b; j = pw - 1 - a; break;
    case 6: i = ph - 1 - b; j = a; break;
    case 7: i = ph - 1 - b; j = pw - 1 - a; break;
    }
}

constexpr int FIRST_THREADS = 256;
constexpr int FIRST_TW = 32, FIRST_TH = 8;       // output pixels of a block: 32 columns x 8 rows
constexpr int FIRST_PW = FIRST_TW + 2 + 2;       // patch row pitch in floats (34 used)

// One block = a 32 x 8 patch of one padded tile.  The 34 x 10 x 3 source patch is gathered ONCE (reflect-101, TTA variant, u8 -> float
// x 1/255) into shared memory -- the first version gathered it per output thread, 36 times per pixel, and was bound by that index
// arithmetic (1.6 TB/s of stores).  Thread (pl, quarter) then computes 16 output channels of 4 pixels: every weight float4 fetched
// from shared memory feeds 16 FMAs.  Summation order per output = the naive layer's (bias, then input-channel major, tap minor).
__global__ void __launch_bounds__(FIRST_THREADS) conv_first_kernel(const FirstArgs a)
{
    __shared__ float4 sw[27 * 16];  // weights re-ordered to [c][tap][oc]: 4 consecutive output channels are one float4
    __shared__ float sb[64];
    __shared__ float sp[3][FIRST_TH + 2][FIRST_PW];
    for (int i = threadIdx.x; i < 27 * 64; i += FIRST_THREADS) {
        const int oc = i & 63, ct = i >> 6;  // ct = c*9 + tap
        reinterpret_cast<float*>(sw)[ct * 64 + oc] = a.w[oc * 27 + ct];
    }
    if (threadIdx.x < 64) sb[threadIdx.x] = a.bias[threadIdx.x];

    const int b = blockIdx.y;
    const TileDesc td = a.tiles[b];
    const int tiles_x = (a.W + FIRST_TW - 1) / FIRST_TW;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * FIRST_TW, y0 = ty * FIRST_TH;
    for (int i = threadIdx.x; i < (FIRST_TH + 2) * (FIRST_TW + 2); i += FIRST_THREADS) {
        const int py = i / (FIRST_TW + 2), px = i - py * (FIRST_TW + 2);
        const int aa = y0 + py - 1, bb = x0 + px - 1;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (aa >= 0 && aa < a.H && bb >= 0 && bb < a.W) {  // zero padding of the conv at the padded-tile border
            int ii, jj;
            variant_to_base(td.variant, aa, bb, td.base_h, td.base_w, ii, jj);
            const int X = reflect101(td.in_x0 + jj, a.img_w);
            const int Y = reflect101(td.in_y0 + ii, a.img_h);
            const uint8_t* src = a.img + ((size_t)Y * a.img_w + X) * a.img_c;
            v0 = (float)src[0] * (1 / 255.f);
            v1 = (float)src[1] * (1 / 255.f);
            v2 = (float)src[2] * (1 / 255.f);
        }
        sp[0][py][px] = v0;
        sp[1][py][px] = v1;
        sp[2][py][px] = v2;
    }
    __syncthreads();

    const int quarter = threadIdx.x & 3, pl = threadIdx.x >> 2;  // pixels pl + 64 i: row 2 i + (pl >> 5), column pl & 31
    const int col = pl & 31, row0 = pl >> 5;
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 16; j++) acc[i][j] = sb[quarter * 16 + j];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const int ct = c * 9 + ky * 3 + kx;
                float4 w[4];
#pragma unroll
                for (int j4 = 0; j4 < 4; j4++) w[j4] = sw[ct * 16 + quarter * 4 + j4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float v = sp[c][2 * i + row0 + ky][col + kx];
#pragma unroll
                    for (int j4 = 0; j4 < 4; j4++) {
                        acc[i][j4 * 4 + 0] = fmaf(v, w[j4].x, acc[i][j4 * 4 + 0]);
                        acc[i][j4 * 4 + 1] = fmaf(v, w[j4].y, acc[i][j4 * 4 + 1]);
                        acc[i][j4 * 4 + 2] = fmaf(v, w[j4].z, acc[i][j4 * 4 + 2]);
                        acc[i][j4 * 4 + 3] = fmaf(v, w[j4].w, acc[i][j4 * 4 + 3]);
                    }
                }
            }
    const int x = x0 + col;
    if (x >= a.W) return;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int y = y0 + 2 * i + row0;
        if (y >= a.H) continue;
        const size_t pix = ((size_t)b * a.H + y) * a.W + x;
        float4* f = reinterpret_cast<float4*>(a.f32 + pix * 64 + quarter * 16);  // fp32 `fea` for the trunk add (x4.param:994-995)
#pragma unroll
        for (int j4 = 0; j4 < 4; j4++) f[j4] = make_float4(acc[i][j4 * 4], acc[i][j4 * 4 + 1], acc[i][j4 * 4 + 2], acc[i][j4 * 4 + 3]);
        uint4* d = reinterpret_cast<uint4*>(a.d16 + pix * 64 + quarter * 16);
        uint32_t ext[4] = {0, 0, 0, 0};  // int8 extensions of the 16 fp16 values (ConvArgs::ext_in), 4 per word
#pragma unroll
        for (int j8 = 0; j8 < 2; j8++) {
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float v0 = acc[i][j8 * 8 + 2 * k], v1 = acc[i][j8 * 8 + 2 * k + 1];
                __half2 hh = __floats2half2_rn(v0, v1);
                pk[k] = *reinterpret_cast<uint32_t*>(&hh);
                const float2 g = __half22float2(hh);
                const uint32_t q = (uint32_t)(res_ext_encode(v0, g.x) & 255) | ((uint32_t)(res_ext_encode(v1, g.y) & 255) << 8);  // floor / saturating, see kernels.h
                ext[j8 * 2 + (k >> 1)] |= q << (16 * (k & 1));
            }
            d[j8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        *reinterpret_cast<uint4*>(a.ext + pix * 64 + quarter * 16) = make_uint4(ext[0], ext[1], ext[2], ext[3]);
    }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ in, int in_vec_stride, uint4* __restrict__ out, int B, int H, int W)
{
    // one thread = one 16-B vector (8 channels) of one output pixel; 8 vectors per output pixel (64 ch).
    // in_vec_stride = 16-B vectors per source pixel (8 for a 64-ch tensor, 24 for the 192-ch dense buffer).
    const long long n = (long long)B * (2 * H) * (2 * W) * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i & 7);
        long long p = i >> 3;
        const int x = (int)(p % (2 * W)); p /= (2 * W);
        const int y = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        out[i] = __ldg(in + (((size_t)b * H + (y >> 1)) * W + (x >> 1)) * in_vec_stride + v);
    }
}

__global__ void __launch_bounds__(256) tta_gather_kernel(const TtaGatherArgs a)
{
    const int g = blockIdx.z;
    const TileDesc td = a.tiles[g * 8];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= td.crop_h || j >= td.crop_w) return;
    const int oh = td.base_h * 4, ow = td.base_w * 4, ps = a.prepad_s;
    const float4* s = reinterpret_cast<const float4*>(a.tta_out) + (size_t)g * 8 * a.slot_stride;
    const size_t ss = a.slot_stride;
    const float4 v0 = s[0 * ss + (size_t)(i + ps) * ow + (j + ps)];
    const float4 v1 = s[1 * ss + (size_t)(oh - 1 - i - ps) * ow + (j + ps)];
    const float4 v2 = s[2 * ss + (size_t)(i + ps) * ow + (ow - 1 - ps - j)];
    const float4 v3 = s[3 * ss + (size_t)(oh - 1 - i - ps) * ow + (ow - 1 - ps - j)];
    const float4 v4 = s[4 * ss + (size_t)(j + ps) * oh + (i + ps)];
    const float4 v5 = s[5 * ss + (size_t)(ow - 1 - j - ps) * oh + (i + ps)];
    const float4 v6 = s[6 * ss + (size_t)(j + ps) * oh + (oh - 1 - i - ps)];
    const float4 v7 = s[7 * ss + (size_t)(ow - 1 - j - ps) * oh + (oh - 1 - i - ps)];
    float o[3];
    o[0] = (v0.x + v1.x + v2.x + v3.x + v4.x + v5.x + v6.x + v7.x) / 8;
    o[1] = (v0.y + v1.y + v2.y + v3.y + v4.y + v5.y + v6.y + v7.y) / 8;
    o[2] = (v0.z + v1.z + v2.z + v3.z + v4.z + v5.z + v6.z + v7.z) / 8;
    uint8_t* dst = a.img_out + ((size_t)(td.out_y0 + i) * a.img_w + td.out_x0 + j) * a.img_c;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int iv = (int)(o[c] * 255.f + 0.5f);
        dst[c] = (uint8_t)min(max(iv, 0), 255);
    }
}

// cubic coefficients + source offset of output coordinate d (scale 1/4), with ncnn's border folding (interp.cpp:205-246)
__device__ __forceinline__ int cubic_coeffs4(int d, int w, float (&a)[4])
{
    float fx = (float)((d + 0.5) * 0.25 - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    const float A = -0.75f;
    const float fx0 = fx + 1, fx1 = fx, fx2 = 1 - fx;
    a[0] = A * fx0 * fx0 * fx0 - 5 * A * fx0 * fx0 + 8 * A * fx0 - 4 * A;
    a[1] = (A + 2) * fx1 * fx1 * fx1 - (A + 3) * fx1 * fx1 + 1;
    a[2] = (A + 2) * fx2 * fx2 * fx2 - (A + 3) * fx2 * fx2 + 1;
    a[3] = 1.f - a[0] - a[1] - a[2];
    if (sx <= -1) { sx = 1; a[0] = 1.f - a[3]; a[1] = a[3]; a[2] = 0.f; a[3] = 0.f; }
    if (sx == 0) { sx = 1; a[0] = a[0] + a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = 0.f; }
    if (sx == w - 2) { sx = w - 3; a[3] = a[2] + a[3]; a[2] = a[1]; a[1] = a[0]; a[0] = 0.f; }
    if (sx >= w - 1) { sx = w - 3; a[3] = 1.f - a[0]; a[2] = a[0]; a[1] = 0.f; a[0] = 0.f; }
    return sx;
}

__global__ void __launch_bounds__(256) alpha_bicubic4x_kernel(const AlphaArgs a)
{
    const int X = blockIdx.x * blockDim.x + threadIdx.x;  // output pixel
    const int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= a.img_w * 4 || Y >= a.img_h * 4) return;
    const int T = a.tilesize;
    // the tile this output pixel belongs to, and the tile's un-padded source region
    const int xi = (X >> 2) / T, yi = (Y >> 2) / T;
    const int tidx = yi * ((a.img_w + T - 1) / T) + xi;
    if (tidx < a.tile_first || tidx >= a.tile_first + a.tile_count) return;  // another handle's tile (rsr_process_tiles)
    const int tx0 = xi * T, ty0 = yi * T;
    const int tw = min(tx0 + T, a.img_w) - tx0, th = min(ty0 + T, a.img_h) - ty0;
    float ca[4], cb[4];
    const int sx = cubic_coeffs4(X - tx0 * 4, tw, ca);
    const int sy = cubic_coeffs4(Y - ty0 * 4, th, cb);
    float rows[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        // rows sy-1 .. sy+2, columns sx-1 .. sx+2 of the tile; zero-weight taps may point outside: clamp the address only
        const int yy = min(max(sy - 1 + k, 0), th - 1);
        const uint8_t* S = a.img + ((size_t)(ty0 + yy) * a.img_w + tx0) * 4 + 3;
        const float s0 = (float)S[(size_t)min(max(sx - 1, 0), tw - 1) * 4];
        const float s1 = (float)S[(size_t)min(max(sx, 0), tw - 1) * 4];
        const float s2 = (float)S[(size_t)min(max(sx + 1, 0), tw - 1) * 4];
        const float s3 = (float)S[(size_t)min(max(sx + 2, 0), tw - 1) * 4];
        // same operation order as ncnn (mul, then three fused multiply-adds), explicit so that nvcc cannot re-associate
        rows[k] = __fmaf_rn(s3, ca[3], __fmaf_rn(s2, ca[2], __fmaf_rn(s1, ca[1], __fmul_rn(s0, ca[0]))));
    }
    const float v = __fmaf_rn(rows[3], cb[3], __fmaf_rn(rows[2], cb[2], __fmaf_rn(rows[1], cb[1], __fmul_rn(rows[0], cb[0]))));
    const int iv = (int)v;
    a.img_out[((size_t)Y * a.img_w * 4 + X) * 4 + 3] = (uint8_t)min(max(iv, 0), 255);
}

// grid / block of conv_first_kernel (shared with the host emulation in tests/emu)
inline void conv_first_launch_dims(const FirstArgs& a, dim3& grid, dim3& block)
{
    grid = dim3((unsigned)(((a.W + FIRST_TW - 1) / FIRST_TW) * ((a.H + FIRST_TH - 1) / FIRST_TH)), a.B);
    block = dim3(FIRST_THREADS);
}

}  // namespace

#ifndef RSR_HOST_EMULATION  // tests/emu includes this file on the host: the <<<>>> launchers are device-build only
cudaError_t launch_alpha_bicubic4x(const AlphaArgs& a, cudaStream_t stream)
{
    dim3 block(32, 8);
    dim3 grid((a.img_w * 4 + 31) / 32, (a.img_h * 4 + 7) / 8);
    alpha_bicubic4x_kernel<<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_conv_first(const FirstArgs& a, cudaStream_t stream)
{
    dim3 grid, block;
    conv_first_launch_dims(a, grid, block);
    conv_first_kernel<<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_upsample2x(const __half* in, int in_cs, __half* out, int B, int H, int W, cudaStream_t stream)
{
    const long long n = (long long)B * 4 * H * W * 8;
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    upsample2x_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), in_cs / 8, reinterpret_cast<uint4*>(out), B, H, W);
    return cudaGetLastError();
}

cudaError_t launch_tta_gather(const TtaGatherArgs& a, int max_crop_w, int max_crop_h, cudaStream_t stream)
{
    dim3 block(32, 8);
    dim3 grid((max_crop_w + 31) / 32, (max_crop_h + 7) / 8, a.ngroups);
    tta_gather_kernel<<<grid, block, 0, stream>>>(a);
    return cudaGetLastError();
}
#endif  // RSR_HOST_EMULATION

}  // namespace rsr
