// fake_conv.cpp -- launch_conv3x3 of the functional fake CUDA build (tests/test_engine_cpu_twin.py): a CPU model of the tensor-core
// kernel's CONTRACT, not of its implementation -- fp16 operands read through the tensor maps the engine built (K-block k of the input =
// map k), fp32 accumulation, and the epilogue of each ConvMode with the rounding points DESIGN.md section 2 commits to (fp16 stores,
// fp16 + int8 residual carrier, fp32 fea add, x255 + 0.5 truncation).  What it validates is the engine's DATA FLOW: which plane feeds which
// conv, the X0 rotation, shape-group offsets, TTA slots, tile crops.  The kernel itself is only validated on hardware (tests -m gpu).
// (Own translation unit: unlike the emulated kernels it may be compiled with FMA contraction -- the accumulation order of the tensor
// cores is not part of the contract.)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "fake_map.h"
#include "kernels.h"

extern void fake_cuda_check(const void* p, size_t n, const char* what);

namespace rsr {

namespace {
FakeMap get(const CUtensorMap* m, const char* what)
{
    FakeMap f;
    memcpy(&f, m, sizeof f);
    if (f.magic != FAKE_MAP_MAGIC || f.rank != 4) { fprintf(stderr, "fake conv: %s is not an encoded 4-D tensor map\n", what); abort(); }
    fake_cuda_check(f.base, f.extent, what);
    return f;
}
// element (c, x, y, b) of a [B][H][W][C] plane
inline uint8_t* at(const FakeMap& f, int c, int x, int y, int b)
{
    return (uint8_t*)f.base + (size_t)c * f.elem_bytes + (size_t)x * f.strides[0] + (size_t)y * f.strides[1] + (size_t)b * f.strides[2];
}
inline float h2f(uint16_t h) { __half v; memcpy(&v, &h, 2); return __half2float(v); }
inline uint16_t f2h(float f) { __half v = __float2half_rn(f); uint16_t h; memcpy(&h, &v, 2); return h; }
// the residual carrier of kernels.h (res_ext_encode / res_ext_decode), restated
inline float carrier_decode(uint16_t h, int8_t q) { return __int_as_float(__float_as_int(h2f(h)) + (int)q * 32); }
inline uint8_t carrier_encode(float v, uint16_t h) { return (uint8_t)std::min((__float_as_int(v) - __float_as_int(h2f(h))) >> 5, 127); }
}  // namespace

cudaError_t launch_conv3x3(const ConvMaps* maps, const ConvArgs& a, int cout_pad, cudaStream_t)
{
    const int cin = a.cin, nk = cin / 32;
    const int ng = a.ngroups > 0 ? a.ngroups : 1;
    const int grp = cout_pad, per_ky = cout_pad == 16 ? 64 : 3 * cout_pad;
    // weights: plain row-major [rows][64] fp16 (csrc/model.h): row ((chunk * 3 + ky) * per_ky + kx * grp + n), column = channel in the chunk
    FakeMap wm;
    memcpy(&wm, &maps->w, sizeof wm);
    if (wm.magic != FAKE_MAP_MAGIC || wm.rank != 2) { fprintf(stderr, "fake conv: weight map\n"); abort(); }
    const uint16_t* wrows = (const uint16_t*)wm.base;
    std::vector<float> wf((size_t)9 * cin * cout_pad);  // [ky][kx][ci][n]
    for (int ky = 0; ky < 3; ky++)
        for (int kx = 0; kx < 3; kx++)
            for (int ci = 0; ci < cin; ci++)
                for (int n = 0; n < cout_pad; n++)
                    wf[(((size_t)ky * 3 + kx) * cin + ci) * cout_pad + n] = h2f(wrows[(size_t)(((ci >> 6) * 3 + ky) * per_ky + kx * grp + n) * 64 + (ci & 63)]);
    fake_cuda_check(a.bias, (size_t)cout_pad * 4, "conv bias");
    for (int g = 0; g < ng; g++) {
        int H = a.H, W = a.W, B = a.B;
        long long px_off = 0, aux = 0;
        int desc0 = 0;
        if (a.ngroups > 0) { const ConvGroup& cg = a.grp[g]; H = cg.H; W = cg.W; B = cg.B; px_off = cg.px_off; aux = cg.aux_off; desc0 = cg.desc0; }
        FakeMap in[CONV_MAX_KBLOCKS];
        for (int k = 0; k < nk; k++) in[k] = get(g == 0 ? &maps->m[k] : &maps->mg[g - 1][k], "conv activation map");
        FakeMap o{}, rx{}, re{}, rrx{}, rre{}, reo{};
        if (a.mode == CONV_STORE16 || a.mode == CONV_RDB) o = get(&maps->o[g], "conv output map");
        if (a.mode == CONV_RDB) {
            rx = get(&maps->rx[g], "x0 row map"); re = get(&maps->re[g], "x0 extension row map"); reo = get(&maps->reo[g], "output extension row map");
            if (a.rrdb_end) { rrx = get(&maps->rrx[g], "RRDB row map"); rre = get(&maps->rre[g], "RRDB extension row map"); }
        }
        std::vector<float> xin((size_t)(H + 2) * (W + 2) * cin);
        for (int b = 0; b < B; b++) {
            std::fill(xin.begin(), xin.end(), 0.f);  // zero padding = the TMA's out-of-bounds fill
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    for (int ci = 0; ci < cin; ci++) {
                        const FakeMap& f = in[ci >> 5];
                        const int c = f.dims[0] == 64 ? ci : (ci & 31);  // K-blocks 0, 1: the halves of a 64-channel plane; >= 2: one 32-channel plane each
                        xin[((size_t)(y + 1) * (W + 2) + x + 1) * cin + ci] = h2f(*(const uint16_t*)at(f, c, x, y, b));
                    }
            // rows are independent: spread them over the host cores (the outputs of different rows never overlap)
            const int nthr = std::max(1, std::min((int)std::thread::hardware_concurrency(), H));
            std::vector<std::thread> pool;
            for (int tid = 0; tid < nthr; tid++)
              pool.emplace_back([&, tid] {
                std::vector<float> acc(cout_pad);
            for (int y = tid; y < H; y += nthr)
                for (int x = 0; x < W; x++) {
                    for (int n = 0; n < cout_pad; n++) acc[n] = 0.f;
                    for (int ky = 0; ky < 3; ky++)
                        for (int kx = 0; kx < 3; kx++) {
                            const float* xi = &xin[((size_t)(y + ky) * (W + 2) + x + kx) * cin];
                            const float* wk = &wf[((size_t)ky * 3 + kx) * cin * cout_pad];
                            for (int ci = 0; ci < cin; ci++) {
                                const float v = xi[ci];
                                const float* wr = wk + (size_t)ci * cout_pad;
                                for (int n = 0; n < cout_pad; n++) acc[n] += v * wr[n];
                            }
                        }
                    const size_t px = (size_t)px_off + ((size_t)b * H + y) * W + x;
                    if (a.mode == CONV_STORE16) {
                        for (int n = 0; n < cout_pad; n++) {
                            float v = acc[n] + a.bias[n];
                            if (a.act == 2) v = v > 0.f ? v : v * a.slope;
                            *(uint16_t*)at(o, n, x, y, b) = f2h(v);
                        }
                    } else if (a.mode == CONV_RDB) {
                        for (int n = 0; n < 64; n++) {
                            float t = (acc[n] + a.bias[n]) * 0.2f + carrier_decode(*(const uint16_t*)at(rx, n, x, y, b), *(const int8_t*)at(re, n, x, y, b));
                            if (a.rrdb_end) t = t * 0.2f + carrier_decode(*(const uint16_t*)at(rrx, n, x, y, b), *(const int8_t*)at(rre, n, x, y, b));
                            const uint16_t h = f2h(t);
                            *(uint16_t*)at(o, n, x, y, b) = h;
                            *at(reo, n, x, y, b) = carrier_encode(t, h);
                        }
                    } else if (a.mode == CONV_TRUNK) {
                        fake_cuda_check(a.f32 + px * 64, 256, "trunk fea operand");
                        fake_cuda_check(a.out16 + px * a.out_cs + a.out_co, 128, "trunk output");
                        for (int n = 0; n < 64; n++) reinterpret_cast<uint16_t*>(a.out16)[px * a.out_cs + a.out_co + n] = f2h(acc[n] + a.bias[n] + a.f32[px * 64 + n]);
                    } else {  // the last conv: crop + post-processing (u8) or one fp32 plane per TTA variant
                        fake_cuda_check(a.tiles + desc0 + b, sizeof(TileDesc), "last conv tile descriptor");
                        const TileDesc td = a.tiles[desc0 + b];
                        if (a.mode == CONV_LAST_U8) {
                            const int ii = y - a.prepad_s, jx = x - a.prepad_s;
                            if (ii >= 0 && ii < td.crop_h && jx >= 0 && jx < td.crop_w) {
                                uint8_t* dst = a.img_out + ((size_t)(td.out_y0 + ii) * a.img_w + td.out_x0 + jx) * a.img_c;
                                fake_cuda_check(dst, 3, "last conv output pixel");
                                for (int j = 0; j < 3; j++) {
                                    const int iv = (int)((acc[j] + a.bias[j]) * 255.f + 0.5f);
                                    dst[j] = (uint8_t)std::min(std::max(iv, 0), 255);
                                }
                            }
                        } else {
                            float* dst = a.tta_out + 4 * ((size_t)aux + (size_t)td.tta_slot * ((size_t)H * W) + (size_t)y * W + x);
                            fake_cuda_check(dst, 16, "last conv TTA plane");
                            dst[0] = acc[0] + a.bias[0]; dst[1] = acc[1] + a.bias[1]; dst[2] = acc[2] + a.bias[2]; dst[3] = 0.f;
                        }
                    }
                }
              });
            for (auto& t : pool) t.join();
        }
    }
    return cudaSuccess;
}

}  // namespace rsr
