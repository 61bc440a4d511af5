// fake_cuda.cpp -- test infrastructure: a CPU stand-in for the CUDA runtime / driver entry points and kernel launchers that
// csrc/engine.cu uses, so that the engine's HOST code (model upload, workspace sizing, call planner, tensor-map encodes, launch
// arguments, workspace pool, locking) runs under AddressSanitizer / ThreadSanitizer without a GPU (tests/test_engine_host_sanitizers.py).
// "Device" memory is host memory with every allocation tracked; kernels compute nothing, but every launcher and every tensor-map
// encode CHECKS that the ranges it would touch lie inside a live allocation and obey the TMA encoding rules (16-B aligned base,
// strides multiple of 16, box <= 256 per dimension, inner box extent within the swizzle span).  A violation aborts the process.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <map>
#include <mutex>

#include "fake_map.h"
#include "kernels.h"

namespace {
std::mutex g_mu;
std::map<uintptr_t, size_t> g_alloc;  // base -> bytes
long long g_launches = 0, g_maps = 0;

[[noreturn]] void die(const char* what, const void* p, size_t n)
{
    fprintf(stderr, "fake_cuda: %s: [%p, +%zu) is not inside a live device allocation\n", what, p, n);
    abort();
}
void check(const void* p, size_t n, const char* what)
{
    if (n == 0) return;
    std::lock_guard<std::mutex> g(g_mu);
    const uintptr_t a = (uintptr_t)p;
    auto it = g_alloc.upper_bound(a);
    if (it == g_alloc.begin()) die(what, p, n);
    --it;
    if (a < it->first || a + n > it->first + it->second) die(what, p, n);
}
void check_map(const CUtensorMap* m, const char* what)
{
    FakeMap f;
    memcpy(&f, m, sizeof f);
    if (f.magic != FAKE_MAP_MAGIC) { fprintf(stderr, "fake_cuda: %s: tensor map was never encoded\n", what); abort(); }
    check(f.base, f.extent, what);
}
}  // namespace

extern "C" {
cudaError_t cudaMalloc(void** p, size_t n)
{
    void* q = malloc(n ? n : 1);
    if (!q) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> g(g_mu);
    g_alloc[(uintptr_t)q] = n;
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> g(g_mu);
        if (!g_alloc.erase((uintptr_t)p)) { fprintf(stderr, "fake_cuda: cudaFree of an unknown pointer %p\n", p); abort(); }
    }
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k)
{
    if (k == cudaMemcpyHostToDevice) check(d, n, "cudaMemcpy H2D destination");
    if (k == cudaMemcpyDeviceToHost) check(s, n, "cudaMemcpy D2H source");
    memcpy(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t)
{
    for (size_t y = 0; y < h; y++) {
        if (k == cudaMemcpyDeviceToHost) check((const uint8_t*)s + y * sp, w, "cudaMemcpy2D D2H source");
        memcpy((uint8_t*)d + y * dp, (const uint8_t*)s + y * sp, w);
    }
    return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { check(d, n, "cudaMemset"); memset(d, v, n); return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { memset(p, 0, sizeof *p); p->major = 10; strcpy(p->name, "fake B200"); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)calloc(1, sizeof(double)); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    *(double*)e = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*(double*)b - *(double*)a); return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake error"; }

static CUresult fake_encode(CUtensorMap* m, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                            const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il, CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                            CUtensorMapFloatOOBfill)
{
    const size_t es = dt == CU_TENSOR_MAP_DATA_TYPE_UINT8 ? 1 : (dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT16 ? 2 : 4);
    if (rank < 1 || rank > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return CUDA_ERROR_INVALID_VALUE;
    if ((uintptr_t)base & 15) return CUDA_ERROR_INVALID_VALUE;
    uint64_t extent = dims[0] * es;
    for (cuuint32_t i = 0; i < rank; i++) {
        if (dims[i] == 0 || dims[i] > (1ull << 32) || box[i] == 0 || box[i] > 256 || estr[i] < 1 || estr[i] > 8) return CUDA_ERROR_INVALID_VALUE;
        if (i > 0) {
            if ((strides[i - 1] & 15) || strides[i - 1] >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;
        }
    }
    // bytes from base to the last element: sum over dimensions of (dim - 1) * stride + one innermost row
    extent = dims[0] * es;
    for (cuuint32_t i = 1; i < rank; i++) extent += (dims[i] - 1) * strides[i - 1];
    const size_t inner = box[0] * es;
    const size_t span = sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : (size_t)1 << 30;
    if ((inner & 15) || inner > span) return CUDA_ERROR_INVALID_VALUE;
    check(base, extent, "cuTensorMapEncodeTiled tensor");
    FakeMap f{};
    f.magic = FAKE_MAP_MAGIC; f.base = base; f.extent = extent; f.elem_bytes = (uint32_t)es; f.rank = rank;
    for (cuuint32_t i = 0; i < rank; i++) { f.dims[i] = dims[i]; if (i) f.strides[i - 1] = strides[i - 1]; }
    memset(m, 0, sizeof *m);
    memcpy(m, &f, sizeof f);
    std::lock_guard<std::mutex> g(g_mu);
    g_maps++;
    return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char* sym, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* st)
{
    if (strcmp(sym, "cuTensorMapEncodeTiled")) return cudaErrorInvalidValue;
    *fn = (void*)&fake_encode;
    if (st) *st = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
long long fake_cuda_launches() { std::lock_guard<std::mutex> g(g_mu); return g_launches; }
long long fake_cuda_maps() { std::lock_guard<std::mutex> g(g_mu); return g_maps; }
long long fake_cuda_live_allocations() { std::lock_guard<std::mutex> g(g_mu); return (long long)g_alloc.size(); }
}  // extern "C"

void fake_cuda_check(const void* p, size_t n, const char* what) { check(p, n, what); }

// ---- the kernel launchers (csrc/kernels.h): argument checks instead of kernels --------------------------------------------------------
// (with -DFAKE_CUDA_FUNCTIONAL the launchers come from fake_kernels.cpp instead: the CUDA-core kernels run under the host emulation and
// the conv kernel is replaced by a CPU model of its contract, so that whole images go through the engine)
#ifndef FAKE_CUDA_FUNCTIONAL
namespace rsr {

static void count_launch() { std::lock_guard<std::mutex> g(g_mu); g_launches++; }

cudaError_t launch_conv_first(const FirstArgs& a, cudaStream_t)
{
    const size_t px = (size_t)a.B * a.H * a.W;
    check(a.img, (size_t)a.img_w * a.img_h * a.img_c, "conv_first source image");
    check(a.tiles, sizeof(TileDesc) * a.B, "conv_first tile descriptors");
    check(a.w, 64 * 27 * 4, "conv_first weights");
    check(a.bias, 64 * 4, "conv_first bias");
    check(a.f32, px * 64 * 4, "conv_first fea plane");
    check(a.d16, px * 64 * 2, "conv_first x0 plane");
    check(a.ext, px * 64, "conv_first extension plane");
    for (int b = 0; b < a.B; b++) {
        const TileDesc& t = a.tiles[b];
        const bool tr = t.variant >= 4;
        if ((tr ? t.base_w : t.base_h) != a.H || (tr ? t.base_h : t.base_w) != a.W) { fprintf(stderr, "fake_cuda: conv_first tile shape != group shape\n"); abort(); }
    }
    count_launch();
    return cudaSuccess;
}

cudaError_t launch_conv3x3(const ConvMaps* maps, const ConvArgs& a, int cout_pad, cudaStream_t)
{
    if (cout_pad != 16 && cout_pad != 32 && cout_pad != 64) { fprintf(stderr, "fake_cuda: conv cout_pad %d\n", cout_pad); abort(); }
    check(a.bias, (size_t)cout_pad * 4, "conv bias");
    if (a.cin % 32 || a.cin < 64 || a.cin > 32 * CONV_MAX_KBLOCKS) { fprintf(stderr, "fake_cuda: conv cin %d\n", a.cin); abort(); }
    const int ng = a.ngroups > 0 ? a.ngroups : 1;
    if (ng > 4) { fprintf(stderr, "fake_cuda: conv ngroups %d\n", ng); abort(); }
    check_map(&maps->w, "conv weight map");
    const int band = 2 * conv_bh(cout_pad);
    int t0 = 0;
    for (int g = 0; g < ng; g++) {
        for (int k = 0; k < a.cin / 32; k++) check_map(g == 0 ? &maps->m[k] : &maps->mg[g - 1][k], "conv activation map");
        if (a.mode == CONV_STORE16 || a.mode == CONV_RDB) check_map(&maps->o[g], "conv output map");
        if (a.mode == CONV_RDB) {
            check_map(&maps->rx[g], "conv5 x0 row map");
            check_map(&maps->re[g], "conv5 x0 extension row map");
            check_map(&maps->reo[g], "conv5 output extension row map");
            if (a.rrdb_end) { check_map(&maps->rrx[g], "conv5 RRDB row map"); check_map(&maps->rre[g], "conv5 RRDB extension row map"); }
        }
        if (a.ngroups > 0) {  // the group table must describe consecutive tile ranges of the launch
            const ConvGroup& cg = a.grp[g];
            const int tx = (cg.W + CONV_BW - 1) / CONV_BW, per = tx * ((cg.H + band - 1) / band);
            if (cg.tiles_x != tx || cg.tiles_per_img != per || cg.tile_start != t0 || cg.B < 1) { fprintf(stderr, "fake_cuda: conv group table\n"); abort(); }
            t0 += per * cg.B;
            const size_t px = (size_t)cg.B * cg.H * cg.W;
            if (a.mode == CONV_TRUNK) {
                check(a.f32 + cg.px_off * 64, px * 64 * 4, "trunk fea operand");
                check(a.out16 + cg.px_off * a.out_cs + a.out_co, px * 64 * 2, "trunk output");
            }
            if (a.mode == CONV_LAST_U8 || a.mode == CONV_LAST_F32) check(a.tiles + cg.desc0, sizeof(TileDesc) * cg.B, "last conv tile descriptors");
            if (a.mode == CONV_LAST_F32) check(a.tta_out + cg.aux_off * 4, (size_t)cg.B * cg.H * cg.W * 16, "last conv TTA planes");
        }
    }
    if (a.ngroups > 0 && t0 != a.ntiles_multi) { fprintf(stderr, "fake_cuda: conv ntiles_multi\n"); abort(); }
    if (a.mode == CONV_LAST_U8) check(a.img_out, (size_t)a.img_w * a.img_h * a.img_c, "last conv output image");
    count_launch();
    return cudaSuccess;
}

cudaError_t launch_upsample2x(const __half* in, int in_cs, __half* out, int B, int H, int W, cudaStream_t)
{
    check(in, (size_t)B * H * W * in_cs * 2, "upsample source");
    check(out, (size_t)B * H * W * 4 * 64 * 2, "upsample destination");
    count_launch();
    return cudaSuccess;
}

cudaError_t launch_tta_gather(const TtaGatherArgs& a, int max_crop_w, int max_crop_h, cudaStream_t)
{
    check(a.tiles, sizeof(TileDesc) * 8 * a.ngroups, "tta_gather tile descriptors");
    check(a.tta_out, (size_t)a.ngroups * 8 * a.slot_stride * 16, "tta_gather planes");
    check(a.img_out, (size_t)a.img_w * a.img_h * a.img_c, "tta_gather output image");
    for (int g = 0; g < a.ngroups; g++) {
        const TileDesc& t = a.tiles[g * 8];
        if (t.crop_w > max_crop_w || t.crop_h > max_crop_h || t.out_x0 + t.crop_w > a.img_w || t.out_y0 + t.crop_h > a.img_h) { fprintf(stderr, "fake_cuda: tta_gather crop\n"); abort(); }
    }
    count_launch();
    return cudaSuccess;
}

cudaError_t launch_alpha_bicubic4x(const AlphaArgs& a, cudaStream_t)
{
    check(a.img, (size_t)a.img_w * a.img_h * 4, "alpha source");
    check(a.img_out, (size_t)a.img_w * a.img_h * 64, "alpha destination");
    count_launch();
    return cudaSuccess;
}

}  // namespace rsr
#endif  // FAKE_CUDA_FUNCTIONAL
