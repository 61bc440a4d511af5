// engine.cu -- the tile engine behind the C ABI (include/realsr_b200.h): B200-native replacement of
// RealSR::load / RealSR::process (/root/reference/src/realsr.cpp:38-143, :145-523).
//
// What the reference does per image (realsr.cpp:170-516): loop over tiles, one blocking Vulkan submit per tile,
// ~1000 dispatches per tile through a graph interpreter.  Here: the graph is lowered once to a static schedule of
// 1 + 15*nb + 5 fused kernels; tiles of equal padded shape are batched into ONE launch per layer (tile geometry is
// unchanged -- it is part of the numerical contract), everything is queued on one stream per call without host
// synchronisation, and the source / result images stay resident in HBM for the whole call.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/realsr_b200.h"
#include "kernels.h"
#include "model.h"

using namespace rsr;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// Tensor maps of the activation buffers of one job (one or several same-shape tile groups), built once per job.
// Dense-block activations of a job with npx network-input pixels: the block input x0 lives in one of THREE rotating planes X0[i]
// [npx][64] fp16 (+ its int8 extension E8[i] [npx][64], see ConvArgs::res16): conv5 writes the next block's x0 into another plane
// than it reads (neighbouring CTAs still read this one's halo), and the input of the enclosing RRDB survives until that RRDB's last
// block adds it back -- in the same fp16 + int8 form, so no fp32 copy of the stream exists.  The growth channels x1..x4 live in ONE
// buffer XG as four [npx][32] planes -- one plane per 32-channel K-block, so that a conv reads exactly the channels it needs (ncu,
// 64-channel pair planes: the 96- and 160-channel convs fetched a whole 128-B line per pixel for 64 useful bytes and ran at
// 5.5-6.2 TB/s of DRAM traffic, i.e. HBM-bound).  K-block k of a conv's input lives in map m[k]: k = 0, 1 -> the two halves of
// X0[i] (channel coordinate 32k), k >= 2 -> plane k-2 of XG.  The 64-channel tail buffers use k = 0, 1 only.
struct JobMaps {
    ConvMaps d8[3], d4[3];          // [X0 plane] with 10-row boxes (Cout=32 convs) and 6-row boxes (Cout=64 convs: conv5, trunk)
    ConvMaps u1_4, u2_4, v2_4;      // tail inputs of the Cout=64 convs (up1 reads U1, up2 reads U2, HR conv reads V2): 6-row boxes
    ConvMaps u2_8;                  // last conv (Cout=3, 8-row tiles) reads U2
    CUtensorMap oD[4][4];           // [x1..x4][group]: output maps of the growth convs
    CUtensorMap rX[3][4], rE[3][4]; // [X0 plane][group]: row maps of the residual stream (fp16 plane / int8 extension), conv5's epilogue
    CUtensorMap oV1[4], oV2[4], oU2[4];
};
inline CUtensorMap* map_slot(ConvMaps& cm, int g, int k) { return g == 0 ? &cm.m[k] : &cm.mg[g - 1][k]; }

struct Workspace {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf x0[3], e8[3], xg, f32, u1, v1, u2, v2, tta, tiles, img_in, img_out;
    JobMaps maps;                 // tensor maps of the job being enqueued (kernel parameters are copied at launch)
    TileDesc* h_tiles = nullptr;  // pinned staging for the tile descriptors of one process() call
    size_t h_tiles_cap = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_events;
};

struct DeviceModel {
    std::vector<float*> bias;   // per conv, padded to cout_pad
    std::vector<int> cout_pad;
    std::vector<CUtensorMap> wmap;  // per conv: 2-D map over its weight rows [rows][64 ch], box {32 ch, N/2 rows}, SWIZZLE_64B
    float* w0 = nullptr;        // Conv_0 fp32 weights [64][27]
    void* blob = nullptr;       // one allocation holding everything
};

}  // namespace

struct rsr_handle {
    int gpuid = 0, tta = 0;
    int scale = 4, tilesize = 200, prepadding = 10;
    bool loaded = false;
    Model model;
    DeviceModel dm;
    std::mutex mu;
    std::vector<Workspace*> pool;
    std::string err;
    // options
    long long max_batch = 0;  // 0 = auto
    int merge_shapes = 1;     // one launch per layer over all shape groups of a job (see run_job); 0 = one job per shape group
    int serpentine = 1;  // every second conv walks its work tiles backwards: it starts on the pixels the previous conv touched last, which are still in L2
    int debug_mask = 0;  // timing experiments only (see include/realsr_b200.h)
    int time_convs = 0;
    // stats
    long long launches = 0;
    double last_ms = 0, last_conv_ms = 0;
    long long last_conv_launches = 0;  // tensor-core conv launches of the last call
    long long last_padded = 0;
};

namespace {

int set_err(rsr_handle* h, int code, const std::string& msg)
{
    std::lock_guard<std::mutex> g(h->mu);
    h->err = msg;
    return code;
}

#define CK(call)                                                                                                        \
    do {                                                                                                                \
        cudaError_t e_ = (call);                                                                                        \
        if (e_ != cudaSuccess) {                                                                                        \
            char b_[512];                                                                                               \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);        \
            return set_err(h, e_ == cudaErrorMemoryAllocation ? RSR_ERR_ALLOC : RSR_ERR_CUDA, b_);                      \
        }                                                                                                               \
    } while (0)

int ensure(rsr_handle* h, DevBuf& b, size_t bytes)
{
    if (b.cap >= bytes) return 0;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;  // grow-only with slack; TMA needs 16-B alignment (cudaMalloc gives 256)
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return 0;
}

// weight rows [rows][64 ch] of one conv, box {32 ch, box_rows}: a CTA's half (N/2 rows) of one (32-channel K-block, ky) block
int make_wmap(rsr_handle* h, CUtensorMap* m, const void* base, int rows, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled failed for a weight map");
    return 0;
}

// one fp16 activation plane [B][H][W][C] (C = 64 or 32 channels per pixel); box {32 ch, 32 px, box_rows, 1}, SWIZZLE_64B,
// out-of-bounds = zero fill (the conv's padding)
int make_map(rsr_handle* h, CUtensorMap* m, const void* base, int C, int W, int H, int B, int box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)CONV_P, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) for C=%d W=%d H=%d B=%d", (int)r, C, W, H, B);
        return set_err(h, RSR_ERR_CUDA, b);
    }
    return 0;
}

// output map of a CONV_STORE16 launch of the unified kernel (conv3x3.cu): a 64-channel fp16 plane [B][H][W][64], box = one 30-pixel
// row segment of box_c channels (32: SWIZZLE_64B, 64: SWIZZLE_128B) -- the layout the epilogue stages in shared memory
int make_omap(rsr_handle* h, CUtensorMap* m, const void* base, int W, int H, int B, int box_c, int C = 64)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)CONV_BW, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    box_c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled failed for an output map");
    return 0;
}

// row map of an int8 extension plane [B][H][W][64 B]: box = one 30-pixel row segment, SWIZZLE_64B (conv5's epilogue, conv3x3.cu)
int make_emap(rsr_handle* h, CUtensorMap* m, const void* base, int W, int H, int B)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {64, (cuuint64_t)W * 64, (cuuint64_t)H * W * 64};
    cuuint32_t box[4] = {64, (cuuint32_t)CONV_BW, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(h, RSR_ERR_CUDA, "cuTensorMapEncodeTiled failed for an extension map");
    return 0;
}

Workspace* acquire_ws(rsr_handle* h)
{
    {
        std::lock_guard<std::mutex> g(h->mu);
        if (!h->pool.empty()) {
            Workspace* w = h->pool.back();
            h->pool.pop_back();
            return w;
        }
    }
    Workspace* w = new Workspace();
    if (cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&w->ev0) != cudaSuccess ||
        cudaEventCreate(&w->ev1) != cudaSuccess) {
        delete w;
        return nullptr;
    }
    return w;
}

void release_ws(rsr_handle* h, Workspace* w)
{
    std::lock_guard<std::mutex> g(h->mu);
    h->pool.push_back(w);
}

void free_ws(Workspace* w)
{
    DevBuf* bufs[] = {&w->x0[0], &w->x0[1], &w->x0[2], &w->e8[0], &w->e8[1], &w->e8[2], &w->xg, &w->f32, &w->u1, &w->v1, &w->u2, &w->v2, &w->tta, &w->tiles, &w->img_in, &w->img_out};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    for (auto& e : w->conv_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (w->h_tiles) cudaFreeHost(w->h_tiles);
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    if (w->stream) cudaStreamDestroy(w->stream);
    delete w;
}

struct BatchCtx {
    rsr_handle* h;
    Workspace* ws;
    int tilesize, prepadding;  // snapshot of the handle's parameters taken (under its mutex) when the call started
    long long launches = 0, conv_launches = 0;
    size_t nconv_ev = 0;
};

// residual-stream row maps of a CONV_RDB launch, one per shape group each: block input, RRDB input, output (fp16 plane, extension)
struct RdbMaps { const CUtensorMap *x, *e, *rrx, *rre, *ox, *oe; };

// omaps: the launch's output maps (one per shape group; CONV_STORE16 only)
int conv(BatchCtx& c, const ConvMaps* map, ConvArgs a, int conv_idx, const CUtensorMap* omaps = nullptr, int nomaps = 0, const RdbMaps* rm = nullptr)
{
    rsr_handle* h = c.h;
    const DeviceModel& dm = h->dm;
    a.bias = dm.bias[conv_idx];
    a.cin = h->model.convs[conv_idx].cin;
    a.debug_mask = h->debug_mask;
    a.reverse = h->serpentine ? (int)(c.conv_launches & 1) : 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->time_convs) {
        if (c.nconv_ev >= c.ws->conv_events.size()) {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x));
            CK(cudaEventCreate(&y));
            c.ws->conv_events.push_back({x, y});
        }
        e0 = c.ws->conv_events[c.nconv_ev].first;
        e1 = c.ws->conv_events[c.nconv_ev].second;
        c.nconv_ev++;
        CK(cudaEventRecord(e0, c.ws->stream));
    }
    ConvMaps mm = *map;
    mm.w = dm.wmap[conv_idx];
    for (int g = 0; g < nomaps && g < 4; g++) mm.o[g] = omaps[g];
    if (a.mode == CONV_STORE16 && nomaps < (a.ngroups > 0 ? a.ngroups : 1)) return set_err(h, RSR_ERR_GENERIC, "conv: missing output maps");
    if (a.mode == CONV_RDB) {
        if (!rm) return set_err(h, RSR_ERR_GENERIC, "conv: missing residual maps");
        for (int g = 0; g < (a.ngroups > 0 ? a.ngroups : 1) && g < 4; g++) {
            mm.rx[g] = rm->x[g]; mm.re[g] = rm->e[g]; mm.rrx[g] = rm->rrx[g]; mm.rre[g] = rm->rre[g];
            mm.o[g] = rm->ox[g]; mm.reo[g] = rm->oe[g];
        }
    }
    CK(launch_conv3x3(&mm, a, dm.cout_pad[conv_idx], c.ws->stream));
    if (e1) CK(cudaEventRecord(e1, c.ws->stream));
    c.launches++;
    c.conv_launches++;
    return 0;
}

// ---- one job = up to four same-shape tile groups run through the whole network, ONE launch per layer --------------------------
// A frame's tiles come in up to four padded shapes (interior, right edge, bottom edge, corner).  The shape groups of a job are
// concatenated pixel-major in every activation buffer and each layer runs as one launch over all of them (ConvArgs::grp): the
// per-launch pipeline fill/drain is paid once per layer, not once per layer and shape (measured: -8 % per 1080p frame).  Option
// merge_shapes = 0 plans one job per shape group instead (same code path).
struct Grp {
    size_t off;     // first TileDesc of the group in the call's descriptor array
    int n, H, W;    // tiles, padded tile size
    size_t px_off;  // first pixel of the group in the (LR-resolution) activation buffers of the job
    size_t tta_off; // TTA: float4 offset of the group's slots in the job's TTA buffer (both halves of a chunk share it)
};
struct GatherJob {  // TTA: inverse-transform + average of one chunk of tiles (8 result slots each)
    size_t off;     // the chunk's 8n descriptors in slot order
    int n, mw, mh;
    size_t slot_stride, tta_off;
};
struct MJob {
    std::vector<Grp> gs;
    std::vector<GatherJob> gathers;
    size_t tta_float4 = 0;
};

int run_job(BatchCtx& c, const uint8_t* d_img, int img_w, int img_h, int img_c, uint8_t* d_out, const TileDesc* d_tiles_all,
            const std::vector<Grp>& gs, bool tta_out)
{
    rsr_handle* h = c.h;
    Workspace* ws = c.ws;
    const Model& m = h->model;
    const int G = (int)gs.size();
    if (G < 1 || G > 4) return set_err(h, RSR_ERR_GENERIC, "run_job: 1..4 shape groups per job");
    size_t npx = 0;
    for (const Grp& g : gs) npx += (size_t)g.n * g.H * g.W;
    int r;
    for (int i = 0; i < 3; i++) {
        if ((r = ensure(h, ws->x0[i], npx * 64 * 2))) return r;
        if ((r = ensure(h, ws->e8[i], npx * 64))) return r;
    }
    if ((r = ensure(h, ws->xg, npx * 128 * 2))) return r;
    if ((r = ensure(h, ws->f32, npx * 64 * 4))) return r;
    if ((r = ensure(h, ws->u1, npx * 4 * 64 * 2))) return r;
    if ((r = ensure(h, ws->v1, npx * 4 * 64 * 2))) return r;
    if ((r = ensure(h, ws->u2, npx * 16 * 64 * 2))) return r;
    if ((r = ensure(h, ws->v2, npx * 16 * 64 * 2))) return r;
    cudaStream_t st = ws->stream;
    __half* X0[3] = {(__half*)ws->x0[0].p, (__half*)ws->x0[1].p, (__half*)ws->x0[2].p};
    uint8_t* E8[3] = {(uint8_t*)ws->e8[0].p, (uint8_t*)ws->e8[1].p, (uint8_t*)ws->e8[2].p};
    auto xplane = [&](int k) -> __half* { return (__half*)ws->xg.p + (size_t)(k - 1) * npx * 32; };  // x_k, k = 1..4: [npx][32]
    const int rows8 = conv_bh(32) + 2, rows4 = conv_bh(64) + 2;
    JobMaps* jm = &ws->maps;
    for (int g = 0; g < G; g++) {
        const Grp& gr = gs[g];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < CONV_MAX_KBLOCKS; k++) {
                const __half* base = k < 2 ? X0[i] + gr.px_off * 64 : xplane(k - 1) + gr.px_off * 32;
                const int C = k < 2 ? 64 : 32;
                if ((r = make_map(h, map_slot(jm->d8[i], g, k), base, C, gr.W, gr.H, gr.n, rows8))) return r;
                if ((r = make_map(h, map_slot(jm->d4[i], g, k), base, C, gr.W, gr.H, gr.n, rows4))) return r;
            }
        for (int k = 1; k <= 4; k++)
            if ((r = make_omap(h, &jm->oD[k - 1][g], xplane(k) + gr.px_off * 32, gr.W, gr.H, gr.n, 32, 32))) return r;
        for (int i = 0; i < 3; i++) {
            if ((r = make_omap(h, &jm->rX[i][g], X0[i] + gr.px_off * 64, gr.W, gr.H, gr.n, 64))) return r;
            if ((r = make_emap(h, &jm->rE[i][g], E8[i] + gr.px_off * 64, gr.W, gr.H, gr.n))) return r;
        }
        struct { ConvMaps* cm; void* base; int s; int rows; } tails[4] = {{&jm->u1_4, ws->u1.p, 2, rows4}, {&jm->u2_4, ws->u2.p, 4, rows4},
                                                                          {&jm->v2_4, ws->v2.p, 4, rows4}, {&jm->u2_8, ws->u2.p, 4, rows8}};
        for (auto& t : tails) {
            if ((r = make_map(h, map_slot(*t.cm, g, 0), (__half*)t.base + gr.px_off * (size_t)(t.s * t.s) * 64, 64, t.s * gr.W, t.s * gr.H, gr.n, t.rows)))
                return r;
            *map_slot(*t.cm, g, 1) = *map_slot(*t.cm, g, 0);
        }
        if ((r = make_omap(h, &jm->oV1[g], (__half*)ws->v1.p + gr.px_off * 4 * 64, 2 * gr.W, 2 * gr.H, gr.n, 64))) return r;
        if ((r = make_omap(h, &jm->oV2[g], (__half*)ws->v2.p + gr.px_off * 16 * 64, 4 * gr.W, 4 * gr.H, gr.n, 64))) return r;
        if ((r = make_omap(h, &jm->oU2[g], (__half*)ws->u2.p + gr.px_off * 16 * 64, 4 * gr.W, 4 * gr.H, gr.n, 64))) return r;
    }
    // Conv_0 + preproc, one (cheap, bandwidth-bound) launch per group
    for (const Grp& gr : gs) {
        FirstArgs f{};
        f.img = d_img; f.img_w = img_w; f.img_h = img_h; f.img_c = img_c; f.tiles = d_tiles_all + gr.off;
        f.w = h->dm.w0; f.bias = h->dm.bias[0];
        f.H = gr.H; f.W = gr.W; f.B = gr.n;
        f.d16 = X0[0] + gr.px_off * 64;
        f.ext = E8[0] + gr.px_off * 64;
        f.f32 = (float*)ws->f32.p + gr.px_off * 64;
        CK(launch_conv_first(f, st));
        c.launches++;
    }
    // group table of a launch at scale s (1 for the body, 2 / 4 for the tail) for a conv with cout_pad output channels
    auto fill_groups = [&](ConvArgs& a, int s, int cout_pad) {
        const int band = 2 * conv_bh(cout_pad);
        a.H = a.W = a.B = 0;
        a.ngroups = G;
        int t0 = 0;
        for (int g = 0; g < G; g++) {
            ConvGroup& cg = a.grp[g];
            cg.H = gs[g].H * s; cg.W = gs[g].W * s; cg.B = gs[g].n;
            cg.tiles_x = (cg.W + CONV_BW - 1) / CONV_BW;
            cg.tiles_per_img = cg.tiles_x * ((cg.H + band - 1) / band);
            cg.tile_start = t0;
            cg.desc0 = (int)gs[g].off;
            cg.px_off = (long long)gs[g].px_off * s * s;
            cg.aux_off = (long long)gs[g].tta_off;
            t0 += cg.tiles_per_img * cg.B;
        }
        a.ntiles_multi = t0;
    };
    ConvArgs base{};
    base.f32 = (const float*)ws->f32.p;
    base.tiles = d_tiles_all;
    // X0 plane rotation: `cur` holds the current block's input, `rin` the input of the enclosing RRDB (kept until its last block)
    int cur = 0;
    for (int rrdb = 0; rrdb < m.num_rrdb; rrdb++) {
        const int rin = cur;
        for (int rdb = 0; rdb < 3; rdb++) {
            for (int k = 0; k < 4; k++) {  // x_{k+1} = lrelu(conv([x0..xk])) -> its own 32-channel plane
                ConvArgs a = base;
                fill_groups(a, 1, 32);
                a.mode = CONV_STORE16; a.act = 2; a.slope = m.convs[m.idx_rdb(rrdb, rdb, k)].slope;
                if ((r = conv(c, &jm->d8[cur], a, m.idx_rdb(rrdb, rdb, k), jm->oD[k], G))) return r;
            }
            int nxt = 0;  // the plane that is neither this block's input nor the RRDB's input
            while (nxt == cur || nxt == rin) nxt++;
            ConvArgs a = base;  // x5 -> residual(s) -> x0 of the next block
            fill_groups(a, 1, 64);
            a.mode = CONV_RDB; a.rrdb_end = (rdb == 2);
            a.res16 = X0[cur]; a.ext_in = E8[cur];
            a.rr16 = X0[rin]; a.rr_ext = E8[rin];
            a.out16 = X0[nxt]; a.ext_out = E8[nxt]; a.out_cs = 64; a.out_co = 0;
            const RdbMaps rm{jm->rX[cur], jm->rE[cur], jm->rX[rin], jm->rE[rin], jm->rX[nxt], jm->rE[nxt]};
            if ((r = conv(c, &jm->d4[cur], a, m.idx_rdb(rrdb, rdb, 4), nullptr, 0, &rm))) return r;
            cur = nxt;
        }
    }
    const int tr_out = (cur + 1) % 3;
    {   // trunk conv + fea (BinaryOp ADD), result in another X0 plane
        ConvArgs a = base;
        fill_groups(a, 1, 64);
        a.mode = CONV_TRUNK; a.out16 = X0[tr_out]; a.out_cs = 64; a.out_co = 0;
        if ((r = conv(c, &jm->d4[cur], a, m.idx_trunk()))) return r;
        cur = tr_out;
    }
    // nearest x2 -> conv+lrelu, twice (x4.param:996-999); the first source is plane 0 of the dense buffer
    for (const Grp& gr : gs) {
        CK(launch_upsample2x(X0[cur] + gr.px_off * 64, 64, (__half*)ws->u1.p + gr.px_off * 4 * 64, gr.n, gr.H, gr.W, st));
        c.launches++;
    }
    {
        ConvArgs a = base;
        fill_groups(a, 2, 64);
        a.mode = CONV_STORE16; a.act = 2; a.slope = m.convs[m.idx_up1()].slope;
        if ((r = conv(c, &jm->u1_4, a, m.idx_up1(), jm->oV1, G))) return r;
    }
    for (const Grp& gr : gs) {
        CK(launch_upsample2x((const __half*)ws->v1.p + gr.px_off * 4 * 64, 64, (__half*)ws->u2.p + gr.px_off * 16 * 64, gr.n, 2 * gr.H, 2 * gr.W, st));
        c.launches++;
    }
    {
        ConvArgs a = base;
        fill_groups(a, 4, 64);
        a.mode = CONV_STORE16; a.act = 2; a.slope = m.convs[m.idx_up2()].slope;
        if ((r = conv(c, &jm->u2_4, a, m.idx_up2(), jm->oV2, G))) return r;
        a.slope = m.convs[m.idx_hr()].slope;  // u2 is dead after up2: it takes the HR conv's output
        if ((r = conv(c, &jm->v2_4, a, m.idx_hr(), jm->oU2, G))) return r;
    }
    {   // last conv (64->3, N padded to 16) + postproc / TTA store
        ConvArgs a = base;
        fill_groups(a, 4, 16);
        a.mode = tta_out ? CONV_LAST_F32 : CONV_LAST_U8;
        a.img_out = d_out; a.img_w = img_w * 4; a.img_h = img_h * 4; a.img_c = img_c;
        a.prepad_s = c.prepadding * 4;
        a.tta_out = (float*)ws->tta.p;
        if ((r = conv(c, &jm->u2_8, a, m.idx_last()))) return r;
    }
    return 0;
}

// The schedule of one process() call, computed on the host before anything is enqueued (no CUDA involved: rsr_debug_plan exposes it to the
// CPU tests): the tile grid of the reference, the tiles grouped by padded shape, the groups packed into jobs.
struct CallPlan {
    std::vector<TileDesc> flat;   // every descriptor the kernels index, in device order
    std::vector<MJob> mjobs;      // every job = up to max_groups shape groups, one launch per layer
    size_t max_tta_bytes = 0;
    long long padded = 0;         // network-input pixels of the call
};

// tiles [tile_first, tile_first + tile_count) of the reference's row-major tile grid; `done` (optional) receives their descriptors
void plan_call(const rsr_handle* h, int w, int hh, int T, int P, int tile_first, int tile_count, CallPlan& pl, std::vector<TileDesc>* done)
{
    std::vector<TileDesc>& flat = pl.flat;
    std::vector<MJob>& mjobs = pl.mjobs;
    size_t& max_tta_bytes = pl.max_tta_bytes;
    long long& padded = pl.padded;
    const int xtiles = (w + T - 1) / T, ytiles = (hh + T - 1) / T;
    // tile grid: reference src/realsr.cpp:170-171,:176-187,:235-241 (GPU) == :537-553 (CPU)
    std::map<std::pair<int, int>, std::vector<TileDesc>> groups;
    for (int yi = 0; yi < ytiles; yi++)
        for (int xi = 0; xi < xtiles; xi++) {
            const int tidx = yi * xtiles + xi;
            if (tidx < tile_first || tidx >= tile_first + tile_count) continue;
            const int tw = std::min((xi + 1) * T, w) - xi * T, th = std::min((yi + 1) * T, hh) - yi * T;
            TileDesc t{};
            t.in_x0 = xi * T - P; t.in_y0 = yi * T - P;
            t.out_x0 = xi * T * 4; t.out_y0 = yi * T * 4;
            t.crop_w = tw * 4; t.crop_h = th * 4;
            t.variant = 0; t.base_h = th + 2 * P; t.base_w = tw + 2 * P; t.tta_slot = 0;
            groups[{t.base_h, t.base_w}].push_back(t);
            if (done) done->push_back(t);
        }
    // Every batch of the call is planned first, so that all tile descriptors go to the device in ONE copy from pinned memory
    // and the batches then run back to back on the stream with no host synchronisation in between (same-stream order makes
    // the reuse of the activation buffers safe).  Larger shapes first: the workspace is sized once.
    {
        const double budget = 16e9;  // workspace bytes per job
        const int per_tile = h->tta ? 8 : 1;   // network passes per tile
        const size_t need = h->tta ? 2 : 1;    // group slots per chunk: TTA runs variants 0..3 as H x W and 4..7 as W x H
        const size_t max_groups = h->merge_shapes ? 4 : need;
        MJob cur;
        double cur_bytes = 0;
        size_t cur_px = 0;
        auto flush = [&]() {
            if (!cur.gs.empty()) {
                max_tta_bytes = std::max(max_tta_bytes, cur.tta_float4 * sizeof(float4));
                mjobs.push_back(cur);
            }
            cur = MJob();
            cur_bytes = 0; cur_px = 0;
        };
        for (auto g = groups.rbegin(); g != groups.rend(); ++g) {  // larger shapes first: the workspace is sized once
            const int H = g->first.first, W = g->first.second;
            const std::vector<TileDesc>& tl = g->second;
            // workspace per padded pixel: 3 x (x0 fp16 + ext) + x1..x4 fp16 + fea fp32 + (u1,v1) 4x64x2 fp16 + (u2,v2) 16x64x2 fp16 [+ TTA 16 float4 x2]
            const double bpt = (double)per_tile * H * W * (576.0 + 256.0 + 256.0 + 1024.0 + 4096.0 + (h->tta ? 512.0 : 0.0));
            size_t pos = 0;
            while (pos < tl.size()) {
                long long cap = (long long)((budget - cur_bytes) / bpt);
                if (h->max_batch > 0) cap = std::min(cap, std::max(1LL, h->max_batch / per_tile));
                if (cur.gs.size() + need > max_groups || cap < 1) {
                    if (cur.gs.empty()) cap = 1;  // a single tile larger than the budget still has to run
                    else { flush(); continue; }
                }
                const size_t take = std::min(tl.size() - pos, (size_t)cap);
                if (!h->tta) {
                    cur.gs.push_back({flat.size(), (int)take, H, W, cur_px, 0});
                    flat.insert(flat.end(), tl.begin() + pos, tl.begin() + pos + take);
                    cur_px += take * (size_t)H * W;
                } else {
                    // 8 dihedral variants per tile (src/realsr.cpp:242-318): 0..3 keep the shape, 4..7 run transposed
                    const size_t slot_stride = (size_t)16 * H * W;  // float4 per result slot
                    for (int half = 0; half < 2; half++) {
                        cur.gs.push_back({flat.size(), (int)take * 4, half ? W : H, half ? H : W, cur_px, cur.tta_float4});
                        for (size_t gi = 0; gi < take; gi++)
                            for (int v = half * 4; v < half * 4 + 4; v++) {
                                TileDesc t = tl[pos + gi];
                                t.variant = v;
                                t.tta_slot = (int)(gi * 8 + v);
                                flat.push_back(t);
                            }
                        cur_px += take * 4 * (size_t)H * W;
                    }
                    GatherJob gj{flat.size(), (int)take, 0, 0, slot_stride, cur.tta_float4};
                    for (size_t gi = 0; gi < take; gi++) {
                        gj.mw = std::max(gj.mw, tl[pos + gi].crop_w);
                        gj.mh = std::max(gj.mh, tl[pos + gi].crop_h);
                        for (int v = 0; v < 8; v++) {
                            TileDesc t = tl[pos + gi];
                            t.variant = v;
                            t.tta_slot = (int)(gi * 8 + v);
                            flat.push_back(t);
                        }
                    }
                    cur.gathers.push_back(gj);
                    cur.tta_float4 += take * 8 * slot_stride;
                }
                cur_bytes += take * bpt;
                padded += (long long)take * per_tile * H * W;
                pos += take;
            }
            if (!h->merge_shapes) flush();
        }
        flush();
    }
}

int process_on_device(rsr_handle* h, Workspace* ws, const uint8_t* d_in, int w, int hh, int ch, uint8_t* d_out, long long* padded_out,
                      long long* launches_out, int tile_first, int tile_count, std::vector<TileDesc>* done, size_t* nconv_ev_out, long long* conv_launches_out, int T, int P)
{
    BatchCtx c{h, ws, T, P};
    int r;
    CallPlan pl;
    plan_call(h, w, hh, T, P, tile_first, tile_count, pl, done);
    const std::vector<TileDesc>& flat = pl.flat;
    const std::vector<MJob>& mjobs = pl.mjobs;
    const size_t max_tta_bytes = pl.max_tta_bytes;
    const long long padded = pl.padded;
    if (flat.size() > ws->h_tiles_cap) {
        if (ws->h_tiles) cudaFreeHost(ws->h_tiles);
        ws->h_tiles = nullptr;
        ws->h_tiles_cap = 0;
        CK(cudaMallocHost((void**)&ws->h_tiles, sizeof(TileDesc) * flat.size()));
        ws->h_tiles_cap = flat.size();
    }
    memcpy(ws->h_tiles, flat.data(), sizeof(TileDesc) * flat.size());
    if ((r = ensure(h, ws->tiles, sizeof(TileDesc) * flat.size()))) return r;
    if (max_tta_bytes && (r = ensure(h, ws->tta, max_tta_bytes))) return r;
    CK(cudaMemcpyAsync(ws->tiles.p, ws->h_tiles, sizeof(TileDesc) * flat.size(), cudaMemcpyHostToDevice, ws->stream));
    const TileDesc* d_tiles = (const TileDesc*)ws->tiles.p;
    for (const MJob& mj : mjobs) {
        if ((r = run_job(c, d_in, w, hh, ch, d_out, d_tiles, mj.gs, h->tta != 0))) return r;
        for (const GatherJob& gj : mj.gathers) {
            TtaGatherArgs ga{};
            ga.tta_out = (const float*)((const float4*)ws->tta.p + gj.tta_off); ga.tiles = d_tiles + gj.off; ga.ngroups = gj.n;
            ga.slot_stride = gj.slot_stride; ga.img_out = d_out; ga.img_w = w * 4; ga.img_h = hh * 4; ga.img_c = ch; ga.prepad_s = P * 4;
            CK(launch_tta_gather(ga, gj.mw, gj.mh, ws->stream));
            c.launches++;
        }
    }
    if (ch == 4) {  // alpha: bicubic x4 per un-padded tile region, independent of the network
        AlphaArgs aa{};
        aa.img = d_in; aa.img_out = d_out; aa.img_w = w; aa.img_h = hh; aa.tilesize = T;
        aa.tile_first = tile_first; aa.tile_count = tile_count;
        CK(launch_alpha_bicubic4x(aa, ws->stream));
        c.launches++;
    }
    *padded_out = padded;
    *launches_out = c.launches;
    *conv_launches_out = c.conv_launches;
    *nconv_ev_out = c.nconv_ev;  // the per-conv events are read by the caller AFTER it has synchronised the stream
    return 0;
}

int check_args(rsr_handle* h, const void* in, int w, int hh, int c, const void* out)
{
    if (!h) return RSR_ERR_GENERIC;
    if (!h->loaded) return set_err(h, RSR_ERR_GENERIC, "rsr_process: no model loaded");
    if (!in || !out || w <= 0 || hh <= 0) return set_err(h, RSR_ERR_GENERIC, "rsr_process: bad image arguments");
    if (c != 3 && c != 4) return set_err(h, RSR_ERR_GENERIC, "rsr_process: only 3-channel RGB and 4-channel RGBA images are supported");
    return 0;
}

}  // namespace

// ================================================= C ABI ===========================================================
extern "C" {

rsr_handle* rsr_create(int gpuid, int tta_mode, int num_threads)
{
    (void)num_threads;
    rsr_handle* h = new (std::nothrow) rsr_handle();
    if (!h) return nullptr;
    h->gpuid = gpuid;
    h->tta = tta_mode ? 1 : 0;
    return h;
}

void rsr_destroy(rsr_handle* h)
{
    if (!h) return;
    if (h->gpuid >= 0) cudaSetDevice(h->gpuid);
    for (Workspace* w : h->pool) free_ws(w);
    if (h->dm.blob) cudaFree(h->dm.blob);
    delete h;
}

int rsr_load(rsr_handle* h, const char* parampath, const char* modelpath)
{
    if (!h || !parampath || !modelpath) return RSR_ERR_GENERIC;
    if (h->gpuid < 0) return set_err(h, RSR_ERR_NO_DEVICE, "gpuid -1 (ncnn CPU path) is not provided by librealsr_b200; use a CUDA device ordinal");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || h->gpuid >= ndev)
        return set_err(h, RSR_ERR_NO_DEVICE, "invalid gpu device " + std::to_string(h->gpuid) + " (" + std::to_string(ndev) + " CUDA devices visible)");
    CK(cudaSetDevice(h->gpuid));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->gpuid));
    if (prop.major != 10) return set_err(h, RSR_ERR_NO_DEVICE, std::string("device ") + prop.name + " is not sm_100 (this library is B200-only)");
    Model m;
    int r = load_model(parampath, modelpath, m);
    if (r) return set_err(h, RSR_ERR_MODEL, m.error);
    // pack + upload: one blob, 1024-B aligned sub-allocations (bulk copies need 16 B; keep slabs nicely aligned)
    const int nconv = (int)m.convs.size();
    std::vector<std::vector<uint16_t>> packed_rows(nconv);
    std::vector<size_t> off_b(nconv, 0), off_r(nconv, 0);
    std::vector<int> nrows(nconv, 0);
    DeviceModel dm;
    dm.cout_pad.assign(nconv, 0);
    size_t total = 0;
    auto bump = [&](size_t bytes) { size_t o = total; total += (bytes + 1023) & ~(size_t)1023; return o; };
    const size_t off_w0 = bump(64 * 27 * 4);
    for (int i = 0; i < nconv; i++) {
        const ConvLayer& c = m.convs[i];
        dm.cout_pad[i] = c.cout <= 16 ? 16 : (c.cout <= 32 ? 32 : 64);
        if (i > 0) {
            if (c.cout > 64 || (c.cin % 32) || c.cin > 32 * CONV_MAX_KBLOCKS) return set_err(h, RSR_ERR_MODEL, "unsupported conv shape in " + c.name);
            nrows[i] = pack_conv_weight_rows(c, dm.cout_pad[i], packed_rows[i]);
            off_r[i] = bump(packed_rows[i].size() * 2);
        }
        off_b[i] = bump(64 * 4);
    }
    if (h->dm.blob) { cudaFree(h->dm.blob); h->dm = DeviceModel(); }
    CK(cudaMalloc(&dm.blob, total));
    std::vector<uint8_t> host(total, 0);
    memcpy(host.data() + off_w0, m.convs[0].weight.data(), 64 * 27 * 4);
    for (int i = 0; i < nconv; i++) {
        if (i > 0) memcpy(host.data() + off_r[i], packed_rows[i].data(), packed_rows[i].size() * 2);
        memcpy(host.data() + off_b[i], m.convs[i].bias.data(), m.convs[i].cout * 4);
    }
    CK(cudaMemcpy(dm.blob, host.data(), total, cudaMemcpyHostToDevice));
    dm.w0 = (float*)((uint8_t*)dm.blob + off_w0);
    dm.bias.assign(nconv, nullptr);
    dm.wmap.resize(nconv);
    for (int i = 1; i < nconv; i++) {  // box = a CTA's half of one (K-block, ky) weight block: N/2 rows (N = 192 / 96 / 64)
        const int rr = make_wmap(h, &dm.wmap[i], (uint8_t*)dm.blob + off_r[i], nrows[i], dm.cout_pad[i] == 64 ? 96 : (dm.cout_pad[i] == 32 ? 48 : 32));
        if (rr) return rr;
    }
    for (int i = 0; i < nconv; i++) dm.bias[i] = (float*)((uint8_t*)dm.blob + off_b[i]);
    h->model = std::move(m);
    h->dm = dm;
    h->loaded = true;
    return 0;
}

int rsr_set(rsr_handle* h, int scale, int tilesize, int prepadding)
{
    if (!h) return RSR_ERR_GENERIC;
    if (scale != 4) return set_err(h, RSR_ERR_GENERIC, "invalid scale argument (only 4 is supported, like the reference)");
    if (tilesize < 32) return set_err(h, RSR_ERR_GENERIC, "invalid tilesize argument (must be >= 32)");
    if (prepadding < 0 || prepadding > 64) return set_err(h, RSR_ERR_GENERIC, "invalid prepadding");
    // RealSR::process() const forwards its public members on every call, possibly from several proc threads at once
    // (host/realsr.h, reference src/main.cpp:811-828): the write is serialised and every call works on its own snapshot
    std::lock_guard<std::mutex> g(h->mu);
    h->scale = scale;
    h->tilesize = tilesize;
    h->prepadding = prepadding;
    return 0;
}

static int process_impl(rsr_handle* h, const uint8_t* in, int w, int hh, int c, uint8_t* out, bool host_buffers, int tile_first = 0,
                        int tile_count = -1)
{
    int r = check_args(h, in, w, hh, c, out);
    if (r) return r;
    int T, P;
    {
        std::lock_guard<std::mutex> g(h->mu);
        T = h->tilesize;
        P = h->prepadding;
    }
    const int ntiles_img = ((w + T - 1) / T) * ((hh + T - 1) / T);
    const bool whole = tile_count < 0;
    if (whole) { tile_first = 0; tile_count = ntiles_img; }
    if (tile_first < 0 || tile_count < 0 || tile_first + tile_count > ntiles_img)
        return set_err(h, RSR_ERR_GENERIC, "rsr_process_tiles: tile range outside the image's tile grid");
    if (tile_count == 0) return 0;
    CK(cudaSetDevice(h->gpuid));
    Workspace* ws = acquire_ws(h);
    if (!ws) return set_err(h, RSR_ERR_CUDA, "cannot create stream/events");
    const size_t in_bytes = (size_t)w * hh * c, out_bytes = in_bytes * 16;
    const uint8_t* d_in = in;
    uint8_t* d_out = out;
    long long padded = 0, launches = 0;
    auto body = [&]() -> int {
        int rr;
        if (host_buffers) {
            if ((rr = ensure(h, ws->img_in, in_bytes))) return rr;
            if ((rr = ensure(h, ws->img_out, out_bytes))) return rr;
            d_in = (const uint8_t*)ws->img_in.p;
            d_out = (uint8_t*)ws->img_out.p;
        }
        CK(cudaEventRecord(ws->ev0, ws->stream));
        if (host_buffers) CK(cudaMemcpyAsync(ws->img_in.p, in, in_bytes, cudaMemcpyHostToDevice, ws->stream));
        std::vector<TileDesc> done;
        size_t nconv_ev = 0;
        long long conv_launches = 0;
        if ((rr = process_on_device(h, ws, d_in, w, hh, c, d_out, &padded, &launches, tile_first, tile_count, whole ? nullptr : &done, &nconv_ev, &conv_launches, T, P))) return rr;
        if (host_buffers) {
            if (whole) {
                CK(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ws->stream));
            } else {  // only this handle's tile regions travel back: other handles fill the rest of the caller's buffer
                const size_t pitch = (size_t)w * 4 * c;
                for (const TileDesc& t : done) {
                    const size_t off = (size_t)t.out_y0 * pitch + (size_t)t.out_x0 * c;
                    CK(cudaMemcpy2DAsync(out + off, pitch, d_out + off, pitch, (size_t)t.crop_w * c, (size_t)t.crop_h, cudaMemcpyDeviceToHost, ws->stream));
                }
            }
        }
        CK(cudaEventRecord(ws->ev1, ws->stream));
        CK(cudaStreamSynchronize(ws->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));
        double conv_ms = 0;
        for (size_t i = 0; i < nconv_ev; i++) {  // time_convs: every pair has completed (stream synchronised above); no silent zeros
            float t = 0;
            CK(cudaEventElapsedTime(&t, ws->conv_events[i].first, ws->conv_events[i].second));
            conv_ms += t;
        }
        std::lock_guard<std::mutex> g(h->mu);
        h->last_ms = ms;
        if (h->time_convs) h->last_conv_ms = conv_ms;
        h->last_conv_launches = conv_launches;
        h->last_padded = padded;
        h->launches += launches;
        return 0;
    };
    r = body();
    if (r) cudaStreamSynchronize(ws->stream);
    release_ws(h, ws);
    return r;
}

int rsr_process(rsr_handle* h, const uint8_t* in, int w, int h_, int c, uint8_t* out) { return process_impl(h, in, w, h_, c, out, true); }

int rsr_process_device(rsr_handle* h, const uint8_t* d_in, int w, int h_, int c, uint8_t* d_out)
{
    return process_impl(h, d_in, w, h_, c, d_out, false);
}

int rsr_tile_count(const rsr_handle* h, int w, int h_)
{
    if (!h || w <= 0 || h_ <= 0) return 0;
    int T;
    {
        std::lock_guard<std::mutex> g(const_cast<rsr_handle*>(h)->mu);
        T = h->tilesize;
    }
    if (T <= 0) return 0;
    return ((w + T - 1) / T) * ((h_ + T - 1) / T);
}

int rsr_process_tiles(rsr_handle* h, const uint8_t* in, int w, int h_, int c, uint8_t* out, int tile_first, int tile_count)
{
    if (h && tile_count < 0) return set_err(h, RSR_ERR_GENERIC, "rsr_process_tiles: negative tile_count");
    return process_impl(h, in, w, h_, c, out, true, tile_first, tile_count);
}

const char* rsr_last_error(const rsr_handle* h)
{
    if (!h) return "null handle";
    // several proc threads share one handle (main.cpp:811-828): hand out a per-thread copy taken under the lock, so that another thread's
    // failure cannot reallocate the string this caller is still reading (valid until this thread's next rsr_last_error call)
    static thread_local std::string copy;
    std::lock_guard<std::mutex> g(const_cast<rsr_handle*>(h)->mu);
    copy = h->err;
    return copy.c_str();
}
// statistics: written by every finishing call under the handle mutex, so the getters take it too (proc threads share a handle)
#define RSR_LOCKED_GET(h, expr, none)                                   \
    do {                                                                \
        if (!(h)) return (none);                                        \
        std::lock_guard<std::mutex> g(const_cast<rsr_handle*>(h)->mu);  \
        return (expr);                                                  \
    } while (0)
long long rsr_kernel_launches(const rsr_handle* h) { RSR_LOCKED_GET(h, h->launches, 0LL); }
double rsr_last_ms(const rsr_handle* h, int what) { RSR_LOCKED_GET(h, what == 1 ? h->last_conv_ms : h->last_ms, 0.0); }
long long rsr_last_padded_pixels(const rsr_handle* h) { RSR_LOCKED_GET(h, h->last_padded, 0LL); }
long long rsr_last_conv_launches(const rsr_handle* h) { RSR_LOCKED_GET(h, h->last_conv_launches, 0LL); }

int rsr_set_option(rsr_handle* h, const char* key, long long value)
{
    if (!h || !key) return RSR_ERR_GENERIC;
    if (strcmp(key, "max_batch") && strcmp(key, "serpentine") && strcmp(key, "merge_shapes") && strcmp(key, "debug_mask") && strcmp(key, "time_convs"))
        return set_err(h, RSR_ERR_GENERIC, std::string("unknown option ") + key);
    std::lock_guard<std::mutex> g(h->mu);  // (calls already running keep reading their options unlocked: set options between calls)
    if (!strcmp(key, "max_batch")) h->max_batch = value;
    else if (!strcmp(key, "serpentine")) h->serpentine = value ? 1 : 0;
    else if (!strcmp(key, "merge_shapes")) h->merge_shapes = value ? 1 : 0;
    else if (!strcmp(key, "debug_mask")) h->debug_mask = (int)value;
    else if (!strcmp(key, "time_convs")) h->time_convs = (int)value;
    return 0;
}

int rsr_check_model(const char* parampath, const char* modelpath, int* num_rrdb, int* num_convs, char* err, int errlen)
{
    if (!parampath || !modelpath) return RSR_ERR_GENERIC;
    Model m;
    const int r = load_model(parampath, modelpath, m);
    if (err && errlen > 0) snprintf(err, (size_t)errlen, "%s", m.error.c_str());
    if (r) return RSR_ERR_MODEL;
    if (num_rrdb) *num_rrdb = m.num_rrdb;
    if (num_convs) *num_convs = (int)m.convs.size();
    return 0;
}

long long rsr_debug_plan(rsr_handle* h, int w, int hh, int tile_first, int tile_count, long long* out, long long capacity)
{
    if (!h || w <= 0 || hh <= 0 || tile_first < 0) return RSR_ERR_GENERIC;
    int T, P;
    {
        std::lock_guard<std::mutex> g(h->mu);
        T = h->tilesize;
        P = h->prepadding;
    }
    if (tile_count < 0) tile_count = ((w + T - 1) / T) * ((hh + T - 1) / T);
    CallPlan pl;
    plan_call(h, w, hh, T, P, tile_first, tile_count, pl, nullptr);
    std::vector<long long> v = {(long long)pl.mjobs.size(), (long long)pl.flat.size(), pl.padded, (long long)pl.max_tta_bytes};
    for (const MJob& j : pl.mjobs) {
        v.push_back((long long)j.gs.size());
        v.push_back((long long)j.gathers.size());
        v.push_back((long long)j.tta_float4);
        for (const Grp& g : j.gs) {
            const long long e[6] = {(long long)g.off, g.n, g.H, g.W, (long long)g.px_off, (long long)g.tta_off};
            v.insert(v.end(), e, e + 6);
        }
        for (const GatherJob& g : j.gathers) {
            const long long e[6] = {(long long)g.off, g.n, g.mw, g.mh, (long long)g.slot_stride, (long long)g.tta_off};
            v.insert(v.end(), e, e + 6);
        }
    }
    for (const TileDesc& t : pl.flat) {
        const long long e[10] = {t.in_x0, t.in_y0, t.out_x0, t.out_y0, t.crop_w, t.crop_h, t.variant, t.base_h, t.base_w, t.tta_slot};
        v.insert(v.end(), e, e + 10);
    }
    if (out && capacity >= (long long)v.size()) memcpy(out, v.data(), v.size() * sizeof(long long));
    return (long long)v.size();
}

long long rsr_debug_pack_rows(const char* parampath, const char* modelpath, int conv_index, uint16_t* packed, long long capacity)
{
    if (!parampath || !modelpath) return RSR_ERR_GENERIC;
    Model m;
    if (load_model(parampath, modelpath, m)) return RSR_ERR_MODEL;
    if (conv_index < 1 || conv_index >= (int)m.convs.size()) return RSR_ERR_GENERIC;
    const ConvLayer& c = m.convs[conv_index];
    std::vector<uint16_t> out;
    const int rows = pack_conv_weight_rows(c, c.cout <= 16 ? 16 : (c.cout <= 32 ? 32 : 64), out);
    if (packed) {
        if ((long long)out.size() > capacity) return RSR_ERR_GENERIC;
        memcpy(packed, out.data(), out.size() * 2);
    }
    return rows;
}

int rsr_debug_conv(rsr_handle* h, int conv_index, const uint16_t* in16, int H, int W, float* out32)
{
    if (!h || !h->loaded) return RSR_ERR_GENERIC;
    if (conv_index < 1 || conv_index >= (int)h->model.convs.size()) return set_err(h, RSR_ERR_GENERIC, "bad conv index");
    CK(cudaSetDevice(h->gpuid));
    const ConvLayer& cl = h->model.convs[conv_index];
    const int npad = h->dm.cout_pad[conv_index];
    const size_t npx = (size_t)H * W;
    __half *d_in = nullptr, *d_out = nullptr;
    float* d_f = nullptr;
    TileDesc* d_td = nullptr;
    // the engine's dense-buffer layout: plane 0 = channels 0..63 [px][64], then one [px][32] plane per further 32-channel K-block
    const int nkb = cl.cin / 32;
    CK(cudaMalloc(&d_in, npx * (size_t)std::max(cl.cin, 64) * 2));
    CK(cudaMalloc(&d_out, npx * 64 * 2));
    CK(cudaMalloc(&d_f, npx * 16));
    CK(cudaMalloc(&d_td, sizeof(TileDesc)));
    {
        std::vector<uint16_t> planes(npx * (size_t)std::max(cl.cin, 64), 0);
        for (size_t p = 0; p < npx; p++)
            for (int ci = 0; ci < cl.cin; ci++) {
                const size_t dst = ci < 64 ? p * 64 + ci : npx * 64 + (size_t)((ci - 64) >> 5) * npx * 32 + p * 32 + (ci & 31);
                planes[dst] = in16[p * cl.cin + ci];
            }
        CK(cudaMemcpy(d_in, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
    }
    TileDesc td{};
    CK(cudaMemcpy(d_td, &td, sizeof td, cudaMemcpyHostToDevice));
    ConvMaps map;
    int r = 0;
    for (int k = 0; k < nkb && k < CONV_MAX_KBLOCKS; k++)
        if ((r = make_map(h, &map.m[k], k < 2 ? d_in : d_in + npx * 64 + (size_t)(k - 2) * npx * 32, k < 2 ? 64 : 32, W, H, 1, conv_bh(npad) + 2))) return r;
    map.w = h->dm.wmap[conv_index];
    ConvArgs a{};
    a.bias = h->dm.bias[conv_index]; a.cin = cl.cin;
    a.H = H; a.W = W; a.B = 1;
    a.act = cl.act; a.slope = cl.slope;
    a.out16 = d_out; a.out_cs = npad; a.out_co = 0;
    a.tiles = d_td; a.tta_out = d_f;
    a.mode = npad == 16 ? CONV_LAST_F32 : CONV_STORE16;
    if (npad != 16 && (r = make_omap(h, &map.o[0], d_out, W, H, 1, npad, npad))) return r;
    a.debug_mask = h->debug_mask;
    long long* d_ts = nullptr;
    uint8_t* d_ext = nullptr;
    if (getenv("RSR_DBG_RDB") && npad == 64) {  // bring-up: time the Cout=64 conv with its real (CONV_RDB) epilogue; value 2 = at an RRDB end
        CK(cudaMalloc(&d_ext, npx * 64 * 2));
        CK(cudaMemset(d_ext, 0, npx * 64 * 2));
        a.mode = CONV_RDB;
        a.res16 = d_in; a.ext_in = d_ext;
        a.rr16 = d_in; a.rr_ext = d_ext;
        a.ext_out = d_ext + npx * 64;
        a.rrdb_end = atoi(getenv("RSR_DBG_RDB")) > 1;
        if ((r = make_omap(h, &map.rx[0], d_in, W, H, 1, 64)) || (r = make_emap(h, &map.re[0], d_ext, W, H, 1)) ||
            (r = make_emap(h, &map.reo[0], d_ext + npx * 64, W, H, 1)))
            return r;
        map.rrx[0] = map.rx[0];
        map.rre[0] = map.re[0];
    }
    if (getenv("RSR_DBG_TS")) {
        CK(cudaMalloc(&d_ts, 64 * 8 * sizeof(long long)));
        CK(cudaMemset(d_ts, 0, 64 * 8 * sizeof(long long)));
        a.dbg_ts = d_ts;
    }
    for (int rep = 0; rep < (d_ts ? 2 : 1); rep++) CK(launch_conv3x3(&map, a, npad, 0));
    CK(cudaDeviceSynchronize());
    if (d_ts) {  // bring-up: per-tile pipeline timestamps of cluster 0 (clock64 of the leader SM)
        std::vector<long long> ts(64 * 8);
        CK(cudaMemcpy(ts.data(), d_ts, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        const long long t0 = ts[0];
        fprintf(stderr, "tile: mma_wait_start mma_acc_ready mma_a0_ready mma_issued | epi_wait_start epi_acc_full epi_handback [epi_done]   (clk since first)\n");
        for (int i = 0; i < 64 && ts[i * 8 + 0]; i++)
            fprintf(stderr, "%3d: %8lld %8lld %8lld %8lld | %8lld %8lld %8lld %8lld\n", i, ts[i * 8 + 0] - t0, ts[i * 8 + 1] - t0, ts[i * 8 + 2] - t0,
                    ts[i * 8 + 3] - t0, ts[i * 8 + 4] - t0, ts[i * 8 + 5] - t0, ts[i * 8 + 6] - t0, ts[i * 8 + 7] ? ts[i * 8 + 7] - t0 : 0LL);
        cudaFree(d_ts);
    }
    if (d_ext) cudaFree(d_ext);
    if (npad == 16) {
        std::vector<float> tmp(npx * 4);
        CK(cudaMemcpy(tmp.data(), d_f, npx * 16, cudaMemcpyDeviceToHost));
        for (size_t p = 0; p < npx; p++)
            for (int j = 0; j < cl.cout; j++) out32[p * cl.cout + j] = tmp[p * 4 + j];
    } else {
        std::vector<__half> tmp(npx * npad);
        CK(cudaMemcpy(tmp.data(), d_out, npx * npad * 2, cudaMemcpyDeviceToHost));
        for (size_t p = 0; p < npx; p++)
            for (int j = 0; j < cl.cout; j++) out32[p * cl.cout + j] = __half2float(tmp[p * npad + j]);
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_f); cudaFree(d_td);
    {
        std::lock_guard<std::mutex> g(h->mu);
        h->launches += 1;
    }
    return 0;
}

}  // extern "C"
