/*
 * realsr_b200.h -- C ABI of the B200-native RealSR 4x tile engine (librealsr_b200.so).
 *
 * This is the drop-in boundary for the reference's tile engine, `class RealSR`
 * (/root/reference/src/realsr.h:13-42, implemented in /root/reference/src/realsr.cpp).  Each entry point names
 * the reference interface it replaces.  Plain pointers and sizes only; no C++ or torch types cross the boundary.
 * The C++ wrapper `class RealSR` with the reference's exact signatures lives in
 * realsr-ncnn-vulkan_b200/host/realsr.h; the Python mirror in realsr-ncnn-vulkan_b200/realsr.py.
 *
 * Error convention (superset of the reference, which always returns 0 and only logs: realsr.cpp:142,:522,:837):
 * 0 on success, negative on failure; the message is available through rsr_last_error().  ncnn-style codes are
 * kept where they apply: -1 generic / bad argument, -100 allocation failure.  Nothing throws, nothing exits.
 *
 * Threading: rsr_process() is re-entrant on one handle, like `RealSR::process() const`, which the reference calls
 * concurrently from its proc threads (src/main.cpp:811-828).  Each call takes a private stream + workspace from a
 * per-handle pool; weights are immutable after rsr_load().
 */
#ifndef REALSR_B200_H
#define REALSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RSR_API __attribute__((visibility("default")))
#else
#define RSR_API
#endif

typedef struct rsr_handle rsr_handle;

/* RSR_ERR_* : return codes */
#define RSR_OK 0
#define RSR_ERR_GENERIC (-1)
#define RSR_ERR_MODEL (-2)    /* .param/.bin unreadable, or not the RRDBNet x4 graph */
#define RSR_ERR_CUDA (-3)     /* a CUDA call failed; rsr_last_error() has the CUDA message */
#define RSR_ERR_NO_DEVICE (-4) /* gpuid invalid, or gpuid == -1 (the ncnn CPU path is not part of this library) */
#define RSR_ERR_ALLOC (-100)

/* Replaces RealSR::RealSR(int gpuid, bool tta_mode, int num_threads)  (realsr.h:16, realsr.cpp:13-23).
 * gpuid is a CUDA device ordinal.  num_threads is accepted for signature parity and ignored (it only sizes the
 * reference's OpenMP CPU path).  Returns NULL only on host allocation failure; an invalid device is reported by
 * rsr_load().  */
RSR_API rsr_handle* rsr_create(int gpuid, int tta_mode, int num_threads);

/* Replaces RealSR::~RealSR()  (realsr.cpp:25-36). */
RSR_API void rsr_destroy(rsr_handle* h);

/* Replaces RealSR::load(parampath, modelpath)  (realsr.h:22, realsr.cpp:38-143): reads the reference's
 * x4.param / x4.bin unchanged, checks the graph, packs and uploads the weights. */
RSR_API int rsr_load(rsr_handle* h, const char* parampath, const char* modelpath);

/* Replaces the public data members scale / tilesize / prepadding (realsr.h:31-33), which the reference sets after
 * load() (src/main.cpp:788-790).  scale must be 4 (src/main.cpp:533-537), tilesize >= 32, prepadding >= 0. */
RSR_API int rsr_set(rsr_handle* h, int scale, int tilesize, int prepadding);

/* Replaces RealSR::process(const ncnn::Mat& in, ncnn::Mat& out) const  (realsr.h:25, realsr.cpp:145-523).
 *   in  : w*h*c bytes, tightly packed HWC (row stride w*c), RGB order -- what the reference's Linux build passes
 *         (ncnn::Mat(w, h, pixels, elemsize=c, elempack=c), src/main.cpp:275); owned by the caller.
 *   out : (4w)*(4h)*c bytes, tightly packed HWC, allocated by the caller before the call (src/main.cpp:276) and
 *         written in place.
 *   c   : 3 (RGB) or 4 (RGBA: the alpha plane is up-scaled bicubically per tile like realsr.cpp:431-442, not by the network).
 * Host pointers.  Blocks until `out` is complete.
 * Images with w or h <= prepadding (10): the reference's reflect-101 border (ncnn copy_make_border, one fold) then indexes outside
 * the image -- its result is undefined; this library clamps the folded coordinate instead, so such images are processed without error
 * but are outside the parity contract (the smallest well-defined sizes, 12x11, are golden-tested). */
RSR_API int rsr_process(rsr_handle* h, const uint8_t* in, int w, int h_, int c, uint8_t* out);

/* Same computation with `in` / `out` already resident in device memory of the handle's GPU (no H2D/D2H copies).
 * Used by bench.py to separate the compute stage from the PCIe copies; not part of the reference API. */
RSR_API int rsr_process_device(rsr_handle* h, const uint8_t* d_in, int w, int h_, int c, uint8_t* d_out);

/* Tile-granular sharding of ONE image over several handles (one per GPU) -- SURVEY 8e; the reference only shards whole
 * images over GPUs (src/main.cpp:811-828).  Tiles are numbered row-major over the reference's tile grid
 * (yi * xtiles + xi, src/realsr.cpp:170-171,:176-187 with the handle's current tilesize).
 *   rsr_tile_count   : number of tiles of a w x h image.
 *   rsr_process_tiles: rsr_process restricted to tiles [tile_first, tile_first + tile_count).  `in` is the WHOLE image;
 *                      only the output regions of those tiles are written to `out` (the whole 4w x 4h x c buffer), so
 *                      concurrent calls on different handles with disjoint ranges assemble the result in place with no
 *                      exchange step.  The union over a partition of the grid is byte-identical to rsr_process. */
RSR_API int rsr_tile_count(const rsr_handle* h, int w, int h_);
RSR_API int rsr_process_tiles(rsr_handle* h, const uint8_t* in, int w, int h_, int c, uint8_t* out, int tile_first, int tile_count);

/* Last error message of this handle ("" if none).  Thread-safe: the returned pointer is a per-thread copy, valid until this thread's
 * next rsr_last_error() call. */
RSR_API const char* rsr_last_error(const rsr_handle* h);

/* ---- introspection used by the tests and the benchmark (no reference counterpart) ------------------------------ */

/* Number of kernels launched by rsr_process* calls on this handle so far. */
RSR_API long long rsr_kernel_launches(const rsr_handle* h);
/* Device time (ms, CUDA events on the call's stream) of the last rsr_process* call: whole call, and the part spent
 * in the tensor-core conv kernels (only measured when rsr_set_option("time_convs", 1) is on). */
RSR_API double rsr_last_ms(const rsr_handle* h, int what /*0 total, 1 conv kernels*/);
/* Network-input ("padded LR") pixels processed by the last call, x8 in TTA mode: the roofline work unit. */
RSR_API long long rsr_last_padded_pixels(const rsr_handle* h);
/* Tensor-core conv kernel launches of the last rsr_process* call (the launches rsr_last_ms(h, 1) sums over). */
RSR_API long long rsr_last_conv_launches(const rsr_handle* h);
/* Engine options: "max_batch" (tiles per job, 0 = auto), "merge_shapes" (1 default: one launch per layer over all shape groups of a
 * frame; 0: one job per shape group), "serpentine" (1 default: every second conv walks its tiles backwards, L2 reuse),
 * "time_convs" (CUDA events around every conv launch), "debug_mask" (timing experiments only, results become garbage: 1 = skip MMAs, 2 = skip
 * activation loads, 4 = skip weight loads, 8 = skip epilogue stores).  Unknown keys return RSR_ERR_GENERIC.  Set options between calls:
 * a rsr_process* call that is already running may or may not see the new value. */
RSR_API int rsr_set_option(rsr_handle* h, const char* key, long long value);

/* Test hook: runs ONE 3x3 convolution of the loaded model (index in file order, >= 1) on the tensor-core kernel.
 *   in16  : host fp16 [H][W][cin], out32: host fp32 [H][W][cout] = bias + conv (+ leaky relu if the layer has it),
 *           rounded to fp16 exactly as the engine stores it. */
RSR_API int rsr_debug_conv(rsr_handle* h, int conv_index, const uint16_t* in16, int H, int W, float* out32);

/* Host-only checks (no GPU needed): parse + verify a .param/.bin pair exactly as rsr_load() does.
 * On success returns 0 and fills num_rrdb / num_convs; on failure returns RSR_ERR_MODEL and writes the reason to err. */
RSR_API int rsr_check_model(const char* parampath, const char* modelpath, int* num_rrdb, int* num_convs, char* err, int errlen);
/* Host-only: conv #conv_index's weights (file order, >= 1) as the plain row-major [rows][64] fp16 matrix the conv kernel fetches by TMA
 * (rows grouped per (64-channel plane, ky): Cout=64 -> 192 rows, Cout<=32 -> 96, Cout<=16 -> 64; see model.h).
 * Returns the number of rows (packed may be NULL), or a negative code. */
RSR_API long long rsr_debug_pack_rows(const char* parampath, const char* modelpath, int conv_index, uint16_t* packed, long long capacity);

/* Host-only: the schedule rsr_process* would run for a w x h image with the handle's current tile size, prepadding, TTA mode and options
 * (tiles [tile_first, tile_first + tile_count) of the row-major tile grid; tile_count < 0 = all): serialised as 64-bit integers
 *   njobs, ndescs, padded_pixels, tta_bytes,
 *   per job: ngroups, ngathers, tta_float4, ngroups x {first_desc, ntiles, H, W, first_pixel, tta_offset}, ngathers x {first_desc, ntiles, max_crop_w, max_crop_h, slot_stride, tta_offset},
 *   ndescs x {in_x0, in_y0, out_x0, out_y0, crop_w, crop_h, variant, base_h, base_w, tta_slot}.
 * Returns the number of integers (out may be NULL), or a negative code.  No model and no GPU needed. */
RSR_API long long rsr_debug_plan(rsr_handle* h, int w, int h_, int tile_first, int tile_count, long long* out, long long capacity);

#ifdef __cplusplus
}
#endif
#endif /* REALSR_B200_H */
