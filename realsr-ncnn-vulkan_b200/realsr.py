"""Python mirror of the reference's tile-engine class (reference: /root/reference/src/realsr.h:13-42).

Same names, argument meaning and call order as the C++ class the reference's main.cpp drives
(src/main.cpp:778-791: construct, load, then set scale/tilesize/prepadding; :325: process(in, out)).  All work
happens in librealsr_b200.so (hand-written CUDA for sm_100a) through ctypes; there is NO CPU fallback: if the
library is missing or no B200 is visible the calls fail loudly.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class RealSRError(RuntimeError):
    pass


def lib_path():
    """librealsr_b200.so next to this file; RSR_LIB overrides it (A/B of two builds of the library, tools/ab_builds.sh)"""
    return os.environ.get("RSR_LIB") or os.path.join(_HERE, "librealsr_b200.so")


def load_library():
    """Loads librealsr_b200.so (built in-tree by csrc/build.sh / __graft_entry__.build()). Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise RealSRError("librealsr_b200.so not found at %s -- build it with `python __graft_entry__.py build` "
                          "(there is no CPU fallback)" % p)
    lib = ctypes.CDLL(p)
    vp, ci, cc, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_longlong
    lib.rsr_create.restype = vp
    lib.rsr_create.argtypes = [ci, ci, ci]
    lib.rsr_destroy.restype = None
    lib.rsr_destroy.argtypes = [vp]
    lib.rsr_load.restype = ci
    lib.rsr_load.argtypes = [vp, cc, cc]
    lib.rsr_set.restype = ci
    lib.rsr_set.argtypes = [vp, ci, ci, ci]
    lib.rsr_process.restype = ci
    lib.rsr_process.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.rsr_process_device.restype = ci
    lib.rsr_process_device.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.rsr_tile_count.restype = ci
    lib.rsr_tile_count.argtypes = [vp, ci, ci]
    lib.rsr_process_tiles.restype = ci
    lib.rsr_process_tiles.argtypes = [vp, vp, ci, ci, ci, vp, ci, ci]
    lib.rsr_last_error.restype = cc
    lib.rsr_last_error.argtypes = [vp]
    lib.rsr_kernel_launches.restype = ll
    lib.rsr_kernel_launches.argtypes = [vp]
    lib.rsr_last_ms.restype = ctypes.c_double
    lib.rsr_last_ms.argtypes = [vp, ci]
    lib.rsr_last_padded_pixels.restype = ll
    lib.rsr_last_padded_pixels.argtypes = [vp]
    lib.rsr_last_conv_launches.restype = ll
    lib.rsr_last_conv_launches.argtypes = [vp]
    lib.rsr_set_option.restype = ci
    lib.rsr_set_option.argtypes = [vp, cc, ll]
    lib.rsr_debug_conv.restype = ci
    lib.rsr_debug_conv.argtypes = [vp, ci, vp, ci, ci, vp]
    lib.rsr_debug_plan.restype = ll
    lib.rsr_debug_plan.argtypes = [vp, ci, ci, ci, ci, vp, ll]
    _LIB = lib
    return lib


class RealSR:
    """RealSR(gpuid, tta_mode=False, num_threads=1) -- reference: src/realsr.h:16.

    Public data members `scale`, `tilesize`, `prepadding` are plain attributes set after load(), exactly like the
    reference (src/realsr.h:31-33, src/main.cpp:788-790)."""

    def __init__(self, gpuid, tta_mode=False, num_threads=1):
        self._lib = load_library()
        self._h = self._lib.rsr_create(int(gpuid), 1 if tta_mode else 0, int(num_threads))
        if not self._h:
            raise RealSRError("rsr_create failed")
        self.scale = 4
        self.tilesize = 200
        self.prepadding = 10
        self.tta_mode = bool(tta_mode)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.rsr_destroy(h)

    def last_error(self):
        return self._lib.rsr_last_error(self._h).decode()

    def load(self, parampath, modelpath):
        """int RealSR::load(parampath, modelpath) -- reference: src/realsr.cpp:38-143. Returns 0 on success."""
        return self._lib.rsr_load(self._h, os.fsencode(parampath), os.fsencode(modelpath))

    def _push_params(self):
        r = self._lib.rsr_set(self._h, int(self.scale), int(self.tilesize), int(self.prepadding))
        if r:
            raise RealSRError(self.last_error())

    def process(self, inimage, outimage):
        """int RealSR::process(const ncnn::Mat& in, ncnn::Mat& out) const -- reference: src/realsr.cpp:145-523.

        inimage: uint8 ndarray (h, w, c), c = 3 (RGB) or 4 (RGBA), C-contiguous; outimage: preallocated uint8 ndarray (4h, 4w, c),
        written in place (the reference's caller allocates it too, src/main.cpp:276).  Returns 0 on success,
        a negative code otherwise (message in last_error())."""
        if inimage.dtype != np.uint8 or inimage.ndim != 3 or not inimage.flags.c_contiguous:
            raise RealSRError("inimage must be a C-contiguous uint8 HWC array")
        h, w, c = inimage.shape
        if outimage.dtype != np.uint8 or outimage.shape != (h * 4, w * 4, c) or not outimage.flags.c_contiguous:
            raise RealSRError("outimage must be a preallocated C-contiguous uint8 array of shape (4h, 4w, c)")
        self._push_params()
        return self._lib.rsr_process(self._h, inimage.ctypes.data, w, h, c, outimage.ctypes.data)

    def tile_count(self, w, h):
        """number of tiles of a w x h image at the current tilesize (reference tile grid, src/realsr.cpp:170-171)"""
        self._push_params()
        return int(self._lib.rsr_tile_count(self._h, int(w), int(h)))

    def process_tiles(self, inimage, outimage, tile_first, tile_count):
        """process() restricted to tiles [tile_first, tile_first + tile_count) (row-major over the tile grid): only their
        regions of `outimage` are written.  Building block of sharded.process_sharded (one image over several GPUs)."""
        if inimage.dtype != np.uint8 or inimage.ndim != 3 or not inimage.flags.c_contiguous:
            raise RealSRError("inimage must be a C-contiguous uint8 HWC array")
        h, w, c = inimage.shape
        if outimage.dtype != np.uint8 or outimage.shape != (h * 4, w * 4, c) or not outimage.flags.c_contiguous:
            raise RealSRError("outimage must be a preallocated C-contiguous uint8 array of shape (4h, 4w, c)")
        self._push_params()
        return self._lib.rsr_process_tiles(self._h, inimage.ctypes.data, w, h, c, outimage.ctypes.data, int(tile_first), int(tile_count))

    def process_ptr(self, in_ptr, w, h, c, out_ptr):
        """process() on raw HOST pointers (e.g. pinned torch tensors): same contract as the C ABI rsr_process."""
        self._push_params()
        return self._lib.rsr_process(self._h, ctypes.c_void_p(in_ptr), w, h, c, ctypes.c_void_p(out_ptr))

    def process_device(self, d_in_ptr, w, h, c, d_out_ptr):
        """Same computation on buffers already resident on the handle's GPU (raw device pointers)."""
        self._push_params()
        return self._lib.rsr_process_device(self._h, ctypes.c_void_p(d_in_ptr), w, h, c, ctypes.c_void_p(d_out_ptr))

    # ---- introspection (no reference counterpart) ----
    def set_option(self, key, value):
        return self._lib.rsr_set_option(self._h, key.encode(), int(value))

    def kernel_launches(self):
        return int(self._lib.rsr_kernel_launches(self._h))

    def last_ms(self, what=0):
        return float(self._lib.rsr_last_ms(self._h, what))

    def last_padded_pixels(self):
        return int(self._lib.rsr_last_padded_pixels(self._h))

    def last_conv_launches(self):
        return int(self._lib.rsr_last_conv_launches(self._h))

    def plan(self, w, h, tile_first=0, tile_count=-1):
        """the schedule process() would run for a w x h image with the current parameters / options (host only, no model or GPU needed):
        {"padded": px, "tta_bytes": n, "descs": (n, 10) int array of tile descriptors [in_x0, in_y0, out_x0, out_y0, crop_w, crop_h, variant,
        base_h, base_w, tta_slot], "jobs": [{"groups": [dict(off, n, H, W, px_off, tta_off)], "gathers": [dict(off, n, mw, mh, slot_stride,
        tta_off)], "tta_float4": n}]}"""
        self._push_params()
        n = self._lib.rsr_debug_plan(self._h, w, h, tile_first, tile_count, None, 0)
        if n < 0:
            raise RealSRError("rsr_debug_plan failed: %d" % n)
        buf = np.zeros(n, dtype=np.int64)
        assert self._lib.rsr_debug_plan(self._h, w, h, tile_first, tile_count, buf.ctypes.data, n) == n
        v = buf.tolist()
        njobs, ndesc, padded, tta_bytes = v[:4]
        pos, jobs = 4, []
        for _ in range(njobs):
            ng, nga, tta4 = v[pos:pos + 3]
            pos += 3
            groups = [dict(zip(("off", "n", "H", "W", "px_off", "tta_off"), v[pos + 6 * i:pos + 6 * i + 6])) for i in range(ng)]
            pos += 6 * ng
            gathers = [dict(zip(("off", "n", "mw", "mh", "slot_stride", "tta_off"), v[pos + 6 * i:pos + 6 * i + 6])) for i in range(nga)]
            pos += 6 * nga
            jobs.append({"groups": groups, "gathers": gathers, "tta_float4": tta4})
        descs = np.array(v[pos:pos + 10 * ndesc], dtype=np.int64).reshape(ndesc, 10)
        assert pos + 10 * ndesc == n
        return {"padded": padded, "tta_bytes": tta_bytes, "descs": descs, "jobs": jobs}

    def debug_conv(self, conv_index, x16, cout):
        """runs one conv layer of the loaded model on the tensor-core kernel: x16 (H, W, cin) float16 -> (H, W, cout) float32"""
        x16 = np.ascontiguousarray(x16, dtype=np.float16)
        H, W, _ = x16.shape
        out = np.zeros((H, W, cout), dtype=np.float32)
        r = self._lib.rsr_debug_conv(self._h, conv_index, x16.ctypes.data, H, W, out.ctypes.data)
        if r:
            raise RealSRError(self.last_error())
        return out
