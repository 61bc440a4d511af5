// realsr.h -- `class RealSR` with the reference's exact public interface (reference: src/realsr.h:13-42), implemented
// as a thin header-only wrapper over the C ABI of librealsr_b200.so (include/realsr_b200.h).
//
// Drop-in use: the reference's main.cpp needs no change around the tile engine --
//     realsr[i] = new RealSR(gpuid[i], tta_mode, num_threads);      // src/main.cpp:782
//     realsr[i]->load(paramfullpath, modelfullpath);                // :784
//     realsr[i]->scale = scale; ->tilesize = ...; ->prepadding = ...  // :788-790
//     realsr->process(v.inimage, v.outimage);                       // :325, concurrently from several proc threads
// Differences, all supersets: load()/process() return a negative code on failure instead of always 0
// (realsr.cpp:142,:522), gpuid is a CUDA ordinal, and gpuid == -1 (ncnn CPU mode) is an error: process_cpu() is not
// provided by the B200 library.
#ifndef REALSR_H
#define REALSR_H

#include <string>

#ifndef RSR_USE_REAL_NCNN
#include "mat.h"
#else
#include "mat.h"  // the real ncnn header when it is first on the include path
#endif

#include "../../include/realsr_b200.h"

class RealSR
{
public:
    RealSR(int gpuid, bool tta_mode = false, int num_threads = 1) : scale(4), tilesize(200), prepadding(10)
    {
        handle = rsr_create(gpuid, tta_mode ? 1 : 0, num_threads);
    }
    ~RealSR() { rsr_destroy(handle); }

    int load(const std::string& parampath, const std::string& modelpath) { return rsr_load(handle, parampath.c_str(), modelpath.c_str()); }

    // in: ncnn::Mat(w, h, pixels, elemsize = channels, elempack = channels); out: ncnn::Mat(4w, 4h, channels, channels)
    int process(const ncnn::Mat& inimage, ncnn::Mat& outimage) const
    {
        // the reference reads its public members at call time (realsr.cpp:158-160): forward them on every call.
        // rsr_set only stores three ints under the handle; concurrent callers pass identical values.
        int r = rsr_set(handle, scale, tilesize, prepadding);
        if (r) return r;
        if (!outimage.data || outimage.w != inimage.w * scale || outimage.h != inimage.h * scale || outimage.elempack != inimage.elempack) return -1;
        return rsr_process(handle, (const unsigned char*)inimage.data, inimage.w, inimage.h, inimage.elempack, (unsigned char*)outimage.data);
    }

    // The reference's ncnn CPU path (src/realsr.cpp:525-838) is the test oracle of this repo, not part of the product.
    int process_cpu(const ncnn::Mat&, ncnn::Mat&) const { return RSR_ERR_NO_DEVICE; }

    const char* last_error() const { return rsr_last_error(handle); }

public:
    // realsr parameters
    int scale;
    int tilesize;
    int prepadding;

private:
    RealSR(const RealSR&);
    RealSR& operator=(const RealSR&);
    rsr_handle* handle;
};

#endif  // REALSR_H
