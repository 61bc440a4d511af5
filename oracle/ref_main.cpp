// oracle/ref_main.cpp -- driver for the REAL reference CPU path. TEST INFRASTRUCTURE ONLY.
//
// Links the reference's own RealSR::process_cpu (taken verbatim at build time from
// /root/reference/src/realsr.cpp:525-838 into process_cpu.inc by oracle/build_ref.py) against the vendored
// ncnn b4ba207 built CPU-only.  The small class below only re-declares the members process_cpu uses
// (reference: src/realsr.h:13-42) and restates the CPU-relevant part of RealSR::load
// (src/realsr.cpp:43-47 option flags, :75-76 load_param/load_model, :129-140 bicubic layer for alpha).
//
// usage: oracle_ref <x4.param> <x4.bin> <in.raw> <w> <h> <c> <out.raw> <tilesize> <tta 0|1> <threads> [prepadding]
//   in.raw  : w*h*c bytes, tightly packed HWC (RGB / RGBA order, as the Linux build of the reference sees them)
//   out.raw : 4w*4h*c bytes
// prints one JSON line {"seconds": process_cpu wall time, ...} to stdout.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>

#include "net.h"
#include "layer.h"
#include "cpu.h"

class RealSR
{
public:
    RealSR(bool _tta_mode, int num_threads)
    {
        net.opt.num_threads = num_threads;
        bicubic_4x = 0;
        tta_mode = _tta_mode;
    }
    ~RealSR()
    {
        if (bicubic_4x)
        {
            bicubic_4x->destroy_pipeline(net.opt);
            delete bicubic_4x;
        }
    }
    int load(const std::string& parampath, const std::string& modelpath)
    {
        net.opt.use_vulkan_compute = false;
        net.opt.use_fp16_packed = true;
        net.opt.use_fp16_storage = false;
        net.opt.use_fp16_arithmetic = false;
        net.opt.use_int8_storage = true;
        if (net.load_param(parampath.c_str())) return -1;
        if (net.load_model(modelpath.c_str())) return -2;
        bicubic_4x = ncnn::create_layer("Interp");
        ncnn::ParamDict pd;
        pd.set(0, 3); // bicubic
        pd.set(1, 4.f);
        pd.set(2, 4.f);
        bicubic_4x->load_param(pd);
        bicubic_4x->create_pipeline(net.opt);
        return 0;
    }
    int process_cpu(const ncnn::Mat& inimage, ncnn::Mat& outimage) const;

public:
    int scale;
    int tilesize;
    int prepadding;

private:
    ncnn::Net net;
    ncnn::Layer* bicubic_4x;
    bool tta_mode;
};

#include "process_cpu.inc"

int main(int argc, char** argv)
{
    if (argc < 11)
    {
        fprintf(stderr, "usage: %s param bin in.raw w h c out.raw tilesize tta threads [prepadding]\n", argv[0]);
        return 2;
    }
    const int w = atoi(argv[4]), h = atoi(argv[5]), c = atoi(argv[6]);
    const int tilesize = atoi(argv[8]), tta = atoi(argv[9]), threads = atoi(argv[10]);
    const int prepadding = argc > 11 ? atoi(argv[11]) : 10;

    std::vector<unsigned char> in((size_t)w * h * c);
    FILE* fp = fopen(argv[3], "rb");
    if (!fp || fread(in.data(), 1, in.size(), fp) != in.size())
    {
        fprintf(stderr, "cannot read %s\n", argv[3]);
        return 3;
    }
    fclose(fp);

    ncnn::set_omp_num_threads(threads);
    RealSR realsr(tta != 0, threads);
    if (realsr.load(argv[1], argv[2]))
    {
        fprintf(stderr, "model load failed\n");
        return 4;
    }
    realsr.scale = 4;
    realsr.tilesize = tilesize;
    realsr.prepadding = prepadding;

    // same Mat construction as the reference's load thread (src/main.cpp:275-276)
    ncnn::Mat inimage(w, h, (void*)in.data(), (size_t)c, c);
    ncnn::Mat outimage(w * 4, h * 4, (size_t)c, c);

    auto t0 = std::chrono::steady_clock::now();
    realsr.process_cpu(inimage, outimage);
    auto t1 = std::chrono::steady_clock::now();
    double sec = std::chrono::duration<double>(t1 - t0).count();

    fp = fopen(argv[7], "wb");
    if (!fp || fwrite(outimage.data, 1, (size_t)w * 4 * h * 4 * c, fp) != (size_t)w * 4 * h * 4 * c)
    {
        fprintf(stderr, "cannot write %s\n", argv[7]);
        return 5;
    }
    fclose(fp);
    printf("{\"seconds\": %.6f, \"w\": %d, \"h\": %d, \"c\": %d, \"tilesize\": %d, \"tta\": %d, \"threads\": %d, \"out_mp\": %.6f}\n",
           sec, w, h, c, tilesize, tta, threads, (double)w * 4 * h * 4 / 1e6);
    return 0;
}
