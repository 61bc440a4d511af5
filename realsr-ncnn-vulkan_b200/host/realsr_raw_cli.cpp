// realsr_raw_cli.cpp -- minimal caller written exactly like the reference's proc stage (src/main.cpp:778-791,:311-331):
// constructs RealSR, loads a model, sets the public parameters, runs process() from `jobs` threads on one object.
// Raw u8 HWC files in/out (the codecs and the full CLI are "next" rows).
//   realsr_raw_cli <model-dir> <in.raw> <w> <h> <out.raw> [-t tilesize] [-g gpu] [-x] [-j jobs]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "realsr.h"

int main(int argc, char** argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s model-dir in.raw w h out.raw [-t tilesize] [-g gpu] [-x] [-j jobs]\n", argv[0]);
        return 2;
    }
    const std::string model = argv[1];
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    int tilesize = 200, gpu = 0, tta = 0, jobs = 1;
    for (int i = 6; i < argc; i++) {
        if (!strcmp(argv[i], "-t") && i + 1 < argc) tilesize = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) gpu = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-j") && i + 1 < argc) jobs = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-x")) tta = 1;
    }
    const int channels = 3, scale = 4;
    std::vector<unsigned char> pixeldata((size_t)w * h * channels);
    FILE* fp = fopen(argv[2], "rb");
    if (!fp || fread(pixeldata.data(), 1, pixeldata.size(), fp) != pixeldata.size()) {
        fprintf(stderr, "decode image %s failed\n", argv[2]);
        return 3;
    }
    fclose(fp);

    RealSR realsr(gpu, tta != 0, 1);
    if (realsr.load(model + "/x4.param", model + "/x4.bin")) {
        fprintf(stderr, "load failed: %s\n", realsr.last_error());
        return 4;
    }
    realsr.scale = scale;
    realsr.tilesize = tilesize;
    realsr.prepadding = 10;

    std::vector<ncnn::Mat> outs;
    for (int j = 0; j < jobs; j++) outs.push_back(ncnn::Mat(w * scale, h * scale, (size_t)channels, channels));
    const ncnn::Mat inimage(w, h, (void*)pixeldata.data(), (size_t)channels, channels);
    std::vector<int> rc(jobs, 0);
    std::vector<std::thread> th;
    for (int j = 0; j < jobs; j++) th.emplace_back([&, j] { rc[j] = realsr.process(inimage, outs[j]); });
    for (auto& t : th) t.join();
    for (int j = 0; j < jobs; j++) {
        if (rc[j]) {
            fprintf(stderr, "process failed (%d): %s\n", rc[j], realsr.last_error());
            return 5;
        }
        if (j && memcmp(outs[j].data, outs[0].data, (size_t)w * h * channels * scale * scale)) {
            fprintf(stderr, "job %d produced a different image\n", j);
            return 6;
        }
    }
    fp = fopen(argv[5], "wb");
    if (!fp || fwrite(outs[0].data, 1, (size_t)w * h * channels * scale * scale, fp) != (size_t)w * h * channels * scale * scale) {
        fprintf(stderr, "encode image %s failed\n", argv[5]);
        return 7;
    }
    fclose(fp);
    fprintf(stderr, "%s -> %s done\n", argv[2], argv[5]);
    return 0;
}
