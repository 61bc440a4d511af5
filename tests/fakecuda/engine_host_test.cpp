// Drives the REAL engine host code (csrc/engine.cu + csrc/model.cpp, compiled as C++) against fake_cuda.cpp:
//   engine_host_test <models-dir> sweep     one thread, many image sizes / tile sizes / TTA / RGBA / tile ranges / options  (ASan + UBSan build)
//   engine_host_test <models-dir> threads   several proc threads on ONE handle + rsr_set / rsr_set_option / rsr_last_error  (TSan build)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/realsr_b200.h"
extern "C" long long fake_cuda_launches();
extern "C" long long fake_cuda_maps();
extern "C" long long fake_cuda_live_allocations();

#define REQUIRE(c)                                                                  \
    do {                                                                            \
        if (!(c)) { fprintf(stderr, "%s:%d: REQUIRE(%s) failed\n", __FILE__, __LINE__, #c); exit(1); } \
    } while (0)

static rsr_handle* make(const std::string& models, const char* model, int gpu, int tta)
{
    rsr_handle* h = rsr_create(gpu, tta, 1);
    REQUIRE(h);
    const std::string p = models + "/" + model + "/x4.param", b = models + "/" + model + "/x4.bin";
    if (rsr_load(h, p.c_str(), b.c_str())) { fprintf(stderr, "load: %s\n", rsr_last_error(h)); exit(1); }
    return h;
}

static void run(rsr_handle* h, int w, int hh, int c, int tile)
{
    REQUIRE(rsr_set(h, 4, tile, 10) == 0);
    std::vector<uint8_t> in((size_t)w * hh * c, 7), out((size_t)w * hh * c * 16, 0);
    const long long l0 = rsr_kernel_launches(h);
    if (rsr_process(h, in.data(), w, hh, c, out.data())) { fprintf(stderr, "process %dx%dx%d t%d: %s\n", w, hh, c, tile, rsr_last_error(h)); exit(1); }
    REQUIRE(rsr_kernel_launches(h) > l0);
    REQUIRE(rsr_last_padded_pixels(h) >= (long long)w * hh);
}

int main(int argc, char** argv)
{
    if (argc < 3) return 64;
    const std::string models = argv[1], mode = argv[2];
    if (mode == "image" && argc == 11) {  // engine_host_test <models> image <model> <in.raw> w h c tile tta <out.raw>   (functional build)
        const int w = atoi(argv[5]), hh = atoi(argv[6]), c = atoi(argv[7]), tile = atoi(argv[8]), tta = atoi(argv[9]);
        rsr_handle* h = make(models, argv[3], 0, tta);
        REQUIRE(rsr_set(h, 4, tile, 10) == 0);
        if (const char* o = getenv("RSR_TWIN_OPTS")) {  // key=value[,key=value...]
            std::string s = o;
            for (size_t p = 0; p < s.size();) {
                const size_t e = s.find(',', p), q = s.find('=', p);
                REQUIRE(q != std::string::npos);
                REQUIRE(rsr_set_option(h, s.substr(p, q - p).c_str(), atoll(s.c_str() + q + 1)) == 0);
                p = e == std::string::npos ? s.size() : e + 1;
            }
        }
        std::vector<uint8_t> in((size_t)w * hh * c), out((size_t)w * hh * c * 16, 0);
        FILE* fp = fopen(argv[4], "rb");
        REQUIRE(fp && fread(in.data(), 1, in.size(), fp) == in.size());
        fclose(fp);
        if (getenv("RSR_TWIN_SHARDS")) {  // the image in tile ranges, as sharded.process_sharded does with one handle per GPU
            const int n = rsr_tile_count(h, w, hh), parts = atoi(getenv("RSR_TWIN_SHARDS"));
            for (int k = 0; k < parts; k++) {
                const int first = (int)((long long)n * k / parts), last = (int)((long long)n * (k + 1) / parts);
                if (last > first && rsr_process_tiles(h, in.data(), w, hh, c, out.data(), first, last - first)) { fprintf(stderr, "process_tiles: %s\n", rsr_last_error(h)); return 1; }
            }
        } else if (rsr_process(h, in.data(), w, hh, c, out.data())) { fprintf(stderr, "process: %s\n", rsr_last_error(h)); return 1; }
        fp = fopen(argv[10], "wb");
        REQUIRE(fp && fwrite(out.data(), 1, out.size(), fp) == out.size());
        fclose(fp);
        printf("image ok: %lld launches\n", rsr_kernel_launches(h));
        rsr_destroy(h);
        return 0;
    }
    if (mode == "sweep") {
        for (int tta = 0; tta < 2; tta++) {
            rsr_handle* h = make(models, tta ? "models-DF2K" : "models-DF2K_JPEG", tta, tta);
            const int sizes[][2] = {{1, 1}, {12, 11}, {64, 64}, {61, 45}, {200, 200}, {201, 399}, {256, 256}, {500, 33}, {640, 360}};
            for (auto& s : sizes)
                for (int tile : {32, 64, 200, 400})
                    for (int c : {3, 4}) {
                        if (tta && (long long)s[0] * s[1] > 300 * 300) continue;
                        run(h, s[0], s[1], c, tile);
                    }
            if (!tta) {  // the BASELINE configurations themselves (workspace up to the 16 GB job budget: virtual memory only, nothing touches it)
                run(h, 1920, 1080, 3, 400);
                REQUIRE(rsr_last_padded_pixels(h) == 2302800 && rsr_last_conv_launches(h) == 350);
                run(h, 3840, 2160, 3, 200);
                REQUIRE(rsr_last_padded_pixels(h) == 10091200);
            } else {
                run(h, 1280, 720, 3, 200);
                REQUIRE(rsr_last_padded_pixels(h) == 9088000);
            }
            // options that change the packing, a device-resident call, tile ranges of the sharded path, a model reload
            REQUIRE(rsr_set_option(h, "merge_shapes", 0) == 0);
            run(h, 333, 222, 3, 100);
            REQUIRE(rsr_set_option(h, "merge_shapes", 1) == 0 && rsr_set_option(h, "max_batch", 3) == 0 && rsr_set_option(h, "serpentine", 0) == 0);
            run(h, 333, 222, 4, 100);
            REQUIRE(rsr_set_option(h, "max_batch", 0) == 0 && rsr_set_option(h, "time_convs", 1) == 0);
            run(h, 100, 140, 3, 200);
            REQUIRE(rsr_last_conv_launches(h) == 350 * (tta ? 1 : 1) || rsr_last_conv_launches(h) > 0);
            REQUIRE(rsr_set_option(h, "time_convs", 0) == 0 && rsr_set_option(h, "nope", 1) != 0);
            REQUIRE(rsr_set(h, 4, 64, 10) == 0);
            const int w = 150, hh = 130, n = rsr_tile_count(h, w, hh);
            REQUIRE(n == 9);
            std::vector<uint8_t> in((size_t)w * hh * 4, 9), out((size_t)w * hh * 64, 0);
            for (int first = 0; first < n; first += 4) REQUIRE(rsr_process_tiles(h, in.data(), w, hh, 4, out.data(), first, first + 4 <= n ? 4 : n - first) == 0);
            REQUIRE(rsr_process_tiles(h, in.data(), w, hh, 4, out.data(), 7, 5) != 0);  // beyond the grid: an error, not an overrun
            REQUIRE(rsr_process(h, nullptr, w, hh, 3, out.data()) != 0 && rsr_process(h, in.data(), w, hh, 2, out.data()) != 0);
            rsr_destroy(h);
        }
        REQUIRE(fake_cuda_live_allocations() == 0);  // rsr_destroy returned every device allocation
        printf("sweep ok: %lld launches, %lld tensor maps checked\n", fake_cuda_launches(), fake_cuda_maps());
        return 0;
    }
    if (mode == "threads") {
        rsr_handle* h = make(models, "models-DF2K", 0, 0);
        std::atomic<int> bad{0};
        std::vector<std::thread> th;
        for (int t = 0; t < 6; t++)
            th.emplace_back([&, t] {
                for (int it = 0; it < 6; it++) {
                    const int w = 40 + 37 * t + it, hh = 30 + 11 * it + t, c = (t & 1) ? 4 : 3;
                    std::vector<uint8_t> in((size_t)w * hh * c, (uint8_t)t), out((size_t)w * hh * c * 16);
                    // like RealSR::process() const (host/realsr.h): forward the public members, then process -- from every proc thread
                    if (rsr_set(h, 4, 64, 10) || rsr_process(h, in.data(), w, hh, c, out.data())) bad++;
                    if (it == 3 && rsr_process(h, in.data(), w, hh, 7, out.data()) == 0) bad++;  // a failing call per thread ...
                    if (strlen(rsr_last_error(h)) > 200) bad++;                                  // ... and readers of the message
                    (void)rsr_kernel_launches(h);
                }
            });
        for (auto& x : th) x.join();
        REQUIRE(bad == 0);
        rsr_destroy(h);
        printf("threads ok: %lld launches\n", fake_cuda_launches());
        return 0;
    }
    return 64;
}
