// bench_png.cpp -- throughput of the C++ command line's PNG writer (host/png_codec.h) on one SR-like image, CPU only:
//   python tools/bench_png.py            (writes /tmp/enc.raw: 3840x2160 bicubic-upscaled smooth noise, builds and runs this file)
//   ./bench_png <threads>
#include <chrono>
#include <thread>
#include "png_codec.h"
int main(int argc, char** argv) {
    const int w = 3840, h = 2160, c = 3; const int th = atoi(argv[1]);
    std::vector<uint8_t> px((size_t)w * h * c);
    FILE* fp = fopen("/tmp/enc.raw", "rb"); if (!fp || fread(px.data(), 1, px.size(), fp) != px.size()) return 2; fclose(fp);
    for (int rep = 0; rep < 5; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        bool ok = rsrhost::png_encode_file("/dev/shm/enc_b.png", px.data(), w, h, c, 1, th);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.3f s  %.1f MP/s ok=%d\n", th, dt, w * h / 1e6 / dt, ok);
    }
}
