// png_tool.cpp -- exercises host/png_codec.h, host/jpeg_codec.h and host/webp_codec.h (by file extension) from the command line (used by tests/test_cpp_cli.py):
//   png_tool decode in.png out.raw     -> prints "w h c", writes w*h*c bytes
//   png_tool encode in.raw w h c out.png [zlib-level [threads]]
#include <stdio.h>
#include <stdlib.h>

#include "jpeg_codec.h"
#include "png_codec.h"
#include "webp_codec.h"

int main(int argc, char** argv)
{
    if (argc == 4 && std::string(argv[1]) == "decode") {
        int w, h, c;
        std::vector<uint8_t> px;
        std::string why;
        const std::string in = argv[2];
        const bool jpg = in.size() > 4 && (in.substr(in.size() - 4) == ".jpg" || in.substr(in.size() - 5) == ".jpeg");
        const bool webp = in.size() > 5 && in.substr(in.size() - 5) == ".webp";
        if (!(webp ? rsrhost::webp_decode_file(in, w, h, c, px, &why)
                   : jpg ? rsrhost::jpeg_decode_file(in, w, h, c, px, &why) : rsrhost::png_decode_file(in, w, h, c, px, &why))) {
            fprintf(stderr, "decode failed: %s\n", why.c_str());
            return 1;
        }
        FILE* fp = fopen(argv[3], "wb");
        if (!fp || fwrite(px.data(), 1, px.size(), fp) != px.size()) return 2;
        fclose(fp);
        printf("%d %d %d\n", w, h, c);
        return 0;
    }
    if (argc >= 7 && argc <= 9 && std::string(argv[1]) == "encode") {
        const int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]);
        std::vector<uint8_t> px((size_t)w * h * c);
        FILE* fp = fopen(argv[2], "rb");
        if (!fp || fread(px.data(), 1, px.size(), fp) != px.size()) return 2;
        fclose(fp);
        const std::string outp = argv[6];
        if (outp.size() > 5 && outp.substr(outp.size() - 5) == ".webp") return rsrhost::webp_encode_file(outp, px.data(), w, h, c) ? 0 : 1;
        if (outp.size() > 4 && outp.substr(outp.size() - 4) == ".jpg") return rsrhost::jpeg_encode_file(outp, px.data(), w, h, c, argc >= 8 ? atoi(argv[7]) : 100) ? 0 : 1;
        return rsrhost::png_encode_file(outp, px.data(), w, h, c, argc >= 8 ? atoi(argv[7]) : 1, argc >= 9 ? atoi(argv[8]) : 1) ? 0 : 1;
    }
    fprintf(stderr, "usage: png_tool decode in.png out.raw | png_tool encode in.raw w h c out.png [level]\n");
    return 64;
}
