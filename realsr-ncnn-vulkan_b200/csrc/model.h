// model.h -- loads the reference's shipped model files unchanged and lowers the graph to a static schedule.
//
// File formats (reference, read-only spec):
//   x4.param : ncnn text graph, /root/reference/src/ncnn/src/net.cpp:1374-1589, key=value dictionary
//              paramdict.cpp:246-340 (ids <= -23300 are arrays "n,v0,v1,...").
//   x4.bin   : per Convolution, in layer order: u32 tag (0x01306B47 => fp16 payload padded to 4 B, 0 => raw fp32),
//              weights OIHW, then Cout fp32 biases; modelbin.cpp:88-296, layer/convolution.cpp:65-99.
// The shipped graphs (models-DF2K, models-DF2K_JPEG) are RRDBNet(nf=64, gc=32, nb=23) with two nearest-x2
// upsamples: 351 Convolution / 276 Split / 276 Concat / 92 Eltwise / 2 Interp / 1 BinaryOp / 1 Input.  The loader
// parses the full text graph, then *verifies* it has exactly that topology (any nb) before lowering it; anything
// else is rejected with an error (generic graphs are a "next" row).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace rsr {

struct ConvLayer {
    std::string name;
    int cin = 0, cout = 0;
    int act = 0;        // 0 none, 2 leaky relu (ncnn activation_type)
    float slope = 0.f;
    std::vector<float> weight;  // OIHW fp32 (decoded)
    std::vector<uint16_t> weight_h;  // OIHW fp16 bit patterns exactly as stored in the .bin
    std::vector<float> bias;
};

struct Model {
    int num_rrdb = 0;            // 23 for the shipped models
    std::vector<ConvLayer> convs;  // file order: Conv_0, nb x 3 x {c1..c5}, trunk, up1, up2, hr, last
    std::string error;

    // index helpers into convs
    int idx_first() const { return 0; }
    int idx_rdb(int rrdb, int rdb, int k) const { return 1 + (rrdb * 3 + rdb) * 5 + k; }
    int idx_trunk() const { return 1 + num_rrdb * 15; }
    int idx_up1() const { return idx_trunk() + 1; }
    int idx_up2() const { return idx_trunk() + 2; }
    int idx_hr() const { return idx_trunk() + 3; }
    int idx_last() const { return idx_trunk() + 4; }
};

// returns 0 on success; on failure m.error holds a one-line reason and a negative code is returned
int load_model(const std::string& parampath, const std::string& binpath, Model& m);

// Weight rows for the conv kernel (conv3x3.cu): plain row-major [rows][64] fp16 (TMA applies the swizzle), rows
// grouped per (64-channel plane, ky): Cout=64 -> 192 rows [W(ky,0) | W(ky,1) | W(ky,2)]; Cout<=32 -> 96 rows
// [W(ky,0) | W(ky,1) | W(ky,2)] (32 each); Cout<=16 -> 64 rows [16 | 16 | 16 | 0].  Returns the number of rows.
int pack_conv_weight_rows(const ConvLayer& c, int cout_pad, std::vector<uint16_t>& out);

}  // namespace rsr
