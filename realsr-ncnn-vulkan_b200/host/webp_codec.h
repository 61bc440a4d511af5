// webp_codec.h -- WebP reader (lossless VP8L here, lossy VP8 in vp8_codec.h, alpha planes of lossy files) and lossless writer for
// the C++ command line (host/realsr_cli.cpp).
//
// The reference links libwebp: it decodes any WebP input and always SAVES lossless (src/webp_image.h:60-77,
// WebPEncodeLosslessRGB / RGBA).  libwebp is not copied; this is an independent implementation of the "WebP Lossless
// Bitstream Specification":
//   decode: RIFF / VP8L (also inside a VP8X container), all four transforms (predictor with its 14 modes, cross-colour,
//           subtract-green, colour indexing with pixel packing), meta prefix codes, colour cache, LZ77 backward
//           references with the 2-D distance map; lossy files: the VP8 key frame (vp8_codec.h) plus, inside a VP8X container, the
//           ALPH chunk (raw or VP8L-coded alpha with the horizontal / vertical / gradient prediction filters);
//   encode: subtract-green + predictor transform (mode chosen per 16x16 block by the residual entropy proxy) and one set of
//           length-limited Huffman codes; backward references only as run-length copies of the previous pixel (flat areas
//           of drawings) -- super-resolved photographs have few longer exact repeats.
// Host plumbing only: no pixel of the super-resolution path is computed here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "vp8_codec.h"

namespace rsrhost {

namespace vp8l {

struct BitReader {
    const uint8_t* p;
    size_t n, pos = 0;  // pos in bits
    bool eos = false;
    uint32_t get(int nbits)
    {
        uint32_t v = 0;
        for (int i = 0; i < nbits; i++) {
            const size_t byte = pos >> 3;
            if (byte >= n) { eos = true; return v; }
            v |= (uint32_t)((p[byte] >> (pos & 7)) & 1) << i;
            pos++;
        }
        return v;
    }
};

// canonical prefix code, decoded bit by bit (first bit read = most significant bit of the code)
struct Code {
    std::vector<int> count, first_code, first_sym, syms;
    int single = -1;  // a code with exactly one used symbol consumes no bits
    bool build(const std::vector<int>& len)
    {
        int used = 0, last = -1;
        for (size_t i = 0; i < len.size(); i++)
            if (len[i]) { used++; last = (int)i; }
        single = -1;
        if (used == 0) { single = 0; return true; }
        if (used == 1) { single = last; return true; }
        count.assign(16, 0);
        for (int l : len) {
            if (l < 0 || l > 15) return false;
            if (l) count[l]++;
        }
        // completeness (Kraft equality)
        long long kraft = 0;
        for (int l = 1; l <= 15; l++) kraft += (long long)count[l] << (15 - l);
        if (kraft != (1LL << 15)) return false;
        first_code.assign(16, 0);
        first_sym.assign(16, 0);
        int code = 0, k = 0;
        for (int l = 1; l <= 15; l++) {
            first_code[l] = code;
            first_sym[l] = k;
            code = (code + count[l]) << 1;
            k += count[l];
        }
        syms.clear();
        for (int l = 1; l <= 15; l++)
            for (size_t i = 0; i < len.size(); i++)
                if (len[i] == l) syms.push_back((int)i);
        return true;
    }
    int read(BitReader& br) const
    {
        if (single >= 0) return single;
        int code = 0;
        for (int l = 1; l <= 15; l++) {
            code = (code << 1) | (int)br.get(1);
            if (count[l] && code - first_code[l] < count[l] && code >= first_code[l]) return syms[first_sym[l] + code - first_code[l]];
        }
        return -1;
    }
};

static const int kCodeLengthOrder[19] = {17, 18, 0, 1, 2, 3, 4, 5, 16, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};

inline bool read_code(BitReader& br, int alphabet, Code& c)
{
    std::vector<int> len((size_t)alphabet, 0);
    if (br.get(1)) {  // simple code: one or two symbols
        const int nsym = (int)br.get(1) + 1;
        const int first8 = (int)br.get(1);
        const int s0 = (int)br.get(first8 ? 8 : 1);
        if (s0 >= alphabet) return false;
        len[(size_t)s0] = 1;
        if (nsym == 2) {
            const int s1 = (int)br.get(8);
            if (s1 >= alphabet) return false;
            len[(size_t)s1] = 1;
        }
        if (nsym == 1 || std::count_if(len.begin(), len.end(), [](int l) { return l != 0; }) == 1) {
            c.single = s0;
            return !br.eos;
        }
        return c.build(len) && !br.eos;
    }
    std::vector<int> cl(19, 0);
    const int ncl = 4 + (int)br.get(4);
    if (ncl > 19) return false;
    for (int i = 0; i < ncl; i++) cl[(size_t)kCodeLengthOrder[i]] = (int)br.get(3);
    Code clc;
    if (!clc.build(cl)) return false;
    int max_symbol = alphabet;
    if (br.get(1)) {
        const int nb = 2 + 2 * (int)br.get(3);
        max_symbol = 2 + (int)br.get(nb);
        if (max_symbol > alphabet) return false;
    }
    int prev = 8, i = 0;
    while (i < alphabet) {
        if (max_symbol-- == 0) break;
        const int s = clc.read(br);
        if (s < 0 || br.eos) return false;
        if (s < 16) {
            len[(size_t)i++] = s;
            if (s) prev = s;
        } else {
            const int extra = s == 16 ? 2 : (s == 17 ? 3 : 7), base = s == 16 ? 3 : (s == 17 ? 3 : 11);
            const int rep = base + (int)br.get(extra);
            if (i + rep > alphabet) return false;
            const int v = s == 16 ? prev : 0;
            for (int k = 0; k < rep; k++) len[(size_t)i++] = v;
        }
    }
    return c.build(len) && !br.eos;
}

struct Group { Code c[5]; };

static const int8_t kDistMap[120][2] = {
    {0, 1},  {1, 0},  {1, 1},  {-1, 1}, {0, 2},  {2, 0},  {1, 2},  {-1, 2}, {2, 1},  {-2, 1}, {2, 2},  {-2, 2}, {0, 3},  {3, 0},  {1, 3},
    {-1, 3}, {3, 1},  {-3, 1}, {2, 3},  {-2, 3}, {3, 2},  {-3, 2}, {0, 4},  {4, 0},  {1, 4},  {-1, 4}, {4, 1},  {-4, 1}, {3, 3},  {-3, 3},
    {2, 4},  {-2, 4}, {4, 2},  {-4, 2}, {0, 5},  {3, 4},  {-3, 4}, {4, 3},  {-4, 3}, {5, 0},  {1, 5},  {-1, 5}, {5, 1},  {-5, 1}, {2, 5},
    {-2, 5}, {5, 2},  {-5, 2}, {4, 4},  {-4, 4}, {3, 5},  {-3, 5}, {5, 3},  {-5, 3}, {0, 6},  {6, 0},  {1, 6},  {-1, 6}, {6, 1},  {-6, 1},
    {2, 6},  {-2, 6}, {6, 2},  {-6, 2}, {4, 5},  {-4, 5}, {5, 4},  {-5, 4}, {3, 6},  {-3, 6}, {6, 3},  {-6, 3}, {0, 7},  {7, 0},  {1, 7},
    {-1, 7}, {5, 5},  {-5, 5}, {7, 1},  {-7, 1}, {4, 6},  {-4, 6}, {6, 4},  {-6, 4}, {2, 7},  {-2, 7}, {7, 2},  {-7, 2}, {3, 7},  {-3, 7},
    {7, 3},  {-7, 3}, {5, 6},  {-5, 6}, {6, 5},  {-6, 5}, {8, 0},  {4, 7},  {-4, 7}, {7, 4},  {-7, 4}, {8, 1},  {8, 2},  {6, 6},  {-6, 6},
    {8, 3},  {5, 7},  {-5, 7}, {7, 5},  {-7, 5}, {8, 4},  {6, 7},  {-6, 7}, {7, 6},  {-7, 6}, {8, 5},  {7, 7},  {-7, 7}, {8, 6},  {8, 7}};
// note: the table above is (dx, dy) pairs as (xi, yi) in the specification's order

inline int prefix_value(BitReader& br, int prefix)
{
    if (prefix < 4) return prefix + 1;
    const int extra = (prefix - 2) >> 1;
    const int offset = (2 + (prefix & 1)) << extra;
    return offset + (int)br.get(extra) + 1;
}

// an entropy-coded ARGB image of xs x ys pixels; top = the main image (may carry a colour cache and meta prefix codes)
inline bool decode_image(BitReader& br, int xs, int ys, bool top, std::vector<uint32_t>& out)
{
    int cache_bits = 0;
    if (br.get(1)) {
        cache_bits = (int)br.get(4);
        if (cache_bits < 1 || cache_bits > 11) return false;
    }
    const int cache_size = cache_bits ? 1 << cache_bits : 0;
    std::vector<uint32_t> meta;
    int prefix_bits = 0, meta_w = 0, ngroups = 1;
    if (top && br.get(1)) {
        prefix_bits = (int)br.get(3) + 2;
        meta_w = (xs + (1 << prefix_bits) - 1) >> prefix_bits;
        const int meta_h = (ys + (1 << prefix_bits) - 1) >> prefix_bits;
        if (!decode_image(br, meta_w, meta_h, false, meta)) return false;
        for (uint32_t& m : meta) {
            m = (m >> 8) & 0xffff;
            ngroups = std::max(ngroups, (int)m + 1);
        }
    }
    std::vector<Group> groups((size_t)ngroups);
    const int alpha_sizes[5] = {256 + 24 + cache_size, 256, 256, 256, 40};
    for (Group& g : groups)
        for (int k = 0; k < 5; k++)
            if (!read_code(br, alpha_sizes[k], g.c[k])) return false;
    std::vector<uint32_t> cache((size_t)cache_size, 0);
    out.assign((size_t)xs * ys, 0);
    const size_t total = (size_t)xs * ys;
    size_t i = 0, cached_upto = 0;
    auto cache_fill = [&](size_t upto) {
        if (!cache_size) return;
        for (; cached_upto < upto; cached_upto++) cache[(0x1e35a7bdu * out[cached_upto]) >> (32 - cache_bits)] = out[cached_upto];
    };
    while (i < total) {
        const int x = (int)(i % (size_t)xs), y = (int)(i / (size_t)xs);
        const Group& g = groups[meta.empty() ? 0 : meta[(size_t)(y >> prefix_bits) * meta_w + (x >> prefix_bits)]];
        const int s = g.c[0].read(br);
        if (s < 0 || br.eos) return false;
        if (s < 256) {
            const int r = g.c[1].read(br), b = g.c[2].read(br), a = g.c[3].read(br);
            if (r < 0 || b < 0 || a < 0) return false;
            out[i++] = ((uint32_t)a << 24) | ((uint32_t)r << 16) | ((uint32_t)s << 8) | (uint32_t)b;
        } else if (s < 256 + 24) {
            const int len = prefix_value(br, s - 256);
            const int ds = g.c[4].read(br);
            if (ds < 0) return false;
            const int dcode = prefix_value(br, ds);
            long long dist;
            if (dcode > 120) dist = dcode - 120;
            else {
                const int dx = kDistMap[dcode - 1][0], dy = kDistMap[dcode - 1][1];
                dist = (long long)dx + (long long)dy * xs;
                if (dist < 1) dist = 1;
            }
            if ((long long)i < dist || i + (size_t)len > total) return false;
            for (int k = 0; k < len; k++, i++) out[i] = out[i - (size_t)dist];
        } else {
            const int idx = s - 280;
            if (idx >= cache_size) return false;
            cache_fill(i);
            out[i++] = cache[(size_t)idx];
        }
        if (br.eos) return false;
    }
    (void)cache_fill;
    return true;
}

inline uint32_t add_px(uint32_t a, uint32_t b) { return (((a & 0xff00ff00u) + (b & 0xff00ff00u)) & 0xff00ff00u) | (((a & 0x00ff00ffu) + (b & 0x00ff00ffu)) & 0x00ff00ffu); }
inline uint32_t avg2(uint32_t a, uint32_t b) { return (((a ^ b) & 0xfefefefeu) >> 1) + (a & b); }
inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
inline uint32_t select_px(uint32_t L, uint32_t T, uint32_t TL)
{
    int pl = 0, pt = 0;
    for (int sh = 0; sh < 32; sh += 8) {
        const int l = (L >> sh) & 255, t = (T >> sh) & 255, tl = (TL >> sh) & 255;
        const int p = l + t - tl;
        pl += abs(p - l);
        pt += abs(p - t);
    }
    return pl < pt ? L : T;
}
inline uint32_t clamp_add_sub_full(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r = 0;
    for (int sh = 0; sh < 32; sh += 8) r |= (uint32_t)clamp255((int)((a >> sh) & 255) + (int)((b >> sh) & 255) - (int)((c >> sh) & 255)) << sh;
    return r;
}
inline uint32_t clamp_add_sub_half(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int sh = 0; sh < 32; sh += 8) {
        const int x = (a >> sh) & 255, y = (b >> sh) & 255;
        r |= (uint32_t)clamp255(x + (x - y) / 2) << sh;
    }
    return r;
}
inline uint32_t predict(int mode, uint32_t L, uint32_t T, uint32_t TR, uint32_t TL)
{
    switch (mode) {
    case 0: return 0xff000000u;
    case 1: return L;
    case 2: return T;
    case 3: return TR;
    case 4: return TL;
    case 5: return avg2(avg2(L, TR), T);
    case 6: return avg2(L, TL);
    case 7: return avg2(L, T);
    case 8: return avg2(TL, T);
    case 9: return avg2(T, TR);
    case 10: return avg2(avg2(L, TL), avg2(T, TR));
    case 11: return select_px(L, T, TL);
    case 12: return clamp_add_sub_full(L, T, TL);
    case 13: return clamp_add_sub_half(avg2(L, T), TL);
    default: return 0xff000000u;
    }
}
// the predictor of pixel (x, y) with the border rules of the specification
inline uint32_t predict_at(const uint32_t* px, int w, int x, int y, int mode)
{
    if (y == 0) return x == 0 ? 0xff000000u : px[x - 1];
    if (x == 0) return px[(size_t)(y - 1) * w];
    const uint32_t* row = px + (size_t)y * w;
    const uint32_t* up = row - w;
    const uint32_t TR = x + 1 < w ? up[x + 1] : row[0];  // rightmost column: the leftmost pixel of the current row
    return predict(mode, row[x - 1], up[x], TR, up[x - 1]);
}

}  // namespace vp8l

namespace vp8l {

// the image stream after the 5-byte VP8L header (also the payload of a compressed ALPH chunk): transforms + entropy-coded ARGB
inline bool decode_stream(BitReader& br, int w, int h, std::vector<uint32_t>& px, const char** why)
{
    auto fail = [&](const char* m) { *why = m; return false; };
    // transforms, in bitstream order
    struct Xf { int type, bits, xs; std::vector<uint32_t> data; };
    std::vector<Xf> xfs;
    int xs = w;
    bool seen[4] = {};
    while (br.get(1)) {
        Xf t;
        t.type = (int)br.get(2);
        if (seen[t.type]) return fail("repeated transform");
        seen[t.type] = true;
        t.bits = 0;
        t.xs = xs;
        if (t.type == 0 || t.type == 1) {
            t.bits = (int)br.get(3) + 2;
            const int bw = (xs + (1 << t.bits) - 1) >> t.bits, bh = (h + (1 << t.bits) - 1) >> t.bits;
            if (!decode_image(br, bw, bh, false, t.data)) return fail("corrupt transform image");
        } else if (t.type == 3) {
            const int ncol = (int)br.get(8) + 1;
            if (!decode_image(br, ncol, 1, false, t.data)) return fail("corrupt colour table");
            for (int i = 1; i < ncol; i++) t.data[(size_t)i] = add_px(t.data[(size_t)i], t.data[(size_t)i - 1]);
            t.bits = ncol <= 2 ? 3 : (ncol <= 4 ? 2 : (ncol <= 16 ? 1 : 0));
            xs = (xs + (1 << t.bits) - 1) >> t.bits;
        }
        xfs.push_back(std::move(t));
        if (br.eos) return fail("truncated bitstream");
    }
    if (!decode_image(br, xs, h, true, px)) return fail("corrupt image data");
    // inverse transforms, last one first
    for (size_t ti = xfs.size(); ti-- > 0;) {
        const Xf& t = xfs[ti];
        const int tw = t.xs;  // width of the image this transform produces
        if (t.type == 2) {
            for (uint32_t& p : px) {
                const uint32_t g = (p >> 8) & 255;
                p = (p & 0xff00ff00u) | ((((p >> 16) & 255) + g) & 255) << 16 | (((p & 255) + g) & 255);
            }
        } else if (t.type == 1) {
            const int bw = (tw + (1 << t.bits) - 1) >> t.bits;
            for (int y = 0; y < h; y++)
                for (int x = 0; x < tw; x++) {
                    const uint32_t e = t.data[(size_t)(y >> t.bits) * bw + (x >> t.bits)];
                    const int8_t g2r = (int8_t)(e & 255), g2b = (int8_t)((e >> 8) & 255), r2b = (int8_t)((e >> 16) & 255);
                    uint32_t& p = px[(size_t)y * tw + x];
                    const int8_t g = (int8_t)((p >> 8) & 255);
                    int r = (p >> 16) & 255, b = p & 255;
                    r = (r + ((g2r * g) >> 5)) & 255;
                    b = (b + ((g2b * g) >> 5)) & 255;
                    b = (b + ((r2b * (int8_t)r) >> 5)) & 255;
                    p = (p & 0xff00ff00u) | ((uint32_t)r << 16) | (uint32_t)b;
                }
        } else if (t.type == 0) {
            const int bw = (tw + (1 << t.bits) - 1) >> t.bits;
            for (int y = 0; y < h; y++)
                for (int x = 0; x < tw; x++) {
                    const int mode = (int)((t.data[(size_t)(y >> t.bits) * bw + (x >> t.bits)] >> 8) & 15);
                    uint32_t& p = px[(size_t)y * tw + x];
                    p = add_px(p, predict_at(px.data(), tw, x, y, mode));
                }
        } else {  // colour indexing: unpack (t.bits > 0) and look up
            const int per = 1 << t.bits, nb = 8 >> t.bits;  // pixels per packed pixel, bits per index
            const int pw = (tw + per - 1) >> t.bits;
            std::vector<uint32_t> un((size_t)tw * h);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < tw; x++) {
                    const uint32_t packed = (px[(size_t)y * pw + (x >> t.bits)] >> 8) & 255;
                    const uint32_t idx = t.bits ? (packed >> ((x & (per - 1)) * nb)) & ((1u << nb) - 1) : packed;
                    un[(size_t)y * tw + x] = idx < t.data.size() ? t.data[idx] : 0;
                }
            px.swap(un);
        }
    }
    if (px.size() != (size_t)w * h) return fail("transform geometry");
    return true;
}

// ALPH chunk of a lossy file (WebP container spec): 1 header byte (bits 0-1 compression: 0 raw / 1 VP8L stream with the alpha in
// green, bits 2-3 prediction filter), then the plane; the filter is undone row by row
inline bool decode_alpha(const uint8_t* d, size_t n, int w, int h, std::vector<uint8_t>& a, const char** why)
{
    auto fail = [&](const char* m) { *why = m; return false; };
    if (n < 1) return fail("empty ALPH chunk");
    const int method = d[0] & 3, filter = (d[0] >> 2) & 3;
    a.assign((size_t)w * h, 255);
    if (method == 0) {
        if (n - 1 < (size_t)w * h) return fail("truncated ALPH chunk");
        memcpy(a.data(), d + 1, (size_t)w * h);
    } else if (method == 1) {
        BitReader br;
        br.p = d + 1;
        br.n = n - 1;
        std::vector<uint32_t> px;
        if (!decode_stream(br, w, h, px, why)) return false;
        for (size_t i = 0; i < px.size(); i++) a[i] = (uint8_t)(px[i] >> 8);
    } else
        return fail("unknown alpha compression");
    if (filter)
        for (int y = 0; y < h; y++) {
            uint8_t* row = &a[(size_t)y * w];
            const uint8_t* prev = y ? row - w : nullptr;
            if (filter == 1 || !prev) {  // horizontal (also the first row of the other two): left neighbour, first pixel from above
                int pred = prev ? prev[0] : 0;
                for (int x = 0; x < w; x++) pred = row[x] = (uint8_t)(row[x] + pred);
            } else if (filter == 2) {    // vertical
                for (int x = 0; x < w; x++) row[x] = (uint8_t)(row[x] + prev[x]);
            } else {                     // gradient: clip(left + top - top-left)
                int left = prev[0], tl = prev[0];
                for (int x = 0; x < w; x++) {
                    const int top = prev[x];
                    const int g = left + top - tl;
                    left = row[x] = (uint8_t)(row[x] + (g < 0 ? 0 : (g > 255 ? 255 : g)));
                    tl = top;
                }
            }
        }
    return true;
}

}  // namespace vp8l

inline bool webp_decode_file(const std::string& path, int& w, int& h, int& c, std::vector<uint8_t>& pixels, std::string* why = nullptr)
{
    using namespace vp8l;
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return fail("cannot open");
    std::vector<uint8_t> f;
    {
        uint8_t buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) f.insert(f.end(), buf, buf + n);
        fclose(fp);
    }
    if (f.size() < 20 || memcmp(f.data(), "RIFF", 4) || memcmp(f.data() + 8, "WEBP", 4)) return fail("not a WebP file");
    size_t pos = 12;
    const uint8_t *vp8l = nullptr, *vp8 = nullptr, *alph = nullptr;
    size_t vp8l_len = 0, vp8_len = 0, alph_len = 0;
    while (pos + 8 <= f.size()) {
        const uint32_t len = (uint32_t)f[pos + 4] | ((uint32_t)f[pos + 5] << 8) | ((uint32_t)f[pos + 6] << 16) | ((uint32_t)f[pos + 7] << 24);
        if (len > f.size() - pos - 8) return fail("truncated chunk");
        if (!memcmp(&f[pos], "VP8L", 4)) { vp8l = &f[pos + 8]; vp8l_len = len; break; }
        if (!memcmp(&f[pos], "VP8 ", 4)) { vp8 = &f[pos + 8]; vp8_len = len; break; }
        if (!memcmp(&f[pos], "ALPH", 4)) { alph = &f[pos + 8]; alph_len = len; }
        if (!memcmp(&f[pos], "ANIM", 4)) return fail("animated WebP is not supported");
        pos += 8 + (size_t)len + (len & 1);
    }
    const char* err = "corrupt WebP";
    if (vp8) {  // lossy: VP8 key frame [+ alpha plane]
        vp8::Frame fr;
        if (!vp8::decode_frame(vp8, vp8_len, fr, &err)) return fail(err);
        w = fr.width;
        h = fr.height;
        std::vector<uint8_t> a;
        if (alph && !decode_alpha(alph, alph_len, w, h, a, &err)) return fail(err);
        c = alph ? 4 : 3;
        pixels.resize((size_t)w * h * c);
        vp8::frame_to_rgb(fr, pixels.data(), c);
        if (alph)
            for (size_t i = 0; i < a.size(); i++) pixels[i * 4 + 3] = a[i];
        return true;
    }
    if (!vp8l || vp8l_len < 5 || vp8l[0] != 0x2f) return fail("no VP8 / VP8L bitstream");
    BitReader br;
    br.p = vp8l + 1;
    br.n = vp8l_len - 1;
    w = (int)br.get(14) + 1;
    h = (int)br.get(14) + 1;
    const bool alpha_hint = br.get(1) != 0;
    if (br.get(3) != 0) return fail("unknown VP8L version");
    std::vector<uint32_t> px;
    if (!decode_stream(br, w, h, px, &err)) return fail(err);
    bool has_alpha = false;
    if (alpha_hint)
        for (uint32_t p : px)
            if ((p >> 24) != 255) { has_alpha = true; break; }
    c = has_alpha ? 4 : 3;
    pixels.resize((size_t)w * h * c);
    for (size_t i = 0; i < px.size(); i++) {
        pixels[i * c] = (uint8_t)(px[i] >> 16);
        pixels[i * c + 1] = (uint8_t)(px[i] >> 8);
        pixels[i * c + 2] = (uint8_t)px[i];
        if (c == 4) pixels[i * c + 3] = (uint8_t)(px[i] >> 24);
    }
    return true;
}

namespace vp8l {

struct BitWriter {
    std::vector<uint8_t> out;
    uint64_t acc = 0;
    int nb = 0;
    void put(uint32_t v, int n)
    {
        acc |= (uint64_t)v << nb;
        nb += n;
        while (nb >= 8) { out.push_back((uint8_t)acc); acc >>= 8; nb -= 8; }
    }
    void flush() { if (nb) { out.push_back((uint8_t)acc); acc = 0; nb = 0; } }
};

// Huffman code lengths (<= 15) for a histogram: standard two-queue construction, then lengths clipped with the Kraft sum
// repaired (overlong codes shortened, the cheapest symbols lengthened)
inline void huff_lengths(const std::vector<uint32_t>& hist, std::vector<int>& len)
{
    const int n = (int)hist.size();
    len.assign((size_t)n, 0);
    std::vector<int> used;
    for (int i = 0; i < n; i++)
        if (hist[(size_t)i]) used.push_back(i);
    if (used.empty()) return;
    if (used.size() == 1) { len[(size_t)used[0]] = 1; return; }
    struct Node { uint64_t w; int l, r; };
    std::vector<Node> nodes;
    std::sort(used.begin(), used.end(), [&](int a, int b) { return hist[(size_t)a] < hist[(size_t)b]; });
    for (int s : used) nodes.push_back({hist[(size_t)s], -1, -1});
    size_t a = 0, b = used.size();  // leaves queue [a, used.size()), internal nodes queue [b, nodes.size())
    const size_t nleaf = used.size();
    auto pop = [&]() -> int {
        if (a < nleaf && (b >= nodes.size() || nodes[a].w <= nodes[b].w)) return (int)a++;
        return (int)b++;
    };
    while ((nleaf - a) + (nodes.size() - b) > 1) {
        const int x = pop(), y = pop();
        nodes.push_back({nodes[(size_t)x].w + nodes[(size_t)y].w, x, y});
    }
    std::vector<int> depth(nodes.size(), 0);
    for (size_t i = nodes.size(); i-- > nleaf;) {
        depth[(size_t)nodes[i].l] = depth[i] + 1;
        depth[(size_t)nodes[i].r] = depth[i] + 1;
    }
    for (size_t i = 0; i < nleaf; i++) len[(size_t)used[i]] = depth[i];
    // limit to 15 bits
    bool over = false;
    for (int s : used)
        if (len[(size_t)s] > 15) { len[(size_t)s] = 15; over = true; }
    if (over) {
        long long kraft = 0;
        for (int s : used) kraft += 1LL << (15 - len[(size_t)s]);
        // too many short codes after clipping: lengthen the rarest symbols that are shorter than 15
        for (size_t i = 0; kraft > (1LL << 15) && i < used.size();) {
            const int s = used[i];
            if (len[(size_t)s] < 15) {
                kraft -= 1LL << (15 - len[(size_t)s] - 1);
                len[(size_t)s]++;
            } else i++;
        }
        // leftover space: give it back, always to the most frequent symbol whose code can still be shortened within it.  Every
        // term of the Kraft sum is a multiple of the smallest one, so the deficit is too: the loop ends at exactly 2^15.
        long long deficit = (1LL << 15) - kraft;
        while (deficit > 0) {
            bool moved = false;
            for (size_t i = used.size(); i-- > 0;) {
                const int s = used[i];
                const long long inc = 1LL << (15 - len[(size_t)s]);
                if (len[(size_t)s] > 1 && inc <= deficit) {
                    len[(size_t)s]--;
                    deficit -= inc;
                    moved = true;
                    break;
                }
            }
            if (!moved) break;
        }
    }
}

struct EncCode {
    std::vector<int> len;
    std::vector<uint32_t> code;  // bit-reversed: ready to be emitted LSB first
    void assign()
    {
        code.assign(len.size(), 0);
        int count[16] = {}, next[16] = {};
        for (int l : len) count[l]++;
        count[0] = 0;
        int c = 0;
        for (int l = 1; l <= 15; l++) { c = (c + count[l - 1]) << 1; next[l] = c; }
        for (size_t i = 0; i < len.size(); i++)
            if (len[i]) {
                uint32_t v = (uint32_t)next[len[i]]++, r = 0;
                for (int k = 0; k < len[i]; k++) r |= ((v >> k) & 1) << (len[i] - 1 - k);
                code[i] = r;
            }
    }
};

// writes one prefix code (normal form, code lengths sent with a fixed 4-bit code-length code; one-symbol alphabets in simple form)
inline void write_code(BitWriter& bw, EncCode& ec, const std::vector<uint32_t>& hist)
{
    huff_lengths(hist, ec.len);
    int used = 0, last = 0;
    for (size_t i = 0; i < ec.len.size(); i++)
        if (ec.len[i]) { used++; last = (int)i; }
    if (used <= 1 && last < 256) {  // simple code, one symbol (symbol 0 when the alphabet is unused): the symbol costs no bits
        bw.put(1, 1); bw.put(0, 1);
        if (last < 2) { bw.put(0, 1); bw.put((uint32_t)last, 1); }
        else { bw.put(1, 1); bw.put((uint32_t)last, 8); }
        std::fill(ec.len.begin(), ec.len.end(), 0);
        ec.code.assign(ec.len.size(), 0);
        return;
    }
    if (used <= 1) { ec.len[0] = 1; ec.len[(size_t)last] = 1; }  // a lone symbol >= 256 has no simple form: pair it with symbol 0
    bw.put(0, 1);         // normal code
    bw.put(19 - 4, 4);    // all 19 code-length code lengths follow
    for (int i = 0; i < 19; i++) bw.put(kCodeLengthOrder[i] < 16 ? 4 : 0, 3);  // symbols 0..15: 4 bits each, no repeat codes
    bw.put(0, 1);         // max_symbol = alphabet size
    for (size_t i = 0; i < ec.len.size(); i++) {  // canonical 4-bit code of value v is v itself, emitted MSB first
        uint32_t v = (uint32_t)ec.len[i], r = 0;
        for (int k = 0; k < 4; k++) r |= ((v >> k) & 1) << (3 - k);
        bw.put(r, 4);
    }
    ec.assign();
}

}  // namespace vp8l

// Writes 8-bit RGB (c = 3) or RGBA (c = 4) as lossless WebP.
inline bool webp_encode_file(const std::string& path, const uint8_t* pixels, int w, int h, int c)
{
    using namespace vp8l;
    if ((c != 3 && c != 4) || w <= 0 || h <= 0 || w > 16384 || h > 16384) return false;
    const size_t n = (size_t)w * h;
    std::vector<uint32_t> px(n);
    for (size_t i = 0; i < n; i++)
        px[i] = ((uint32_t)(c == 4 ? pixels[i * c + 3] : 255) << 24) | ((uint32_t)pixels[i * c] << 16) | ((uint32_t)pixels[i * c + 1] << 8) | pixels[i * c + 2];
    // subtract green (applied first by the encoder, undone last by the decoder)
    for (uint32_t& p : px) {
        const uint32_t g = (p >> 8) & 255;
        p = (p & 0xff00ff00u) | ((((p >> 16) & 255) - g) & 255) << 16 | (((p & 255) - g) & 255);
    }
    // predictor transform: per 16x16 block the mode with the smallest sum of absolute (signed) residuals
    const int bits = 4, bw_ = (w + 15) >> 4, bh_ = (h + 15) >> 4;
    std::vector<uint32_t> modes((size_t)bw_ * bh_, 0), res(n);
    static const int cand[] = {1, 2, 5, 7, 11, 12, 13};
    for (int by = 0; by < bh_; by++)
        for (int bx = 0; bx < bw_; bx++) {
            int best = 11;
            unsigned long long best_cost = ~0ULL;
            for (int m : cand) {
                unsigned long long cost = 0;
                for (int y = by * 16; y < std::min(h, by * 16 + 16); y++)
                    for (int x = bx * 16; x < std::min(w, bx * 16 + 16); x++) {
                        const uint32_t p = px[(size_t)y * w + x], q = predict_at(px.data(), w, x, y, m);
                        for (int sh = 0; sh < 32; sh += 8) {
                            const int d = (int)(((p >> sh) - (q >> sh)) & 255);
                            cost += d < 128 ? d : 256 - d;
                        }
                    }
                if (cost < best_cost) { best_cost = cost; best = m; }
            }
            modes[(size_t)by * bw_ + bx] = 0xff000000u | ((uint32_t)best << 8);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int m = (int)((modes[(size_t)(y >> 4) * bw_ + (x >> 4)] >> 8) & 15);
            const uint32_t p = px[(size_t)y * w + x], q = predict_at(px.data(), w, x, y, m);
            // per-byte p - q: the 0xff filler bytes absorb the borrows so that they cannot run into the next channel
            res[(size_t)y * w + x] = ((0x00ff00ffu + (p & 0xff00ff00u) - (q & 0xff00ff00u)) & 0xff00ff00u) |
                                     ((0xff00ff00u + (p & 0x00ff00ffu) - (q & 0x00ff00ffu)) & 0x00ff00ffu);
        }
    // every entropy-coded image below: no colour cache, no meta codes; literals plus run-length backward references
    // (distance = 1 pixel: after prediction, flat areas are runs of identical residuals)
    BitWriter bw;
    auto length_prefix = [](int len, int& prefix, int& extra_bits, int& extra) {  // inverse of vp8l::prefix_value
        if (len <= 4) { prefix = len - 1; extra_bits = 0; extra = 0; return; }
        const int v = len - 1;
        int hb = 31;
        while (!((v >> hb) & 1)) hb--;
        const int second = (v >> (hb - 1)) & 1;
        prefix = 2 * hb + second;
        extra_bits = hb - 1;
        extra = v & ((1 << extra_bits) - 1);
    };
    auto write_image = [&](const std::vector<uint32_t>& img, bool top) {
        bw.put(0, 1);            // no colour cache
        if (top) bw.put(0, 1);   // no meta prefix codes
        // tokens: (position, run length); run length 0 = literal
        struct Tok { uint32_t pos; int run; };
        std::vector<Tok> toks;
        toks.reserve(img.size());
        for (size_t i = 0; i < img.size();) {
            size_t j = i;
            if (i > 0 && img[i] == img[i - 1])
                while (j < img.size() && j - i < 4096 && img[j] == img[i - 1]) j++;
            if (j - i >= 3) { toks.push_back({(uint32_t)i, (int)(j - i)}); i = j; }
            else { toks.push_back({(uint32_t)i, 0}); i++; }
        }
        std::vector<uint32_t> hg(256 + 24, 0), hr(256, 0), hb(256, 0), ha(256, 0), hd(40, 0);
        for (const Tok& t : toks) {
            if (t.run) {
                int pf, eb, ex;
                length_prefix(t.run, pf, eb, ex);
                hg[256 + (size_t)pf]++;
                hd[1]++;  // distance code 2 = (dx 1, dy 0) = the previous pixel; its prefix symbol is 1, no extra bits
            } else {
                const uint32_t p = img[t.pos];
                hg[(p >> 8) & 255]++; hr[(p >> 16) & 255]++; hb[p & 255]++; ha[p >> 24]++;
            }
        }
        EncCode cg, cr, cb, ca, cd;
        write_code(bw, cg, hg); write_code(bw, cr, hr); write_code(bw, cb, hb); write_code(bw, ca, ha); write_code(bw, cd, hd);
        for (const Tok& t : toks) {
            if (t.run) {
                int pf, eb, ex;
                length_prefix(t.run, pf, eb, ex);
                if (cg.len[256 + (size_t)pf]) bw.put(cg.code[256 + (size_t)pf], cg.len[256 + (size_t)pf]);
                if (eb) bw.put((uint32_t)ex, eb);
                if (cd.len[1]) bw.put(cd.code[1], cd.len[1]);
            } else {
                const uint32_t p = img[t.pos];
                const uint32_t g = (p >> 8) & 255, r = (p >> 16) & 255, b = p & 255, a = p >> 24;
                if (cg.len[g]) bw.put(cg.code[g], cg.len[g]);
                if (cr.len[r]) bw.put(cr.code[r], cr.len[r]);
                if (cb.len[b]) bw.put(cb.code[b], cb.len[b]);
                if (ca.len[a]) bw.put(ca.code[a], ca.len[a]);
            }
        }
    };
    bw.put((uint32_t)(w - 1), 14); bw.put((uint32_t)(h - 1), 14);
    bw.put(c == 4 ? 1 : 0, 1);
    bw.put(0, 3);
    bw.put(1, 1); bw.put(2, 2);  // transform: subtract green
    bw.put(1, 1); bw.put(0, 2);  // transform: predictor
    bw.put((uint32_t)(bits - 2), 3);
    write_image(modes, false);
    bw.put(0, 1);                // no more transforms
    write_image(res, true);
    bw.flush();
    const uint32_t payload = (uint32_t)bw.out.size() + 1;  // + signature byte
    const uint32_t padded = payload + (payload & 1);
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    auto le32 = [&](uint32_t v) { uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)}; return fwrite(b, 1, 4, fp) == 4; };
    bool ok = fwrite("RIFF", 1, 4, fp) == 4 && le32(4 + 8 + padded) && fwrite("WEBPVP8L", 1, 8, fp) == 8 && le32(payload);
    const uint8_t sig = 0x2f;
    ok = ok && fwrite(&sig, 1, 1, fp) == 1 && fwrite(bw.out.data(), 1, bw.out.size(), fp) == bw.out.size();
    if (ok && (payload & 1)) { const uint8_t z = 0; ok = fwrite(&z, 1, 1, fp) == 1; }
    return (fclose(fp) == 0) && ok;
}

}  // namespace rsrhost
