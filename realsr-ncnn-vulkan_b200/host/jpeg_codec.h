// jpeg_codec.h -- JPEG reader (sequential + progressive) / baseline writer for the C++ command line (host/realsr_cli.cpp).
//
// The reference reads jpg through its vendored stb_image and writes with stbi_write_jpg(..., quality 100)
// (src/main.cpp:243-263, :397); those files are not copied.  This is an independent implementation of ITU-T T.81:
//   decode: sequential (SOF0, SOF1; interleaved or one scan per component) and progressive (SOF2: spectral selection and
//           successive approximation, Annex G) DCT, Huffman, 8-bit, 1 or 3 components (YCbCr, JFIF), sampling factors up
//           to 2x2 with the triangle ("fancy") chroma upsampling libjpeg applies, restart intervals; arithmetic-coded and
//           lossless files are refused with a message;
//   encode: baseline, 4:4:4, standard Huffman tables, the Annex K quantisation tables scaled libjpeg-style by a quality.
// Host plumbing only: no pixel of the super-resolution path is computed here.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace rsrhost {

namespace jpg {

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    // canonical code: for length l (1..16) codes [mincode[l], maxcode[l]] map to vals[valptr[l] + code - mincode[l]]
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
    bool present = false;
    void build(const uint8_t counts[16], const uint8_t* symbols, int nsym)
    {
        memcpy(vals, symbols, (size_t)nsym);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k;
            mincode[l] = code;
            code += counts[l - 1];
            k += counts[l - 1];
            maxcode[l] = counts[l - 1] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc = 0;
    int nbits = 0;
    bool hit_marker = false;
    void fill()
    {
        while (nbits <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    if (p < end && *p == 0x00) p++;        // stuffed zero
                    else { hit_marker = true; p--; b = 0; }  // a marker: feed zeros from here on
                }
            }
            acc |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int get(int n)
    {
        if (n == 0) return 0;
        if (nbits < n) fill();
        const int v = (int)(acc >> (32 - n));
        acc <<= n;
        nbits -= n;
        return v;
    }
    void reset() { acc = 0; nbits = 0; hit_marker = false; }
};

inline int decode_symbol(BitReader& br, const Huff& h)
{
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.get(1);
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
inline int extend(int v, int n) { return n && v < (1 << (n - 1)) ? v - (1 << n) + 1 : v; }

// separable float IDCT of one dequantised block (natural order), output clamped to 0..255 with the +128 level shift
inline void idct8x8(const float* in, uint8_t* out, int stride)
{
    static float c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; x++)
            for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? sqrtf(0.125f) : 0.5f) * cosf((2 * x + 1) * u * 3.14159265358979323846f / 16.f);
        init = true;
    }
    float tmp[64];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
            float s = 0;
            for (int u = 0; u < 8; u++) s += c[x][u] * in[y * 8 + u];
            tmp[y * 8 + x] = s;
        }
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++) {
            float s = 0;
            for (int v = 0; v < 8; v++) s += c[y][v] * tmp[v * 8 + x];
            const int iv = (int)lrintf(s + 128.f);
            out[y * stride + x] = (uint8_t)std::min(255, std::max(0, iv));
        }
}
inline void fdct8x8(const float* in, float* out)
{
    static float c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; x++)
            for (int u = 0; u < 8; u++) c[x][u] = (u == 0 ? sqrtf(0.125f) : 0.5f) * cosf((2 * x + 1) * u * 3.14159265358979323846f / 16.f);
        init = true;
    }
    float tmp[64];
    for (int y = 0; y < 8; y++)
        for (int u = 0; u < 8; u++) {
            float s = 0;
            for (int x = 0; x < 8; x++) s += c[x][u] * in[y * 8 + x];
            tmp[y * 8 + u] = s;
        }
    for (int u = 0; u < 8; u++)
        for (int v = 0; v < 8; v++) {
            float s = 0;
            for (int y = 0; y < 8; y++) s += c[y][v] * tmp[y * 8 + u];
            out[v * 8 + u] = s;
        }
}

}  // namespace jpg

// Decodes a Huffman-coded 8-bit JPEG file (sequential or progressive) to 8-bit RGB (c = 3; gray is expanded like the
// reference's load stage does).  All scans are decoded into per-block coefficient storage first (T.81 Annex G for the
// progressive passes), dequantisation + IDCT follow once at the end.
inline bool jpeg_decode_file(const std::string& path, int& w, int& h, int& c, std::vector<uint8_t>& pixels, std::string* why = nullptr)
{
    using namespace jpg;
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
    if (f.size() < 4 || f[0] != 0xFF || f[1] != 0xD8) return fail("not a JPEG");
    uint16_t qt[4][64] = {};
    bool qt_ok[4] = {};
    Huff hdc[4], hac[4];
    struct Comp {
        int id, hs, vs, tq, td, ta;
        int bw, bh;    // plane size in pixels, padded to whole MCUs
        int nbx, nby;  // blocks that carry image data (non-interleaved scans visit only these)
        std::vector<int16_t> coef;  // [bh/8][bw/8][64], natural order
        std::vector<uint8_t> plane;
        int pred;
    } comp[3];
    int ncomp = 0, restart = 0, hmax = 1, vmax = 1;
    bool have_sof = false, progressive = false, adobe = false, saw_scan = false;
    int adobe_transform = -1;
    size_t pos = 2;
    while (pos + 4 <= f.size()) {
        if (f[pos] != 0xFF) { pos++; continue; }  // tolerate stray bytes between segments
        while (pos < f.size() && f[pos] == 0xFF) pos++;  // fill bytes
        if (pos >= f.size()) break;
        const int marker = f[pos++];
        if (marker == 0x00 || marker == 0xD8 || (marker >= 0xD0 && marker <= 0xD7) || marker == 0x01) continue;
        if (marker == 0xD9) break;
        if (pos + 2 > f.size()) return fail("truncated segment");
        const size_t len = ((size_t)f[pos] << 8) | f[pos + 1];
        if (len < 2 || pos + len > f.size()) return fail("truncated segment");
        const uint8_t* d = &f[pos + 2];
        const size_t n = len - 2;
        if (marker == 0xDB) {  // DQT
            size_t i = 0;
            while (i < n) {
                const int pq = d[i] >> 4, tq = d[i] & 15;
                i++;
                if (tq > 3 || i + (pq ? 128 : 64) > n) return fail("bad DQT");
                for (int k = 0; k < 64; k++) {
                    qt[tq][kZigzag[k]] = pq ? (uint16_t)((d[i] << 8) | d[i + 1]) : d[i];
                    i += pq ? 2 : 1;
                }
                qt_ok[tq] = true;
            }
        } else if (marker == 0xC4) {  // DHT
            size_t i = 0;
            while (i + 17 <= n) {
                const int tc = d[i] >> 4, th = d[i] & 15;
                if (tc > 1 || th > 3) return fail("bad DHT");
                int total = 0;
                for (int k = 0; k < 16; k++) total += d[i + 1 + k];
                if (total > 256 || i + 17 + total > n) return fail("bad DHT");
                (tc ? hac[th] : hdc[th]).build(&d[i + 1], &d[i + 17], total);
                i += 17 + total;
            }
        } else if (marker == 0xC0 || marker == 0xC1 || marker == 0xC2) {  // SOF0 / SOF1 / SOF2
            if (have_sof) return fail("more than one frame");
            if (n < 6 || d[0] != 8) return fail("only 8-bit JPEG is supported");
            progressive = marker == 0xC2;
            h = (d[1] << 8) | d[2];
            w = (d[3] << 8) | d[4];
            ncomp = d[5];
            if ((ncomp != 1 && ncomp != 3) || n < (size_t)(6 + 3 * ncomp) || w <= 0 || h <= 0) return fail("unsupported frame header");
            for (int k = 0; k < ncomp; k++) {
                comp[k].id = d[6 + 3 * k];
                comp[k].hs = d[7 + 3 * k] >> 4;
                comp[k].vs = d[7 + 3 * k] & 15;
                comp[k].tq = d[8 + 3 * k];
                if (comp[k].hs < 1 || comp[k].hs > 2 || comp[k].vs < 1 || comp[k].vs > 2 || comp[k].tq > 3) return fail("unsupported sampling factors");
                if (ncomp == 1) comp[k].hs = comp[k].vs = 1;  // a single component is never subsampled against itself
                hmax = std::max(hmax, comp[k].hs);
                vmax = std::max(vmax, comp[k].vs);
            }
            const int mx = (w + 8 * hmax - 1) / (8 * hmax), my = (h + 8 * vmax - 1) / (8 * vmax);
            for (int k = 0; k < ncomp; k++) {
                comp[k].bw = mx * comp[k].hs * 8;
                comp[k].bh = my * comp[k].vs * 8;
                const int cw = (w * comp[k].hs + hmax - 1) / hmax, chh = (h * comp[k].vs + vmax - 1) / vmax;
                comp[k].nbx = (cw + 7) / 8;
                comp[k].nby = (chh + 7) / 8;
                comp[k].coef.assign((size_t)(comp[k].bw / 8) * (comp[k].bh / 8) * 64, 0);
            }
            have_sof = true;
        } else if (marker >= 0xC3 && marker <= 0xCF && marker != 0xC4 && marker != 0xC8 && marker != 0xCC) return fail("unsupported JPEG process (lossless / arithmetic)");
        else if (marker == 0xDD) { if (n >= 2) restart = (d[0] << 8) | d[1]; }
        else if (marker == 0xEE && n >= 12 && !memcmp(d, "Adobe", 5)) { adobe = true; adobe_transform = d[11]; }
        else if (marker == 0xDA) {  // SOS: scan header, then the entropy-coded segment
            if (!have_sof) return fail("SOS before SOF");
            const int ns = n ? d[0] : 0;
            if (ns < 1 || ns > ncomp || n < (size_t)(1 + 2 * ns + 3)) return fail("bad scan header");
            int sc[3];
            for (int k = 0; k < ns; k++) {
                int ci = -1;
                for (int j = 0; j < ncomp; j++)
                    if (comp[j].id == d[1 + 2 * k]) ci = j;
                if (ci < 0) return fail("bad scan component");
                sc[k] = ci;
                comp[ci].td = d[2 + 2 * k] >> 4;
                comp[ci].ta = d[2 + 2 * k] & 15;
                if (comp[ci].td > 3 || comp[ci].ta > 3) return fail("bad table index");
            }
            const int Ss = d[1 + 2 * ns], Se = d[2 + 2 * ns], Ah = d[3 + 2 * ns] >> 4, Al = d[3 + 2 * ns] & 15;
            if (!progressive && (Ss != 0 || Se != 63 || Ah || Al)) return fail("bad sequential scan parameters");
            if (progressive && (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13)) return fail("bad progressive scan parameters");
            for (int k = 0; k < ns; k++) {
                const Comp& cc = comp[sc[k]];
                if ((Ss == 0 && Ah == 0 && !hdc[cc.td].present) || ((Se > 0) && !hac[cc.ta].present)) return fail("missing Huffman table");
            }
            BitReader br;
            br.p = &f[pos + len];
            br.end = f.data() + f.size();
            for (int k = 0; k < ncomp; k++) comp[k].pred = 0;
            int eobrun = 0;
            // one block of the scan; returns false on a corrupt stream
            auto decode_block = [&](Comp& cc, int16_t* blk) -> bool {
                if (!progressive) {
                    const int t = decode_symbol(br, hdc[cc.td]);
                    if (t < 0 || t > 11) return false;
                    cc.pred += extend(br.get(t), t);
                    blk[0] = (int16_t)cc.pred;
                    for (int i = 1; i < 64;) {
                        const int rs = decode_symbol(br, hac[cc.ta]);
                        if (rs < 0) return false;
                        const int r = rs >> 4, s = rs & 15;
                        if (s == 0) {
                            if (r == 15) { i += 16; continue; }
                            break;
                        }
                        i += r;
                        if (i > 63) return false;
                        blk[kZigzag[i]] = (int16_t)extend(br.get(s), s);
                        i++;
                    }
                    return true;
                }
                if (Ss == 0) {  // DC scan
                    if (Ah == 0) {
                        const int t = decode_symbol(br, hdc[cc.td]);
                        if (t < 0 || t > 11) return false;
                        cc.pred += extend(br.get(t), t);
                        blk[0] = (int16_t)(cc.pred * (1 << Al));
                    } else if (br.get(1)) {
                        blk[0] = (int16_t)(blk[0] | (1 << Al));
                    }
                    return true;
                }
                if (Ah == 0) {  // AC first pass (G.1.2.2)
                    if (eobrun > 0) { eobrun--; return true; }
                    for (int k = Ss; k <= Se;) {
                        const int rs = decode_symbol(br, hac[cc.ta]);
                        if (rs < 0) return false;
                        const int r = rs >> 4, s = rs & 15;
                        if (s == 0) {
                            if (r < 15) {
                                eobrun = (1 << r) - 1;
                                if (r) eobrun += br.get(r);
                                break;
                            }
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) return false;
                        blk[kZigzag[k]] = (int16_t)(extend(br.get(s), s) * (1 << Al));
                        k++;
                    }
                    return true;
                }
                // AC refinement pass (G.1.2.3): correction bits for already-nonzero coefficients interleave with the
                // placement of newly nonzero ones
                const int p1 = 1 << Al, m1 = -(1 << Al);
                int k = Ss;
                if (eobrun == 0) {
                    while (k <= Se) {
                        const int rs = decode_symbol(br, hac[cc.ta]);
                        if (rs < 0) return false;
                        int r = rs >> 4;
                        const int s = rs & 15;
                        int val = 0;
                        if (s == 0) {
                            if (r < 15) {
                                eobrun = 1 << r;
                                if (r) eobrun += br.get(r);
                                break;
                            }
                        } else {
                            if (s != 1) return false;
                            val = br.get(1) ? p1 : m1;
                        }
                        while (k <= Se) {
                            int16_t& cf = blk[kZigzag[k]];
                            if (cf != 0) {
                                if (br.get(1) && (cf & p1) == 0) cf = (int16_t)(cf + (cf > 0 ? p1 : m1));
                            } else {
                                if (r == 0) {
                                    if (val) cf = (int16_t)val;
                                    k++;
                                    break;
                                }
                                r--;
                            }
                            k++;
                        }
                    }
                }
                if (eobrun > 0) {
                    for (; k <= Se; k++) {
                        int16_t& cf = blk[kZigzag[k]];
                        if (cf != 0 && br.get(1) && (cf & p1) == 0) cf = (int16_t)(cf + (cf > 0 ? p1 : m1));
                    }
                    eobrun--;
                }
                return true;
            };
            auto restart_if_due = [&](int& left) {
                if (restart && left == 0) {  // RSTn expected: byte-align, skip to it, reset the predictors
                    br.reset();
                    while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) br.p++;
                    if (br.p + 1 < br.end) br.p += 2;
                    for (int k = 0; k < ncomp; k++) comp[k].pred = 0;
                    eobrun = 0;
                    left = restart;
                }
            };
            int left = restart;
            if (ns == 1) {  // non-interleaved: the component's own block raster
                Comp& cc = comp[sc[0]];
                const int rowblocks = cc.bw / 8;
                for (int y = 0; y < cc.nby; y++)
                    for (int x = 0; x < cc.nbx; x++) {
                        restart_if_due(left);
                        if (!decode_block(cc, &cc.coef[((size_t)y * rowblocks + x) * 64])) return fail("corrupt entropy-coded data");
                        if (restart) left--;
                    }
            } else {
                const int mx = comp[0].bw / (8 * comp[0].hs), my = comp[0].bh / (8 * comp[0].vs);
                for (int yy = 0; yy < my; yy++)
                    for (int xx = 0; xx < mx; xx++) {
                        restart_if_due(left);
                        for (int k = 0; k < ns; k++) {
                            Comp& cc = comp[sc[k]];
                            const int rowblocks = cc.bw / 8;
                            for (int by = 0; by < cc.vs; by++)
                                for (int bx = 0; bx < cc.hs; bx++)
                                    if (!decode_block(cc, &cc.coef[((size_t)(yy * cc.vs + by) * rowblocks + xx * cc.hs + bx) * 64]))
                                        return fail("corrupt entropy-coded data");
                        }
                        if (restart) left--;
                    }
            }
            saw_scan = true;
            // continue at the next marker after the entropy-coded segment
            size_t q = (size_t)(br.p - f.data());
            if (q < pos + len) q = pos + len;
            while (q + 1 < f.size() && !(f[q] == 0xFF && f[q + 1] != 0x00 && !(f[q + 1] >= 0xD0 && f[q + 1] <= 0xD7) && f[q + 1] != 0xFF)) q++;
            pos = q;
            continue;
        }
        pos += len;
    }
    if (!have_sof || !saw_scan) return fail("no image data");
    // dequantise + IDCT every block
    for (int k = 0; k < ncomp; k++) {
        Comp& cc = comp[k];
        if (!qt_ok[cc.tq]) return fail("missing quantisation table");
        cc.plane.assign((size_t)cc.bw * cc.bh, 0);
        const int rowblocks = cc.bw / 8;
        for (int y = 0; y < cc.bh / 8; y++)
            for (int x = 0; x < rowblocks; x++) {
                const int16_t* cf = &cc.coef[((size_t)y * rowblocks + x) * 64];
                float blk[64];
                for (int i = 0; i < 64; i++) blk[i] = (float)cf[i] * qt[cc.tq][i];
                idct8x8(blk, &cc.plane[(size_t)y * 8 * cc.bw + x * 8], cc.bw);
            }
        cc.coef.clear();
        cc.coef.shrink_to_fit();
    }
    // colour: full-resolution planes (triangle upsampling for subsampled chroma), then YCbCr -> RGB (JFIF)
    c = 3;
    pixels.resize((size_t)w * h * 3);
    std::vector<uint8_t> full[3];
    for (int k = 0; k < ncomp; k++) {
        const int sx = hmax / comp[k].hs, sy = vmax / comp[k].vs;
        full[k].resize((size_t)w * h);
        const int cw = (w + sx - 1) / sx, chh = (h + sy - 1) / sy;  // meaningful samples of the component
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                if (sx == 1 && sy == 1) { full[k][(size_t)y * w + x] = comp[k].plane[(size_t)y * comp[k].bw + x]; continue; }
                // libjpeg's "fancy" upsampling: 3/4 nearest + 1/4 next-nearest sample in each subsampled direction
                const int x0 = x / sx, y0 = y / sy;
                const int x1 = sx == 2 ? std::min(cw - 1, std::max(0, (x & 1) ? x0 + 1 : x0 - 1)) : x0;
                const int y1 = sy == 2 ? std::min(chh - 1, std::max(0, (y & 1) ? y0 + 1 : y0 - 1)) : y0;
                const uint8_t* P = comp[k].plane.data();
                const int bw = comp[k].bw;
                int v;
                if (sx == 2 && sy == 2)
                    v = (9 * P[(size_t)y0 * bw + x0] + 3 * P[(size_t)y0 * bw + x1] + 3 * P[(size_t)y1 * bw + x0] + P[(size_t)y1 * bw + x1] + 8) >> 4;
                else if (sx == 2)
                    v = (3 * P[(size_t)y0 * bw + x0] + P[(size_t)y0 * bw + x1] + 2) >> 2;
                else
                    v = (3 * P[(size_t)y0 * bw + x0] + P[(size_t)y1 * bw + x0] + 2) >> 2;
                full[k][(size_t)y * w + x] = (uint8_t)v;
            }
    }
    const bool rgb_direct = ncomp == 3 && ((adobe && adobe_transform == 0) || (comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B'));
    for (size_t i = 0; i < (size_t)w * h; i++) {
        if (ncomp == 1) {
            pixels[i * 3] = pixels[i * 3 + 1] = pixels[i * 3 + 2] = full[0][i];
        } else if (rgb_direct) {
            pixels[i * 3] = full[0][i]; pixels[i * 3 + 1] = full[1][i]; pixels[i * 3 + 2] = full[2][i];
        } else {
            const float Y = full[0][i], cb = full[1][i] - 128.f, cr = full[2][i] - 128.f;
            const int r = (int)lrintf(Y + 1.402f * cr), g = (int)lrintf(Y - 0.344136f * cb - 0.714136f * cr), b = (int)lrintf(Y + 1.772f * cb);
            pixels[i * 3] = (uint8_t)std::min(255, std::max(0, r));
            pixels[i * 3 + 1] = (uint8_t)std::min(255, std::max(0, g));
            pixels[i * 3 + 2] = (uint8_t)std::min(255, std::max(0, b));
        }
    }
    return true;
}

// Writes 8-bit RGB (c = 3; c = 4 drops alpha -- the CLI renames RGBA results to png before it gets here) as baseline 4:4:4.
inline bool jpeg_encode_file(const std::string& path, const uint8_t* pixels, int w, int h, int c, int quality = 100)
{
    using namespace jpg;
    if ((c != 3 && c != 4) || w <= 0 || h <= 0 || w > 65535 || h > 65535) return false;
    static const uint8_t lum_q[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,  14, 13, 16, 24, 40,  57,
                                      69, 56, 14, 17, 22,  29,  51,  87,  80, 62, 18, 22, 37,  56,  68,  109, 103, 77, 24, 35, 55,  64,
                                      81, 104, 113, 92, 49, 64,  78,  87,  103, 121, 120, 101, 72, 92, 95,  98,  112, 100, 103, 99};
    static const uint8_t chr_q[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                      99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                      99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    // Annex K.3 typical Huffman tables
    static const uint8_t dc_lum_n[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
    static const uint8_t dc_chr_n[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
    static const uint8_t dc_v[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
    static const uint8_t ac_lum_n[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
    static const uint8_t ac_lum_v[162] = {
        0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
        0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
        0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
        0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
        0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
        0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
        0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
        0xfa};
    static const uint8_t ac_chr_n[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
    static const uint8_t ac_chr_v[162] = {
        0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
        0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
        0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
        0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
        0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
        0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
        0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
        0xfa};
    quality = std::min(100, std::max(1, quality));
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    uint8_t q[2][64];
    for (int i = 0; i < 64; i++) {
        q[0][i] = (uint8_t)std::min(255, std::max(1, (lum_q[i] * scale + 50) / 100));
        q[1][i] = (uint8_t)std::min(255, std::max(1, (chr_q[i] * scale + 50) / 100));
    }
    // code tables: ehufco / ehufsi per symbol
    struct Enc { uint16_t code[256]; uint8_t len[256]; };
    auto make = [](const uint8_t* counts, const uint8_t* vals, Enc& e) {
        memset(&e, 0, sizeof e);
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            for (int i = 0; i < counts[l - 1]; i++, k++) {
                e.code[vals[k]] = (uint16_t)code++;
                e.len[vals[k]] = (uint8_t)l;
            }
            code <<= 1;
        }
    };
    Enc edc[2], eac[2];
    make(dc_lum_n, dc_v, edc[0]); make(dc_chr_n, dc_v, edc[1]);
    make(ac_lum_n, ac_lum_v, eac[0]); make(ac_chr_n, ac_chr_v, eac[1]);

    std::vector<uint8_t> out;
    auto put16 = [&](int v) { out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v); };
    auto seg = [&](int marker, const std::vector<uint8_t>& body) {
        out.push_back(0xFF); out.push_back((uint8_t)marker);
        put16((int)body.size() + 2);
        out.insert(out.end(), body.begin(), body.end());
    };
    out.push_back(0xFF); out.push_back(0xD8);
    seg(0xE0, {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int t = 0; t < 2; t++) {
        std::vector<uint8_t> b(1, (uint8_t)t);
        for (int k = 0; k < 64; k++) b.push_back(q[t][kZigzag[k]]);
        seg(0xDB, b);
    }
    {
        std::vector<uint8_t> b = {8, (uint8_t)(h >> 8), (uint8_t)h, (uint8_t)(w >> 8), (uint8_t)w, 3, 1, 0x11, 0, 2, 0x11, 1, 3, 0x11, 1};
        seg(0xC0, b);
    }
    auto dht = [&](int tc_th, const uint8_t* counts, const uint8_t* vals, int nv) {
        std::vector<uint8_t> b(1, (uint8_t)tc_th);
        b.insert(b.end(), counts, counts + 16);
        b.insert(b.end(), vals, vals + nv);
        seg(0xC4, b);
    };
    dht(0x00, dc_lum_n, dc_v, 12); dht(0x10, ac_lum_n, ac_lum_v, 162);
    dht(0x01, dc_chr_n, dc_v, 12); dht(0x11, ac_chr_n, ac_chr_v, 162);
    seg(0xDA, {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    uint32_t acc = 0;
    int nb = 0;
    auto putbits = [&](int code, int len) {
        acc = (acc << len) | ((uint32_t)code & ((1u << len) - 1));
        nb += len;
        while (nb >= 8) {
            const uint8_t byte = (uint8_t)(acc >> (nb - 8));
            out.push_back(byte);
            if (byte == 0xFF) out.push_back(0);
            nb -= 8;
        }
    };
    int pred[3] = {0, 0, 0};
    for (int by = 0; by < h; by += 8)
        for (int bx = 0; bx < w; bx += 8) {
            float blk[3][64];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    const uint8_t* p = pixels + ((size_t)std::min(h - 1, by + y) * w + std::min(w - 1, bx + x)) * c;
                    const float r = p[0], g = p[1], b = p[2];
                    blk[0][y * 8 + x] = 0.299f * r + 0.587f * g + 0.114f * b - 128.f;
                    blk[1][y * 8 + x] = -0.168736f * r - 0.331264f * g + 0.5f * b;
                    blk[2][y * 8 + x] = 0.5f * r - 0.418688f * g - 0.081312f * b;
                }
            for (int k = 0; k < 3; k++) {
                const int t = k ? 1 : 0;
                float dct[64];
                fdct8x8(blk[k], dct);
                int zz[64];
                for (int i = 0; i < 64; i++) zz[i] = (int)lrintf(dct[kZigzag[i]] / q[t][kZigzag[i]]);
                // DC
                int diff = zz[0] - pred[k];
                pred[k] = zz[0];
                int mag = diff < 0 ? -diff : diff, nbits = 0;
                while (mag) { nbits++; mag >>= 1; }
                putbits(edc[t].code[nbits], edc[t].len[nbits]);
                if (nbits) putbits(diff < 0 ? diff - 1 : diff, nbits);
                // AC
                int run = 0;
                for (int i = 1; i < 64; i++) {
                    if (zz[i] == 0) { run++; continue; }
                    while (run > 15) { putbits(eac[t].code[0xF0], eac[t].len[0xF0]); run -= 16; }
                    int v = zz[i], m = v < 0 ? -v : v, s = 0;
                    while (m) { s++; m >>= 1; }
                    putbits(eac[t].code[(run << 4) | s], eac[t].len[(run << 4) | s]);
                    putbits(v < 0 ? v - 1 : v, s);
                    run = 0;
                }
                if (run) putbits(eac[t].code[0], eac[t].len[0]);
            }
        }
    if (nb) putbits(0x7F, 8 - nb);  // pad the last byte with ones
    out.push_back(0xFF); out.push_back(0xD9);
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
    return (fclose(fp) == 0) && ok;
}

}  // namespace rsrhost
