// vp8_codec.h -- lossy WebP (VP8 key frame) decoder for the C++ command line (host/realsr_cli.cpp, through host/webp_codec.h).
//
// The reference links libwebp and accepts any WebP input (src/webp_image.h:10-49, WebPDecodeRGB / WebPDecodeRGBA).  libwebp is not
// copied; this is an independent implementation of RFC 6386 ("VP8 Data Format and Decoding Guide") restricted to what a still WebP
// holds -- one key frame -- plus the two steps WebP adds on top: chroma upsampling with the 9-3-3-1 kernel and the fixed-point
// YUV -> RGB conversion libwebp applies in WebPDecodeRGB, so that the pixels handed to the network are the ones the reference's
// decoder produces (checked bit for bit against libwebp through Pillow in tests/test_cpp_cli.py).
//   boolean entropy decoder (RFC 7), frame header incl. segmentation, loop-filter deltas, quantiser indices and token probability
//   updates (RFC 9, 19.2), per-macroblock modes with the key-frame contexts (RFC 11), DCT tokens with band / neighbour contexts over
//   up to 8 token partitions (RFC 13), dequantisation, inverse WHT / DCT (RFC 14), intra prediction 16x16 / chroma / 4x4 with the
//   frame-edge rules (RFC 12), simple and normal loop filters with per-segment and per-mode strengths (RFC 15).
// Constant tables: vp8_tables.h (generated, see tools/gen_vp8_tables.py).  Host plumbing only: no pixel of the super-resolution path
// is computed here.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vp8_tables.h"

namespace rsrhost {
namespace vp8 {

struct BoolDec {  // RFC 6386 section 7.3
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    uint32_t value = 0, range = 255;
    int bit_count = 0;
    void init(const uint8_t* data, size_t n)
    {
        p = data;
        end = data + n;
        value = 0;
        for (int i = 0; i < 2; i++) value = (value << 8) | (p < end ? *p++ : 0);
        range = 255;
        bit_count = 0;
    }
    int get(int prob)
    {
        const uint32_t split = 1 + (((range - 1) * (uint32_t)prob) >> 8);
        const uint32_t big = split << 8;
        int bit;
        if (value >= big) { bit = 1; range -= split; value -= big; }
        else { bit = 0; range = split; }
        while (range < 128) {
            value <<= 1;
            range <<= 1;
            if (++bit_count == 8) {
                bit_count = 0;
                value |= (p < end ? *p++ : 0);
            }
        }
        return bit;
    }
    uint32_t literal(int n)
    {
        uint32_t v = 0;
        while (n-- > 0) v = (v << 1) | (uint32_t)get(128);
        return v;
    }
    int signed_literal(int n)
    {
        const int v = (int)literal(n);
        return get(128) ? -v : v;
    }
};

// sub-block prediction modes, in the order the mode-probability table is indexed
enum { B_DC = 0, B_TM, B_VE, B_HE, B_RD, B_VR, B_LD, B_VL, B_HD, B_HU, NUM_BMODES };
enum { DC_PRED = 0, TM_PRED = 1, V_PRED = 2, H_PRED = 3, B_PRED = 4 };  // 16x16 / chroma modes (same numbering as their 4x4 counterparts)

static const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static const uint8_t kBands[17] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7, 0};
static const uint8_t kCat3[] = {173, 148, 140, 0}, kCat4[] = {176, 155, 140, 135, 0}, kCat5[] = {180, 157, 141, 134, 130, 0},
                     kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129, 0};
static const uint8_t* const kCat3456[4] = {kCat3, kCat4, kCat5, kCat6};

inline int clip(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

struct Frame {
    int width = 0, height = 0, mbw = 0, mbh = 0;
    int ys = 0, cs = 0;  // strides of the (macroblock-aligned) planes
    std::vector<uint8_t> Y, U, V;
};

struct Decoder {
    BoolDec br;                 // first partition
    std::vector<BoolDec> tok;   // token partitions
    uint8_t probs[4][8][3][11];
    // segment header
    bool seg_enabled = false, seg_update_map = false, seg_abs = false;
    int seg_quant[4] = {0, 0, 0, 0}, seg_lf[4] = {0, 0, 0, 0};
    uint8_t seg_probs[3] = {255, 255, 255};
    // filter header
    bool lf_simple = false, lf_delta = false;
    int lf_level = 0, lf_sharp = 0, ref_delta[4] = {0, 0, 0, 0}, mode_delta[4] = {0, 0, 0, 0};
    // quantisers per segment: [0] DC, [1] AC
    int y1q[4][2], y2q[4][2], uvq[4][2];
    bool use_skip = false;
    int skip_prob = 0;
    struct FInfo { int limit, ilevel, hev; };
    FInfo finfo[4][2];

    const char* err = nullptr;
    bool fail(const char* m) { err = m; return false; }

    bool parse_header(const uint8_t* data, size_t n, Frame& f)
    {
        if (n < 10) return fail("truncated VP8 frame header");
        const uint32_t tag = data[0] | (data[1] << 8) | (data[2] << 16);
        if (tag & 1) return fail("VP8: not a key frame");
        if (((tag >> 1) & 7) > 3) return fail("VP8: unknown profile");
        const uint32_t part0 = tag >> 5;
        if (data[3] != 0x9d || data[4] != 0x01 || data[5] != 0x2a) return fail("VP8: bad start code");
        f.width = (data[6] | (data[7] << 8)) & 0x3fff;
        f.height = (data[8] | (data[9] << 8)) & 0x3fff;
        if (f.width == 0 || f.height == 0) return fail("VP8: empty frame");
        if (part0 > n - 10) return fail("VP8: truncated first partition");
        f.mbw = (f.width + 15) >> 4;
        f.mbh = (f.height + 15) >> 4;
        br.init(data + 10, part0);
        br.get(128);  // colour space
        br.get(128);  // clamping type (the reconstruction clamps either way)
        seg_enabled = br.get(128);
        if (seg_enabled) {
            seg_update_map = br.get(128);
            if (br.get(128)) {  // update segment feature data
                seg_abs = br.get(128);
                for (int i = 0; i < 4; i++) seg_quant[i] = br.get(128) ? br.signed_literal(7) : 0;
                for (int i = 0; i < 4; i++) seg_lf[i] = br.get(128) ? br.signed_literal(6) : 0;
            }
            if (seg_update_map)
                for (int i = 0; i < 3; i++) seg_probs[i] = br.get(128) ? (uint8_t)br.literal(8) : 255;
        }
        lf_simple = br.get(128);
        lf_level = (int)br.literal(6);
        lf_sharp = (int)br.literal(3);
        lf_delta = br.get(128);
        if (lf_delta && br.get(128)) {  // delta update
            for (int i = 0; i < 4; i++)
                if (br.get(128)) ref_delta[i] = br.signed_literal(6);
            for (int i = 0; i < 4; i++)
                if (br.get(128)) mode_delta[i] = br.signed_literal(6);
        }
        const int nparts = 1 << br.literal(2);
        // token partitions: 3-byte sizes of all but the last, then the data
        const uint8_t* sz = data + 10 + part0;
        const uint8_t* end = data + n;
        if ((size_t)(end - sz) < (size_t)3 * (nparts - 1)) return fail("VP8: truncated partition table");
        const uint8_t* pd = sz + 3 * (nparts - 1);
        tok.resize(nparts);
        for (int i = 0; i < nparts; i++) {
            size_t len = i + 1 < nparts ? (size_t)(sz[3 * i] | (sz[3 * i + 1] << 8) | (sz[3 * i + 2] << 16)) : (size_t)(end - pd);
            if (len > (size_t)(end - pd)) len = (size_t)(end - pd);  // libwebp tolerates a short last partition the same way
            tok[i].init(pd, len);
            pd += len;
        }
        // quantiser indices
        const int yac = (int)br.literal(7);
        const int ydc = br.get(128) ? br.signed_literal(4) : 0;
        const int y2dc = br.get(128) ? br.signed_literal(4) : 0;
        const int y2ac = br.get(128) ? br.signed_literal(4) : 0;
        const int uvdc = br.get(128) ? br.signed_literal(4) : 0;
        const int uvac = br.get(128) ? br.signed_literal(4) : 0;
        for (int s = 0; s < 4; s++) {
            int q = yac;
            if (seg_enabled) q = seg_abs ? seg_quant[s] : yac + seg_quant[s];
            y1q[s][0] = kVp8DcQ[clip(q + ydc, 127)];
            y1q[s][1] = kVp8AcQ[clip(q, 127)];
            y2q[s][0] = kVp8DcQ[clip(q + y2dc, 127)] * 2;
            y2q[s][1] = (kVp8AcQ[clip(q + y2ac, 127)] * 101581) >> 16;  // x 155 / 100
            if (y2q[s][1] < 8) y2q[s][1] = 8;
            uvq[s][0] = kVp8DcQ[clip(q + uvdc, 117)];
            uvq[s][1] = kVp8AcQ[clip(q + uvac, 127)];
        }
        br.get(128);  // refresh_entropy_probs: irrelevant for a single frame
        memcpy(probs, kVp8CoeffProbs, sizeof probs);
        const uint8_t* up = kVp8CoeffUpdateProbs;
        for (int t = 0; t < 4; t++)
            for (int b = 0; b < 8; b++)
                for (int c = 0; c < 3; c++)
                    for (int p = 0; p < 11; p++)
                        if (br.get(*up++)) probs[t][b][c][p] = (uint8_t)br.literal(8);
        use_skip = br.get(128);
        skip_prob = use_skip ? (int)br.literal(8) : 0;
        // loop-filter strengths per segment and per "has 4x4 modes" (RFC 15.1 / 9.6)
        for (int s = 0; s < 4; s++) {
            int base = lf_level;
            if (seg_enabled) base = seg_abs ? seg_lf[s] : lf_level + seg_lf[s];
            for (int i4 = 0; i4 < 2; i4++) {
                int level = base;
                if (lf_delta) {
                    level += ref_delta[0];
                    if (i4) level += mode_delta[0];
                }
                level = clip(level, 63);
                FInfo& fi = finfo[s][i4];
                fi.limit = 0;
                fi.ilevel = 0;
                fi.hev = 0;
                if (level > 0) {
                    int il = level;
                    if (lf_sharp > 0) {
                        il >>= (lf_sharp > 4) ? 2 : 1;
                        if (il > 9 - lf_sharp) il = 9 - lf_sharp;
                    }
                    if (il < 1) il = 1;
                    fi.ilevel = il;
                    fi.limit = 2 * level + il;
                    fi.hev = level >= 40 ? 2 : (level >= 15 ? 1 : 0);
                }
            }
        }
        return true;
    }

    // one block's tokens (RFC 13.2-13.5); returns the index after the last decoded coefficient (0 = immediate end of block)
    int coeffs(BoolDec& t, int type, int ctx, const int* dq, int n, int16_t* out)
    {
        const uint8_t* p = probs[type][kBands[n]][ctx];
        for (; n < 16; n++) {
            if (!t.get(p[0])) return n;  // end of block
            while (!t.get(p[1])) {       // zeros
                p = probs[type][kBands[++n]][0];
                if (n == 16) return 16;
            }
            int v, next;
            if (!t.get(p[2])) { v = 1; next = 1; }
            else {
                if (!t.get(p[3])) {
                    if (!t.get(p[4])) v = 2;
                    else v = 3 + t.get(p[5]);
                } else if (!t.get(p[6])) {
                    if (!t.get(p[7])) v = 5 + t.get(159);
                    else { v = 7 + 2 * t.get(165); v += t.get(145); }
                } else {
                    const int b1 = t.get(p[8]);
                    const int b0 = t.get(p[9 + b1]);
                    const int cat = 2 * b1 + b0;
                    v = 0;
                    for (const uint8_t* tab = kCat3456[cat]; *tab; ++tab) v += v + t.get(*tab);
                    v += 3 + (8 << cat);
                }
                next = 2;
            }
            const int sgn = t.get(128);
            out[kZigzag[n]] = (int16_t)((sgn ? -v : v) * dq[n > 0]);
            p = probs[type][kBands[n + 1]][next];
        }
        return 16;
    }
};

// ---- inverse transforms (RFC 14.3, 14.4) -------------------------------------------------------------------------------------
inline void inverse_wht(const int16_t* in, int16_t* out /* DC of 16 blocks, stride 16 */)
{
    int tmp[16];
    for (int i = 0; i < 4; i++) {
        const int a0 = in[0 + i] + in[12 + i], a1 = in[4 + i] + in[8 + i], a2 = in[4 + i] - in[8 + i], a3 = in[0 + i] - in[12 + i];
        tmp[0 + i] = a0 + a1;
        tmp[8 + i] = a0 - a1;
        tmp[4 + i] = a3 + a2;
        tmp[12 + i] = a3 - a2;
    }
    for (int i = 0; i < 4; i++) {
        const int dc = tmp[0 + i * 4] + 3;
        const int a0 = dc + tmp[3 + i * 4], a1 = tmp[1 + i * 4] + tmp[2 + i * 4], a2 = tmp[1 + i * 4] - tmp[2 + i * 4], a3 = dc - tmp[3 + i * 4];
        out[0] = (int16_t)((a0 + a1) >> 3);
        out[16] = (int16_t)((a3 + a2) >> 3);
        out[32] = (int16_t)((a0 - a1) >> 3);
        out[48] = (int16_t)((a3 - a2) >> 3);
        out += 64;
    }
}

inline int mul1(int a) { return ((a * 20091) >> 16) + a; }
inline int mul2(int a) { return (a * 35468) >> 16; }

// adds the inverse DCT of in[16] to the 4x4 block at dst
inline void inverse_dct_add(const int16_t* in, uint8_t* dst, int stride)
{
    int tmp[16];
    int* t = tmp;
    for (int i = 0; i < 4; i++) {  // vertical pass
        const int a = in[0] + in[8], b = in[0] - in[8];
        const int c = mul2(in[4]) - mul1(in[12]), d = mul1(in[4]) + mul2(in[12]);
        t[0] = a + d;
        t[1] = b + c;
        t[2] = b - c;
        t[3] = a - d;
        t += 4;
        in++;
    }
    t = tmp;
    for (int i = 0; i < 4; i++) {  // horizontal pass
        const int dc = t[0] + 4;
        const int a = dc + t[8], b = dc - t[8];
        const int c = mul2(t[4]) - mul1(t[12]), d = mul1(t[4]) + mul2(t[12]);
        dst[0] = clip8(dst[0] + ((a + d) >> 3));
        dst[1] = clip8(dst[1] + ((b + c) >> 3));
        dst[2] = clip8(dst[2] + ((b - c) >> 3));
        dst[3] = clip8(dst[3] + ((a - d) >> 3));
        t++;
        dst += stride;
    }
}

// ---- intra prediction (RFC 12) -----------------------------------------------------------------------------------------------
// The working buffer of one macroblock holds the block plus one row above (with 4 more pixels to the right for the luma sub-blocks)
// and one column to the left, filled with neighbouring pixels or the frame-edge constants 127 (above) / 129 (left).
inline int avg2(int a, int b) { return (a + b + 1) >> 1; }
inline int avg3(int a, int b, int c) { return (a + 2 * b + c + 2) >> 2; }

inline void pred_block(uint8_t* d, int s, int size, int mode, bool has_top, bool has_left)
{
    const uint8_t* top = d - s;
    if (mode == DC_PRED) {
        int dc;
        const int sh = size == 16 ? 4 : (size == 8 ? 3 : 2);
        if (size == 4 || (has_top && has_left)) {  // 4x4 blocks always use both edges (constants at the frame border)
            dc = size;
            for (int i = 0; i < size; i++) dc += top[i] + d[i * s - 1];
            dc >>= sh + 1;
        } else if (has_top || has_left) {
            dc = size >> 1;
            for (int i = 0; i < size; i++) dc += has_top ? top[i] : d[i * s - 1];
            dc >>= sh;
        } else
            dc = 128;
        for (int y = 0; y < size; y++) memset(d + y * s, dc, size);
    } else if (mode == TM_PRED) {
        const int tl = top[-1];
        for (int y = 0; y < size; y++) {
            const int l = d[y * s - 1];
            for (int x = 0; x < size; x++) d[y * s + x] = clip8(top[x] + l - tl);
        }
    } else if (mode == V_PRED) {
        for (int y = 0; y < size; y++) memcpy(d + y * s, top, size);
    } else {  // H_PRED
        for (int y = 0; y < size; y++) memset(d + y * s, d[y * s - 1], size);
    }
}

inline void pred_sub(uint8_t* d, int s, int mode)
{
    const uint8_t* top = d - s;
    const int X = top[-1];
    const int A = top[0], B = top[1], C = top[2], D = top[3], E = top[4], F = top[5], G = top[6], H = top[7];
    const int I = d[-1], J = d[s - 1], K = d[2 * s - 1], L = d[3 * s - 1];
#define DST(x, y) d[(x) + (y) * s]
    switch (mode) {
    case B_DC: pred_block(d, s, 4, DC_PRED, true, true); break;
    case B_TM: pred_block(d, s, 4, TM_PRED, true, true); break;
    case B_VE: {
        const uint8_t v[4] = {(uint8_t)avg3(X, A, B), (uint8_t)avg3(A, B, C), (uint8_t)avg3(B, C, D), (uint8_t)avg3(C, D, E)};
        for (int y = 0; y < 4; y++) memcpy(d + y * s, v, 4);
        break;
    }
    case B_HE:
        memset(d, avg3(X, I, J), 4);
        memset(d + s, avg3(I, J, K), 4);
        memset(d + 2 * s, avg3(J, K, L), 4);
        memset(d + 3 * s, avg3(K, L, L), 4);
        break;
    case B_RD:
        DST(0, 3) = avg3(J, K, L);
        DST(0, 2) = DST(1, 3) = avg3(I, J, K);
        DST(0, 1) = DST(1, 2) = DST(2, 3) = avg3(X, I, J);
        DST(0, 0) = DST(1, 1) = DST(2, 2) = DST(3, 3) = avg3(A, X, I);
        DST(1, 0) = DST(2, 1) = DST(3, 2) = avg3(B, A, X);
        DST(2, 0) = DST(3, 1) = avg3(C, B, A);
        DST(3, 0) = avg3(D, C, B);
        break;
    case B_VR:
        DST(0, 0) = DST(1, 2) = avg2(X, A);
        DST(1, 0) = DST(2, 2) = avg2(A, B);
        DST(2, 0) = DST(3, 2) = avg2(B, C);
        DST(3, 0) = avg2(C, D);
        DST(0, 3) = avg3(K, J, I);
        DST(0, 2) = avg3(J, I, X);
        DST(0, 1) = DST(1, 3) = avg3(I, X, A);
        DST(1, 1) = DST(2, 3) = avg3(X, A, B);
        DST(2, 1) = DST(3, 3) = avg3(A, B, C);
        DST(3, 1) = avg3(B, C, D);
        break;
    case B_LD:
        DST(0, 0) = avg3(A, B, C);
        DST(1, 0) = DST(0, 1) = avg3(B, C, D);
        DST(2, 0) = DST(1, 1) = DST(0, 2) = avg3(C, D, E);
        DST(3, 0) = DST(2, 1) = DST(1, 2) = DST(0, 3) = avg3(D, E, F);
        DST(3, 1) = DST(2, 2) = DST(1, 3) = avg3(E, F, G);
        DST(3, 2) = DST(2, 3) = avg3(F, G, H);
        DST(3, 3) = avg3(G, H, H);
        break;
    case B_VL:
        DST(0, 0) = avg2(A, B);
        DST(1, 0) = DST(0, 2) = avg2(B, C);
        DST(2, 0) = DST(1, 2) = avg2(C, D);
        DST(3, 0) = DST(2, 2) = avg2(D, E);
        DST(0, 1) = avg3(A, B, C);
        DST(1, 1) = DST(0, 3) = avg3(B, C, D);
        DST(2, 1) = DST(1, 3) = avg3(C, D, E);
        DST(3, 1) = DST(2, 3) = avg3(D, E, F);
        DST(3, 2) = avg3(E, F, G);
        DST(3, 3) = avg3(F, G, H);
        break;
    case B_HD:
        DST(0, 0) = DST(2, 1) = avg2(I, X);
        DST(0, 1) = DST(2, 2) = avg2(J, I);
        DST(0, 2) = DST(2, 3) = avg2(K, J);
        DST(0, 3) = avg2(L, K);
        DST(3, 0) = avg3(A, B, C);
        DST(2, 0) = avg3(X, A, B);
        DST(1, 0) = DST(3, 1) = avg3(I, X, A);
        DST(1, 1) = DST(3, 2) = avg3(J, I, X);
        DST(1, 2) = DST(3, 3) = avg3(K, J, I);
        DST(1, 3) = avg3(L, K, J);
        break;
    default:  // B_HU
        DST(0, 0) = avg2(I, J);
        DST(2, 0) = DST(0, 1) = avg2(J, K);
        DST(2, 1) = DST(0, 2) = avg2(K, L);
        DST(1, 0) = avg3(I, J, K);
        DST(3, 0) = DST(1, 1) = avg3(J, K, L);
        DST(3, 1) = DST(1, 2) = avg3(K, L, L);
        DST(3, 2) = DST(2, 2) = DST(0, 3) = DST(1, 3) = DST(2, 3) = DST(3, 3) = L;
        break;
    }
#undef DST
}

// ---- loop filter (RFC 15) ----------------------------------------------------------------------------------------------------
inline int sclip1(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }   // signed 8-bit saturation
inline int sclip2(int v) { return v < -16 ? -16 : (v > 15 ? 15 : v); }
inline int iabs(int v) { return v < 0 ? -v : v; }

inline void filter2(uint8_t* p, int step)
{
    const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
    const int a = 3 * (q0 - p0) + sclip1(p1 - q1);
    const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3);
    p[-step] = clip8(p0 + a2);
    p[0] = clip8(q0 - a1);
}
inline void filter4(uint8_t* p, int step)
{
    const int p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step];
    const int a = 3 * (q0 - p0);
    const int a1 = sclip2((a + 4) >> 3), a2 = sclip2((a + 3) >> 3);
    const int a3 = (a1 + 1) >> 1;
    p[-2 * step] = clip8(p1 + a3);
    p[-step] = clip8(p0 + a2);
    p[0] = clip8(q0 - a1);
    p[step] = clip8(q1 - a3);
}
inline void filter6(uint8_t* p, int step)
{
    const int p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step], q2 = p[2 * step];
    const int a = sclip1(3 * (q0 - p0) + sclip1(p1 - q1));
    const int a1 = (27 * a + 63) >> 7, a2 = (18 * a + 63) >> 7, a3 = (9 * a + 63) >> 7;
    p[-3 * step] = clip8(p2 + a3);
    p[-2 * step] = clip8(p1 + a2);
    p[-step] = clip8(p0 + a1);
    p[0] = clip8(q0 - a1);
    p[step] = clip8(q1 - a2);
    p[2 * step] = clip8(q2 - a3);
}
inline bool hev(const uint8_t* p, int step, int t) { return iabs(p[-2 * step] - p[-step]) > t || iabs(p[step] - p[0]) > t; }
inline bool needs_filter(const uint8_t* p, int step, int t) { return 4 * iabs(p[-step] - p[0]) + iabs(p[-2 * step] - p[step]) <= t; }
inline bool needs_filter2(const uint8_t* p, int step, int t, int it)
{
    const int p3 = p[-4 * step], p2 = p[-3 * step], p1 = p[-2 * step], p0 = p[-step], q0 = p[0], q1 = p[step], q2 = p[2 * step], q3 = p[3 * step];
    if (4 * iabs(p0 - q0) + iabs(p1 - q1) > t) return false;
    return iabs(p3 - p2) <= it && iabs(p2 - p1) <= it && iabs(p1 - p0) <= it && iabs(q3 - q2) <= it && iabs(q2 - q1) <= it && iabs(q1 - q0) <= it;
}
// an edge of `size` pixels: hstride across it, vstride along it
inline void edge_simple(uint8_t* p, int hs, int vs, int size, int thresh)
{
    const int t2 = 2 * thresh + 1;
    for (int i = 0; i < size; i++, p += vs)
        if (needs_filter(p, hs, t2)) filter2(p, hs);
}
inline void edge_mb(uint8_t* p, int hs, int vs, int size, int thresh, int it, int hv)
{
    const int t2 = 2 * thresh + 1;
    for (int i = 0; i < size; i++, p += vs)
        if (needs_filter2(p, hs, t2, it)) {
            if (hev(p, hs, hv)) filter2(p, hs);
            else filter6(p, hs);
        }
}
inline void edge_inner(uint8_t* p, int hs, int vs, int size, int thresh, int it, int hv)
{
    const int t2 = 2 * thresh + 1;
    for (int i = 0; i < size; i++, p += vs)
        if (needs_filter2(p, hs, t2, it)) {
            if (hev(p, hs, hv)) filter2(p, hs);
            else filter4(p, hs);
        }
}

// ---- the frame ---------------------------------------------------------------------------------------------------------------
inline bool decode_frame(const uint8_t* data, size_t n, Frame& f, const char** why)
{
    Decoder d;
    if (!d.parse_header(data, n, f)) { *why = d.err; return false; }
    const int mbw = f.mbw, mbh = f.mbh;
    f.ys = mbw * 16;
    f.cs = mbw * 8;
    f.Y.assign((size_t)f.ys * mbh * 16, 0);
    f.U.assign((size_t)f.cs * mbh * 8, 0);
    f.V.assign((size_t)f.cs * mbh * 8, 0);
    // per-macroblock state kept for the loop filter; contexts of the row above / the macroblock to the left
    struct MBInfo { uint8_t seg, i4, inner; };
    std::vector<MBInfo> info((size_t)mbw * mbh);
    std::vector<uint8_t> top_modes((size_t)mbw * 4, B_DC);  // sub-block modes of the row above
    std::vector<uint8_t> top_nz((size_t)mbw * 9, 0);        // per macroblock: 4 Y, 2 U, 2 V columns + Y2
    // working buffer: 1 + 16 rows of 1 + 16 + 4 pixels (luma), 1 + 8 rows of 1 + 8 (chroma)
    const int BS = 32;
    uint8_t ybuf[17 * BS], ubuf[9 * BS], vbuf[9 * BS];
    for (int my = 0; my < mbh; my++) {
        uint8_t left_modes[4] = {B_DC, B_DC, B_DC, B_DC};
        uint8_t left_nz[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        BoolDec& t = d.tok[my & ((int)d.tok.size() - 1)];
        for (int mx = 0; mx < mbw; mx++) {
            MBInfo& mi = info[(size_t)my * mbw + mx];
            // -- modes (RFC 9.3, 11.2-11.4; key frame) --
            int seg = 0;
            if (d.seg_update_map) seg = !d.br.get(d.seg_probs[0]) ? d.br.get(d.seg_probs[1]) : 2 + d.br.get(d.seg_probs[2]);
            const bool skip_flag = d.use_skip ? d.br.get(d.skip_prob) : false;
            int ymode;
            uint8_t bm[16];
            if (!d.br.get(145)) {
                ymode = B_PRED;
                for (int by = 0; by < 4; by++)
                    for (int bx = 0; bx < 4; bx++) {
                        const int A = by == 0 ? top_modes[mx * 4 + bx] : bm[(by - 1) * 4 + bx];
                        const int Lm = bx == 0 ? left_modes[by] : bm[by * 4 + bx - 1];
                        const uint8_t* p = &kVp8BModeProbs[(A * NUM_BMODES + Lm) * 9];
                        int m;
                        if (!d.br.get(p[0])) m = B_DC;
                        else if (!d.br.get(p[1])) m = B_TM;
                        else if (!d.br.get(p[2])) m = B_VE;
                        else if (!d.br.get(p[3])) {
                            if (!d.br.get(p[4])) m = B_HE;
                            else m = !d.br.get(p[5]) ? B_RD : B_VR;
                        } else if (!d.br.get(p[6])) m = B_LD;
                        else if (!d.br.get(p[7])) m = B_VL;
                        else m = !d.br.get(p[8]) ? B_HD : B_HU;
                        bm[by * 4 + bx] = (uint8_t)m;
                    }
            } else {
                ymode = d.br.get(156) ? (d.br.get(128) ? TM_PRED : H_PRED) : (d.br.get(163) ? V_PRED : DC_PRED);
                memset(bm, ymode, 16);  // the implied sub-block modes share the numbering (DC, TM, V -> VE, H -> HE)
            }
            for (int i = 0; i < 4; i++) {
                top_modes[mx * 4 + i] = bm[12 + i];
                left_modes[i] = bm[i * 4 + 3];
            }
            const int uvmode = !d.br.get(142) ? DC_PRED : (!d.br.get(114) ? V_PRED : (d.br.get(183) ? TM_PRED : H_PRED));

            // -- residual tokens (RFC 13) --
            int16_t coef[25 * 16];
            memset(coef, 0, sizeof coef);
            uint8_t* tnz = &top_nz[(size_t)mx * 9];
            bool any = false;
            if (!skip_flag) {
                int first = 0, ytype = 3;
                if (ymode != B_PRED) {
                    int16_t dc[16] = {0};
                    const int nz = d.coeffs(t, 1, tnz[8] + left_nz[8], d.y2q[seg], 0, dc);
                    tnz[8] = left_nz[8] = nz > 0;
                    inverse_wht(dc, coef);
                    first = 1;
                    ytype = 0;
                }
                for (int by = 0; by < 4; by++)
                    for (int bx = 0; bx < 4; bx++) {
                        const int nz = d.coeffs(t, ytype, tnz[bx] + left_nz[by], d.y1q[seg], first, coef + (by * 4 + bx) * 16);
                        tnz[bx] = left_nz[by] = nz > first;
                    }
                for (int c = 0; c < 2; c++)  // U then V: 2 x 2 blocks each
                    for (int by = 0; by < 2; by++)
                        for (int bx = 0; bx < 2; bx++) {
                            const int nz = d.coeffs(t, 2, tnz[4 + 2 * c + bx] + left_nz[4 + 2 * c + by], d.uvq[seg], 0, coef + (16 + c * 4 + by * 2 + bx) * 16);
                            tnz[4 + 2 * c + bx] = left_nz[4 + 2 * c + by] = nz > 0;
                        }
            } else {
                memset(tnz, 0, 8);
                memset(left_nz, 0, 8);
                if (ymode != B_PRED) tnz[8] = left_nz[8] = 0;
            }
            for (int i = 0; i < 24 * 16 && !any; i++) any = coef[i] != 0;  // incl. a DC that only the WHT produced
            mi.seg = (uint8_t)seg;
            mi.i4 = ymode == B_PRED;
            mi.inner = mi.i4 || any;

            // -- prediction + reconstruction (RFC 12, 14.5) --
            uint8_t* yd = ybuf + BS + 1;
            uint8_t* ud = ubuf + BS + 1;
            uint8_t* vd = vbuf + BS + 1;
            uint8_t* Yp = &f.Y[(size_t)my * 16 * f.ys + mx * 16];
            uint8_t* Up = &f.U[(size_t)my * 8 * f.cs + mx * 8];
            uint8_t* Vp = &f.V[(size_t)my * 8 * f.cs + mx * 8];
            // edges: the row above (127 at the top of the frame), the column to the left (129 at the left of the frame)
            if (my > 0) {
                memcpy(yd - BS, Yp - f.ys, 16);
                memcpy(ud - BS, Up - f.cs, 8);
                memcpy(vd - BS, Vp - f.cs, 8);
                if (mx + 1 < mbw) memcpy(yd - BS + 16, Yp - f.ys + 16, 4);
                else memset(yd - BS + 16, Yp[-f.ys + 15], 4);  // right-most macroblock: the last pixel above, repeated
            } else {
                memset(yd - BS, 127, 20);
                memset(ud - BS, 127, 8);
                memset(vd - BS, 127, 8);
            }
            for (int i = 0; i < 16; i++) yd[i * BS - 1] = mx > 0 ? Yp[(size_t)i * f.ys - 1] : 129;
            for (int i = 0; i < 8; i++) {
                ud[i * BS - 1] = mx > 0 ? Up[(size_t)i * f.cs - 1] : 129;
                vd[i * BS - 1] = mx > 0 ? Vp[(size_t)i * f.cs - 1] : 129;
            }
            const int corner_y = my > 0 ? (mx > 0 ? Yp[-f.ys - 1] : 129) : 127;
            yd[-BS - 1] = (uint8_t)corner_y;
            ud[-BS - 1] = (uint8_t)(my > 0 ? (mx > 0 ? Up[-f.cs - 1] : 129) : 127);
            vd[-BS - 1] = (uint8_t)(my > 0 ? (mx > 0 ? Vp[-f.cs - 1] : 129) : 127);
            if (ymode == B_PRED) {
                // sub-blocks of rows 1..3 in the last column see the same 4 above-right pixels as the first row (RFC 12.3)
                for (int r = 1; r < 4; r++) memcpy(yd + (r * 4 - 1) * BS + 16, yd - BS + 16, 4);
                for (int by = 0; by < 4; by++)
                    for (int bx = 0; bx < 4; bx++) {
                        uint8_t* b = yd + by * 4 * BS + bx * 4;
                        pred_sub(b, BS, bm[by * 4 + bx]);
                        inverse_dct_add(coef + (by * 4 + bx) * 16, b, BS);
                    }
            } else {
                pred_block(yd, BS, 16, ymode, my > 0, mx > 0);
                for (int by = 0; by < 4; by++)
                    for (int bx = 0; bx < 4; bx++) inverse_dct_add(coef + (by * 4 + bx) * 16, yd + by * 4 * BS + bx * 4, BS);
            }
            pred_block(ud, BS, 8, uvmode, my > 0, mx > 0);
            pred_block(vd, BS, 8, uvmode, my > 0, mx > 0);
            for (int by = 0; by < 2; by++)
                for (int bx = 0; bx < 2; bx++) {
                    inverse_dct_add(coef + (16 + by * 2 + bx) * 16, ud + by * 4 * BS + bx * 4, BS);
                    inverse_dct_add(coef + (20 + by * 2 + bx) * 16, vd + by * 4 * BS + bx * 4, BS);
                }
            for (int i = 0; i < 16; i++) memcpy(Yp + (size_t)i * f.ys, yd + i * BS, 16);
            for (int i = 0; i < 8; i++) {
                memcpy(Up + (size_t)i * f.cs, ud + i * BS, 8);
                memcpy(Vp + (size_t)i * f.cs, vd + i * BS, 8);
            }
        }
    }
    // -- loop filter, macroblocks in raster order: left edge, inner vertical edges, top edge, inner horizontal edges (RFC 15.2-15.3)
    if (d.lf_level > 0) {
        for (int my = 0; my < mbh; my++)
            for (int mx = 0; mx < mbw; mx++) {
                const MBInfo& mi = info[(size_t)my * mbw + mx];
                const Decoder::FInfo& fi = d.finfo[mi.seg][mi.i4];
                const int limit = fi.limit, il = fi.ilevel, hv = fi.hev;
                if (limit == 0) continue;
                uint8_t* Yp = &f.Y[(size_t)my * 16 * f.ys + mx * 16];
                uint8_t* Up = &f.U[(size_t)my * 8 * f.cs + mx * 8];
                uint8_t* Vp = &f.V[(size_t)my * 8 * f.cs + mx * 8];
                const int ys = f.ys, cs = f.cs;
                if (d.lf_simple) {
                    if (mx > 0) edge_simple(Yp, 1, ys, 16, limit + 4);
                    if (mi.inner)
                        for (int k = 4; k < 16; k += 4) edge_simple(Yp + k, 1, ys, 16, limit);
                    if (my > 0) edge_simple(Yp, ys, 1, 16, limit + 4);
                    if (mi.inner)
                        for (int k = 4; k < 16; k += 4) edge_simple(Yp + (size_t)k * ys, ys, 1, 16, limit);
                } else {
                    if (mx > 0) {
                        edge_mb(Yp, 1, ys, 16, limit + 4, il, hv);
                        edge_mb(Up, 1, cs, 8, limit + 4, il, hv);
                        edge_mb(Vp, 1, cs, 8, limit + 4, il, hv);
                    }
                    if (mi.inner) {
                        for (int k = 4; k < 16; k += 4) edge_inner(Yp + k, 1, ys, 16, limit, il, hv);
                        edge_inner(Up + 4, 1, cs, 8, limit, il, hv);
                        edge_inner(Vp + 4, 1, cs, 8, limit, il, hv);
                    }
                    if (my > 0) {
                        edge_mb(Yp, ys, 1, 16, limit + 4, il, hv);
                        edge_mb(Up, cs, 1, 8, limit + 4, il, hv);
                        edge_mb(Vp, cs, 1, 8, limit + 4, il, hv);
                    }
                    if (mi.inner) {
                        for (int k = 4; k < 16; k += 4) edge_inner(Yp + (size_t)k * ys, ys, 1, 16, limit, il, hv);
                        edge_inner(Up + (size_t)4 * cs, cs, 1, 8, limit, il, hv);
                        edge_inner(Vp + (size_t)4 * cs, cs, 1, 8, limit, il, hv);
                    }
                }
            }
    }
    return true;
}

// ---- YUV 4:2:0 -> RGB the way WebPDecodeRGB does it: 9-3-3-1 chroma upsampling, 14-bit fixed-point BT.601 ----------------------
inline int mult_hi(int v, int c) { return (v * c) >> 8; }
inline uint8_t yuv_clip(int v) { return (v & ~16383) == 0 ? (uint8_t)(v >> 6) : (v < 0 ? 0 : 255); }
inline void yuv_to_rgb(int y, int u, int v, uint8_t* rgb)
{
    rgb[0] = yuv_clip(mult_hi(y, 19077) + mult_hi(v, 26149) - 14234);
    rgb[1] = yuv_clip(mult_hi(y, 19077) - mult_hi(u, 6419) - mult_hi(v, 13320) + 8708);
    rgb[2] = yuv_clip(mult_hi(y, 19077) + mult_hi(u, 33050) - 17685);
}

// one output row: `near` = the chroma row this luma row sits closer to, `far` = the other one (3:1 vertically)
inline void upsample_row(const uint8_t* y, const uint8_t* nu, const uint8_t* nv, const uint8_t* fu, const uint8_t* fv, int w, uint8_t* dst, int step)
{
    const int last = (w - 1) >> 1;
    int tlu = nu[0], tlv = nv[0], lu = fu[0], lv = fv[0];
    yuv_to_rgb(y[0], (3 * tlu + lu + 2) >> 2, (3 * tlv + lv + 2) >> 2, dst);
    for (int x = 1; x <= last; x++) {
        const int tu = nu[x], tv = nv[x], cu = fu[x], cv = fv[x];
        const int au = tlu + tu + lu + cu + 8, av = tlv + tv + lv + cv + 8;
        const int d12u = (au + 2 * (tu + lu)) >> 3, d12v = (av + 2 * (tv + lv)) >> 3;
        const int d03u = (au + 2 * (tlu + cu)) >> 3, d03v = (av + 2 * (tlv + cv)) >> 3;
        yuv_to_rgb(y[2 * x - 1], (d12u + tlu) >> 1, (d12v + tlv) >> 1, dst + (size_t)(2 * x - 1) * step);
        yuv_to_rgb(y[2 * x], (d03u + tu) >> 1, (d03v + tv) >> 1, dst + (size_t)(2 * x) * step);
        tlu = tu; tlv = tv; lu = cu; lv = cv;
    }
    if (!(w & 1)) yuv_to_rgb(y[w - 1], (3 * tlu + lu + 2) >> 2, (3 * tlv + lv + 2) >> 2, dst + (size_t)(w - 1) * step);
}

// f -> interleaved pixels with `step` bytes per pixel (3 = RGB, 4 = RGBA with byte 3 left alone)
inline void frame_to_rgb(const Frame& f, uint8_t* out, int step)
{
    const int w = f.width, h = f.height, ch = (h + 1) >> 1;
    for (int y = 0; y < h; y++) {
        int near, far;
        if (y == 0) near = far = 0;
        else if (y & 1) { near = (y - 1) >> 1; far = std::min(near + 1, ch - 1); }
        else { near = y >> 1; far = near - 1; }
        upsample_row(&f.Y[(size_t)y * f.ys], &f.U[(size_t)near * f.cs], &f.V[(size_t)near * f.cs], &f.U[(size_t)far * f.cs], &f.V[(size_t)far * f.cs], w,
                     out + (size_t)y * w * step, step);
    }
}

}  // namespace vp8
}  // namespace rsrhost
