// png_codec.h -- minimal PNG reader / writer on top of zlib, for the C++ command line (host/realsr_cli.cpp).
//
// The reference decodes with its vendored stb_image (src/main.cpp:243-263: auto channel count, gray -> RGB,
// gray+alpha -> RGBA) and encodes with stb_image_write (src/main.cpp:383); those files are the reference's and are not
// copied here.  This is an independent implementation of the PNG specification (ISO/IEC 15948): chunk parsing with
// CRC check, zlib inflate, the five scanline filters, bit depths 1/2/4/8/16, colour types 0/2/3/4/6, tRNS, and Adam7
// interlace on the read side; 8-bit RGB / RGBA on the write side, in a fast mode built for 33-megapixel results (SIMD Sub / Up /
// Paeth selection per row + an own run-length / Huffman deflate, rows compressed in parallel bands; zlib is only used for Adler-32 and
// CRC-32 there) or with per-row filter selection and zlib's full deflate.  Host plumbing only: no pixel of the super-resolution path is
// computed here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace rsrhost {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put_be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

inline int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the scanline filters in place; `raw` holds h rows of (1 filter byte + stride bytes)
inline bool png_unfilter(uint8_t* raw, size_t stride, int h, int bpp /* bytes per complete pixel, >= 1 */)
{
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (int y = 0; y < h; y++) {
        uint8_t* row = raw + (size_t)y * (stride + 1);
        const int ft = row[0];
        uint8_t* cur = row + 1;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
            int add;
            switch (ft) {
            case 0: add = 0; break;
            case 1: add = a; break;
            case 2: add = b; break;
            case 3: add = (a + b) >> 1; break;
            case 4: add = paeth(a, b, c); break;
            default: return false;
            }
            cur[i] = (uint8_t)(cur[i] + add);
        }
        prev = cur;
    }
    return true;
}

// Decodes `path`.  On success: w, h, c in {3, 4} (gray -> RGB, gray+alpha / tRNS -> RGBA, like the reference's load stage),
// pixels tightly packed HWC 8 bit.
inline bool png_decode_file(const std::string& path, int& w, int& h, int& c, std::vector<uint8_t>& pixels, std::string* why = nullptr)
{
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
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 8 || memcmp(f.data(), sig, 8)) return fail("not a PNG");
    size_t pos = 8;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_ihdr = false, have_iend = false;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        if (len > f.size() - pos - 12) return fail("truncated chunk");
        const uint8_t* type = &f[pos + 4];
        const uint8_t* data = &f[pos + 8];
        const uint32_t crc = be32(&f[pos + 8 + len]);
        if (crc32(crc32(0L, Z_NULL, 0), type, len + 4) != crc) return fail("chunk CRC mismatch");
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return fail("bad IHDR");
            w = (int)be32(data); h = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (w <= 0 || h <= 0 || (long long)w * h > (1LL << 30) || data[10] || data[11] || interlace > 1) return fail("unsupported IHDR");
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) { have_iend = true; break; }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || !have_iend || idat.empty()) return fail("missing IHDR / IDAT / IEND");
    int nch;
    switch (ctype) {
    case 0: nch = 1; break;
    case 2: nch = 3; break;
    case 3: nch = 1; break;
    case 4: nch = 2; break;
    case 6: nch = 4; break;
    default: return fail("bad colour type");
    }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok) return fail("bad bit depth");
    if (ctype == 3 && plte.size() < 3) return fail("palette image without PLTE");
    const int bits_pp = nch * depth;
    const int bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    // pass geometry: non-interlaced = one pass over everything; Adam7 = 7 sub-images
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    const Pass whole = {0, 0, 1, 1};
    const int npass = interlace ? 7 : 1;
    size_t raw_size = 0;
    for (int p = 0; p < npass; p++) {
        const Pass& ps = interlace ? adam7[p] : whole;
        const int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw > 0 && ph > 0) raw_size += (size_t)ph * (1 + ((size_t)pw * bits_pp + 7) / 8);
    }
    std::vector<uint8_t> raw(raw_size);
    {
        uLongf dl = (uLongf)raw_size;
        const int zr = uncompress(raw.data(), &dl, idat.data(), (uLong)idat.size());
        if (zr != Z_OK || dl != raw_size) return fail("inflate failed");
    }
    const bool has_alpha = ctype == 4 || ctype == 6 || !trns.empty();
    c = has_alpha ? 4 : 3;
    pixels.assign((size_t)w * h * c, 255);
    size_t off = 0;
    for (int p = 0; p < npass; p++) {
        const Pass& ps = interlace ? adam7[p] : whole;
        const int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
        if (!png_unfilter(raw.data() + off, stride, ph, bpp)) return fail("bad filter type");
        for (int y = 0; y < ph; y++) {
            const uint8_t* row = raw.data() + off + (size_t)y * (stride + 1) + 1;
            uint8_t* dst_row = pixels.data() + ((size_t)(ps.y0 + y * ps.dy) * w) * c;
            for (int x = 0; x < pw; x++) {
                // sample k of pixel x -> 8 bit (16 bit: high byte; < 8 bit gray: scaled to 0..255; palette: index)
                int s[4] = {0, 0, 0, 0};
                uint16_t s16[4] = {0, 0, 0, 0};
                for (int k = 0; k < nch; k++) {
                    if (depth == 8) s[k] = row[(size_t)x * nch + k];
                    else if (depth == 16) {
                        s16[k] = (uint16_t)((row[((size_t)x * nch + k) * 2] << 8) | row[((size_t)x * nch + k) * 2 + 1]);
                        s[k] = s16[k] >> 8;
                    } else {
                        const size_t bit = (size_t)x * depth;
                        const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                        s16[k] = (uint16_t)v;
                        s[k] = ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
                    }
                    if (depth == 8) s16[k] = (uint16_t)s[k];
                }
                uint8_t* d = dst_row + (size_t)(ps.x0 + x * ps.dx) * c;
                int r, g, b, a = 255;
                if (ctype == 3) {
                    const size_t idx = (size_t)s[0];
                    if (idx * 3 + 2 >= plte.size()) return fail("palette index out of range");
                    r = plte[idx * 3]; g = plte[idx * 3 + 1]; b = plte[idx * 3 + 2];
                    if (idx < trns.size()) a = trns[idx];
                } else if (ctype == 0 || ctype == 4) {
                    r = g = b = s[0];
                    if (ctype == 4) a = s[1];
                    else if (trns.size() >= 2 && s16[0] == (uint16_t)((trns[0] << 8) | trns[1])) a = 0;
                } else {
                    r = s[0]; g = s[1]; b = s[2];
                    if (ctype == 6) a = s[3];
                    else if (trns.size() >= 6 && s16[0] == (uint16_t)((trns[0] << 8) | trns[1]) && s16[1] == (uint16_t)((trns[2] << 8) | trns[3]) &&
                             s16[2] == (uint16_t)((trns[4] << 8) | trns[5]))
                        a = 0;
                }
                d[0] = (uint8_t)r; d[1] = (uint8_t)g; d[2] = (uint8_t)b;
                if (c == 4) d[3] = (uint8_t)a;
            }
        }
        off += (size_t)ph * (stride + 1);
    }
    return true;
}

// ---- fast deflate for filtered scanlines (the level <= 1 writer) ----------------------------------------------------------------------
// Filtered rows of a photograph hold almost no repeats that an LZ77 search could use except runs of one byte (flat areas), so this
// encoder emits per 128-KB block: literals + run matches (length 3..258 at distance 1) under ONE dynamic Huffman code built from the
// block's histogram (RFC 1951 3.2.7), or a stored block when that would be larger (noise).  Two passes over the block (count, emit),
// no hash chains: ~4x the throughput of zlib's level 1 / Z_RLE at the same size.
namespace fastz {

struct BitWriter {
    uint8_t* p;     // the caller guarantees room (+ 8 bytes of slack: every put stores the whole accumulator)
    uint64_t acc = 0;
    int nb = 0;     // < 8 between calls
    void put(uint32_t v, int n)  // n <= 32; branch-free: store, then advance by the whole bytes
    {
        acc |= (uint64_t)v << nb;
        nb += n;
        memcpy(p, &acc, 8);  // little-endian host
        p += nb >> 3;
        acc >>= nb & ~7;
        nb &= 7;
    }
    void align()
    {
        if (nb > 0) *p++ = (uint8_t)acc;
        acc = 0;
        nb = 0;
    }
};

// Huffman code lengths for n symbols, at most `limit` bits (two-queue construction; over-long codes are clipped and the Kraft sum repaired)
inline void code_lengths(const uint32_t* hist, int n, int limit, uint8_t* len)
{
    memset(len, 0, (size_t)n);
    int used[288], nu = 0;
    for (int i = 0; i < n; i++)
        if (hist[i]) used[nu++] = i;
    if (nu == 0) return;
    if (nu == 1) { len[used[0]] = 1; return; }
    std::sort(used, used + nu, [&](int a, int b) { return hist[a] < hist[b] || (hist[a] == hist[b] && a < b); });
    uint64_t w[576];
    int l[576], r[576], depth[576];
    for (int i = 0; i < nu; i++) { w[i] = hist[used[i]]; l[i] = r[i] = -1; }
    int a = 0, b = nu, nn = nu;
    auto pop = [&]() { return (a < nu && (b >= nn || w[a] <= w[b])) ? a++ : b++; };
    while ((nu - a) + (nn - b) > 1) {
        const int x = pop(), y = pop();
        w[nn] = w[x] + w[y];
        l[nn] = x;
        r[nn] = y;
        nn++;
    }
    depth[nn - 1] = 0;
    for (int i = nn - 1; i >= nu; i--) depth[l[i]] = depth[r[i]] = depth[i] + 1;
    long long kraft = 0;
    for (int i = 0; i < nu; i++) {
        len[used[i]] = (uint8_t)std::min(depth[i], limit);
        kraft += 1LL << (limit - len[used[i]]);
    }
    // Only after clipping can the code be over-subscribed.  used[] is sorted by rising frequency: lengthen the rarest codes that are
    // still short until it fits, then hand the spare room back, always to the most frequent symbol whose step still fits.  Every term of
    // the Kraft sum is a multiple of the smallest one, so the deficit is too and the loop ends at exactly 2^limit (inflate rejects an
    // incomplete literal/length code).
    for (int i = 0; kraft > (1LL << limit) && i < nu;) {
        uint8_t& L = len[used[i]];
        if (L < limit) { kraft -= 1LL << (limit - L - 1); L++; }
        else i++;
    }
    while (kraft < (1LL << limit)) {
        bool moved = false;
        for (int i = nu - 1; i >= 0; i--) {
            uint8_t& L = len[used[i]];
            if (L > 1 && kraft + (1LL << (limit - L)) <= (1LL << limit)) { kraft += 1LL << (limit - L); L--; moved = true; break; }
        }
        if (!moved) break;
    }
}

// canonical codes (RFC 1951 3.2.2), bit-reversed because deflate packs Huffman codes starting from their most significant bit
inline void canonical_codes(const uint8_t* len, int n, uint16_t* code)
{
    int count[16] = {0}, next[16];
    for (int i = 0; i < n; i++) count[len[i]]++;
    count[0] = 0;
    int c = 0;
    for (int b = 1; b < 16; b++) { c = (c + count[b - 1]) << 1; next[b] = c; }
    for (int i = 0; i < n; i++) {
        if (!len[i]) { code[i] = 0; continue; }
        int v = next[len[i]]++, rv = 0;
        for (int k = 0; k < len[i]; k++) rv |= ((v >> k) & 1) << (len[i] - 1 - k);
        code[i] = (uint16_t)rv;
    }
}

struct LenTab {
    uint16_t sym[259];
    uint8_t xbits[259];
    uint16_t xval[259];
    LenTab()
    {
        static const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const int xb[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        for (int L = 3; L <= 258; L++) {
            int k = 28;
            while (base[k] > L) k--;
            if (L == 258) k = 28;
            else if (k == 28) k = 27;  // 227..257 belong to code 284
            sym[L] = (uint16_t)(257 + k);
            xbits[L] = (uint8_t)xb[k];
            xval[L] = (uint16_t)(L - base[k]);
        }
    }
};

// Does a run of >= 3 bytes equal to d[i - 1] start at i (i >= 1, i + 3 <= n)?  One 4-byte compare: in textured areas the branch on it
// is almost never taken, in flat areas almost always -- a byte-by-byte "equal to the previous one" test would mispredict all the time.
inline bool run_starts(const uint8_t* d, size_t i)
{
    uint32_t v;
    memcpy(&v, d + i - 1, 4);
    return v == (v & 0xff) * 0x01010101u;
}
// its length, 3..258, limited to n
inline int run_len(const uint8_t* d, size_t i, size_t n)
{
    const uint8_t v = d[i - 1];
    const size_t lim = std::min(n, i + 258);
    size_t j = i + 3;
    while (j < lim && d[j] == v) j++;
    return (int)(j - i);
}

// appends raw-deflate blocks for d[0..n) to out (never the final block: the caller closes the stream); returns the new end
inline uint8_t* deflate_blocks(const uint8_t* d, size_t n, uint8_t* out)
{
    static const LenTab lt;
    static const uint8_t clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const size_t BLOCK = 128 * 1024;
    for (size_t b0 = 0; b0 < n; b0 += BLOCK) {
        const uint8_t* blk = d + b0;
        const size_t bn = std::min(BLOCK, n - b0);
        // pass 1: histogram of literals / run lengths (runs may look back into the previous block: distance 1 is always valid there)
        uint32_t hist[286], h4[4][256];  // four literal histograms: neighbouring residuals are often equal, one counter would serialise
        memset(hist, 0, sizeof hist);
        memset(h4, 0, sizeof h4);
        uint64_t xbits_total = 0;
        const size_t first = b0 ? 0 : 1;              // blk[-1] is the previous block's last byte; the band's first byte has none
        const size_t safe = bn >= 3 ? bn - 3 : 0;     // run_starts reads blk[i - 1 .. i + 2]
        if (first) h4[0][blk[0]]++;
        size_t i = first;
        for (; i < safe;) {
            if (run_starts(blk, i)) {
                const int r = run_len(blk, i, bn);
                hist[lt.sym[r]]++;
                xbits_total += lt.xbits[r];
                i += (size_t)r;
                continue;
            }
            h4[i & 3][blk[i]]++;
            i++;
        }
        for (; i < bn; i++) h4[i & 3][blk[i]]++;
        for (int k = 0; k < 256; k++) hist[k] = h4[0][k] + h4[1][k] + h4[2][k] + h4[3][k];
        hist[256] = 1;
        uint8_t ll[286], dl[1] = {1};
        uint16_t lc[286];
        code_lengths(hist, 286, 15, ll);
        int nlit = 286;
        while (nlit > 257 && !ll[nlit - 1]) nlit--;
        // code-length alphabet: the lengths with runs of zeros folded into symbols 17 / 18
        uint8_t seq[300], seqx[300];
        int ns = 0;
        uint32_t clh[19];
        memset(clh, 0, sizeof clh);
        {
            uint8_t all[287];
            memcpy(all, ll, (size_t)nlit);
            all[nlit] = dl[0];
            const int na = nlit + 1;
            for (int i = 0; i < na;) {
                if (all[i] == 0) {
                    int z = 1;
                    while (i + z < na && all[i + z] == 0 && z < 138) z++;
                    if (z >= 11) { seq[ns] = 18; seqx[ns++] = (uint8_t)(z - 11); i += z; continue; }
                    if (z >= 3) { seq[ns] = 17; seqx[ns++] = (uint8_t)(z - 3); i += z; continue; }
                }
                seq[ns] = all[i]; seqx[ns++] = 0; i++;
            }
            for (int i = 0; i < ns; i++) clh[seq[i]]++;
        }
        uint8_t cll[19];
        uint16_t clc[19];
        code_lengths(clh, 19, 7, cll);
        int ncl = 19;
        while (ncl > 4 && !cll[clorder[ncl - 1]]) ncl--;
        uint64_t bits = 3 + 5 + 5 + 4 + 3 * (uint64_t)ncl + xbits_total;
        for (int i = 0; i < ns; i++) bits += cll[seq[i]] + (seq[i] == 17 ? 3 : (seq[i] == 18 ? 7 : 0));
        for (int i = 0; i < 286; i++) bits += (uint64_t)hist[i] * ll[i];
        for (int i = 257; i < 286; i++) bits += hist[i];  // one distance bit per match
        BitWriter bw;
        bw.p = out;
        if (bits > 8 * (uint64_t)bn + 40 * ((bn + 65534) / 65535)) {  // stored blocks (at most 65535 bytes each)
            for (size_t o = 0; o < bn; o += 65535) {
                const uint32_t len = (uint32_t)std::min<size_t>(65535, bn - o);
                *out++ = 0;  // BFINAL = 0, BTYPE = 00, padding to the byte boundary (every block here starts aligned)
                *out++ = (uint8_t)len; *out++ = (uint8_t)(len >> 8); *out++ = (uint8_t)~len; *out++ = (uint8_t)(~len >> 8);
                memcpy(out, blk + o, len);
                out += len;
            }
            continue;
        }
        canonical_codes(ll, 286, lc);
        canonical_codes(cll, 19, clc);
        bw.put(0, 1);  // BFINAL
        bw.put(2, 2);  // dynamic Huffman
        bw.put((uint32_t)(nlit - 257), 5);
        bw.put(0, 5);  // one distance code
        bw.put((uint32_t)(ncl - 4), 4);
        for (int i = 0; i < ncl; i++) bw.put(cll[clorder[i]], 3);
        for (int i = 0; i < ns; i++) {
            bw.put(clc[seq[i]], cll[seq[i]]);
            if (seq[i] == 17) bw.put(seqx[i], 3);
            else if (seq[i] == 18) bw.put(seqx[i], 7);
        }
        // pass 2: the symbols
        uint32_t lit[256];  // code | length << 16
        for (int k = 0; k < 256; k++) lit[k] = lc[k] | ((uint32_t)ll[k] << 16);
        if (first) bw.put(lit[blk[0]] & 0xffff, (int)(lit[blk[0]] >> 16));
        for (i = first; i < safe;) {
            if (run_starts(blk, i)) {
                const int r = run_len(blk, i, bn);
                const int sy = lt.sym[r];
                bw.put(lc[sy], ll[sy]);
                bw.put(lt.xval[r], lt.xbits[r] + 1);  // extra bits, then the 1-bit distance code 0 (= distance 1)
                i += (size_t)r;
                continue;
            }
            const uint32_t e = lit[blk[i]];
            bw.put(e & 0xffff, (int)(e >> 16));
            i++;
        }
        for (; i < bn; i++) bw.put(lit[blk[i]] & 0xffff, (int)(lit[blk[i]] >> 16));
        bw.put(lc[256], ll[256]);
        // every block of this encoder starts on a byte boundary: an empty stored block pads (what zlib's sync flush emits)
        bw.put(0, 3);
        bw.align();
        out = bw.p;
        *out++ = 0; *out++ = 0; *out++ = 0xff; *out++ = 0xff;
    }
    return out;
}

inline size_t bound(size_t n) { return n + n / 8 + 64 * (n / (128 * 1024) + 2) + 1024; }

// One scanline through the Sub, Up and Paeth filters (branch-free loops: the compiler turns them into byte-wide SIMD; an AVX2 clone
// is picked at load time where the CPU has it); the candidate with the smallest sum of |residual| lands in out.  Returns its PNG
// filter type.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default"), optimize("O3")))
#endif
inline int filter_row(const uint8_t* cur, const uint8_t* prev, size_t stride, size_t c, uint8_t* out, uint8_t* t1, uint8_t* t2)
{
    for (size_t i = 0; i < c; i++) {  // the first pixel has no left neighbour
        out[i] = cur[i];
        t1[i] = t2[i] = (uint8_t)(cur[i] - prev[i]);
    }
    uint32_t s0 = 0, s1 = 0, s2 = 0;
    for (size_t i = c; i < stride; i++) {
        const uint8_t v = (uint8_t)(cur[i] - cur[i - c]);
        out[i] = v;
        s0 += (uint32_t)abs((int)(int8_t)v);
    }
    for (size_t i = c; i < stride; i++) {
        const uint8_t v = (uint8_t)(cur[i] - prev[i]);
        t1[i] = v;
        s1 += (uint32_t)abs((int)(int8_t)v);
    }
    for (size_t i = c; i < stride; i++) {
        const int a = cur[i - c], b = prev[i], q = prev[i - c];
        const int pa = abs(b - q), pb = abs(a - q), pc = abs(a + b - 2 * q);
        const int pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : q);
        const uint8_t v = (uint8_t)(cur[i] - pred);
        t2[i] = v;
        s2 += (uint32_t)abs((int)(int8_t)v);
    }
    if (s0 <= s1 && s0 <= s2) return 1;
    if (s1 <= s2) { memcpy(out, t1, stride); return 2; }
    memcpy(out, t2, stride);
    return 4;
}

}  // namespace fastz

// Writes 8-bit RGB (c = 3) or RGBA (c = 4), non-interlaced.
//   level <= 1 (default): fast mode -- per row the best of Sub / Up / Paeth, then the run-length + Huffman deflate above: the save
//             stage must keep up with ~10 results of 33 megapixels per second and GPU (BASELINE configs[2]);
//   level >= 2: per-row filter selection (minimum sum of absolute residuals) and deflate at that level.
//   threads  : the rows are cut into that many bands, filtered and deflated concurrently as independent raw deflate
//             streams (each ended with a sync flush) and concatenated into ONE zlib stream (Adler-32 combined), which
//             any PNG reader accepts.
inline bool png_encode_file(const std::string& path, const uint8_t* pixels, int w, int h, int c, int level = 1, int threads = 1)
{
    if ((c != 3 && c != 4) || w <= 0 || h <= 0) return false;
    const size_t stride = (size_t)w * c;
    int nb = std::max(1, std::min(threads, h / 16));
    std::vector<std::vector<uint8_t>> zband((size_t)nb);
    std::vector<uLong> adler((size_t)nb, 1), rawlen((size_t)nb, 0);
    std::vector<int> okb((size_t)nb, 0);
    auto work = [&](int bi) {
        const int y0 = (int)((long long)h * bi / nb), y1 = (int)((long long)h * (bi + 1) / nb);
        std::vector<uint8_t> raw((stride + 1) * (size_t)(y1 - y0));
        std::vector<uint8_t> zero(stride, 0), tmp1(stride), tmp2(stride);
        for (int y = y0; y < y1; y++) {
            const uint8_t* cur = pixels + (size_t)y * stride;
            const uint8_t* prev = y ? cur - stride : zero.data();
            uint8_t* out = raw.data() + (size_t)(y - y0) * (stride + 1);
            if (level <= 1) {  // fast mode: Sub / Up / Paeth, whichever leaves the smallest residuals
                out[0] = (uint8_t)fastz::filter_row(cur, prev, stride, (size_t)c, out + 1, tmp1.data(), tmp2.data());
                continue;
            }
            int best = 4;
            if (level >= 2) {  // choose among None / Sub / Up / Paeth by the minimum sum of absolute (signed) residuals
                unsigned long long best_sum = ~0ULL;
                for (int ft : {0, 1, 2, 4}) {
                    unsigned long long sum = 0;
                    for (size_t i = 0; i < stride; i++) {
                        const int a = i >= (size_t)c ? cur[i - c] : 0, b = prev[i], cc = i >= (size_t)c ? prev[i - c] : 0;
                        const int pred = ft == 0 ? 0 : ft == 1 ? a : ft == 2 ? b : paeth(a, b, cc);
                        const uint8_t v = (uint8_t)(cur[i] - pred);
                        sum += v < 128 ? v : 256 - v;
                        if (sum >= best_sum) break;
                    }
                    if (sum < best_sum) { best_sum = sum; best = ft; }
                }
            }
            out[0] = (uint8_t)best;
            for (size_t i = 0; i < stride; i++) {
                const int a = i >= (size_t)c ? cur[i - c] : 0, b = prev[i], cc = i >= (size_t)c ? prev[i - c] : 0;
                const int pred = best == 0 ? 0 : best == 1 ? a : best == 2 ? b : paeth(a, b, cc);
                out[1 + i] = (uint8_t)(cur[i] - pred);
            }
        }
        adler[(size_t)bi] = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
        rawlen[(size_t)bi] = (uLong)raw.size();
        if (level <= 1) {  // own run-length + Huffman deflate (fastz above); the last band closes the stream with an empty final block
            std::vector<uint8_t>& z = zband[(size_t)bi];
            z.resize(fastz::bound(raw.size()));
            uint8_t* e = fastz::deflate_blocks(raw.data(), raw.size(), z.data());
            if (bi == nb - 1) { *e++ = 1; *e++ = 0; *e++ = 0; *e++ = 0xff; *e++ = 0xff; }
            z.resize((size_t)(e - z.data()));
            okb[(size_t)bi] = 1;
            return;
        }
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, level <= 1 ? 1 : std::min(level, 9), Z_DEFLATED, -15, 9, level <= 1 ? Z_RLE : Z_DEFAULT_STRATEGY) != Z_OK) return;
        std::vector<uint8_t>& z = zband[(size_t)bi];
        z.resize(deflateBound(&zs, (uLong)raw.size()) + 16);
        zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
        zs.next_out = z.data(); zs.avail_out = (uInt)z.size();
        const int flush = bi == nb - 1 ? Z_FINISH : Z_SYNC_FLUSH;
        const int zr = deflate(&zs, flush);
        okb[(size_t)bi] = (flush == Z_FINISH ? zr == Z_STREAM_END : zr == Z_OK) && zs.avail_in == 0;
        z.resize(zs.total_out);
        deflateEnd(&zs);
    };
    if ((stride + 1) * (size_t)((h + nb - 1) / nb) > 0x7fffffffULL) return false;  // one band must fit zlib's 32-bit counters
    if (nb == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int bi = 1; bi < nb; bi++) th.emplace_back(work, bi);
        work(0);
        for (auto& t : th) t.join();
    }
    uLong ad = adler[0];
    for (int bi = 0; bi < nb; bi++) {
        if (!okb[(size_t)bi]) return false;
        if (bi) ad = adler32_combine(ad, adler[(size_t)bi], (z_off_t)rawlen[(size_t)bi]);
    }
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) return false;
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        put_be32(hdr, len);
        memcpy(hdr + 4, type, 4);
        uLong crc = crc32(crc32(0L, Z_NULL, 0), hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t tail[4];
        put_be32(tail, (uint32_t)crc);
        return fwrite(hdr, 1, 8, fp) == 8 && (!len || fwrite(data, 1, len, fp) == len) && fwrite(tail, 1, 4, fp) == 4;
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    uint8_t ihdr[13];
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr + 4, (uint32_t)h);
    ihdr[8] = 8; ihdr[9] = c == 3 ? 2 : 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    bool ok = fwrite(sig, 1, 8, fp) == 8 && chunk("IHDR", ihdr, 13);
    // the zlib stream = 2-byte header + the bands' raw deflate data + Adler-32, cut into IDAT chunks of at most 1 MiB
    static const uint8_t zhdr[2] = {0x78, 0x01};
    ok = ok && chunk("IDAT", zhdr, 2);
    for (int bi = 0; ok && bi < nb; bi++) {
        const std::vector<uint8_t>& z = zband[(size_t)bi];
        for (size_t p = 0; ok && p < z.size(); p += (1u << 20)) ok = chunk("IDAT", z.data() + p, (uint32_t)std::min<size_t>(1u << 20, z.size() - p));
    }
    uint8_t adl[4];
    put_be32(adl, (uint32_t)ad);
    ok = ok && chunk("IDAT", adl, 4) && chunk("IEND", nullptr, 0);
    ok = (fclose(fp) == 0) && ok;
    return ok;
}

}  // namespace rsrhost
