// model.cpp -- .param/.bin reader + RRDB topology check + weight packing.  See model.h for the format references.
#include "model.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <sstream>

namespace rsr {

namespace {

struct RawLayer {
    std::string type, name;
    std::vector<std::string> bottoms, tops;
    std::map<int, std::string> kv;  // scalar params (text) and arrays (text after '=')
    int geti(int id, int def) const
    {
        auto it = kv.find(id);
        return it == kv.end() ? def : atoi(it->second.c_str());
    }
    float getf(int id, float def) const
    {
        auto it = kv.find(id);
        return it == kv.end() ? def : strtof(it->second.c_str(), nullptr);
    }
    std::vector<float> getarr(int id) const  // id given as the positive array id (e.g. 10 for -23310)
    {
        std::vector<float> r;
        auto it = kv.find(-23300 - id);
        if (it == kv.end()) return r;
        const char* s = it->second.c_str();
        int n = atoi(s);
        const char* p = strchr(s, ',');
        for (int i = 0; i < n && p; i++) {
            r.push_back(strtof(p + 1, nullptr));
            p = strchr(p + 1, ',');
        }
        return r;
    }
};

int fail(Model& m, int code, const std::string& why)
{
    m.error = why;
    return code;
}

bool near(float a, float b) { return fabsf(a - b) <= 1e-6f; }

}  // namespace

int load_model(const std::string& parampath, const std::string& binpath, Model& m)
{
    m = Model();
    // ---- text graph ----------------------------------------------------------------------------------------
    FILE* fp = fopen(parampath.c_str(), "rb");
    if (!fp) return fail(m, -1, "cannot open " + parampath);
    std::string text;
    {
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, n);
        fclose(fp);
    }
    std::istringstream is(text);
    int magic = 0, nlayers = 0, nblobs = 0;
    is >> magic;
    if (magic != 7767517) return fail(m, -2, "param: bad magic (expected 7767517)");
    is >> nlayers >> nblobs;
    if (!is || nlayers <= 0 || nlayers > 100000) return fail(m, -2, "param: bad layer count");
    std::string line;
    std::getline(is, line);
    std::vector<RawLayer> layers;
    while ((int)layers.size() < nlayers && std::getline(is, line)) {
        std::istringstream ls(line);
        RawLayer L;
        int nb = 0, nt = 0;
        if (!(ls >> L.type >> L.name >> nb >> nt)) continue;
        for (int i = 0; i < nb; i++) { std::string s; ls >> s; L.bottoms.push_back(s); }
        for (int i = 0; i < nt; i++) { std::string s; ls >> s; L.tops.push_back(s); }
        std::string tok;
        while (ls >> tok) {
            size_t eq = tok.find('=');
            if (eq == std::string::npos) continue;
            L.kv[atoi(tok.substr(0, eq).c_str())] = tok.substr(eq + 1);
        }
        layers.push_back(L);
    }
    if ((int)layers.size() != nlayers) return fail(m, -2, "param: truncated layer list");

    // ---- resolve Split aliases, keep the op sequence ------------------------------------------------------------
    std::map<std::string, std::string> alias;
    auto canon = [&](const std::string& s) {
        auto it = alias.find(s);
        return it == alias.end() ? s : it->second;
    };
    std::vector<const RawLayer*> ops;
    for (const RawLayer& L : layers) {
        if (L.type == "Split") {
            if (L.bottoms.size() != 1) return fail(m, -3, "graph: Split with != 1 input");
            for (const std::string& t : L.tops) alias[t] = canon(L.bottoms[0]);
        } else if (L.type == "Input" || L.type == "Convolution" || L.type == "Concat" || L.type == "Eltwise" ||
                   L.type == "BinaryOp" || L.type == "Interp") {
            ops.push_back(&L);
        } else {
            return fail(m, -3, "graph: unsupported layer type " + L.type + " (only the RRDBNet x4 graph is supported)");
        }
    }

    // ---- verify RRDBNet topology while collecting the convs -------------------------------------------------------
    size_t pos = 0;
    auto bad = [&](const char* what) {
        return fail(m, -3, std::string("graph: not an RRDBNet x4 graph (") + what + " at op " + std::to_string(pos) + ")");
    };
    auto next = [&]() -> const RawLayer* { return pos < ops.size() ? ops[pos++] : nullptr; };
    auto take_conv = [&](const RawLayer* L, const std::string& want_in, int cout, int act) -> bool {
        if (!L || L->type != "Convolution" || L->bottoms.size() != 1 || canon(L->bottoms[0]) != want_in) return false;
        if (L->geti(1, 0) != 3 || L->geti(4, 0) != 1 || L->geti(5, 0) != 1) return false;            // 3x3, pad 1, bias
        if (L->geti(2, 1) != 1 || L->geti(3, 1) != 1 || L->geti(11, 3) != 3 || L->geti(7, 1) != 1) return false;  // dil, stride, kh, group
        if (L->geti(0, 0) != cout) return false;
        ConvLayer c;
        c.name = L->name;
        c.cout = cout;
        const int wsize = L->geti(6, 0);
        if (wsize <= 0 || wsize % (cout * 9)) return false;
        c.cin = wsize / (cout * 9);
        c.act = L->geti(9, 0);
        if (c.act != act) return false;
        if (c.act == 2) {
            std::vector<float> p = L->getarr(10);
            if (p.size() != 1) return false;
            c.slope = p[0];
        }
        m.convs.push_back(c);
        return true;
    };
    auto is_scaled_add = [&](const RawLayer* L, const std::string& branch, const std::string& skip) -> bool {
        if (!L || L->type != "Eltwise" || L->bottoms.size() != 2 || L->geti(0, 0) != 1) return false;  // op 1 = SUM
        std::vector<float> co = L->getarr(1);
        if (co.size() != 2 || !near(co[0], 0.2f) || !near(co[1], 1.0f)) return false;
        return canon(L->bottoms[0]) == branch && canon(L->bottoms[1]) == skip;
    };

    const RawLayer* L = next();
    if (!L || L->type != "Input" || L->tops.size() != 1 || L->tops[0] != "data") return bad("Input 'data'");
    L = next();
    if (!take_conv(L, "data", 64, 0) || m.convs.back().cin != 3) return bad("Conv_0 3->64");
    const std::string fea = L->tops[0];
    std::string t = fea;
    for (;;) {
        // an RRDB starts with a 64->32 conv on t; the trunk conv is 64->64 on t
        if (pos >= ops.size()) return bad("unexpected end");
        if (ops[pos]->type == "Convolution" && ops[pos]->geti(0, 0) == 64) break;
        const std::string rrdb_in = t;
        for (int rdb = 0; rdb < 3; rdb++) {
            const std::string x0 = t;
            std::vector<std::string> xs{x0};
            for (int k = 1; k <= 5; k++) {
                std::string in = x0;
                if (k > 1) {
                    L = next();
                    if (!L || L->type != "Concat" || L->geti(0, 0) != 0 || L->bottoms.size() != (size_t)k) return bad("Concat");
                    for (int i = 0; i < k; i++)
                        if (canon(L->bottoms[i]) != xs[i]) return bad("Concat order");
                    in = L->tops[0];
                }
                L = next();
                if (!take_conv(L, in, k < 5 ? 32 : 64, k < 5 ? 2 : 0)) return bad("RDB conv");
                if (m.convs.back().cin != 64 + 32 * (k - 1)) return bad("RDB conv cin");
                if (k < 5 && !near(m.convs.back().slope, 0.2f)) return bad("leaky relu slope");
                xs.push_back(L->tops[0]);
            }
            L = next();
            if (!is_scaled_add(L, xs[5], x0)) return bad("RDB residual 0.2*x5 + x0");
            t = L->tops[0];
        }
        L = next();
        if (!is_scaled_add(L, t, rrdb_in)) return bad("RRDB residual 0.2*t + r");
        t = L->tops[0];
        m.num_rrdb++;
    }
    if (m.num_rrdb < 1) return bad("no RRDB");
    L = next();
    if (!take_conv(L, t, 64, 0) || m.convs.back().cin != 64) return bad("trunk conv");
    const std::string trunk = L->tops[0];
    L = next();
    if (!L || L->type != "BinaryOp" || L->geti(0, 0) != 0 || L->geti(1, 0) != 0 || L->bottoms.size() != 2) return bad("BinaryOp ADD");
    {
        const std::string b0 = canon(L->bottoms[0]), b1 = canon(L->bottoms[1]);
        if (!((b0 == fea && b1 == trunk) || (b0 == trunk && b1 == fea))) return bad("fea + trunk");
    }
    t = L->tops[0];
    for (int u = 0; u < 2; u++) {
        L = next();
        if (!L || L->type != "Interp" || L->geti(0, 0) != 1 || !near(L->getf(1, 1.f), 2.f) || !near(L->getf(2, 1.f), 2.f) ||
            canon(L->bottoms[0]) != t)
            return bad("Interp nearest x2");
        t = L->tops[0];
        L = next();
        if (!take_conv(L, t, 64, 2) || m.convs.back().cin != 64 || !near(m.convs.back().slope, 0.2f)) return bad("upsample conv");
        t = L->tops[0];
    }
    L = next();
    if (!take_conv(L, t, 64, 2) || m.convs.back().cin != 64 || !near(m.convs.back().slope, 0.2f)) return bad("HR conv");
    t = L->tops[0];
    L = next();
    if (!take_conv(L, t, 3, 0) || m.convs.back().cin != 64 || L->tops[0] != "output") return bad("last conv 64->3 'output'");
    if (pos != ops.size()) return bad("trailing layers");

    // ---- weights ------------------------------------------------------------------------------------------------
    fp = fopen(binpath.c_str(), "rb");
    if (!fp) return fail(m, -4, "cannot open " + binpath);
    for (ConvLayer& c : m.convs) {
        const size_t n = (size_t)c.cout * c.cin * 9;
        uint32_t tag = 0;
        if (fread(&tag, 4, 1, fp) != 1) { fclose(fp); return fail(m, -5, "bin: truncated at " + c.name); }
        c.weight.resize(n);
        c.weight_h.resize(n);
        if (tag == 0x01306B47u) {
            const size_t nbytes = (n * 2 + 3) & ~(size_t)3;
            std::vector<uint16_t> tmp(nbytes / 2);
            if (fread(tmp.data(), 1, nbytes, fp) != nbytes) { fclose(fp); return fail(m, -5, "bin: truncated at " + c.name); }
            memcpy(c.weight_h.data(), tmp.data(), n * 2);
            for (size_t i = 0; i < n; i++) {
                // fp16 -> fp32 (exact)
                const uint16_t h = tmp[i];
                uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31, f = h & 1023, u;
                if (e == 0) {
                    if (f == 0) u = s;
                    else {
                        int sh = 0;
                        while (!(f & 1024)) { f <<= 1; sh++; }
                        u = s | ((uint32_t)(113 - sh) << 23) | ((f & 1023) << 13);
                    }
                } else if (e == 31) u = s | 0x7f800000u | (f << 13);
                else u = s | ((e + 112) << 23) | (f << 13);
                memcpy(&c.weight[i], &u, 4);
            }
        } else {
            fclose(fp);
            char b[64];
            snprintf(b, sizeof b, "bin: unsupported weight tag %08x at ", tag);
            return fail(m, -5, b + c.name + " (the shipped models store fp16)");
        }
        c.bias.resize(c.cout);
        if (fread(c.bias.data(), 4, c.cout, fp) != (size_t)c.cout) { fclose(fp); return fail(m, -5, "bin: truncated bias at " + c.name); }
    }
    {
        const long here = ftell(fp);
        fseek(fp, 0, SEEK_END);
        const long end = ftell(fp);
        fclose(fp);
        if (here != end) return fail(m, -5, "bin: trailing bytes (graph / weight file mismatch)");
    }
    return 0;
}

int pack_conv_weight_rows(const ConvLayer& c, int cout_pad, std::vector<uint16_t>& out)
{
    const int nchunk = (c.cin + 63) / 64;
    const int grp = cout_pad;                               // rows per tap group
    // 3 tap groups; Cout<=16 adds a zero group so that N=64; Cout<=32 runs as N=96 MMAs, no padding
    const int rows_per_ky = cout_pad == 16 ? 64 : 3 * grp;
    const int rows = nchunk * 3 * rows_per_ky;
    out.assign((size_t)rows * 64, 0);
    for (int ch = 0; ch < nchunk; ch++)
        for (int ky = 0; ky < 3; ky++)
            for (int kx = 0; kx < 3; kx++)
                for (int n = 0; n < c.cout; n++) {
                    uint16_t* row = out.data() + ((size_t)(ch * 3 + ky) * rows_per_ky + kx * grp + n) * 64;
                    for (int k = 0; k < 64; k++) {
                        const int ci = ch * 64 + k;
                        if (ci >= c.cin) break;
                        row[k] = c.weight_h[((size_t)n * c.cin + ci) * 9 + ky * 3 + kx];
                    }
                }
    return rows;
}

}  // namespace rsr
