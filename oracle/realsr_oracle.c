/*
 * oracle/realsr_oracle.c -- CPU restatement of the reference's RealSR 4x tile path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (realsr-ncnn-vulkan_b200/) may include, link or call this
 * file.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker.  Parity status: PINNED -- this restatement is checked (tests/test_oracle.py) against the
 * reference itself (oracle/_ref = RealSR::process_cpu + vendored ncnn b4ba207, built by oracle/build_ref.py)
 * and against the committed golden vectors in tests/golden/ that oracle/_ref produced.
 *
 * What is restated (all paths relative to /root/reference):
 *   model files   src/ncnn/src/net.cpp:1374-1589 (text .param), paramdict.cpp:246-340 (k=v, -233xx arrays),
 *                 modelbin.cpp:88-296 (tag 0x01306B47 => fp16 payload, 4-byte aligned; type 1 => raw fp32),
 *                 layer/convolution.cpp:29-99 (param ids 0,1,4,5,6,9,10; weights then bias)
 *   operators     layer/convolution.cpp:133-204 (direct 3x3, zero pad, bias, fused activation),
 *                 layer/fused_activation.h:31-36 (leaky relu: v > 0 ? v : v*slope),
 *                 layer/split.cpp:32-41, layer/concat.cpp:140-175 (axis 0 = channels),
 *                 layer/eltwise.cpp:116-131 (SUM with coeffs: a*c0 + b*c1),
 *                 layer/binaryop.cpp:27-40 (op 0 = ADD), layer/interp.cpp:598-618 (nearest),
 *                 layer/padding.cpp:196-260 (type 2 = reflect-101)
 *   tile engine   src/realsr.cpp:525-838 (RealSR::process_cpu): tile grid :537-553, from_pixels_roi :552-574,
 *                 x(1/255.f) :746, border reflect padding :757-766, crop + v*255.f+0.5f :793-807,
 *                 to_pixels = (int) truncation + clamp (mat_pixel.cpp:158), TTA scatter :618-664 and
 *                 gather/average :691-724.
 * The executor is a plain in-order interpreter over the 7 layer types the shipped graphs use
 * (Input, Convolution, Split, Concat, Eltwise, BinaryOp, Interp); arithmetic is fp32 throughout, like the
 * reference's -g -1 path.  RGBA: the 4th channel is not run through the network but up-scaled bicubically (A = -0.75,
 * ncnn layer/interp.cpp:178-249,:645-668 via realsr.cpp:129-140) and truncated by to_pixels without the +0.5 the colour
 * channels get (realsr.cpp:809-812,:816-833).  The reference's CPU branch is only self-consistent for single-tile RGBA
 * images (for several tiles it copies a halo-sized alpha plane into a tile-sized one, realsr.cpp:752,:811); per tile this
 * restatement follows the GPU shaders instead (realsr_preproc.comp:79-88: alpha of the un-padded tile region only), which
 * coincides with the CPU branch in the single-tile case -- that case is what the golden vectors pin.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { L_INPUT, L_CONV, L_SPLIT, L_CONCAT, L_ELTWISE, L_BINARYOP, L_INTERP };

typedef struct {
    int type;
    int nb, nt;
    int *bottoms, *tops;
    /* convolution */
    int cout, ksize, pad, has_bias, wsize, act_type, cin;
    float act_slope;
    float *weight; /* OIHW fp32 (decoded from fp16) */
    float *bias;
    /* eltwise */
    int op_type, ncoeff;
    float coeff[4];
    /* interp */
    int resize_type;
    float hscale, wscale;
} layer_t;

typedef struct {
    float *data;
    int c, h, w;
    int *rc; /* shared refcount for Split aliases */
} blob_t;

typedef struct rsr_oracle {
    int nlayers, nblobs;
    layer_t *layers;
    char **blob_names;
    int nnames;
    int in_blob, out_blob;
    int *last_use; /* layer index of the last consumer of each blob */
    long nconv;
    char err[256];
} rsr_oracle;

static float half_to_float(uint16_t h)
{
    uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 31, m = h & 1023, u;
    if (e == 0) {
        if (m == 0) u = s;
        else { /* subnormal */
            int sh = 0;
            while (!(m & 1024)) { m <<= 1; sh++; }
            m &= 1023;
            u = s | ((uint32_t)(113 - sh) << 23) | (m << 13);
        }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static int blob_index(rsr_oracle *o, const char *name)
{
    for (int i = 0; i < o->nnames; i++)
        if (!strcmp(o->blob_names[i], name)) return i;
    o->blob_names[o->nnames] = strdup(name);
    return o->nnames++;
}

/* ---- model loading -------------------------------------------------------------------------------------- */

static int parse_param(rsr_oracle *o, const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) { snprintf(o->err, sizeof o->err, "cannot open %s", path); return -1; }
    int magic = 0;
    if (fscanf(fp, "%d", &magic) != 1 || magic != 7767517) { fclose(fp); snprintf(o->err, sizeof o->err, "bad magic"); return -1; }
    if (fscanf(fp, "%d %d", &o->nlayers, &o->nblobs) != 2) { fclose(fp); return -1; }
    o->layers = (layer_t *)calloc(o->nlayers, sizeof(layer_t));
    o->blob_names = (char **)calloc(o->nblobs + 8, sizeof(char *));
    o->in_blob = o->out_blob = -1;
    for (int li = 0; li < o->nlayers; li++) {
        layer_t *L = &o->layers[li];
        char type[64], name[128], tok[256];
        if (fscanf(fp, "%63s %127s %d %d", type, name, &L->nb, &L->nt) != 4) { fclose(fp); snprintf(o->err, sizeof o->err, "truncated param at layer %d", li); return -1; }
        L->bottoms = (int *)calloc(L->nb + 1, sizeof(int));
        L->tops = (int *)calloc(L->nt + 1, sizeof(int));
        for (int i = 0; i < L->nb; i++) { if (fscanf(fp, "%255s", tok) != 1) return -1; L->bottoms[i] = blob_index(o, tok); }
        for (int i = 0; i < L->nt; i++) { if (fscanf(fp, "%255s", tok) != 1) return -1; L->tops[i] = blob_index(o, tok); }
        if (!strcmp(type, "Input")) L->type = L_INPUT;
        else if (!strcmp(type, "Convolution")) L->type = L_CONV;
        else if (!strcmp(type, "Split")) L->type = L_SPLIT;
        else if (!strcmp(type, "Concat")) L->type = L_CONCAT;
        else if (!strcmp(type, "Eltwise")) L->type = L_ELTWISE;
        else if (!strcmp(type, "BinaryOp")) L->type = L_BINARYOP;
        else if (!strcmp(type, "Interp")) L->type = L_INTERP;
        else { fclose(fp); snprintf(o->err, sizeof o->err, "unsupported layer type %s", type); return -1; }
        L->ksize = 1; L->hscale = L->wscale = 1.f;
        /* rest of the line: id=value or -233xx=n,v,v,... (paramdict.cpp:246-340) */
        char line[4096];
        if (!fgets(line, sizeof line, fp)) line[0] = 0;
        for (char *t = strtok(line, " \t\r\n"); t; t = strtok(NULL, " \t\r\n")) {
            int id = 0; char *eq = strchr(t, '=');
            if (!eq) continue;
            id = atoi(t);
            const char *v = eq + 1;
            if (id <= -23300) {
                int aid = -id - 23300, n = atoi(v);
                const char *p = strchr(v, ',');
                float arr[8] = {0};
                for (int k = 0; k < n && k < 8 && p; k++) { arr[k] = strtof(p + 1, NULL); p = strchr(p + 1, ','); }
                if (L->type == L_CONV && aid == 10) L->act_slope = arr[0];
                if (L->type == L_ELTWISE && aid == 1) { L->ncoeff = n; memcpy(L->coeff, arr, sizeof(float) * 4); }
                continue;
            }
            if (L->type == L_CONV) {
                if (id == 0) L->cout = atoi(v);
                else if (id == 1) L->ksize = atoi(v);
                else if (id == 4) L->pad = atoi(v);
                else if (id == 5) L->has_bias = atoi(v);
                else if (id == 6) L->wsize = atoi(v);
                else if (id == 9) L->act_type = atoi(v);
            } else if (L->type == L_ELTWISE || L->type == L_BINARYOP) {
                if (id == 0) L->op_type = atoi(v);
            } else if (L->type == L_INTERP) {
                if (id == 0) L->resize_type = atoi(v);
                else if (id == 1) L->hscale = strtof(v, NULL);
                else if (id == 2) L->wscale = strtof(v, NULL);
            }
        }
        if (L->type == L_INPUT) o->in_blob = L->tops[0];
        if (L->type == L_CONV) { L->cin = L->wsize / (L->cout * L->ksize * L->ksize); o->nconv++; }
    }
    fclose(fp);
    for (int i = 0; i < o->nnames; i++) {
        if (!strcmp(o->blob_names[i], "data")) o->in_blob = i;
        if (!strcmp(o->blob_names[i], "output")) o->out_blob = i;
    }
    if (o->in_blob < 0 || o->out_blob < 0) { snprintf(o->err, sizeof o->err, "no data/output blob"); return -1; }
    o->last_use = (int *)calloc(o->nnames + 1, sizeof(int));
    for (int i = 0; i < o->nnames; i++) o->last_use[i] = -1;
    for (int li = 0; li < o->nlayers; li++)
        for (int i = 0; i < o->layers[li].nb; i++) o->last_use[o->layers[li].bottoms[i]] = li;
    o->last_use[o->out_blob] = o->nlayers + 1;
    return 0;
}

static int load_bin(rsr_oracle *o, const char *path)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) { snprintf(o->err, sizeof o->err, "cannot open %s", path); return -1; }
    for (int li = 0; li < o->nlayers; li++) {
        layer_t *L = &o->layers[li];
        if (L->type != L_CONV) continue;
        uint32_t tag = 0;
        if (fread(&tag, 4, 1, fp) != 1) goto trunc;
        L->weight = (float *)malloc(sizeof(float) * L->wsize);
        if (tag == 0x01306B47u) { /* fp16 payload, padded to 4 bytes (modelbin.cpp:113-144) */
            size_t nbytes = ((size_t)L->wsize * 2 + 3) & ~(size_t)3;
            uint16_t *tmp = (uint16_t *)malloc(nbytes);
            if (fread(tmp, 1, nbytes, fp) != nbytes) { free(tmp); goto trunc; }
            for (int i = 0; i < L->wsize; i++) L->weight[i] = half_to_float(tmp[i]);
            free(tmp);
        } else if (tag == 0) {
            if (fread(L->weight, 4, L->wsize, fp) != (size_t)L->wsize) goto trunc;
        } else { fclose(fp); snprintf(o->err, sizeof o->err, "unsupported weight tag %08x", tag); return -1; }
        L->bias = (float *)calloc(L->cout, sizeof(float));
        if (L->has_bias && fread(L->bias, 4, L->cout, fp) != (size_t)L->cout) goto trunc;
    }
    {
        long pos = ftell(fp);
        fseek(fp, 0, SEEK_END);
        if (ftell(fp) != pos) { fclose(fp); snprintf(o->err, sizeof o->err, "trailing bytes in .bin"); return -1; }
    }
    fclose(fp);
    return 0;
trunc:
    fclose(fp);
    snprintf(o->err, sizeof o->err, "truncated .bin");
    return -1;
}

rsr_oracle *rsr_oracle_load(const char *param, const char *bin)
{
    rsr_oracle *o = (rsr_oracle *)calloc(1, sizeof(rsr_oracle));
    if (parse_param(o, param) || load_bin(o, bin)) {
        fprintf(stderr, "rsr_oracle_load: %s\n", o->err);
        return NULL; /* leaks o on purpose: test infrastructure */
    }
    return o;
}

void rsr_oracle_free(rsr_oracle *o)
{
    if (!o) return;
    for (int li = 0; li < o->nlayers; li++) {
        free(o->layers[li].bottoms); free(o->layers[li].tops);
        free(o->layers[li].weight); free(o->layers[li].bias);
    }
    for (int i = 0; i < o->nnames; i++) free(o->blob_names[i]);
    free(o->blob_names); free(o->layers); free(o->last_use); free(o);
}

int rsr_oracle_num_conv(const rsr_oracle *o) { return (int)o->nconv; }

/* Conv #idx (0-based, file order): returns shapes and pointers to decoded fp32 weights (OIHW) and bias. */
int rsr_oracle_conv_info(const rsr_oracle *o, int idx, int *cin, int *cout, int *act, float *slope,
                         const float **weight, const float **bias)
{
    int k = 0;
    for (int li = 0; li < o->nlayers; li++) {
        const layer_t *L = &o->layers[li];
        if (L->type != L_CONV) continue;
        if (k++ == idx) {
            *cin = L->cin; *cout = L->cout; *act = L->act_type; *slope = L->act_slope;
            *weight = L->weight; *bias = L->bias;
            return 0;
        }
    }
    return -1;
}

/* ---- operators ------------------------------------------------------------------------------------------ */

/* layer/convolution.cpp:133-204: y = b + sum_c sum_k w*x (zero padded), then fused activation. The sum runs
 * input-channel-major, tap-minor, per output element, as the naive layer does. */
void rsr_oracle_conv3x3(const float *in, int cin, int h, int w, const float *weight, const float *bias, int cout,
                        int act_type, float slope, float *out)
{
    const size_t plane = (size_t)h * w;
#pragma omp parallel for schedule(dynamic, 1)
    for (int oc = 0; oc < cout; oc++) {
        float *o = out + plane * oc;
        const float b = bias ? bias[oc] : 0.f;
        for (size_t i = 0; i < plane; i++) o[i] = b;
        for (int ic = 0; ic < cin; ic++) {
            const float *x = in + plane * ic;
            const float *k = weight + ((size_t)oc * cin + ic) * 9;
            for (int ky = 0; ky < 3; ky++) {
                for (int kx = 0; kx < 3; kx++) {
                    const float wv = k[ky * 3 + kx];
                    const int dy = ky - 1, dx = kx - 1;
                    const int y0 = dy < 0 ? 1 : 0, y1 = dy > 0 ? h - 1 : h;
                    const int x0 = dx < 0 ? 1 : 0, x1 = dx > 0 ? w - 1 : w;
                    for (int y = y0; y < y1; y++) {
                        float *orow = o + (size_t)y * w;
                        const float *xrow = x + (size_t)(y + dy) * w + dx;
                        for (int xx = x0; xx < x1; xx++) orow[xx] += xrow[xx] * wv;
                    }
                }
            }
        }
        if (act_type == 1) { for (size_t i = 0; i < plane; i++) o[i] = fmaxf(o[i], 0.f); }
        else if (act_type == 2) { for (size_t i = 0; i < plane; i++) o[i] = o[i] > 0.f ? o[i] : o[i] * slope; }
    }
}

static void blob_release(blob_t *b)
{
    if (b->rc && --*b->rc == 0) { free(b->data); free(b->rc); }
    b->data = NULL; b->rc = NULL;
}

static void blob_create(blob_t *b, int c, int h, int w)
{
    b->c = c; b->h = h; b->w = w;
    b->data = (float *)malloc(sizeof(float) * (size_t)c * h * w);
    b->rc = (int *)malloc(sizeof(int));
    *b->rc = 1;
}

/* Runs the graph on one planar fp32 tile (3 x h x w) -> (3 x 4h x 4w). out must hold 3*16*h*w floats. */
int rsr_oracle_net(rsr_oracle *o, const float *in, int h, int w, float *out)
{
    blob_t *blobs = (blob_t *)calloc(o->nnames + 1, sizeof(blob_t));
    int ret = 0;
    for (int li = 0; li < o->nlayers && !ret; li++) {
        layer_t *L = &o->layers[li];
        blob_t *t = &blobs[L->tops[0]];
        switch (L->type) {
        case L_INPUT:
            blob_create(t, 3, h, w);
            memcpy(t->data, in, sizeof(float) * 3 * (size_t)h * w);
            break;
        case L_CONV: {
            blob_t *b = &blobs[L->bottoms[0]];
            if (L->ksize != 3 || L->pad != 1 || b->c != L->cin) { ret = -2; break; }
            blob_create(t, L->cout, b->h, b->w);
            rsr_oracle_conv3x3(b->data, L->cin, b->h, b->w, L->weight, L->bias, L->cout, L->act_type, L->act_slope, t->data);
            break;
        }
        case L_SPLIT: { /* split.cpp:32-41: tops alias the bottom */
            blob_t *b = &blobs[L->bottoms[0]];
            for (int i = 0; i < L->nt; i++) { blobs[L->tops[i]] = *b; ++*b->rc; }
            break;
        }
        case L_CONCAT: { /* concat.cpp:140-175, axis 0 */
            int c = 0;
            for (int i = 0; i < L->nb; i++) c += blobs[L->bottoms[i]].c;
            blob_t *b0 = &blobs[L->bottoms[0]];
            blob_create(t, c, b0->h, b0->w);
            float *p = t->data;
            for (int i = 0; i < L->nb; i++) {
                blob_t *b = &blobs[L->bottoms[i]];
                size_t n = (size_t)b->c * b->h * b->w;
                memcpy(p, b->data, n * sizeof(float));
                p += n;
            }
            break;
        }
        case L_ELTWISE: { /* eltwise.cpp:116-131: SUM with coeffs */
            blob_t *a = &blobs[L->bottoms[0]], *b = &blobs[L->bottoms[1]];
            if (L->op_type != 1 || L->nb != 2) { ret = -3; break; }
            blob_create(t, a->c, a->h, a->w);
            size_t n = (size_t)a->c * a->h * a->w;
            if (L->ncoeff == 0) { for (size_t i = 0; i < n; i++) t->data[i] = a->data[i] + b->data[i]; }
            else {
                const float k0 = L->coeff[0], k1 = L->coeff[1];
                for (size_t i = 0; i < n; i++) t->data[i] = a->data[i] * k0 + b->data[i] * k1;
            }
            break;
        }
        case L_BINARYOP: { /* binaryop.cpp:27-40, op 0 = ADD */
            blob_t *a = &blobs[L->bottoms[0]], *b = &blobs[L->bottoms[1]];
            if (L->op_type != 0) { ret = -4; break; }
            blob_create(t, a->c, a->h, a->w);
            size_t n = (size_t)a->c * a->h * a->w;
            for (size_t i = 0; i < n; i++) t->data[i] = a->data[i] + b->data[i];
            break;
        }
        case L_INTERP: { /* interp.cpp:598-618 nearest */
            blob_t *b = &blobs[L->bottoms[0]];
            if (L->resize_type != 1) { ret = -5; break; }
            int oh = (int)(b->h * L->hscale), ow = (int)(b->w * L->wscale);
            const float hs = 1.f / L->hscale, ws = 1.f / L->wscale;
            blob_create(t, b->c, oh, ow);
#pragma omp parallel for
            for (int q = 0; q < b->c; q++) {
                const float *ptr = b->data + (size_t)q * b->h * b->w;
                float *op = t->data + (size_t)q * oh * ow;
                for (int y = 0; y < oh; y++) {
                    int iy = (int)(y * hs); if (iy > b->h - 1) iy = b->h - 1;
                    for (int x = 0; x < ow; x++) {
                        int ix = (int)(x * ws); if (ix > b->w - 1) ix = b->w - 1;
                        *op++ = ptr[iy * b->w + ix];
                    }
                }
            }
            break;
        }
        }
        /* light mode: drop blobs whose last consumer has run (net.cpp:917-921) */
        for (int i = 0; i < L->nb; i++)
            if (o->last_use[L->bottoms[i]] == li && blobs[L->bottoms[i]].data) blob_release(&blobs[L->bottoms[i]]);
    }
    if (!ret) {
        blob_t *ob = &blobs[o->out_blob];
        if (!ob->data || ob->c != 3 || ob->h != 4 * h || ob->w != 4 * w) ret = -6;
        else memcpy(out, ob->data, sizeof(float) * 3 * 16 * (size_t)h * w);
    }
    for (int i = 0; i < o->nnames; i++) if (blobs[i].data) blob_release(&blobs[i]);
    free(blobs);
    return ret;
}

/* ---- bicubic x4 of the alpha plane (layer/interp.cpp:178-249 coefficients, :251-330 two-pass resize) ------------ */

static void interpolate_cubic(float fx, float *coeffs)
{
    const float A = -0.75f;
    float fx0 = fx + 1, fx1 = fx, fx2 = 1 - fx;
    coeffs[0] = A * fx0 * fx0 * fx0 - 5 * A * fx0 * fx0 + 8 * A * fx0 - 4 * A;
    coeffs[1] = (A + 2) * fx1 * fx1 * fx1 - (A + 3) * fx1 * fx1 + 1;
    coeffs[2] = (A + 2) * fx2 * fx2 * fx2 - (A + 3) * fx2 * fx2 + 1;
    coeffs[3] = 1.f - coeffs[0] - coeffs[1] - coeffs[2];
}

void rsr_oracle_cubic_coeffs(int w, int outw, int *xofs, float *alpha)
{
    double scale = (double)w / outw;
    for (int dx = 0; dx < outw; dx++) {
        float fx = (float)((dx + 0.5) * scale - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        float *a = alpha + dx * 4;
        interpolate_cubic(fx, a);
        if (sx <= -1) { sx = 1; a[0] = 1.f - a[3]; a[1] = a[3]; a[2] = 0.f; a[3] = 0.f; }
        if (sx == 0) { sx = 1; a[0] = a[0] + a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = 0.f; }
        if (sx == w - 2) { sx = w - 3; a[3] = a[2] + a[3]; a[2] = a[1]; a[1] = a[0]; a[0] = 0.f; }
        if (sx >= w - 1) { sx = w - 3; a[3] = 1.f - a[0]; a[2] = a[0]; a[1] = 0.f; a[0] = 0.f; }
        xofs[dx] = sx;
    }
}

/* src: w x h floats; dst: 4w x 4h floats */
void rsr_oracle_bicubic4x(const float *src, int w, int h, float *dst)
{
    const int ow = w * 4, oh = h * 4;
    int *xofs = (int *)malloc(sizeof(int) * (ow + oh));
    int *yofs = xofs + ow;
    float *alpha = (float *)malloc(sizeof(float) * 4 * (ow + oh));
    float *beta = alpha + 4 * ow;
    rsr_oracle_cubic_coeffs(w, ow, xofs, alpha);
    rsr_oracle_cubic_coeffs(h, oh, yofs, beta);
    float *rows = (float *)malloc(sizeof(float) * 4 * ow);
    for (int dy = 0; dy < oh; dy++) {
        const int sy = yofs[dy];
        for (int k = 0; k < 4; k++) { /* horizontal pass of source rows sy-1 .. sy+2 */
            const float *S = src + (size_t)(sy - 1 + k) * w;
            float *r = rows + k * ow;
            for (int dx = 0; dx < ow; dx++) {
                const float *Sp = S + xofs[dx];
                const float *a = alpha + dx * 4;
                r[dx] = Sp[-1] * a[0] + Sp[0] * a[1] + Sp[1] * a[2] + Sp[2] * a[3];
            }
        }
        const float *b = beta + dy * 4;
        for (int dx = 0; dx < ow; dx++)
            dst[(size_t)dy * ow + dx] = rows[dx] * b[0] + rows[ow + dx] * b[1] + rows[2 * ow + dx] * b[2] + rows[3 * ow + dx] * b[3];
    }
    free(rows); free(alpha); free(xofs);
}

/* ---- tile engine (src/realsr.cpp:525-838) ---------------------------------------------------------------- */

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* padding.cpp:196-260 type 2 (reflect-101) on one plane */
static void reflect_pad(const float *src, int sh, int sw, float *dst, int top, int bottom, int left, int right)
{
    const int dw = sw + left + right, dh = sh + top + bottom;
    for (int y = 0; y < dh; y++) {
        int sy = y - top;
        if (sy < 0) sy = -sy;
        if (sy >= sh) sy = 2 * (sh - 1) - sy;
        for (int x = 0; x < dw; x++) {
            int sx = x - left;
            if (sx < 0) sx = -sx;
            if (sx >= sw) sx = 2 * (sw - 1) - sx;
            dst[(size_t)y * dw + x] = src[(size_t)sy * sw + sx];
        }
    }
}

int rsr_oracle_process(rsr_oracle *o, const uint8_t *pixeldata, int w, int h, int channels, uint8_t *outdata,
                       int tilesize, int prepadding, int tta_mode, int num_threads)
{
    if (channels != 3 && channels != 4) return -10;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
    const int scale = 4, T = tilesize, P = prepadding;
    const int xtiles = (w + T - 1) / T, ytiles = (h + T - 1) / T;
    for (int yi = 0; yi < ytiles; yi++) {
        const int tile_h_nopad = imin((yi + 1) * T, h) - yi * T;
        const int in_y0 = imax(yi * T - P, 0), in_y1 = imin((yi + 1) * T + P, h);
        for (int xi = 0; xi < xtiles; xi++) {
            const int tile_w_nopad = imin((xi + 1) * T, w) - xi * T;
            const int in_x0 = imax(xi * T - P, 0), in_x1 = imin((xi + 1) * T + P, w);
            const int cw = in_x1 - in_x0, ch = in_y1 - in_y0;
            /* from_pixels_roi + x (1/255.f) */
            float *crop = (float *)malloc(sizeof(float) * 3 * (size_t)cw * ch);
            for (int q = 0; q < 3; q++)
                for (int y = 0; y < ch; y++)
                    for (int x = 0; x < cw; x++)
                        crop[((size_t)q * ch + y) * cw + x] =
                            (float)pixeldata[((size_t)(in_y0 + y) * w + in_x0 + x) * channels + q] * (1 / 255.f);
            /* reflect padding only where the tile touches the image border */
            const int pad_top = imax(P - yi * T, 0);
            const int pad_bottom = imax(imin((yi + 1) * T + P - h, P), 0);
            const int pad_left = imax(P - xi * T, 0);
            const int pad_right = imax(imin((xi + 1) * T + P - w, P), 0);
            const int pw = cw + pad_left + pad_right, ph = ch + pad_top + pad_bottom;
            float *in0 = (float *)malloc(sizeof(float) * 3 * (size_t)pw * ph);
            for (int q = 0; q < 3; q++)
                reflect_pad(crop + (size_t)q * cw * ch, ch, cw, in0 + (size_t)q * pw * ph, pad_top, pad_bottom, pad_left, pad_right);
            free(crop);
            const int ow = pw * scale, oh = ph * scale;
            const int tw = tile_w_nopad * scale, th = tile_h_nopad * scale;
            float *res = (float *)malloc(sizeof(float) * 3 * (size_t)tw * th);
            if (!tta_mode) {
                float *ot = (float *)malloc(sizeof(float) * 3 * (size_t)ow * oh);
                int r = rsr_oracle_net(o, in0, ph, pw, ot);
                if (r) { free(ot); free(in0); free(res); return r; }
                for (int q = 0; q < 3; q++)
                    for (int i = 0; i < th; i++)
                        for (int j = 0; j < tw; j++)
                            res[((size_t)q * th + i) * tw + j] =
                                ot[((size_t)q * oh + i + P * scale) * ow + j + P * scale] * 255.f + 0.5f;
                free(ot);
            } else {
                /* realsr.cpp:618-664: 8 dihedral variants; 4..7 are transposed (w,h swapped) */
                float *tin[8], *tout[8];
                const size_t npl = (size_t)pw * ph;
                tin[0] = in0;
                for (int k = 1; k < 8; k++) tin[k] = (float *)malloc(sizeof(float) * 3 * npl);
                for (int q = 0; q < 3; q++) {
                    const float *p0 = tin[0] + q * npl;
                    for (int i = 0; i < ph; i++)
                        for (int j = 0; j < pw; j++) {
                            const float v = p0[(size_t)i * pw + j];
                            tin[1][q * npl + (size_t)(ph - 1 - i) * pw + j] = v;
                            tin[2][q * npl + (size_t)i * pw + (pw - 1 - j)] = v;
                            tin[3][q * npl + (size_t)(ph - 1 - i) * pw + (pw - 1 - j)] = v;
                            tin[4][q * npl + (size_t)j * ph + i] = v;
                            tin[5][q * npl + (size_t)(pw - 1 - j) * ph + i] = v;
                            tin[6][q * npl + (size_t)j * ph + (ph - 1 - i)] = v;
                            tin[7][q * npl + (size_t)(pw - 1 - j) * ph + (ph - 1 - i)] = v;
                        }
                }
                int r = 0;
                for (int k = 0; k < 8; k++) {
                    tout[k] = (float *)malloc(sizeof(float) * 3 * 16 * npl);
                    if (!r) r = k < 4 ? rsr_oracle_net(o, tin[k], ph, pw, tout[k]) : rsr_oracle_net(o, tin[k], pw, ph, tout[k]);
                }
                /* realsr.cpp:691-724 */
                const size_t opl = (size_t)ow * oh;
                const int ps = P * scale;
                if (!r)
                    for (int q = 0; q < 3; q++)
                        for (int i = 0; i < th; i++)
                            for (int j = 0; j < tw; j++) {
                                const float v0 = tout[0][q * opl + (size_t)(i + ps) * ow + (j + ps)];
                                const float v1 = tout[1][q * opl + (size_t)(oh - 1 - i - ps) * ow + (j + ps)];
                                const float v2 = tout[2][q * opl + (size_t)(i + ps) * ow + (ow - 1 - ps - j)];
                                const float v3 = tout[3][q * opl + (size_t)(oh - 1 - i - ps) * ow + (ow - 1 - ps - j)];
                                const float v4 = tout[4][q * opl + (size_t)(j + ps) * oh + (i + ps)];
                                const float v5 = tout[5][q * opl + (size_t)(ow - 1 - j - ps) * oh + (i + ps)];
                                const float v6 = tout[6][q * opl + (size_t)(j + ps) * oh + (oh - 1 - i - ps)];
                                const float v7 = tout[7][q * opl + (size_t)(ow - 1 - j - ps) * oh + (oh - 1 - i - ps)];
                                const float v = (v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7) / 8;
                                res[((size_t)q * th + i) * tw + j] = v * 255.f + 0.5f;
                            }
                for (int k = 1; k < 8; k++) free(tin[k]);
                for (int k = 0; k < 8; k++) free(tout[k]);
                if (r) { free(in0); free(res); return r; }
            }
            free(in0);
            /* alpha: bicubic x4 of the un-padded tile region, raw 0..255 values, no +0.5 (see header) */
            float *ares = NULL;
            if (channels == 4) {
                float *asrc = (float *)malloc(sizeof(float) * (size_t)tile_w_nopad * tile_h_nopad);
                for (int y = 0; y < tile_h_nopad; y++)
                    for (int x = 0; x < tile_w_nopad; x++)
                        asrc[(size_t)y * tile_w_nopad + x] = (float)pixeldata[((size_t)(yi * T + y) * w + xi * T + x) * 4 + 3];
                ares = (float *)malloc(sizeof(float) * (size_t)tw * th);
                rsr_oracle_bicubic4x(asrc, tile_w_nopad, tile_h_nopad, ares);
                free(asrc);
            }
            /* to_pixels: (int) truncation, clamp to [0,255] (mat_pixel.cpp:158,:140-206) */
            for (int i = 0; i < th; i++) {
                uint8_t *dst = outdata + ((size_t)(yi * scale * T + i) * w * scale + (size_t)xi * scale * T) * channels;
                for (int j = 0; j < tw; j++) {
                    for (int q = 0; q < 3; q++) {
                        int v = (int)res[((size_t)q * th + i) * tw + j];
                        dst[j * channels + q] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                    }
                    if (channels == 4) {
                        int v = (int)ares[(size_t)i * tw + j];
                        dst[j * 4 + 3] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                    }
                }
            }
            free(ares);
            free(res);
        }
    }
    return 0;
}

#ifdef RSR_ORACLE_MAIN
/* usage: realsr_oracle <x4.param> <x4.bin> <in.raw> <w> <h> <c> <out.raw> <tilesize> <tta> <threads> [prepadding] */
#include <time.h>
int main(int argc, char **argv)
{
    if (argc < 11) { fprintf(stderr, "usage: %s param bin in.raw w h c out.raw tilesize tta threads [prepadding]\n", argv[0]); return 2; }
    int w = atoi(argv[4]), h = atoi(argv[5]), c = atoi(argv[6]);
    int tilesize = atoi(argv[8]), tta = atoi(argv[9]), threads = atoi(argv[10]);
    int prepadding = argc > 11 ? atoi(argv[11]) : 10;
    size_t n = (size_t)w * h * c;
    uint8_t *in = (uint8_t *)malloc(n), *out = (uint8_t *)malloc(n * 16);
    FILE *fp = fopen(argv[3], "rb");
    if (!fp || fread(in, 1, n, fp) != n) { fprintf(stderr, "cannot read %s\n", argv[3]); return 3; }
    fclose(fp);
    rsr_oracle *o = rsr_oracle_load(argv[1], argv[2]);
    if (!o) return 4;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int r = rsr_oracle_process(o, in, w, h, c, out, tilesize, prepadding, tta, threads);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (r) { fprintf(stderr, "process failed %d\n", r); return 5; }
    fp = fopen(argv[7], "wb");
    if (!fp || fwrite(out, 1, n * 16, fp) != n * 16) return 6;
    fclose(fp);
    printf("{\"seconds\": %.6f, \"w\": %d, \"h\": %d, \"c\": %d, \"tilesize\": %d, \"tta\": %d, \"threads\": %d, \"out_mp\": %.6f}\n",
           (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), w, h, c, tilesize, tta, threads, (double)w * 4 * h * 4 / 1e6);
    rsr_oracle_free(o);
    return 0;
}
#endif
