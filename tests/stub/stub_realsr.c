/* TEST INFRASTRUCTURE ONLY -- never built into or shipped with the product.
 * A stand-in for librealsr_b200.so that implements the C ABI entry points the C++ host code binds (include/realsr_b200.h)
 * with a nearest-neighbour x4 "network", so that the host side (host/realsr.h wrapper, host/realsr_cli.cpp thread
 * pipeline, file planning, PNG codec) can be tested on a machine without a GPU.  tests/test_cpp_cli.py builds it
 * into a temporary directory under the library's name and links the CLI against that copy. */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/realsr_b200.h"

struct rsr_handle {
    int gpuid, loaded, scale, tilesize, prepadding;
    char err[128];
    pthread_mutex_t mu; /* like the real engine: RealSR::process() const forwards the public members through rsr_set on every call,
                           concurrently from all proc threads of a device (host/realsr.h) -- the write is serialised */
};

rsr_handle* rsr_create(int gpuid, int tta_mode, int num_threads)
{
    (void)tta_mode; (void)num_threads;
    rsr_handle* h = (rsr_handle*)calloc(1, sizeof *h);
    h->gpuid = gpuid; h->scale = 4; h->tilesize = 200; h->prepadding = 10;
    pthread_mutex_init(&h->mu, NULL);
    return h;
}
void rsr_destroy(rsr_handle* h) { if (h) { pthread_mutex_destroy(&h->mu); free(h); } }
int rsr_load(rsr_handle* h, const char* parampath, const char* modelpath)
{
    (void)modelpath;
    if (strstr(parampath, "missing")) { strcpy(h->err, "stub: cannot open param file"); return -1; }
    h->loaded = 1;
    return 0;
}
int rsr_set(rsr_handle* h, int scale, int tilesize, int prepadding)
{
    pthread_mutex_lock(&h->mu);
    h->scale = scale; h->tilesize = tilesize; h->prepadding = prepadding;
    pthread_mutex_unlock(&h->mu);
    return scale == 4 ? 0 : -1;
}
int rsr_process(rsr_handle* h, const uint8_t* in, int w, int hh, int c, uint8_t* out)
{
    if (!h->loaded) { strcpy(h->err, "stub: no model loaded"); return -1; }
    for (int y = 0; y < hh * 4; y++)
        for (int x = 0; x < w * 4; x++) memcpy(out + ((size_t)y * w * 4 + x) * c, in + ((size_t)(y / 4) * w + x / 4) * c, (size_t)c);
    return 0;
}
const char* rsr_last_error(const rsr_handle* h) { return h->err; }
