// mat.h -- the slice of ncnn::Mat the RealSR boundary uses (reference: ncnn/src/mat.h:62-341).
//
// The reference's caller (src/main.cpp:275-276) builds
//     ncnn::Mat inimage (w, h, (void*)pixeldata, (size_t)channels, channels);   // non-owning view, refcount == 0
//     ncnn::Mat outimage(w * scale, h * scale, (size_t)channels, channels);     // owning, ref-counted
// and passes both by reference to RealSR::process.  This header provides exactly those two constructors with the
// same field names, copy / release semantics (mat.h:1173-1198) and allocation rule (mat.cpp:403-436: payload rounded
// to 4 bytes, int refcount stored behind it, 64-byte aligned), so code written against ncnn::Mat for this path
// compiles unchanged.  When the real ncnn headers are on the include path, include them *instead* of this file:
// realsr.h only needs the members data / w / h / elempack.
#ifndef RSR_NCNN_MAT_SHIM_H
#define RSR_NCNN_MAT_SHIM_H

#include <stddef.h>
#include <stdlib.h>

namespace ncnn {

class Allocator;

class Mat
{
public:
    Mat() : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0) {}
    // owning image: w x h elements of `elemsize` bytes (elempack channels interleaved)
    Mat(int _w, int _h, size_t _elemsize, int _elempack, Allocator* _allocator = 0)
        : data(0), refcount(0), elemsize(0), elempack(0), allocator(0), dims(0), w(0), h(0), d(0), c(0), cstep(0)
    {
        create(_w, _h, _elemsize, _elempack, _allocator);
    }
    // external (non-owning) image
    Mat(int _w, int _h, void* _data, size_t _elemsize, int _elempack, Allocator* _allocator = 0)
        : data(_data), refcount(0), elemsize(_elemsize), elempack(_elempack), allocator(_allocator), dims(2), w(_w), h(_h), d(1), c(1)
    {
        cstep = (size_t)w * h;
    }
    Mat(const Mat& m)
        : data(m.data), refcount(m.refcount), elemsize(m.elemsize), elempack(m.elempack), allocator(m.allocator), dims(m.dims), w(m.w),
          h(m.h), d(m.d), c(m.c), cstep(m.cstep)
    {
        addref();
    }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m)
    {
        if (this == &m) return *this;
        if (m.refcount) __sync_fetch_and_add(m.refcount, 1);
        release();
        data = m.data; refcount = m.refcount; elemsize = m.elemsize; elempack = m.elempack; allocator = m.allocator;
        dims = m.dims; w = m.w; h = m.h; d = m.d; c = m.c; cstep = m.cstep;
        return *this;
    }
    void create(int _w, int _h, size_t _elemsize, int _elempack, Allocator* _allocator = 0)
    {
        release();
        elemsize = _elemsize; elempack = _elempack; allocator = _allocator;
        dims = 2; w = _w; h = _h; d = 1; c = 1;
        cstep = (size_t)w * h;
        const size_t total = cstep * elemsize;
        if (total > 0) {
            const size_t totalsize = (total + 3) & ~(size_t)3;
            void* p = 0;
            if (posix_memalign(&p, 64, totalsize + sizeof(int) + 64)) p = 0;
            data = p;
            if (data) {
                refcount = (int*)((unsigned char*)data + totalsize);
                *refcount = 1;
            }
        }
    }
    void addref()
    {
        if (refcount) __sync_fetch_and_add(refcount, 1);
    }
    void release()
    {
        if (refcount && __sync_fetch_and_add(refcount, -1) == 1) free(data);
        data = 0; refcount = 0; elemsize = 0; elempack = 0; dims = 0; w = h = d = c = 0; cstep = 0;
    }
    bool empty() const { return data == 0 || cstep * c == 0; }
    size_t total() const { return cstep * c; }

    void* data;
    int* refcount;
    size_t elemsize;
    int elempack;
    Allocator* allocator;
    int dims;
    int w, h, d, c;
    size_t cstep;
};

}  // namespace ncnn

#endif  // RSR_NCNN_MAT_SHIM_H
