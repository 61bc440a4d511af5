// pre-included when csrc/engine.cu and csrc/kernels.h are compiled as plain C++ for the host-sanitizer build: the few device intrinsics
// the inline helpers of kernels.h name (those helpers are never called on the host)
#pragma once
#include <string.h>
#include <algorithm>
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) { (void)b; (void)s; return a; }
using std::max;
using std::min;
