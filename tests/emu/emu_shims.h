// TEST INFRASTRUCTURE -- host shims so that the device headers of julia-raytracer_b200/csrc can be
// compiled by g++ and single-stepped on the CPU. This lets the no-GPU test tier check the wide-BVH
// builder + traversal and the device shading code against the oracle. It is NOT a CPU fallback:
// nothing in the product links or loads it.
#pragma once
#include <cuda_runtime.h>  // vector types + make_float4 & co (host-side headers)
#include <math.h>
#include <stdint.h>
#include <string.h>

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
