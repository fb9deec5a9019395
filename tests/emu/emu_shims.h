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

// ---- single-lane "warps" so that the wavefront kernels (ballot / match_any / shuffle queue appends)
// can be stepped slot by slot on the host: every emulated thread is its own warp of one lane.
struct emu_dim { unsigned x, y, z; };
static thread_local emu_dim emu_threadIdx = {0, 0, 0}, emu_blockIdx = {0, 0, 0}, emu_blockDim = {1, 1, 1};
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __match_any_sync(unsigned, int) { return 1u; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __global__
#define __global__
#endif

#define JT_EMU_COUNT 1
struct jt_emu_counts_t { unsigned long long wide_nodes, wide_prims, wide_instances, wide_xforms; };
static thread_local jt_emu_counts_t jt_emu_counts = {0, 0, 0, 0};

static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {
  unsigned long long v = ((unsigned long long)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
  return r;
}
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) {}
#define JT_FETCH_THRESHOLD 1 /* single-lane warps */
#define JT_SUSPEND_MIN_QUEUE 0
#define JT_SHADE_BLOCK 1 /* one emulated thread per block: the shade kernel maps BLOCKS to material keys */
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
