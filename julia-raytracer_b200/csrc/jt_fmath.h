// jt_fmath.h -- the shared numerical contract between the CUDA path and the CPU oracle.
//
// The reference calls Julia Base's Float32 sin/cos/atan/acos/exp/log (pure-Julia libm, <=1 ulp,
// not bit-identical to glibc or to CUDA's libdevice: "parity unpinned"). A 1-ulp
// difference in one of them can flip a discrete decision inside the render loop (which BSDF
// lobe, Russian roulette, which CDF element), so the GPU kernels and the oracle both evaluate
// the SAME elementary functions, written here from IEEE-754 basic operations plus explicit
// fmaf only. They give bit-identical results on the host (g++ -ffp-contract=off) and on the
// device (nvcc -fmad=false; fmaf() is always a single FFMA). Accuracy (checked in
// tests/test_fmath.py against float64 libm): <= 3 ulp on the ranges the render loop uses.
//
// Nothing in this file restates reference code; it has no counterpart in /root/reference.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define JT_HD __host__ __device__ __forceinline__
#else
#define JT_HD inline
#endif

JT_HD uint32_t jt_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
JT_HD float jt_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

#define JT_PIF 3.14159274101257324f /* Float32(pi), src/math.jl:13 */

// ---- sin / cos ---------------------------------------------------------------------------
// Cody-Waite reduction by pi/2 (three float terms, products kept exact by fmaf), then
// minimax polynomials on [-pi/4, pi/4]. Valid for |x| < ~1e5 (the loop uses |x| <= 4*pi).
JT_HD float jt_sin_poly(float y) {
  float z = y * y;
  float p = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
  p = fmaf(p, z, -1.6666654611e-1f);
  return fmaf(p * z, y, y);
}
JT_HD float jt_cos_poly(float y) {
  float z = y * y;
  float p = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  p = fmaf(p, z, 4.166664568298827e-2f);
  return fmaf(p, z * z, fmaf(-0.5f, z, 1.0f));
}
JT_HD float jt_reduce_pio2(float x, int* q) {
  float fq = rintf(x * 0.636619746685028076f);
  *q = (int)fq;
  float y = fmaf(fq, -1.57079637050628662f, x);
  y = fmaf(fq, 4.371138828673793e-8f, y);
  y = fmaf(fq, 1.7763568394002505e-15f, y);
  return y;
}
JT_HD float jt_sinf(float x) {
  int q;
  float y = jt_reduce_pio2(x, &q);
  float r = (q & 1) ? jt_cos_poly(y) : jt_sin_poly(y);
  return (q & 2) ? -r : r;
}
JT_HD float jt_cosf(float x) {
  int q;
  float y = jt_reduce_pio2(x, &q);
  float r = (q & 1) ? jt_sin_poly(y) : jt_cos_poly(y);
  return ((q + 1) & 2) ? -r : r;
}

// ---- atan / atan2 ------------------------------------------------------------------------
JT_HD float jt_atanf(float xx) {
  float x = fabsf(xx);
  float y;
  if (x > 2.414213562373095f) {  // tan(3pi/8)
    y = 1.57079637050628662f;
    x = -1.0f / x;
  } else if (x > 0.4142135623730950f) {  // tan(pi/8)
    y = 0.785398185253143311f;
    x = (x - 1.0f) / (x + 1.0f);
  } else {
    y = 0.0f;
  }
  float z = x * x;
  float p = fmaf(8.05374449538e-2f, z, -1.38776856032e-1f);
  p = fmaf(p, z, 1.99777106478e-1f);
  p = fmaf(p, z, -3.33329491539e-1f);
  y = y + fmaf(p * z, x, x);
  return (jt_f2u(xx) >> 31) ? -y : y;
}
// atan(y, x) with Julia/C argument order; result in [-pi, pi]
JT_HD float jt_atan2f(float y, float x) {
  if (x != x || y != y) return x + y;
  if (x == 0.0f) {
    if (y == 0.0f) {
      // atan(+-0, +0) = +-0 ; atan(+-0, -0) = +-pi
      float r = (jt_f2u(x) >> 31) ? JT_PIF : 0.0f;
      return (jt_f2u(y) >> 31) ? -r : r;
    }
    return y > 0.0f ? 1.57079637050628662f : -1.57079637050628662f;
  }
  float a = jt_atanf(y / x);
  if (x < 0.0f) {
    a = (jt_f2u(y) >> 31) ? a - JT_PIF : a + JT_PIF;
  }
  return a;
}

// ---- acos (argument already clamped to [-1,1] by every caller) -----------------------------
JT_HD float jt_asin_core(float a) {  // |a| <= 0.5
  float z = a * a;
  float p = fmaf(4.2163199048e-2f, z, 2.4181311049e-2f);
  p = fmaf(p, z, 4.5470025998e-2f);
  p = fmaf(p, z, 7.4953002686e-2f);
  p = fmaf(p, z, 1.6666752422e-1f);
  return fmaf(p * z, a, a);
}
JT_HD float jt_acosf(float x) {
  if (x > 0.5f) {
    return 2.0f * jt_asin_core(sqrtf(0.5f * (1.0f - x)));
  }
  if (x < -0.5f) {
    return JT_PIF - 2.0f * jt_asin_core(sqrtf(0.5f * (1.0f + x)));
  }
  return 1.57079637050628662f - jt_asin_core(x);
}

// ---- exp -----------------------------------------------------------------------------------
JT_HD float jt_pow2i(int n) {  // 2^n for n in [-126, 127]
  return jt_u2f((uint32_t)(n + 127) << 23);
}
JT_HD float jt_expf(float x) {
  if (x != x) return x;
  if (x > 88.7228317f) return INFINITY;
  if (x < -103.972084f) return 0.0f;
  float fn = floorf(fmaf(1.44269504088896341f, x, 0.5f));
  float r = fmaf(fn, -0.693359375f, x);
  r = fmaf(fn, 2.12194440e-4f, r);
  float z = r * r;
  float p = fmaf(1.9875691500e-4f, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float v = fmaf(p, z, r) + 1.0f;
  int n = (int)fn;
  int n1 = n / 2;
  return (v * jt_pow2i(n1)) * jt_pow2i(n - n1);
}

// ---- log -----------------------------------------------------------------------------------
JT_HD float jt_logf(float x) {
  if (x != x) return x;
  if (x < 0.0f) return NAN;
  if (x == 0.0f) return -INFINITY;
  if (x == INFINITY) return x;
  int e = 0;
  uint32_t u = jt_f2u(x);
  if ((u >> 23) == 0) {  // subnormal: scale up by 2^24
    x = x * 16777216.0f;
    u = jt_f2u(x);
    e = -24;
  }
  e += (int)(u >> 23) - 126;
  float m = jt_u2f((u & 0x007fffffu) | 0x3f000000u);  // [0.5, 1)
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = (m + m) - 1.0f;
  } else {
    m = m - 1.0f;
  }
  float z = m * m;
  float p = fmaf(7.0376836292e-2f, m, -1.1514610310e-1f);
  p = fmaf(p, m, 1.1676998740e-1f);
  p = fmaf(p, m, -1.2420140846e-1f);
  p = fmaf(p, m, 1.4249322787e-1f);
  p = fmaf(p, m, -1.6668057665e-1f);
  p = fmaf(p, m, 2.0000714765e-1f);
  p = fmaf(p, m, -2.4999993993e-1f);
  p = fmaf(p, m, 3.3333331174e-1f);
  float fe = (float)e;
  float y = (p * m) * z;
  y = fmaf(-2.12194440e-4f, fe, y);
  y = fmaf(-0.5f, z, y);
  float r = m + y;
  return fmaf(0.693359375f, fe, r);
}
