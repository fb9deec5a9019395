// jt_dev_math.cuh -- device vector math for the render kernels.
//
// Numerics contract: this translation unit is compiled with -fmad=false, IEEE division and
// square root (nvcc defaults), no flush-to-zero. Every expression below therefore rounds
// exactly like the reference's un-fused Float32 Julia code (no muladd/@fastmath
// anywhere in the reference); an FMA appears only where fmaf() is written out (box tests of the
// wide BVH, where exactness is not part of the contract, and jt_fmath.h).
//
// Reference semantics covered here: src/math.jl (dot :69, normalize :71, transform_* :80-129,
// inverse :95-110, reflect :131, refract :133), Julia's NaN-propagating min/max and clamp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jt_fmath.h"
#include "jt_internal.h"
#include "jt_rng.h"

#undef JT_DEV
#define JT_DEV __device__ __forceinline__

struct f3 {
  float x, y, z;
};
struct f2 {
  float x, y;
};

JT_DEV f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
JT_DEV f3 ld3(const float* p) { return f3{p[0], p[1], p[2]}; }
JT_DEV f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
JT_DEV f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
JT_DEV f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
JT_DEV f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
JT_DEV f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
JT_DEV f3 operator*(float s, f3 a) { return f3{s * a.x, s * a.y, s * a.z}; }
JT_DEV f3 operator/(f3 a, float s) { return f3{a.x / s, a.y / s, a.z / s}; }
JT_DEV bool operator==(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
JT_DEV bool is_zero3(f3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
JT_DEV bool finite3(f3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }

// Julia min/max: NaN-propagating, and -0.0 < +0.0
JT_DEV float jl_min(float a, float b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a < b) return a;
  if (b < a) return b;
  return (__float_as_uint(a) >> 31) ? a : b;
}
JT_DEV float jl_max(float a, float b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a > b) return a;
  if (b > a) return b;
  return (__float_as_uint(a) >> 31) ? b : a;
}
JT_DEV float jl_clamp(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }
JT_DEV int jl_clampi(int x, int lo, int hi) { return x > hi ? hi : (x < lo ? lo : x); }
JT_DEV float max3(f3 a) { return jl_max(jl_max(a.x, a.y), a.z); }

JT_DEV float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
JT_DEV f3 cross3(f3 a, f3 b) {
  return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
JT_DEV f3 normalize3(f3 a) {
  float l = sqrtf(dot3(a, a));
  return l != 0.0f ? a / l : a;
}
JT_DEV float length3(f3 a) { return sqrtf(dot3(a, a)); }

// frames are 12 floats: x, y, z, o columns
JT_DEV f3 xform_point(const float* f, f3 p) {
  return f3{((f[0] * p.x + f[3] * p.y) + f[6] * p.z) + f[9], ((f[1] * p.x + f[4] * p.y) + f[7] * p.z) + f[10],
            ((f[2] * p.x + f[5] * p.y) + f[8] * p.z) + f[11]};
}
JT_DEV f3 xform_vector(const float* f, f3 v) {
  return f3{(f[0] * v.x + f[3] * v.y) + f[6] * v.z, (f[1] * v.x + f[4] * v.y) + f[7] * v.z,
            (f[2] * v.x + f[5] * v.y) + f[8] * v.z};
}
JT_DEV f3 xform_direction(const float* f, f3 v) { return normalize3(xform_vector(f, v)); }
// rigid inverse applied to a direction: transpose(rotation) * v  (inverse(frame) non_rigid=false)
JT_DEV f3 xform_vector_transposed(const float* f, f3 v) {
  // rows of the transpose are the columns x, y, z: minv[1] = (x.x, y.x, z.x) ...
  // Mat3f * v = m1*v1 + m2*v2 + m3*v3 with m_k the k-th column of the transposed matrix
  return f3{(f[0] * v.x + f[1] * v.y) + f[2] * v.z, (f[3] * v.x + f[4] * v.y) + f[5] * v.z,
            (f[6] * v.x + f[7] * v.y) + f[8] * v.z};
}

JT_DEV f3 reflect3(f3 w, f3 n) { return -w + (2.0f * dot3(n, w)) * n; }
JT_DEV f3 refract3(f3 w, f3 n, float inv_eta) {
  float cosine = dot3(n, w);
  float k = 1.0f + (inv_eta * inv_eta) * (cosine * cosine - 1.0f);
  if (k < 0.0f) return f3{0.0f, 0.0f, 0.0f};
  return (-w) * inv_eta + (inv_eta * cosine - sqrtf(k)) * n;
}
JT_DEV f3 orthonormalize3(f3 a, f3 b) { return normalize3(a - b * dot3(a, b)); }

struct DRay {
  f3 o, d;
  float tmin, tmax;
};
#define JT_RAY_EPS 0.0001f /* src/geometry.jl:34 */

struct DHit {
  float t, u, v;
  int inst, elem;  // 0-based, -1 = miss
};
