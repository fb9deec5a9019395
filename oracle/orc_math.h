// ORACLE -- TEST INFRASTRUCTURE ONLY. Never linked, imported or called by the product path.
//
// CPU restatement of the reference's vector math, geometry and sampling primitives:
//   /root/reference/src/math.jl, src/geometry.jl, src/sampling.jl, src/color.jl
// with Julia's semantics restated by hand: Float32 arithmetic un-fused and
// left-to-right, NaN-propagating min/max, Float64 promotion where a Float64 literal appears,
// 1-based indices kept in the data and shifted only at the array access.
//
// Parity status: UNPINNED against the reference executable (Julia is not installable in this
// environment, the reference ships no tests or golden vectors). The oracle is pinned by
// analytic known-answer tests and by RMSE against the reference's shipped renders
// (tests/test_oracle_*.py); elementary functions come from the shared jt_fmath.h contract.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

#include "../julia-raytracer_b200/csrc/jt_fmath.h"

namespace orc {

static const float pif = JT_PIF;  // src/math.jl:13
static const float FLT_INF = std::numeric_limits<float>::infinity();

// Julia `min`/`max` on floats: NaN-propagating, -0.0 < +0.0 (Q4)
inline float jmin(float a, float b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a < b) return a;
  if (b < a) return b;
  return std::signbit(a) ? a : b;
}
inline float jmax(float a, float b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a > b) return a;
  if (b > a) return b;
  return std::signbit(a) ? b : a;
}
// Julia clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)); NaN passes through
inline float jclamp(float x, float lo, float hi) { return x > hi ? hi : (x < lo ? lo : x); }
inline int64_t jclampi(int64_t x, int64_t lo, int64_t hi) { return x > hi ? hi : (x < lo ? lo : x); }

struct V2 { float x, y; };
struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V4 { float x, y, z, w; };

inline V3 v3(float a, float b, float c) { return V3{a, b, c}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline V4 operator+(V4 a, V4 b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator*(V4 a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V2 operator+(V2 a, V2 b) { return V2{a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, float s) { return V2{a.x * s, a.y * s}; }
inline V3 xyz(V4 a) { return V3{a.x, a.y, a.z}; }

inline bool is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
inline bool all_finite(V3 a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }
// maximum(::SVector{3}) = max(max(a,b),c)
inline float maximum(V3 a) { return jmax(jmax(a.x, a.y), a.z); }
inline float maximum(V4 a) { return jmax(jmax(jmax(a.x, a.y), a.z), a.w); }
inline float minimum(V3 a) { return jmin(jmin(a.x, a.y), a.z); }
inline V3 vmin(V3 a, V3 b) { return V3{jmin(a.x, b.x), jmin(a.y, b.y), jmin(a.z, b.z)}; }
inline V3 vmax(V3 a, V3 b) { return V3{jmax(a.x, b.x), jmax(a.y, b.y), jmax(a.z, b.z)}; }

// src/math.jl:69  dot(a,b) = sum(a .* b) = (a1b1 + a2b2) + a3b3
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// src/math.jl:71-78
inline V3 normalize(V3 a) {
  float l = sqrtf(dot(a, a));
  return l != 0.0f ? a / l : a;
}
// src/math.jl:115-116
inline V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float math_length(V3 a) { return sqrtf(dot(a, a)); }       // src/math.jl:142
inline float distance_squared(V3 a, V3 b) { return dot(a - b, a - b); }  // src/math.jl:144

struct Frame { V3 x, y, z, o; };  // src/math.jl:46
struct Mat3 { V3 x, y, z; };      // src/math.jl:63 (columns)

// src/math.jl:80-81
inline V3 transform_point(const Frame& f, V3 p) { return ((f.x * p.x + f.y * p.y) + f.z * p.z) + f.o; }
// src/math.jl:83
inline V3 transform_vector(const Frame& f, V3 b) { return (f.x * b.x + f.y * b.y) + f.z * b.z; }
// src/math.jl:105 (Mat3f * Vec3f)
inline V3 mul(const Mat3& m, V3 f) { return (m.x * f.x + m.y * f.y) + m.z * f.z; }
inline V3 transform_direction(const Frame& f, V3 b) { return normalize(transform_vector(f, b)); }  // :87
inline V3 transform_direction(const Mat3& m, V3 b) { return normalize(mul(m, b)); }                // :129
// src/math.jl:124-125 -- rigid formula always (Q15)
inline V3 transform_normal(const Frame& f, V3 b) { return normalize(transform_vector(f, b)); }
inline V3 orthonormalize(V3 a, V3 b) { return normalize(a - b * dot(a, b)); }  // src/math.jl:127

inline Mat3 transpose(const Mat3& m) {  // src/math.jl:121-122
  return Mat3{V3{m.x.x, m.y.x, m.z.x}, V3{m.x.y, m.y.y, m.z.y}, V3{m.x.z, m.y.z, m.z.z}};
}
inline float determinant(const Mat3& m) { return dot(m.x, cross(m.y, m.z)); }  // :118
inline Mat3 adjoint(const Mat3& m) {                                           // :112-113
  return transpose(Mat3{cross(m.y, m.z), cross(m.z, m.x), cross(m.x, m.y)});
}
inline Mat3 inverse(const Mat3& m) {  // :107  adjoint(m) * (1 / det)
  Mat3 a = adjoint(m);
  float s = 1.0f / determinant(m);
  return Mat3{a.x * s, a.y * s, a.z * s};
}
// src/math.jl:95-103
inline Frame inverse(const Frame& f, bool non_rigid) {
  Mat3 rot{f.x, f.y, f.z};
  Mat3 minv = non_rigid ? inverse(rot) : transpose(rot);
  V3 t = -mul(minv, f.o);
  return Frame{minv.x, minv.y, minv.z, t};
}

// lerp, src/math.jl:89-93:  a * (1 - u) + b * u
inline V3 lerp(V3 a, V3 b, float u) { return a * (1.0f - u) + b * u; }
inline V4 lerp(V4 a, V4 b, float u) { return a * (1.0f - u) + b * u; }

// src/math.jl:131   -w + 2 * dot(n, w) * n
inline V3 reflect(V3 w, V3 n) { return -w + (2.0f * dot(n, w)) * n; }
// src/math.jl:133-140
inline V3 refract(V3 w, V3 n, float inv_eta) {
  float cosine = dot(n, w);
  float k = 1.0f + (inv_eta * inv_eta) * (cosine * cosine - 1.0f);
  if (k < 0.0f) return V3{0, 0, 0};
  return (-w) * inv_eta + (inv_eta * cosine - sqrtf(k)) * n;
}

// ---- geometry.jl -------------------------------------------------------------------------
struct Bbox { V3 mn, mx; };
inline Bbox empty_bbox() {  // src/geometry.jl:26-29: typemax / typemin of Float32 = +-Inf
  return Bbox{V3{FLT_INF, FLT_INF, FLT_INF}, V3{-FLT_INF, -FLT_INF, -FLT_INF}};
}
static const float ray_eps = 0.0001f;  // src/geometry.jl:34
struct Ray {
  V3 o, d;
  float tmin, tmax;
};
inline Ray make_ray(V3 o, V3 d) { return Ray{o, d, ray_eps, FLT_INF}; }  // src/geometry.jl:43

struct PrimIsec {  // src/geometry.jl:49-56
  V2 uv;
  float distance;
  bool hit;
};
inline PrimIsec no_prim_isec() { return PrimIsec{V2{0, 0}, FLT_INF, false}; }

inline Bbox merge(Bbox a, V3 p) { return Bbox{vmin(a.mn, p), vmax(a.mx, p)}; }           // :88-89
inline Bbox merge(Bbox a, Bbox b) { return Bbox{vmin(a.mn, b.mn), vmax(a.mx, b.mx)}; }   // :91-92
inline V3 center(Bbox b) { return (b.mn + b.mx) / 2.0f; }                                 // :94
inline Bbox triangle_bounds(V3 a, V3 b, V3 c) {  // :64-65  min.(p1,p2,p3) = min(min(p1,p2),p3)
  return Bbox{vmin(vmin(a, b), c), vmax(vmax(a, b), c)};
}
inline Bbox quad_bounds(V3 a, V3 b, V3 c, V3 d) {  // :67-68
  return Bbox{vmin(vmin(vmin(a, b), c), d), vmax(vmax(vmax(a, b), c), d)};
}
inline Bbox transform_bbox(const Frame& f, Bbox b) {  // :70-86
  V3 c[8] = {{b.mn.x, b.mn.y, b.mn.z}, {b.mn.x, b.mn.y, b.mx.z}, {b.mn.x, b.mx.y, b.mn.z},
             {b.mn.x, b.mx.y, b.mx.z}, {b.mx.x, b.mn.y, b.mn.z}, {b.mx.x, b.mn.y, b.mx.z},
             {b.mx.x, b.mx.y, b.mn.z}, {b.mx.x, b.mx.y, b.mx.z}};
  Bbox x = empty_bbox();
  for (int i = 0; i < 8; i++) x = merge(x, transform_point(f, c[i]));
  return x;
}

// src/geometry.jl:96-105 (Q3: Float64 literal promotes t1; Q4: NaN-propagating min/max)
inline bool intersect_bbox(const Ray& ray, V3 dinv, const Bbox& b) {
  V3 it_min = (b.mn - ray.o) * dinv;
  V3 it_max = (b.mx - ray.o) * dinv;
  V3 tmin = vmin(it_min, it_max);
  V3 tmax = vmax(it_min, it_max);
  float t0 = jmax(maximum(tmin), ray.tmin);
  float t1 = jmin(minimum(tmax), ray.tmax);
  double t1d = (double)t1 * 1.00000024;
  return (double)t0 <= t1d;
}

inline Ray transform_ray(const Frame& f, const Ray& r) {  // :107-111
  return Ray{transform_point(f, r.o), transform_vector(f, r.d), r.tmin, r.tmax};
}

// src/geometry.jl:206-236
inline PrimIsec intersect_triangle(const Ray& ray, V3 p1, V3 p2, V3 p3) {
  V3 edge1 = p2 - p1;
  V3 edge2 = p3 - p1;
  V3 pvec = cross(ray.d, edge2);
  float det = dot(edge1, pvec);
  if (det == 0.0f) return no_prim_isec();
  float inv_det = 1.0f / det;
  V3 tvec = ray.o - p1;
  float u = dot(tvec, pvec) * inv_det;
  if (u < 0.0f || u > 1.0f) return no_prim_isec();
  V3 qvec = cross(tvec, edge1);
  float v = dot(ray.d, qvec) * inv_det;
  if (v < 0.0f || u + v > 1.0f) return no_prim_isec();
  float t = dot(edge2, qvec) * inv_det;
  if (t < ray.tmin || t > ray.tmax) return no_prim_isec();
  return PrimIsec{V2{u, v}, t, true};
}
// src/geometry.jl:238-258 (Q11: degeneracy compares positions)
inline PrimIsec intersect_quad(const Ray& ray, V3 p1, V3 p2, V3 p3, V3 p4) {
  if (p3 == p4) return intersect_triangle(ray, p1, p2, p4);
  PrimIsec i1 = intersect_triangle(ray, p1, p2, p4);
  PrimIsec i2 = intersect_triangle(ray, p3, p4, p2);
  if (i2.hit) i2 = PrimIsec{V2{1.0f - i2.uv.x, 1.0f - i2.uv.y}, i2.distance, i2.hit};
  return i1.distance < i2.distance ? i1 : i2;
}

inline V3 triangle_normal(V3 a, V3 b, V3 c) { return normalize(cross(b - a, c - a)); }  // :262
inline float triangle_area(V3 a, V3 b, V3 c) { return math_length(cross(b - a, c - a)) / 2.0f; }  // :264-265
inline V3 quad_normal(V3 a, V3 b, V3 c, V3 d) {  // :267-268
  return normalize(triangle_normal(a, b, d) + triangle_normal(c, d, b));
}
inline float quad_area(V3 a, V3 b, V3 c, V3 d) { return triangle_area(a, b, d) + triangle_area(c, d, b); }  // :270-271

// interpolate_triangle, src/geometry.jl:275-276:  p1*(1-u-v) + p2*u + p3*v, (1-u)-v left to right
inline V3 interp_tri(V3 a, V3 b, V3 c, V2 uv) {
  float w = (1.0f - uv.x) - uv.y;
  return (a * w + b * uv.x) + c * uv.y;
}
inline V2 interp_tri(V2 a, V2 b, V2 c, V2 uv) {
  float w = (1.0f - uv.x) - uv.y;
  return (a * w + b * uv.x) + c * uv.y;
}
inline V4 interp_tri(V4 a, V4 b, V4 c, V2 uv) {
  float w = (1.0f - uv.x) - uv.y;
  return (a * w + b * uv.x) + c * uv.y;
}
// interpolate_quad, src/geometry.jl:278-283
template <class T>
inline T interp_quad(T a, T b, T c, T d, V2 uv) {
  if (uv.x + uv.y <= 1.0f) return interp_tri(a, b, d, uv);
  return interp_tri(c, d, b, V2{1.0f - uv.x, 1.0f - uv.y});
}

// src/geometry.jl:285-316
inline void triangle_tangents_fromuv(V3 p1, V3 p2, V3 p3, V2 uv1, V2 uv2, V2 uv3, V3* tu, V3* tv) {
  V3 p = p2 - p1, q = p3 - p1;
  V2 s{uv2.x - uv1.x, uv3.x - uv1.x};
  V2 t{uv2.y - uv1.y, uv3.y - uv1.y};
  float div = s.x * t.y - s.y * t.x;
  if (div != 0.0f) {
    *tu = V3{t.y * p.x - t.x * q.x, t.y * p.y - t.x * q.y, t.y * p.z - t.x * q.z} / div;
    *tv = V3{s.x * q.x - s.y * p.x, s.x * q.y - s.y * p.y, s.x * q.z - s.y * p.z} / div;
  } else {
    *tu = V3{1, 0, 0};
    *tv = V3{0, 1, 0};
  }
}

// ---- sampling.jl ---------------------------------------------------------------------------
inline V2 sample_disk(V2 ruv) {  // :12-16
  float r = sqrtf(ruv.y);
  float phi = (2.0f * pif) * ruv.x;
  return V2{jt_cosf(phi) * r, jt_sinf(phi) * r};
}
inline float sample_hemisphere_cos_pdf(V3 normal, V3 direction) {  // :24-27
  float cosw = dot(normal, direction);
  return cosw <= 0.0f ? 0.0f : cosw / pif;
}
// :29  clamp(trunc(Int, r * size) + 1, 1, size)
inline int64_t sample_uniform(int64_t size, float r) {
  return jclampi((int64_t)(r * (float)size) + 1, 1, size);
}
// :31  Float32(1 / size) with Int/Int -> Float64 division
inline float sample_uniform_pdf(int64_t size) { return (float)(1.0 / (double)size); }
// :42-56 (1-based result, 0 = none)
inline int64_t upper_bound(const float* cdf, int64_t n, float limit) {
  int64_t idx = 0, l = 1, r = n;
  while (l <= r) {
    int64_t m = (l + r) / 2;
    if (cdf[m - 1] > limit) {
      idx = m;
      r = m - 1;
    } else {
      l = m + 1;
    }
  }
  return idx;
}
// :33-37
inline int64_t sample_discrete(const float* cdf, int64_t n, float r) {
  float last = cdf[n - 1];
  r = jclamp(r * last, 0.0f, last - 0.00001f);
  return jclampi(upper_bound(cdf, n, r), 1, n);
}
// :39-40
inline float sample_discrete_pdf(const float* cdf, int64_t idx) {
  return idx == 1 ? cdf[0] : cdf[idx - 1] - cdf[idx - 2];
}
inline V2 sample_triangle(V2 ruv) {  // :58
  return V2{1.0f - sqrtf(ruv.x), ruv.y * sqrtf(ruv.x)};
}

// ---- color.jl ------------------------------------------------------------------------------
// :18-23; Julia ^(::Float32, ::Float32) for a non-integer exponent evaluates
// Float32(exp2(log2(Float64(x)) * Float64(y)))
inline float srgb_to_rgb(float c) {
  if (c <= 0.04045f) return c / 12.92f;
  float b = (c + 0.055f) / 1.055f;
  return (float)exp2(log2((double)b) * (double)2.4f);
}

}  // namespace orc
