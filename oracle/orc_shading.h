// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restatement of /root/reference/src/shading.jl (every lobe reachable from trace.jl's dispatch):
// matte :14-37, glossy :39-101, reflective rough :103-151 / delta :202-225, transparent rough
// :323-401 / delta :403-446, refractive rough :448-534 / delta :536-604, passthrough :636-646,
// transmittance :650-669, Henyey-Greenstein :671-693, fresnel_dielectric :695-714,
// sample_hemisphere_cos :716-722, basis_fromz :724-732, GGX :734-816, reflectivity_to_eta :820-823,
// fresnel_conductor :831-851. gltfpbr is not restated: it throws UndefVarError in the reference
// (SURVEY.md §2.3).
#pragma once
#include "orc_math.h"

namespace orc {

inline bool same_hemisphere(V3 normal, V3 outgoing, V3 incoming) {  // :828-829
  return dot(normal, outgoing) * dot(normal, incoming) >= 0.0f;
}

// :695-714
inline float fresnel_dielectric(float eta, V3 normal, V3 outgoing) {
  float cosw = fabsf(dot(normal, outgoing));
  float sin2 = 1.0f - cosw * cosw;
  float eta2 = eta * eta;
  float cos2t = 1.0f - sin2 / eta2;
  if (cos2t < 0.0f) return 1.0f;
  float t0 = sqrtf(cos2t);
  float t1 = eta * t0;
  float t2 = eta * cosw;
  float rs = (cosw - t1) / (cosw + t1);
  float rp = (t0 - t2) / (t0 + t2);
  return (rs * rs + rp * rp) / 2.0f;
}

// :820-823
inline V3 reflectivity_to_eta(V3 r_) {
  V3 r{jclamp(r_.x, 0.0f, 0.99f), jclamp(r_.y, 0.0f, 0.99f), jclamp(r_.z, 0.0f, 0.99f)};
  return V3{(1.0f + sqrtf(r.x)) / (1.0f - sqrtf(r.x)), (1.0f + sqrtf(r.y)) / (1.0f - sqrtf(r.y)),
            (1.0f + sqrtf(r.z)) / (1.0f - sqrtf(r.z))};
}

// :831-851
inline float fresnel_conductor1(float eta, float etak, float cosw, float cos2, float sin2) {
  float eta2 = eta * eta;
  float etak2 = etak * etak;
  float t0 = (eta2 - etak2) - sin2;
  float a2plusb2 = sqrtf(t0 * t0 + (4.0f * eta2) * etak2);
  float t1 = a2plusb2 + cos2;
  float a = sqrtf((a2plusb2 + t0) / 2.0f);
  float t2 = (2.0f * a) * cosw;
  float rs = (t1 - t2) / (t1 + t2);
  float t3 = cos2 * a2plusb2 + sin2 * sin2;
  float t4 = t2 * sin2;
  float rp = (rs * (t3 - t4)) / (t3 + t4);
  return (rp + rs) / 2.0f;
}
inline V3 fresnel_conductor(V3 eta, V3 etak, V3 normal, V3 outgoing) {
  float cosw = dot(normal, outgoing);
  if (cosw <= 0.0f) return V3{0, 0, 0};
  cosw = jclamp(cosw, -1.0f, 1.0f);
  float cos2 = cosw * cosw;
  float sin2 = jclamp(1.0f - cos2, 0.0f, 1.0f);
  return V3{fresnel_conductor1(eta.x, etak.x, cosw, cos2, sin2),
            fresnel_conductor1(eta.y, etak.y, cosw, cos2, sin2),
            fresnel_conductor1(eta.z, etak.z, cosw, cos2, sin2)};
}

// :724-732
inline Mat3 basis_fromz(V3 v) {
  V3 z = normalize(v);
  float sign = copysignf(1.0f, z.z);
  float a = -1.0f / (sign + z.z);
  float b = (z.x * z.y) * a;
  V3 x{1.0f + ((sign * z.x) * z.x) * a, sign * b, (-sign) * z.x};
  V3 y{b, sign + (z.y * z.y) * a, -z.y};
  return Mat3{x, y, z};
}

// :716-722
inline V3 sample_hemisphere_cos(V3 normal, V2 ruv) {
  float z = sqrtf(ruv.y);
  float r = sqrtf(1.0f - z * z);
  float phi = (2.0f * pif) * ruv.x;
  V3 local{r * jt_cosf(phi), r * jt_sinf(phi), z};
  return transform_direction(basis_fromz(normal), local);
}

// :734-750 (ggx = true always)
inline float microfacet_distribution(float roughness, V3 normal, V3 halfway) {
  float cosine = dot(normal, halfway);
  if (cosine <= 0.0f) return 0.0f;
  float roughness2 = roughness * roughness;
  float cosine2 = cosine * cosine;
  float k = (cosine2 * roughness2 + 1.0f) - cosine2;
  return roughness2 / ((pif * k) * k);
}
// :752-773
inline float microfacet_shadowing1(float roughness, V3 normal, V3 halfway, V3 direction) {
  float cosine = dot(normal, direction);
  float cosineh = dot(halfway, direction);
  if (cosine * cosineh <= 0.0f) return 0.0f;
  float roughness2 = roughness * roughness;
  float cosine2 = cosine * cosine;
  return (2.0f * fabsf(cosine)) /
         (fabsf(cosine) + sqrtf((cosine2 - roughness2 * cosine2) + roughness2));
}
// :775-785
inline float microfacet_shadowing(float roughness, V3 normal, V3 halfway, V3 outgoing, V3 incoming) {
  return microfacet_shadowing1(roughness, normal, halfway, outgoing) *
         microfacet_shadowing1(roughness, normal, halfway, incoming);
}
// :787-803
inline V3 sample_microfacet(float roughness, V3 normal, V2 rn) {
  float phi = (2.0f * pif) * rn.x;
  float theta = jt_atanf(roughness * sqrtf(rn.y / (1.0f - rn.y)));
  float st = jt_sinf(theta), ct = jt_cosf(theta);
  V3 local{jt_cosf(phi) * st, jt_sinf(phi) * st, ct};
  return transform_direction(basis_fromz(normal), local);
}
// :805-816
inline float sample_microfacet_pdf(float roughness, V3 normal, V3 halfway) {
  float cosine = dot(normal, halfway);
  if (cosine < 0.0f) return 0.0f;
  return microfacet_distribution(roughness, normal, halfway) * cosine;
}

inline V3 up(V3 normal, V3 outgoing) { return dot(normal, outgoing) <= 0.0f ? -normal : normal; }

// ---- matte :14-37 ----------------------------------------------------------------------------
inline V3 eval_matte(V3 color, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return V3{0, 0, 0};
  return (color / pif) * fabsf(dot(normal, incoming));
}
inline V3 sample_matte(V3, V3 normal, V3 outgoing, V2 rn) {
  return sample_hemisphere_cos(up(normal, outgoing), rn);
}
inline float sample_matte_pdf(V3, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return 0.0f;
  return sample_hemisphere_cos_pdf(up(normal, outgoing), incoming);
}

// ---- glossy :39-101 --------------------------------------------------------------------------
inline V3 eval_glossy(V3 color, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return V3{0, 0, 0};
  V3 up_normal = up(normal, outgoing);
  float F1 = fresnel_dielectric(ior, up_normal, outgoing);
  V3 halfway = normalize(incoming + outgoing);
  float F = fresnel_dielectric(ior, halfway, incoming);
  float D = microfacet_distribution(roughness, up_normal, halfway);
  float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
  float ni = fabsf(dot(up_normal, incoming));
  float den = (4.0f * dot(up_normal, outgoing)) * dot(up_normal, incoming);
  float spec = ((((1.0f * F) * D) * G) / den) * ni;
  V3 diff = ((color * (1.0f - F1)) / pif) * ni;
  return V3{diff.x + spec, diff.y + spec, diff.z + spec};
}
inline V3 sample_glossy(V3, float ior, float roughness, V3 normal, V3 outgoing, float rnl, V2 rn) {
  V3 up_normal = up(normal, outgoing);
  if (rnl < fresnel_dielectric(ior, up_normal, outgoing)) {
    V3 halfway = sample_microfacet(roughness, up_normal, rn);
    V3 incoming = reflect(outgoing, halfway);
    if (!same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
    return incoming;
  }
  return sample_hemisphere_cos(up_normal, rn);
}
inline float sample_glossy_pdf(V3, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return 0.0f;
  V3 up_normal = up(normal, outgoing);
  V3 halfway = normalize(outgoing + incoming);
  float F = fresnel_dielectric(ior, up_normal, outgoing);
  return (F * sample_microfacet_pdf(roughness, up_normal, halfway)) /
             (4.0f * fabsf(dot(outgoing, halfway))) +
         (1.0f - F) * sample_hemisphere_cos_pdf(up_normal, incoming);
}

// ---- reflective, rough :103-151 ----------------------------------------------------------------
inline V3 eval_reflective(V3 color, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return V3{0, 0, 0};
  V3 up_normal = up(normal, outgoing);
  V3 halfway = normalize(incoming + outgoing);
  V3 F = fresnel_conductor(reflectivity_to_eta(color), V3{0, 0, 0}, halfway, incoming);
  float D = microfacet_distribution(roughness, up_normal, halfway);
  float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
  float den = (4.0f * dot(up_normal, outgoing)) * dot(up_normal, incoming);
  return (((F * D) * G) / den) * fabsf(dot(up_normal, incoming));
}
inline V3 sample_reflective(V3, float roughness, V3 normal, V3 outgoing, V2 rn) {
  V3 up_normal = up(normal, outgoing);
  V3 halfway = sample_microfacet(roughness, up_normal, rn);
  V3 incoming = reflect(outgoing, halfway);
  if (!same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
  return incoming;
}
inline float sample_reflective_pdf(V3, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return 0.0f;
  V3 up_normal = up(normal, outgoing);
  V3 halfway = normalize(outgoing + incoming);
  return sample_microfacet_pdf(roughness, up_normal, halfway) / (4.0f * fabsf(dot(outgoing, halfway)));
}
// ---- reflective, delta :202-225 ----------------------------------------------------------------
inline V3 eval_reflective_delta(V3 color, V3 normal, V3 outgoing, V3 incoming) {
  if (dot(normal, incoming) * dot(normal, outgoing) <= 0.0f) return V3{0, 0, 0};
  V3 up_normal = up(normal, outgoing);
  return fresnel_conductor(reflectivity_to_eta(color), V3{0, 0, 0}, up_normal, outgoing);
}
inline V3 sample_reflective_delta(V3, V3 normal, V3 outgoing) {
  return reflect(outgoing, up(normal, outgoing));
}
inline float sample_reflective_delta_pdf(V3, V3 normal, V3 outgoing, V3 incoming) {
  return dot(normal, incoming) * dot(normal, outgoing) <= 0.0f ? 0.0f : 1.0f;
}

// ---- transparent, rough :323-401 ---------------------------------------------------------------
inline V3 eval_transparent(V3 color, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  V3 up_normal = up(normal, outgoing);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f) {
    V3 halfway = normalize(incoming + outgoing);
    float F = fresnel_dielectric(ior, halfway, outgoing);
    float D = microfacet_distribution(roughness, up_normal, halfway);
    float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float den = (4.0f * dot(up_normal, outgoing)) * dot(up_normal, incoming);
    return ((((V3{1, 1, 1} * F) * D) * G) / den) * fabsf(dot(up_normal, incoming));
  } else {
    V3 reflected = reflect(-incoming, up_normal);
    V3 halfway = normalize(reflected + outgoing);
    float F = fresnel_dielectric(ior, halfway, outgoing);
    float D = microfacet_distribution(roughness, up_normal, halfway);
    float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, reflected);
    float den = (4.0f * dot(up_normal, outgoing)) * dot(up_normal, reflected);
    return ((((color * (1.0f - F)) * D) * G) / den) * fabsf(dot(up_normal, reflected));
  }
}
inline V3 sample_transparent(V3, float ior, float roughness, V3 normal, V3 outgoing, float rnl, V2 rn) {
  V3 up_normal = up(normal, outgoing);
  V3 halfway = sample_microfacet(roughness, up_normal, rn);
  if (rnl < fresnel_dielectric(ior, halfway, outgoing)) {
    V3 incoming = reflect(outgoing, halfway);
    if (!same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
    return incoming;
  } else {
    V3 reflected = reflect(outgoing, halfway);
    V3 incoming = -reflect(reflected, up_normal);
    if (same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
    return incoming;
  }
}
inline float sample_transparent_pdf(V3, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  V3 up_normal = up(normal, outgoing);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f) {
    V3 halfway = normalize(incoming + outgoing);
    return (fresnel_dielectric(ior, halfway, outgoing) * sample_microfacet_pdf(roughness, up_normal, halfway)) /
           (4.0f * fabsf(dot(outgoing, halfway)));
  } else {
    V3 reflected = reflect(-incoming, up_normal);
    V3 halfway = normalize(reflected + outgoing);
    float d = (1.0f - fresnel_dielectric(ior, halfway, outgoing)) *
              sample_microfacet_pdf(roughness, up_normal, halfway);
    return d / (4.0f * fabsf(dot(outgoing, halfway)));
  }
}
// ---- transparent, delta :403-446 ---------------------------------------------------------------
inline V3 eval_transparent_delta(V3 color, float ior, V3 normal, V3 outgoing, V3 incoming) {
  V3 up_normal = up(normal, outgoing);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f)
    return V3{1, 1, 1} * fresnel_dielectric(ior, up_normal, outgoing);
  return color * (1.0f - fresnel_dielectric(ior, up_normal, outgoing));
}
inline V3 sample_transparent_delta(V3, float ior, V3 normal, V3 outgoing, float rnl) {
  V3 up_normal = up(normal, outgoing);
  if (rnl < fresnel_dielectric(ior, up_normal, outgoing)) return reflect(outgoing, up_normal);
  return -outgoing;
}
inline float sample_transparent_delta_pdf(V3, float ior, V3 normal, V3 outgoing, V3 incoming) {
  V3 up_normal = up(normal, outgoing);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f)
    return fresnel_dielectric(ior, up_normal, outgoing);
  return 1.0f - fresnel_dielectric(ior, up_normal, outgoing);
}

// ---- refractive, rough :448-534 ----------------------------------------------------------------
inline V3 eval_refractive(V3, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  float rel_ior = entering ? ior : (1.0f / ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f) {
    V3 halfway = normalize(incoming + outgoing);
    float F = fresnel_dielectric(rel_ior, halfway, outgoing);
    float D = microfacet_distribution(roughness, up_normal, halfway);
    float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float den = fabsf((4.0f * dot(normal, outgoing)) * dot(normal, incoming));
    return ((((V3{1, 1, 1} * F) * D) * G) / den) * fabsf(dot(normal, incoming));
  } else {
    V3 halfway = (-normalize(rel_ior * incoming + outgoing)) * (entering ? 1.0f : -1.0f);
    float F = fresnel_dielectric(rel_ior, halfway, outgoing);
    float D = microfacet_distribution(roughness, up_normal, halfway);
    float G = microfacet_shadowing(roughness, up_normal, halfway, outgoing, incoming);
    float a = fabsf((dot(outgoing, halfway) * dot(incoming, halfway)) /
                    (dot(outgoing, normal) * dot(incoming, normal)));
    float s = rel_ior * dot(halfway, incoming) + dot(halfway, outgoing);
    // x^2.0f0 -> Float32(Float64(x)*Float64(x)) == x*x correctly rounded
    return (((((V3{1, 1, 1} * a) * (1.0f - F)) * D) * G) / (s * s)) * fabsf(dot(normal, incoming));
  }
}
inline V3 sample_refractive(V3, float ior, float roughness, V3 normal, V3 outgoing, float rnl, V2 rn) {
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  V3 halfway = sample_microfacet(roughness, up_normal, rn);
  if (rnl < fresnel_dielectric(entering ? ior : (1.0f / ior), halfway, outgoing)) {
    V3 incoming = reflect(outgoing, halfway);
    if (!same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
    return incoming;
  } else {
    V3 incoming = refract(outgoing, halfway, entering ? (1.0f / ior) : ior);
    if (same_hemisphere(up_normal, outgoing, incoming)) return V3{0, 0, 0};
    return incoming;
  }
}
inline float sample_refractive_pdf(V3, float ior, float roughness, V3 normal, V3 outgoing, V3 incoming) {
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  float rel_ior = entering ? ior : (1.0f / ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f) {
    V3 halfway = normalize(incoming + outgoing);
    return (fresnel_dielectric(rel_ior, halfway, outgoing) *
            sample_microfacet_pdf(roughness, up_normal, halfway)) /
           (4.0f * fabsf(dot(outgoing, halfway)));
  } else {
    V3 halfway = (-normalize(rel_ior * incoming + outgoing)) * (entering ? 1.0f : -1.0f);
    float s = rel_ior * dot(halfway, incoming) + dot(halfway, outgoing);
    return (((1.0f - fresnel_dielectric(rel_ior, halfway, outgoing)) *
             sample_microfacet_pdf(roughness, up_normal, halfway)) *
            fabsf(dot(halfway, incoming))) /
           (s * s);
  }
}
// ---- refractive, delta :536-604 ----------------------------------------------------------------
inline V3 eval_refractive_delta(V3, float ior, V3 normal, V3 outgoing, V3 incoming) {
  if ((double)fabsf(ior - 1.0f) < 1e-3) {  // Float64 literal, :543
    return dot(normal, incoming) * dot(normal, outgoing) <= 0.0f ? V3{1, 1, 1} : V3{0, 0, 0};
  }
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  float rel_ior = entering ? ior : (1.0f / ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f)
    return V3{1, 1, 1} * fresnel_dielectric(rel_ior, up_normal, outgoing);
  return (V3{1, 1, 1} * (1.0f / (rel_ior * rel_ior))) * (1.0f - fresnel_dielectric(rel_ior, up_normal, outgoing));
}
inline V3 sample_refractive_delta(V3, float ior, V3 normal, V3 outgoing, float rnl) {
  if ((double)fabsf(ior - 1.0f) < 1e-3) return -outgoing;  // :572
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  float rel_ior = entering ? ior : (1.0f / ior);
  if (rnl < fresnel_dielectric(rel_ior, up_normal, outgoing)) return reflect(outgoing, up_normal);
  return refract(outgoing, up_normal, 1.0f / rel_ior);
}
inline float sample_refractive_delta_pdf(V3, float ior, V3 normal, V3 outgoing, V3 incoming) {
  if (fabsf(ior - 1.0f) < 0.001f) {  // Float32 literal here, :593
    return dot(normal, incoming) * dot(normal, outgoing) < 0.0f ? 1.0f : 0.0f;
  }
  bool entering = dot(normal, outgoing) >= 0.0f;
  V3 up_normal = entering ? normal : -normal;
  float rel_ior = entering ? ior : (1.0f / ior);
  if (dot(normal, incoming) * dot(normal, outgoing) >= 0.0f)
    return fresnel_dielectric(rel_ior, up_normal, outgoing);
  return 1.0f - fresnel_dielectric(rel_ior, up_normal, outgoing);
}

// ---- passthrough :636-646 ----------------------------------------------------------------------
inline V3 eval_passthrough(V3, V3 normal, V3 outgoing, V3 incoming) {
  return dot(normal, incoming) * dot(normal, outgoing) >= 0.0f ? V3{0, 0, 0} : V3{1, 1, 1};
}
inline V3 sample_passthrough(V3, V3, V3 outgoing) { return -outgoing; }
inline float sample_passthrough_pdf(V3, V3 normal, V3 outgoing, V3 incoming) {
  return dot(normal, incoming) * dot(normal, outgoing) >= 0.0f ? 0.0f : 1.0f;
}

// ---- volumes :650-693 --------------------------------------------------------------------------
inline V3 eval_transmittance(V3 density, float distance) {  // exp.(-density * distance)
  V3 a = (-density) * distance;
  return V3{jt_expf(a.x), jt_expf(a.y), jt_expf(a.z)};
}
// Q6: channel = clamp(trunc(Int, rl*3), 1, 3) on a 0-based value -> {1,1,2}
inline float sample_transmittance(V3 density, float max_distance, float rl, float rd) {
  int64_t channel = jclampi((int64_t)(rl * 3.0f), 1, 3);
  float dc = density[(int)channel - 1];
  float distance = dc == 0.0f ? FLT_INF : (-jt_logf(1.0f - rd)) / dc;
  return jmin(distance, max_distance);
}
inline float sample_transmittance_pdf(V3 density, float distance, float max_distance) {
  if (distance < max_distance) {
    V3 e = (-density) * distance;
    V3 t = density * V3{jt_expf(e.x), jt_expf(e.y), jt_expf(e.z)};
    return ((t.x + t.y) + t.z) / 3.0f;
  }
  V3 e = (-density) * max_distance;
  return ((jt_expf(e.x) + jt_expf(e.y)) + jt_expf(e.z)) / 3.0f;
}
inline float eval_phasefunction(float anisotropy, V3 outgoing, V3 incoming) {  // :671-675
  float cosine = -dot(outgoing, incoming);
  float denom = (1.0f + anisotropy * anisotropy) - (2.0f * anisotropy) * cosine;
  return (1.0f - anisotropy * anisotropy) / (((4.0f * pif) * denom) * sqrtf(denom));
}
inline V3 sample_phasefunction(float anisotropy, V3 outgoing, V2 rn) {  // :677-690
  float cos_theta;
  if (fabsf(anisotropy) < 0.001f) {
    cos_theta = 1.0f - 2.0f * rn.y;
  } else {
    float square = (1.0f - anisotropy * anisotropy) / ((1.0f + anisotropy) - (2.0f * anisotropy) * rn.y);
    cos_theta = ((1.0f + anisotropy * anisotropy) - square * square) / (2.0f * anisotropy);
  }
  float sin_theta = sqrtf(jmax(0.0f, 1.0f - cos_theta * cos_theta));
  float phi = (2.0f * pif) * rn.x;
  V3 local{sin_theta * jt_cosf(phi), sin_theta * jt_sinf(phi), cos_theta};
  return mul(basis_fromz(-outgoing), local);
}
inline float sample_phasefunction_pdf(float anisotropy, V3 outgoing, V3 incoming) {
  return eval_phasefunction(anisotropy, outgoing, incoming);
}

}  // namespace orc
