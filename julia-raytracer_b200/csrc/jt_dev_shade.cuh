// jt_dev_shade.cuh -- surface / material / texture / environment evaluation and the BSDF lobes.
//
// Device implementation of the reference's evaluation functions and lobe math:
//   src/scene.jl:372-928  eval_camera, eval_position, eval_normal, eval_element_normal,
//                         eval_shading_normal, eval_texcoord, eval_color, eval_texture,
//                         lookup_texture, eval_normalmap, eval_element_tangents, eval_material,
//                         eval_environment, is_delta, is_volumetric
//   src/shading.jl        matte, glossy, reflective, transparent, refractive, passthrough,
//                         transmittance, phase function, fresnel, GGX microfacet helpers
//   src/geometry.jl:260-332 normals, areas, interpolation, tangents
// Operation order follows the reference expression by expression (left-to-right, un-fused),
// because image parity with a shared RNG needs bit-identical control flow.
#pragma once
#include "jt_dev_math.cuh"

enum {
  MAT_MATTE = 0, MAT_GLOSSY, MAT_REFLECTIVE, MAT_TRANSPARENT, MAT_REFRACTIVE, MAT_SUBSURFACE,
  MAT_VOLUMETRIC, MAT_GLTFPBR
};
#define JT_MIN_ROUGHNESS (0.03f * 0.03f) /* src/scene.jl:46 */

struct MatPoint {  // MaterialPoint, src/scene.jl:266-277
  int type;
  f3 emission, color;
  float opacity, roughness, ior;
  f3 density, scattering;
  float scanisotropy;
};

struct f4v {
  float x, y, z, w;
};
JT_DEV f4v operator*(f4v a, float s) { return f4v{a.x * s, a.y * s, a.z * s, a.w * s}; }
JT_DEV f4v operator+(f4v a, f4v b) { return f4v{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

// ---- geometry helpers ---------------------------------------------------------------------------
JT_DEV f3 tri_normal(f3 a, f3 b, f3 c) { return normalize3(cross3(b - a, c - a)); }
JT_DEV f3 quad_normal(f3 a, f3 b, f3 c, f3 d) { return normalize3(tri_normal(a, b, d) + tri_normal(c, d, b)); }

// interpolate_triangle: p1*(1-u-v) + p2*u + p3*v
JT_DEV f3 lerp_tri3(f3 a, f3 b, f3 c, float u, float v) {
  float w = (1.0f - u) - v;
  return (a * w + b * u) + c * v;
}
JT_DEV f2 lerp_tri2(f2 a, f2 b, f2 c, float u, float v) {
  float w = (1.0f - u) - v;
  return f2{(a.x * w + b.x * u) + c.x * v, (a.y * w + b.y * u) + c.y * v};
}
JT_DEV f4v lerp_tri4(f4v a, f4v b, f4v c, float u, float v) {
  float w = (1.0f - u) - v;
  return (a * w + b * u) + c * v;
}

struct ElemRef {  // resolved element: vertex indices into the shape's attribute arrays
  const JtShapeRec* sh;
  int4 q;
};
JT_DEV ElemRef elem_ref(const JtDevScene& S, const JtInstanceRec& I, int elem) {
  const JtShapeRec* sh = &S.shapes[I.shape];
  return ElemRef{sh, __ldg(S.elements + sh->elem_off + elem)};
}
JT_DEV f3 vpos(const JtDevScene& S, const JtShapeRec* sh, int v) { return ld3(S.positions + 3 * (size_t)(sh->pos_off + v)); }
JT_DEV f3 vnorm(const JtDevScene& S, const JtShapeRec* sh, int v) { return ld3(S.normals + 3 * (size_t)(sh->norm_off + v)); }
JT_DEV f2 vuv(const JtDevScene& S, const JtShapeRec* sh, int v) {
  const float* p = S.texcoords + 2 * (size_t)(sh->uv_off + v);
  return f2{p[0], p[1]};
}
JT_DEV f4v vcol(const JtDevScene& S, const JtShapeRec* sh, int v) {
  float4 c = __ldg(S.colors + sh->col_off + v);
  return f4v{c.x, c.y, c.z, c.w};
}

// eval_position, src/scene.jl:435-477
JT_DEV f3 eval_position(const JtDevScene& S, const JtInstanceRec& I, const ElemRef& E, float u, float v) {
  f3 local;
  if (E.sh->kind == 1) {
    local = lerp_tri3(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.z), u, v);
  } else {  // interpolate_quad, src/geometry.jl:278-283
    if (u + v <= 1.0f) local = lerp_tri3(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.w), u, v);
    else local = lerp_tri3(vpos(S, E.sh, E.q.z), vpos(S, E.sh, E.q.w), vpos(S, E.sh, E.q.y), 1.0f - u, 1.0f - v);
  }
  return xform_point(I.frame, local);
}

// eval_element_normal, src/scene.jl:578-612 (transform_normal uses the rigid formula, Q15)
JT_DEV f3 eval_element_normal(const JtDevScene& S, const JtInstanceRec& I, const ElemRef& E) {
  f3 n;
  if (E.sh->kind == 1) n = tri_normal(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.z));
  else n = quad_normal(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.z), vpos(S, E.sh, E.q.w));
  return normalize3(xform_vector(I.frame, n));
}

// eval_normal, src/scene.jl:525-576
JT_DEV f3 eval_normal(const JtDevScene& S, const JtInstanceRec& I, const ElemRef& E, float u, float v) {
  if (E.sh->norm_off < 0) return eval_element_normal(S, I, E);
  f3 n;
  if (E.sh->kind == 1) {
    n = lerp_tri3(vnorm(S, E.sh, E.q.x), vnorm(S, E.sh, E.q.y), vnorm(S, E.sh, E.q.z), u, v);
  } else {
    if (u + v <= 1.0f) n = lerp_tri3(vnorm(S, E.sh, E.q.x), vnorm(S, E.sh, E.q.y), vnorm(S, E.sh, E.q.w), u, v);
    else n = lerp_tri3(vnorm(S, E.sh, E.q.z), vnorm(S, E.sh, E.q.w), vnorm(S, E.sh, E.q.y), 1.0f - u, 1.0f - v);
  }
  return normalize3(xform_vector(I.frame, normalize3(n)));
}

// eval_texcoord, src/scene.jl:753-788
JT_DEV f2 eval_texcoord(const JtDevScene& S, const ElemRef& E, float u, float v) {
  if (E.sh->uv_off < 0) return f2{u, v};
  if (E.sh->kind == 1) return lerp_tri2(vuv(S, E.sh, E.q.x), vuv(S, E.sh, E.q.y), vuv(S, E.sh, E.q.z), u, v);
  if (u + v <= 1.0f) return lerp_tri2(vuv(S, E.sh, E.q.x), vuv(S, E.sh, E.q.y), vuv(S, E.sh, E.q.w), u, v);
  return lerp_tri2(vuv(S, E.sh, E.q.z), vuv(S, E.sh, E.q.w), vuv(S, E.sh, E.q.y), 1.0f - u, 1.0f - v);
}

// eval_color, src/scene.jl:690-720
JT_DEV f4v eval_color(const JtDevScene& S, const ElemRef& E, float u, float v) {
  if (E.sh->col_off < 0) return f4v{1.0f, 1.0f, 1.0f, 1.0f};
  if (E.sh->kind == 1) return lerp_tri4(vcol(S, E.sh, E.q.x), vcol(S, E.sh, E.q.y), vcol(S, E.sh, E.q.z), u, v);
  if (u + v <= 1.0f) return lerp_tri4(vcol(S, E.sh, E.q.x), vcol(S, E.sh, E.q.y), vcol(S, E.sh, E.q.w), u, v);
  return lerp_tri4(vcol(S, E.sh, E.q.z), vcol(S, E.sh, E.q.w), vcol(S, E.sh, E.q.y), 1.0f - u, 1.0f - v);
}

// ---- textures -----------------------------------------------------------------------------------
// lookup_texture, src/scene.jl:836-849; byte_to_float + srgb_to_rgb (src/color.jl:12-23) come from
// the 256-entry table the host computed with its own pow (bit-identical to the per-texel call).
JT_DEV f4v lookup_texture(const JtDevScene& S, const JtTextureRec& T, int i, int j, bool as_linear) {
  size_t at = (size_t)T.offset + (size_t)j * (size_t)T.width + (size_t)i;
  if (T.is_float) {
    float4 c = __ldg(S.texels_f + at);
    return f4v{c.x, c.y, c.z, c.w};
  }
  uchar4 b = __ldg(S.texels_b + at);
  if (as_linear && !T.linear)
    return f4v{__ldg(S.srgb_lut + b.x), __ldg(S.srgb_lut + b.y), __ldg(S.srgb_lut + b.z), (float)b.w / 255.0f};
  return f4v{(float)b.x / 255.0f, (float)b.y / 255.0f, (float)b.z / 255.0f, (float)b.w / 255.0f};
}

// mod1(x, 1f0): Julia mod() folded into (0, 1] (Q10)
JT_DEV float mod1_unit(float x) {
  float r = fmodf(x, 1.0f);  // exact remainder, sign of x
  float m = (r == 0.0f) ? 0.0f : ((r < 0.0f) ? r + 1.0f : r);
  return m == 0.0f ? 1.0f : m;
}

// eval_texture, src/scene.jl:790-834 (bilinear; wrap; no caller sets clamp_to_edge / no_interpolation)
JT_DEV f4v eval_texture(const JtDevScene& S, int tex, f2 uv, bool as_linear) {
  if (tex < 0) return f4v{1.0f, 1.0f, 1.0f, 1.0f};
  const JtTextureRec T = S.textures[tex];
  if (T.width == 0 || T.height == 0) return f4v{0.0f, 0.0f, 0.0f, 0.0f};
  float s = mod1_unit(uv.x) * (float)T.width;
  if (s < 0.0f) s += (float)T.width;
  float t = mod1_unit(uv.y) * (float)T.height;
  if (t < 0.0f) t += (float)T.height;
  int i = jl_clampi((int)s, 0, T.width - 1);
  int j = jl_clampi((int)t, 0, T.height - 1);
  int ii = (i + 1) % T.width;
  int jj = (j + 1) % T.height;
  float u = s - (float)i;
  float v = t - (float)j;
  f4v a = lookup_texture(S, T, i, j, as_linear) * (1.0f - u) * (1.0f - v);
  f4v b = lookup_texture(S, T, i, jj, as_linear) * (1.0f - u) * v;
  f4v c = lookup_texture(S, T, ii, j, as_linear) * u * (1.0f - v);
  f4v d = lookup_texture(S, T, ii, jj, as_linear) * u * v;
  return ((a + b) + c) + d;
}

// triangle_tangents_fromuv, src/geometry.jl:285-316
JT_DEV void tangents_fromuv(f3 p1, f3 p2, f3 p3, f2 uv1, f2 uv2, f2 uv3, f3* tu, f3* tv) {
  f3 p = p2 - p1, q = p3 - p1;
  float sx = uv2.x - uv1.x, sy = uv3.x - uv1.x;
  float tx = uv2.y - uv1.y, ty = uv3.y - uv1.y;
  float div = sx * ty - sy * tx;
  if (div != 0.0f) {
    *tu = f3{ty * p.x - tx * q.x, ty * p.y - tx * q.y, ty * p.z - tx * q.z} / div;
    *tv = f3{sx * q.x - sy * p.x, sx * q.y - sy * p.y, sx * q.z - sy * p.z} / div;
  } else {
    *tu = f3{1.0f, 0.0f, 0.0f};
    *tv = f3{0.0f, 1.0f, 0.0f};
  }
}

// eval_normalmap, src/scene.jl:722-751 with eval_element_tangents :851-891
JT_DEV f3 eval_normalmap(const JtDevScene& S, const JtInstanceRec& I, const ElemRef& E, const JtMaterialRec& M,
                         float u, float v) {
  f3 normal = eval_normal(S, I, E, u, v);
  f2 texcoord = eval_texcoord(S, E, u, v);
  f4v tx = eval_texture(S, M.normal_tex, texcoord, false);
  f3 nm = f3{tx.x * 2.0f - 1.0f, tx.y * 2.0f - 1.0f, tx.z * 2.0f - 1.0f};
  f3 tu = f3{0.0f, 0.0f, 0.0f}, tv = f3{0.0f, 0.0f, 0.0f};
  if (E.sh->uv_off >= 0) {
    f3 a, b;
    if (E.sh->kind == 1)
      tangents_fromuv(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.z), vuv(S, E.sh, E.q.x),
                      vuv(S, E.sh, E.q.y), vuv(S, E.sh, E.q.z), &a, &b);
    else  // quad_tangents_fromuv is always called with uv = (0,0): the (p1,p2,p4) half
      tangents_fromuv(vpos(S, E.sh, E.q.x), vpos(S, E.sh, E.q.y), vpos(S, E.sh, E.q.w), vuv(S, E.sh, E.q.x),
                      vuv(S, E.sh, E.q.y), vuv(S, E.sh, E.q.w), &a, &b);
    tu = xform_direction(I.frame, a);
    tv = xform_direction(I.frame, b);
  }
  f3 fx = orthonormalize3(tu, normal);
  f3 fy = normalize3(cross3(normal, tu));
  bool flip_v = dot3(fy, tv) < 0.0f;
  float ny = nm.y * (flip_v ? 1.0f : -1.0f);
  // transform_normal(frame, normalmap) = normalize(x*n.x + y*n.y + z*n.z)
  f3 r = (fx * nm.x + fy * ny) + normal * nm.z;
  return normalize3(r);
}

// eval_shading_normal, src/scene.jl:479-523
JT_DEV f3 eval_shading_normal(const JtDevScene& S, const JtInstanceRec& I, const ElemRef& E,
                              const JtMaterialRec& M, float u, float v, f3 outgoing) {
  f3 normal = (M.normal_tex >= 0) ? eval_normalmap(S, I, E, M, u, v) : eval_normal(S, I, E, u, v);
  if (M.type == MAT_REFRACTIVE) return normal;
  return dot3(normal, outgoing) >= 0.0f ? normal : -normal;
}

// eval_material, src/scene.jl:615-673
JT_DEV MatPoint eval_material(const JtDevScene& S, const ElemRef& E, const JtMaterialRec& M, float u, float v) {
  f2 texcoord = eval_texcoord(S, E, u, v);
  f4v emission_tex = eval_texture(S, M.emission_tex, texcoord, true);
  f4v color_shp = eval_color(S, E, u, v);
  f4v color_tex = eval_texture(S, M.color_tex, texcoord, true);
  f4v roughness_tex = eval_texture(S, M.roughness_tex, texcoord, false);
  f4v scattering_tex = eval_texture(S, M.scattering_tex, texcoord, true);
  MatPoint P;
  P.type = M.type;
  P.emission = f3{M.emission[0] * emission_tex.x, M.emission[1] * emission_tex.y, M.emission[2] * emission_tex.z};
  P.color = f3{(M.color[0] * color_tex.x) * color_shp.x, (M.color[1] * color_tex.y) * color_shp.y,
               (M.color[2] * color_tex.z) * color_shp.z};
  P.opacity = (M.opacity * color_tex.w) * color_shp.w;
  float roughness = M.roughness * roughness_tex.y;
  roughness = roughness * roughness;
  P.ior = M.ior;
  P.scattering = f3{M.scattering[0] * scattering_tex.x, M.scattering[1] * scattering_tex.y,
                    M.scattering[2] * scattering_tex.z};
  P.scanisotropy = M.scanisotropy;
  if (M.type == MAT_REFRACTIVE || M.type == MAT_VOLUMETRIC || M.type == MAT_SUBSURFACE) {
    P.density = f3{(-jt_logf(jl_clamp(P.color.x, 0.0001f, 1.0f))) / M.trdepth,
                   (-jt_logf(jl_clamp(P.color.y, 0.0001f, 1.0f))) / M.trdepth,
                   (-jt_logf(jl_clamp(P.color.z, 0.0001f, 1.0f))) / M.trdepth};
  } else {
    P.density = f3{0.0f, 0.0f, 0.0f};
  }
  if (M.type == MAT_MATTE || M.type == MAT_GLTFPBR || M.type == MAT_GLOSSY) roughness = jl_clamp(roughness, JT_MIN_ROUGHNESS, 1.0f);
  else if (M.type == MAT_VOLUMETRIC) roughness = 0.0f;
  else if (roughness < JT_MIN_ROUGHNESS) roughness = 0.0f;
  P.roughness = roughness;
  return P;
}

// eval_environment, src/scene.jl:893-914
JT_DEV f3 eval_environment(const JtDevScene& S, f3 direction) {
  f3 emission = f3{0.0f, 0.0f, 0.0f};
  for (int e = 0; e < S.num_environments; e++) {
    const JtEnvRec& En = S.environments[e];
    f3 wl = normalize3(xform_vector_transposed(En.frame, direction));
    f2 tc = f2{jt_atan2f(wl.z, wl.x) / (2.0f * JT_PIF), jt_acosf(jl_clamp(wl.y, -1.0f, 1.0f)) / JT_PIF};
    if (tc.x < 0.0f) tc.x = tc.x + 1.0f;
    f4v tx = eval_texture(S, En.emission_tex, tc, false);
    emission = emission + f3{En.emission[0] * tx.x, En.emission[1] * tx.y, En.emission[2] * tx.z};
  }
  return emission;
}

JT_DEV bool is_delta(const MatPoint& m) {  // src/scene.jl:916-920
  return (m.type == MAT_REFLECTIVE && m.roughness == 0.0f) || (m.type == MAT_REFRACTIVE && m.roughness == 0.0f) ||
         (m.type == MAT_TRANSPARENT && m.roughness == 0.0f) || (m.type == MAT_VOLUMETRIC);
}
JT_DEV bool is_volumetric_type(int t) {  // src/scene.jl:922-928
  return t == MAT_REFRACTIVE || t == MAT_VOLUMETRIC || t == MAT_SUBSURFACE;
}

// eval_camera, src/scene.jl:372-411
JT_DEV DRay eval_camera(const JtCameraRec& C, f2 image_uv, f2 lens_uv) {
  float film_x = C.aspect >= 1.0f ? C.film : C.film * C.aspect;
  float film_y = C.aspect >= 1.0f ? C.film / C.aspect : C.film;
  f3 e, d;
  if (!C.orthographic) {
    f3 q = f3{film_x * (0.5f - image_uv.x), film_y * (image_uv.y - 0.5f), C.lens};
    f3 dc = -normalize3(q);
    e = f3{(lens_uv.x * C.aperture) / 2.0f, (lens_uv.y * C.aperture) / 2.0f, 0.0f};
    f3 p = (dc * C.focus) / fabsf(dc.z);
    d = normalize3(p - e);
  } else {
    float scale = 1.0f / C.lens;
    f3 q = f3{(film_x * (0.5f - image_uv.x)) * scale, (film_y * (image_uv.y - 0.5f)) * scale, C.lens};
    e = f3{-q.x, -q.y, 0.0f} + f3{(lens_uv.x * C.aperture) / 2.0f, (lens_uv.y * C.aperture) / 2.0f, 0.0f};
    f3 p = f3{-q.x, -q.y, -C.focus};
    d = normalize3(p - e);
  }
  return DRay{xform_point(C.frame, e), xform_direction(C.frame, d), JT_RAY_EPS, INFINITY};
}

// sample_disk, src/sampling.jl:12-16
JT_DEV f2 sample_disk(f2 ruv) {
  float r = sqrtf(ruv.y);
  float phi = (2.0f * JT_PIF) * ruv.x;
  return f2{jt_cosf(phi) * r, jt_sinf(phi) * r};
}

// sample_camera, src/trace.jl:651-674
JT_DEV DRay sample_camera(const JtCameraRec& C, int i, int j, int w, int h, f2 puv, f2 luv, bool tent) {
  f2 uv;
  if (!tent) {
    uv = f2{((float)i + puv.x) / (float)w, ((float)j + puv.y) / (float)h};
  } else {
    float fx = puv.x < 0.5f ? sqrtf(2.0f * puv.x) - 1.0f : 1.0f - sqrtf(2.0f - 2.0f * puv.x);
    float fy = puv.y < 0.5f ? sqrtf(2.0f * puv.y) - 1.0f : 1.0f - sqrtf(2.0f - 2.0f * puv.y);
    f2 fuv = f2{2.0f * fx + 0.5f, 2.0f * fy + 0.5f};
    uv = f2{((float)i + fuv.x) / (float)w, ((float)j + fuv.y) / (float)h};
  }
  return eval_camera(C, uv, sample_disk(luv));
}

// =================================================================================================
// shading.jl
// =================================================================================================
JT_DEV bool same_hemisphere(f3 n, f3 o, f3 i) { return dot3(n, o) * dot3(n, i) >= 0.0f; }
JT_DEV f3 up_normal_of(f3 n, f3 o) { return dot3(n, o) <= 0.0f ? -n : n; }

JT_DEV float fresnel_dielectric(float eta, f3 normal, f3 outgoing) {  // :695-714
  float cosw = fabsf(dot3(normal, outgoing));
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

JT_DEV float conductor_channel(float eta, float cosw, float cos2, float sin2) {  // etak = 0, :831-851
  float eta2 = eta * eta;
  float etak2 = 0.0f * 0.0f;
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
// fresnel_conductor(reflectivity_to_eta(color), 0, normal, outgoing), :820-851
JT_DEV f3 fresnel_conductor_color(f3 color, f3 normal, f3 outgoing) {
  float cosw = dot3(normal, outgoing);
  if (cosw <= 0.0f) return f3{0.0f, 0.0f, 0.0f};
  cosw = jl_clamp(cosw, -1.0f, 1.0f);
  float cos2 = cosw * cosw;
  float sin2 = jl_clamp(1.0f - cos2, 0.0f, 1.0f);
  float rx = sqrtf(jl_clamp(color.x, 0.0f, 0.99f)), ry = sqrtf(jl_clamp(color.y, 0.0f, 0.99f)),
        rz = sqrtf(jl_clamp(color.z, 0.0f, 0.99f));
  f3 eta = f3{(1.0f + rx) / (1.0f - rx), (1.0f + ry) / (1.0f - ry), (1.0f + rz) / (1.0f - rz)};
  return f3{conductor_channel(eta.x, cosw, cos2, sin2), conductor_channel(eta.y, cosw, cos2, sin2),
            conductor_channel(eta.z, cosw, cos2, sin2)};
}

struct Basis {
  f3 x, y, z;
};
JT_DEV Basis basis_fromz(f3 v) {  // :724-732
  f3 z = normalize3(v);
  float sign = copysignf(1.0f, z.z);
  float a = -1.0f / (sign + z.z);
  float b = (z.x * z.y) * a;
  Basis B;
  B.x = f3{1.0f + ((sign * z.x) * z.x) * a, sign * b, (-sign) * z.x};
  B.y = f3{b, sign + (z.y * z.y) * a, -z.y};
  B.z = z;
  return B;
}
JT_DEV f3 basis_mul(const Basis& B, f3 l) { return (B.x * l.x + B.y * l.y) + B.z * l.z; }

JT_DEV f3 sample_hemisphere_cos(f3 normal, f2 ruv) {  // :716-722
  float z = sqrtf(ruv.y);
  float r = sqrtf(1.0f - z * z);
  float phi = (2.0f * JT_PIF) * ruv.x;
  f3 local = f3{r * jt_cosf(phi), r * jt_sinf(phi), z};
  return normalize3(basis_mul(basis_fromz(normal), local));
}
JT_DEV float sample_hemisphere_cos_pdf(f3 normal, f3 direction) {  // src/sampling.jl:24-27
  float cosw = dot3(normal, direction);
  return cosw <= 0.0f ? 0.0f : cosw / JT_PIF;
}

JT_DEV float microfacet_distribution(float roughness, f3 normal, f3 halfway) {  // GGX, :734-750
  float cosine = dot3(normal, halfway);
  if (cosine <= 0.0f) return 0.0f;
  float roughness2 = roughness * roughness;
  float cosine2 = cosine * cosine;
  float k = (cosine2 * roughness2 + 1.0f) - cosine2;
  return roughness2 / ((JT_PIF * k) * k);
}
JT_DEV float microfacet_shadowing1(float roughness, f3 normal, f3 halfway, f3 direction) {  // :752-773
  float cosine = dot3(normal, direction);
  float cosineh = dot3(halfway, direction);
  if (cosine * cosineh <= 0.0f) return 0.0f;
  float roughness2 = roughness * roughness;
  float cosine2 = cosine * cosine;
  return (2.0f * fabsf(cosine)) / (fabsf(cosine) + sqrtf((cosine2 - roughness2 * cosine2) + roughness2));
}
JT_DEV float microfacet_shadowing(float roughness, f3 n, f3 h, f3 o, f3 i) {  // :775-785
  return microfacet_shadowing1(roughness, n, h, o) * microfacet_shadowing1(roughness, n, h, i);
}
JT_DEV f3 sample_microfacet(float roughness, f3 normal, f2 rn) {  // :787-803
  float phi = (2.0f * JT_PIF) * rn.x;
  float theta = jt_atanf(roughness * sqrtf(rn.y / (1.0f - rn.y)));
  float st = jt_sinf(theta), ct = jt_cosf(theta);
  f3 local = f3{jt_cosf(phi) * st, jt_sinf(phi) * st, ct};
  return normalize3(basis_mul(basis_fromz(normal), local));
}
JT_DEV float sample_microfacet_pdf(float roughness, f3 normal, f3 halfway) {  // :805-816
  float cosine = dot3(normal, halfway);
  if (cosine < 0.0f) return 0.0f;
  return microfacet_distribution(roughness, normal, halfway) * cosine;
}

// ---- eval_bsdfcos dispatch, src/trace.jl:692-755 ---------------------------------------------------
JT_DEV f3 eval_bsdfcos(const MatPoint& m, f3 n, f3 o, f3 i) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f};
  if (m.roughness == 0.0f) return zero;
  float ndi = dot3(n, i), ndo = dot3(n, o);
  switch (m.type) {
    case MAT_MATTE: {  // src/shading.jl:14-19
      if (ndi * ndo <= 0.0f) return zero;
      return (m.color / JT_PIF) * fabsf(ndi);
    }
    case MAT_GLOSSY: {  // :39-59
      if (ndi * ndo <= 0.0f) return zero;
      f3 un = up_normal_of(n, o);
      float F1 = fresnel_dielectric(m.ior, un, o);
      f3 h = normalize3(i + o);
      float F = fresnel_dielectric(m.ior, h, i);
      float D = microfacet_distribution(m.roughness, un, h);
      float G = microfacet_shadowing(m.roughness, un, h, o, i);
      float ni = fabsf(dot3(un, i));
      float spec = ((((1.0f * F) * D) * G) / ((4.0f * dot3(un, o)) * dot3(un, i))) * ni;
      f3 diff = ((m.color * (1.0f - F1)) / JT_PIF) * ni;
      return f3{diff.x + spec, diff.y + spec, diff.z + spec};
    }
    case MAT_REFLECTIVE: {  // :103-119
      if (ndi * ndo <= 0.0f) return zero;
      f3 un = up_normal_of(n, o);
      f3 h = normalize3(i + o);
      f3 F = fresnel_conductor_color(m.color, h, i);
      float D = microfacet_distribution(m.roughness, un, h);
      float G = microfacet_shadowing(m.roughness, un, h, o, i);
      return (((F * D) * G) / ((4.0f * dot3(un, o)) * dot3(un, i))) * fabsf(dot3(un, i));
    }
    case MAT_TRANSPARENT: {  // :323-349
      f3 un = up_normal_of(n, o);
      if (ndi * ndo >= 0.0f) {
        f3 h = normalize3(i + o);
        float F = fresnel_dielectric(m.ior, h, o);
        float D = microfacet_distribution(m.roughness, un, h);
        float G = microfacet_shadowing(m.roughness, un, h, o, i);
        float s = ((((1.0f * F) * D) * G) / ((4.0f * dot3(un, o)) * dot3(un, i))) * fabsf(dot3(un, i));
        return f3{s, s, s};
      } else {
        f3 r = reflect3(-i, un);
        f3 h = normalize3(r + o);
        float F = fresnel_dielectric(m.ior, h, o);
        float D = microfacet_distribution(m.roughness, un, h);
        float G = microfacet_shadowing(m.roughness, un, h, o, r);
        return ((((m.color * (1.0f - F)) * D) * G) / ((4.0f * dot3(un, o)) * dot3(un, r))) * fabsf(dot3(un, r));
      }
    }
    case MAT_REFRACTIVE:
    case MAT_SUBSURFACE: {  // :448-482
      bool entering = ndo >= 0.0f;
      f3 un = entering ? n : -n;
      float rel_ior = entering ? m.ior : (1.0f / m.ior);
      if (ndi * ndo >= 0.0f) {
        f3 h = normalize3(i + o);
        float F = fresnel_dielectric(rel_ior, h, o);
        float D = microfacet_distribution(m.roughness, un, h);
        float G = microfacet_shadowing(m.roughness, un, h, o, i);
        float s = ((((1.0f * F) * D) * G) / fabsf((4.0f * ndo) * ndi)) * fabsf(ndi);
        return f3{s, s, s};
      } else {
        f3 h = (-normalize3(rel_ior * i + o)) * (entering ? 1.0f : -1.0f);
        float F = fresnel_dielectric(rel_ior, h, o);
        float D = microfacet_distribution(m.roughness, un, h);
        float G = microfacet_shadowing(m.roughness, un, h, o, i);
        float a = fabsf((dot3(o, h) * dot3(i, h)) / (dot3(o, n) * dot3(i, n)));
        float q = rel_ior * dot3(h, i) + dot3(h, o);
        float s = (((((1.0f * a) * (1.0f - F)) * D) * G) / (q * q)) * fabsf(ndi);
        return f3{s, s, s};
      }
    }
    default:
      return zero;
  }
}

// ---- sample_bsdfcos dispatch, src/trace.jl:780-849 ---------------------------------------------------
JT_DEV f3 sample_bsdfcos(const MatPoint& m, f3 n, f3 o, float rnl, f2 rn) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f};
  if (m.roughness == 0.0f) return zero;
  switch (m.type) {
    case MAT_MATTE:  // :21-24
      return sample_hemisphere_cos(up_normal_of(n, o), rn);
    case MAT_GLOSSY: {  // :61-81
      f3 un = up_normal_of(n, o);
      if (rnl < fresnel_dielectric(m.ior, un, o)) {
        f3 h = sample_microfacet(m.roughness, un, rn);
        f3 i = reflect3(o, h);
        return same_hemisphere(un, o, i) ? i : zero;
      }
      return sample_hemisphere_cos(un, rn);
    }
    case MAT_REFLECTIVE: {  // :121-135
      f3 un = up_normal_of(n, o);
      f3 h = sample_microfacet(m.roughness, un, rn);
      f3 i = reflect3(o, h);
      return same_hemisphere(un, o, i) ? i : zero;
    }
    case MAT_TRANSPARENT: {  // :351-376
      f3 un = up_normal_of(n, o);
      f3 h = sample_microfacet(m.roughness, un, rn);
      if (rnl < fresnel_dielectric(m.ior, h, o)) {
        f3 i = reflect3(o, h);
        return same_hemisphere(un, o, i) ? i : zero;
      }
      f3 r = reflect3(o, h);
      f3 i = -reflect3(r, un);
      return same_hemisphere(un, o, i) ? zero : i;
    }
    case MAT_REFRACTIVE:
    case MAT_SUBSURFACE: {  // :484-509
      bool entering = dot3(n, o) >= 0.0f;
      f3 un = entering ? n : -n;
      f3 h = sample_microfacet(m.roughness, un, rn);
      if (rnl < fresnel_dielectric(entering ? m.ior : (1.0f / m.ior), h, o)) {
        f3 i = reflect3(o, h);
        return same_hemisphere(un, o, i) ? i : zero;
      }
      f3 i = refract3(o, h, entering ? (1.0f / m.ior) : m.ior);
      return same_hemisphere(un, o, i) ? zero : i;
    }
    default:
      return zero;
  }
}

// ---- sample_bsdfcos_pdf dispatch, src/trace.jl:874-943 ---------------------------------------------------
JT_DEV float sample_bsdfcos_pdf(const MatPoint& m, f3 n, f3 o, f3 i) {
  if (m.roughness == 0.0f) return 0.0f;
  float ndi = dot3(n, i), ndo = dot3(n, o);
  switch (m.type) {
    case MAT_MATTE: {  // :26-37
      if (ndi * ndo <= 0.0f) return 0.0f;
      return sample_hemisphere_cos_pdf(up_normal_of(n, o), i);
    }
    case MAT_GLOSSY: {  // :83-101
      if (ndi * ndo <= 0.0f) return 0.0f;
      f3 un = up_normal_of(n, o);
      f3 h = normalize3(o + i);
      float F = fresnel_dielectric(m.ior, un, o);
      return (F * sample_microfacet_pdf(m.roughness, un, h)) / (4.0f * fabsf(dot3(o, h))) +
             (1.0f - F) * sample_hemisphere_cos_pdf(un, i);
    }
    case MAT_REFLECTIVE: {  // :137-151
      if (ndi * ndo <= 0.0f) return 0.0f;
      f3 un = up_normal_of(n, o);
      f3 h = normalize3(o + i);
      return sample_microfacet_pdf(m.roughness, un, h) / (4.0f * fabsf(dot3(o, h)));
    }
    case MAT_TRANSPARENT: {  // :378-401
      f3 un = up_normal_of(n, o);
      if (ndi * ndo >= 0.0f) {
        f3 h = normalize3(i + o);
        return (fresnel_dielectric(m.ior, h, o) * sample_microfacet_pdf(m.roughness, un, h)) /
               (4.0f * fabsf(dot3(o, h)));
      }
      f3 r = reflect3(-i, un);
      f3 h = normalize3(r + o);
      float d = (1.0f - fresnel_dielectric(m.ior, h, o)) * sample_microfacet_pdf(m.roughness, un, h);
      return d / (4.0f * fabsf(dot3(o, h)));
    }
    case MAT_REFRACTIVE:
    case MAT_SUBSURFACE: {  // :511-534
      bool entering = ndo >= 0.0f;
      f3 un = entering ? n : -n;
      float rel_ior = entering ? m.ior : (1.0f / m.ior);
      if (ndi * ndo >= 0.0f) {
        f3 h = normalize3(i + o);
        return (fresnel_dielectric(rel_ior, h, o) * sample_microfacet_pdf(m.roughness, un, h)) /
               (4.0f * fabsf(dot3(o, h)));
      }
      f3 h = (-normalize3(rel_ior * i + o)) * (entering ? 1.0f : -1.0f);
      float q = rel_ior * dot3(h, i) + dot3(h, o);
      return (((1.0f - fresnel_dielectric(rel_ior, h, o)) * sample_microfacet_pdf(m.roughness, un, h)) *
              fabsf(dot3(h, i))) /
             (q * q);
    }
    default:
      return 0.0f;
  }
}

// ---- delta lobes: eval_delta :757-778, sample_delta :851-872, sample_delta_pdf :945-966 ----------------
JT_DEV f3 eval_delta(const MatPoint& m, f3 n, f3 o, f3 i) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f}, one = f3{1.0f, 1.0f, 1.0f};
  if (m.roughness != 0.0f) return zero;
  float ndi = dot3(n, i), ndo = dot3(n, o);
  switch (m.type) {
    case MAT_REFLECTIVE: {  // src/shading.jl:202-213
      if (ndi * ndo <= 0.0f) return zero;
      return fresnel_conductor_color(m.color, up_normal_of(n, o), o);
    }
    case MAT_TRANSPARENT: {  // :403-416
      f3 un = up_normal_of(n, o);
      if (ndi * ndo >= 0.0f) return one * fresnel_dielectric(m.ior, un, o);
      return m.color * (1.0f - fresnel_dielectric(m.ior, un, o));
    }
    case MAT_REFRACTIVE: {  // :536-562
      if ((double)fabsf(m.ior - 1.0f) < 1e-3) return ndi * ndo <= 0.0f ? one : zero;
      bool entering = ndo >= 0.0f;
      f3 un = entering ? n : -n;
      float rel_ior = entering ? m.ior : (1.0f / m.ior);
      if (ndi * ndo >= 0.0f) return one * fresnel_dielectric(rel_ior, un, o);
      return (one * (1.0f / (rel_ior * rel_ior))) * (1.0f - fresnel_dielectric(rel_ior, un, o));
    }
    case MAT_VOLUMETRIC:  // eval_passthrough :636-637
      return ndi * ndo >= 0.0f ? zero : one;
    default:
      return zero;
  }
}
JT_DEV f3 sample_delta(const MatPoint& m, f3 n, f3 o, float rnl) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f};
  if (m.roughness != 0.0f) return zero;
  switch (m.type) {
    case MAT_REFLECTIVE:  // :215-218
      return reflect3(o, up_normal_of(n, o));
    case MAT_TRANSPARENT: {  // :418-431
      f3 un = up_normal_of(n, o);
      if (rnl < fresnel_dielectric(m.ior, un, o)) return reflect3(o, un);
      return -o;
    }
    case MAT_REFRACTIVE: {  // :564-582
      if ((double)fabsf(m.ior - 1.0f) < 1e-3) return -o;
      bool entering = dot3(n, o) >= 0.0f;
      f3 un = entering ? n : -n;
      float rel_ior = entering ? m.ior : (1.0f / m.ior);
      if (rnl < fresnel_dielectric(rel_ior, un, o)) return reflect3(o, un);
      return refract3(o, un, 1.0f / rel_ior);
    }
    case MAT_VOLUMETRIC:  // sample_passthrough :639
      return -o;
    default:
      return zero;
  }
}
JT_DEV float sample_delta_pdf(const MatPoint& m, f3 n, f3 o, f3 i) {
  if (m.roughness != 0.0f) return 0.0f;
  float ndi = dot3(n, i), ndo = dot3(n, o);
  switch (m.type) {
    case MAT_REFLECTIVE:  // :220-225
      return ndi * ndo <= 0.0f ? 0.0f : 1.0f;
    case MAT_TRANSPARENT: {  // :433-446
      f3 un = up_normal_of(n, o);
      if (ndi * ndo >= 0.0f) return fresnel_dielectric(m.ior, un, o);
      return 1.0f - fresnel_dielectric(m.ior, un, o);
    }
    case MAT_REFRACTIVE: {  // :584-604
      if (fabsf(m.ior - 1.0f) < 0.001f) return ndi * ndo < 0.0f ? 1.0f : 0.0f;
      bool entering = ndo >= 0.0f;
      f3 un = entering ? n : -n;
      float rel_ior = entering ? m.ior : (1.0f / m.ior);
      if (ndi * ndo >= 0.0f) return fresnel_dielectric(rel_ior, un, o);
      return 1.0f - fresnel_dielectric(rel_ior, un, o);
    }
    case MAT_VOLUMETRIC:  // :641-646
      return ndi * ndo >= 0.0f ? 0.0f : 1.0f;
    default:
      return 0.0f;
  }
}

// ---- volumes, src/shading.jl:650-693 and src/trace.jl:1086-1115 ---------------------------------------
JT_DEV f3 eval_transmittance(f3 density, float distance) {
  return f3{jt_expf((-density.x) * distance), jt_expf((-density.y) * distance), jt_expf((-density.z) * distance)};
}
JT_DEV float sample_transmittance(f3 density, float max_distance, float rl, float rd) {
  int channel = jl_clampi((int)(rl * 3.0f), 1, 3);  // Q6: {1,1,2}
  float dc = channel == 1 ? density.x : (channel == 2 ? density.y : density.z);
  float distance = dc == 0.0f ? INFINITY : (-jt_logf(1.0f - rd)) / dc;
  return jl_min(distance, max_distance);
}
JT_DEV float sample_transmittance_pdf(f3 density, float distance, float max_distance) {
  if (distance < max_distance) {
    float a = density.x * jt_expf((-density.x) * distance);
    float b = density.y * jt_expf((-density.y) * distance);
    float c = density.z * jt_expf((-density.z) * distance);
    return ((a + b) + c) / 3.0f;
  }
  float a = jt_expf((-density.x) * max_distance), b = jt_expf((-density.y) * max_distance),
        c = jt_expf((-density.z) * max_distance);
  return ((a + b) + c) / 3.0f;
}
JT_DEV float eval_phasefunction(float g, f3 outgoing, f3 incoming) {
  float cosine = -dot3(outgoing, incoming);
  float denom = (1.0f + g * g) - (2.0f * g) * cosine;
  return (1.0f - g * g) / (((4.0f * JT_PIF) * denom) * sqrtf(denom));
}
JT_DEV f3 sample_phasefunction(float g, f3 outgoing, f2 rn) {
  float cos_theta;
  if (fabsf(g) < 0.001f) {
    cos_theta = 1.0f - 2.0f * rn.y;
  } else {
    float square = (1.0f - g * g) / ((1.0f + g) - (2.0f * g) * rn.y);
    cos_theta = ((1.0f + g * g) - square * square) / (2.0f * g);
  }
  float sin_theta = sqrtf(jl_max(0.0f, 1.0f - cos_theta * cos_theta));
  float phi = (2.0f * JT_PIF) * rn.x;
  f3 local = f3{sin_theta * jt_cosf(phi), sin_theta * jt_sinf(phi), cos_theta};
  return basis_mul(basis_fromz(-outgoing), local);
}
struct VolPoint {  // the MaterialPoint fields the volume code reads
  f3 density, scattering;
  float scanisotropy;
};
JT_DEV f3 eval_scattering(const VolPoint& m, f3 o, f3 i) {
  if (is_zero3(m.density)) return f3{0.0f, 0.0f, 0.0f};
  return (m.scattering * m.density) * eval_phasefunction(m.scanisotropy, o, i);
}
JT_DEV f3 sample_scattering(const VolPoint& m, f3 o, f2 rn) {
  if (is_zero3(m.density)) return f3{0.0f, 0.0f, 0.0f};
  return sample_phasefunction(m.scanisotropy, o, rn);
}
JT_DEV float sample_scattering_pdf(const VolPoint& m, f3 o, f3 i) {
  if (is_zero3(m.density)) return 0.0f;
  return eval_phasefunction(m.scanisotropy, o, i);
}
