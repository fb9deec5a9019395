// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restatement of the evaluation half of /root/reference/src/scene.jl:372-928:
// eval_camera, eval_position / eval_normal / eval_element_normal, eval_shading_position /
// eval_shading_normal, eval_texcoord, eval_color, eval_texture / lookup_texture, eval_normalmap /
// eval_element_tangents, eval_material, eval_environment, is_delta, is_volumetric.
#pragma once
#include "orc_scene.h"

namespace orc {

static const float min_roughness = 0.03f * 0.03f;  // src/scene.jl:46

struct MaterialPoint {  // src/scene.jl:266-277
  int32_t type;
  V3 emission, color;
  float opacity, roughness, metallic, ior;
  V3 density, scattering;
  float scanisotropy, trdepth;
};

// src/scene.jl:372-411
inline Ray eval_camera(const Camera& camera, V2 image_uv, V2 lens_uv) {
  V2 film = camera.aspect >= 1.0f ? V2{camera.film, camera.film / camera.aspect}
                                  : V2{camera.film * camera.aspect, camera.film};
  if (!camera.orthographic) {
    V3 q{film.x * (0.5f - image_uv.x), film.y * (image_uv.y - 0.5f), camera.lens};
    V3 dc = -normalize(q);
    V3 e{(lens_uv.x * camera.aperture) / 2.0f, (lens_uv.y * camera.aperture) / 2.0f, 0.0f};
    V3 p = (dc * camera.focus) / fabsf(dc.z);
    V3 d = normalize(p - e);
    return make_ray(transform_point(camera.frame, e), transform_direction(camera.frame, d));
  } else {
    float scale = 1.0f / camera.lens;
    V3 q{(film.x * (0.5f - image_uv.x)) * scale, (film.y * (image_uv.y - 0.5f)) * scale, camera.lens};
    V3 e = V3{-q.x, -q.y, 0.0f} +
           V3{(lens_uv.x * camera.aperture) / 2.0f, (lens_uv.y * camera.aperture) / 2.0f, 0.0f};
    V3 p{-q.x, -q.y, -camera.focus};
    V3 d = normalize(p - e);
    return make_ray(transform_point(camera.frame, e), transform_direction(camera.frame, d));
  }
}

// src/scene.jl:435-477 (lines / points unreachable)
inline V3 eval_position(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.ntri() != 0) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    return transform_point(inst.frame, interp_tri(s.positions[t[0] - 1], s.positions[t[1] - 1],
                                                  s.positions[t[2] - 1], uv));
  } else if (s.nquad() != 0) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    return transform_point(inst.frame, interp_quad(s.positions[q[0] - 1], s.positions[q[1] - 1],
                                                   s.positions[q[2] - 1], s.positions[q[3] - 1], uv));
  }
  return V3{0, 0, 0};
}
// src/scene.jl:416-433
inline V3 eval_shading_position(const Scene& scene, const Instance& inst, int64_t element, V2 uv,
                                V3 /*outgoing*/) {
  return eval_position(scene, inst, element, uv);
}

// src/scene.jl:578-612
inline V3 eval_element_normal(const Scene& scene, const Instance& inst, int64_t element) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.ntri() != 0) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    return transform_normal(inst.frame, triangle_normal(s.positions[t[0] - 1], s.positions[t[1] - 1],
                                                        s.positions[t[2] - 1]));
  } else if (s.nquad() != 0) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    return transform_normal(inst.frame, quad_normal(s.positions[q[0] - 1], s.positions[q[1] - 1],
                                                    s.positions[q[2] - 1], s.positions[q[3] - 1]));
  }
  return V3{0, 0, 0};
}

// src/scene.jl:525-576
inline V3 eval_normal(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.normals.empty()) return eval_element_normal(scene, inst, element);
  if (s.ntri() != 0) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    return transform_normal(inst.frame, normalize(interp_tri(s.normals[t[0] - 1], s.normals[t[1] - 1],
                                                             s.normals[t[2] - 1], uv)));
  } else if (s.nquad() != 0) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    return transform_normal(inst.frame,
                            normalize(interp_quad(s.normals[q[0] - 1], s.normals[q[1] - 1],
                                                  s.normals[q[2] - 1], s.normals[q[3] - 1], uv)));
  }
  return V3{0, 0, 0};
}

// src/scene.jl:753-788
inline V2 eval_texcoord(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.texcoords.empty()) return uv;
  if (s.ntri() != 0) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    return interp_tri(s.texcoords[t[0] - 1], s.texcoords[t[1] - 1], s.texcoords[t[2] - 1], uv);
  } else if (s.nquad() != 0) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    return interp_quad(s.texcoords[q[0] - 1], s.texcoords[q[1] - 1], s.texcoords[q[2] - 1],
                       s.texcoords[q[3] - 1], uv);
  }
  return V2{0, 0};
}

// src/scene.jl:690-720
inline V4 eval_color(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.colors.empty()) return V4{1, 1, 1, 1};
  if (s.ntri() != 0) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    return interp_tri(s.colors[t[0] - 1], s.colors[t[1] - 1], s.colors[t[2] - 1], uv);
  } else if (s.nquad() != 0) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    return interp_quad(s.colors[q[0] - 1], s.colors[q[1] - 1], s.colors[q[2] - 1], s.colors[q[3] - 1], uv);
  }
  return V4{0, 0, 0, 0};
}

// src/scene.jl:836-849 + src/color.jl:12-23
inline V4 lookup_texture(const Texture& tex, int64_t i, int64_t j, bool as_linear) {
  V4 color;
  if (!tex.pixelsf.empty()) {
    color = tex.pixelsf[j * tex.width + i];
  } else {
    const uint8_t* b = &tex.pixelsb[4 * (j * tex.width + i)];
    color = V4{(float)b[0] / 255.0f, (float)b[1] / 255.0f, (float)b[2] / 255.0f, (float)b[3] / 255.0f};
  }
  if (as_linear && !tex.linear)
    return V4{srgb_to_rgb(color.x), srgb_to_rgb(color.y), srgb_to_rgb(color.z), color.w};
  return color;
}

// Julia mod1(x, 1f0): mod(x,1) with 0 mapped to 1 (Q10); mod(x,y) = rem-based with sign fix
inline float mod1_one(float x) {
  float r = fmodf(x, 1.0f);  // exact
  float m;
  if (r == 0.0f) m = 0.0f;  // copysign(r, y): +0
  else if (r < 0.0f) m = r + 1.0f;
  else m = r;
  return m == 0.0f ? 1.0f : m;
}

// src/scene.jl:790-834 (clamp_to_edge / no_interpolation are never set by any caller)
inline V4 eval_texture(const Texture& tex, V2 uv, bool as_linear) {
  if (tex.width == 0 || tex.height == 0) return V4{0, 0, 0, 0};
  int64_t sx = tex.width, sy = tex.height;
  float s = mod1_one(uv.x) * (float)sx;
  if (s < 0.0f) s += (float)sx;
  float t = mod1_one(uv.y) * (float)sy;
  if (t < 0.0f) t += (float)sy;
  int64_t i = jclampi((int64_t)s, 0, sx - 1);  // trunc(Int, s)
  int64_t j = jclampi((int64_t)t, 0, sy - 1);
  int64_t ii = (i + 1) % sx;
  int64_t jj = (j + 1) % sy;
  float u = s - (float)i;
  float v = t - (float)j;
  return ((lookup_texture(tex, i, j, as_linear) * (1.0f - u) * (1.0f - v) +
           lookup_texture(tex, i, jj, as_linear) * (1.0f - u) * v) +
          lookup_texture(tex, ii, j, as_linear) * u * (1.0f - v)) +
         lookup_texture(tex, ii, jj, as_linear) * u * v;
}
// src/scene.jl:675-688
inline V4 eval_texture(const Scene& scene, int64_t texture, V2 uv, bool ldr_as_linear) {
  if (texture == invalid_id) return V4{1, 1, 1, 1};
  return eval_texture(scene.textures[texture - 1], uv, ldr_as_linear);
}

// src/scene.jl:851-891
inline void eval_element_tangents(const Scene& scene, const Instance& inst, int64_t element, V3* tu,
                                  V3* tv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  if (s.ntri() != 0 && !s.texcoords.empty()) {
    const int64_t* t = &s.triangles[3 * (element - 1)];
    V3 a, b;
    triangle_tangents_fromuv(s.positions[t[0] - 1], s.positions[t[1] - 1], s.positions[t[2] - 1],
                             s.texcoords[t[0] - 1], s.texcoords[t[1] - 1], s.texcoords[t[2] - 1], &a, &b);
    *tu = transform_direction(inst.frame, a);
    *tv = transform_direction(inst.frame, b);
  } else if (s.nquad() != 0 && !s.texcoords.empty()) {
    const int64_t* q = &s.quads[4 * (element - 1)];
    V3 a, b;
    // quad_tangents_fromuv with current_uv = (0,0): always the (p1,p2,p4) triangle (geometry.jl:318-332)
    triangle_tangents_fromuv(s.positions[q[0] - 1], s.positions[q[1] - 1], s.positions[q[3] - 1],
                             s.texcoords[q[0] - 1], s.texcoords[q[1] - 1], s.texcoords[q[3] - 1], &a, &b);
    *tu = transform_direction(inst.frame, a);
    *tv = transform_direction(inst.frame, b);
  } else {
    *tu = V3{0, 0, 0};
    *tv = V3{0, 0, 0};
  }
}

// src/scene.jl:722-751
inline V3 eval_normalmap(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Shape& s = scene.shapes[inst.shape - 1];
  const Material& m = scene.materials[inst.material - 1];
  V3 normal = eval_normal(scene, inst, element, uv);
  V2 texcoord = eval_texcoord(scene, inst, element, uv);
  if (m.normal_tex != invalid_id && (s.ntri() != 0 || s.nquad() != 0)) {
    const Texture& ntex = scene.textures[m.normal_tex - 1];
    V3 nm = xyz(eval_texture(ntex, texcoord, false));
    nm = V3{nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f};
    V3 tu, tv;
    eval_element_tangents(scene, inst, element, &tu, &tv);
    V3 f1 = orthonormalize(tu, normal);
    V3 f2 = normalize(cross(normal, tu));  // uses the un-orthonormalised tu (frame[1] before reassign)
    bool flip_v = dot(f2, tv) < 0.0f;
    float n2 = nm.y * (flip_v ? 1.0f : -1.0f);
    Frame fr{f1, f2, normal, V3{0, 0, 0}};
    normal = transform_normal(fr, V3{nm.x, n2, nm.z});
  }
  return normal;
}

// src/scene.jl:479-523
inline V3 eval_shading_normal(const Scene& scene, const Instance& inst, int64_t element, V2 uv,
                              V3 outgoing) {
  const Shape& s = scene.shapes[inst.shape - 1];
  const Material& m = scene.materials[inst.material - 1];
  if (s.ntri() != 0 || s.nquad() != 0) {
    V3 normal = eval_normal(scene, inst, element, uv);
    if (m.normal_tex != invalid_id) normal = eval_normalmap(scene, inst, element, uv);
    if (m.type == refractive) return normal;
    return dot(normal, outgoing) >= 0.0f ? normal : -normal;
  }
  return V3{0, 0, 0};
}

// src/scene.jl:615-673
inline MaterialPoint eval_material(const Scene& scene, const Instance& inst, int64_t element, V2 uv) {
  const Material& m = scene.materials[inst.material - 1];
  V2 texcoord = eval_texcoord(scene, inst, element, uv);
  V4 emission_tex = eval_texture(scene, m.emission_tex, texcoord, true);
  V4 color_shp = eval_color(scene, inst, element, uv);
  V4 color_tex = eval_texture(scene, m.color_tex, texcoord, true);
  V4 roughness_tex = eval_texture(scene, m.roughness_tex, texcoord, false);
  V4 scattering_tex = eval_texture(scene, m.scattering_tex, texcoord, true);

  MaterialPoint p;
  p.type = m.type;
  p.emission = m.emission * xyz(emission_tex);
  p.color = (m.color * xyz(color_tex)) * xyz(color_shp);
  p.opacity = (m.opacity * color_tex.w) * color_shp.w;
  p.metallic = m.metallic * roughness_tex.z;
  float roughness = m.roughness * roughness_tex.y;
  roughness = roughness * roughness;
  p.ior = m.ior;
  p.scattering = m.scattering * xyz(scattering_tex);
  p.scanisotropy = m.scanisotropy;
  p.trdepth = m.trdepth;
  if (m.type == refractive || m.type == volumetric || m.type == subsurface) {
    V3 c{jclamp(p.color.x, 0.0001f, 1.0f), jclamp(p.color.y, 0.0001f, 1.0f), jclamp(p.color.z, 0.0001f, 1.0f)};
    p.density = V3{-jt_logf(c.x), -jt_logf(c.y), -jt_logf(c.z)} / m.trdepth;
  } else {
    p.density = V3{0, 0, 0};
  }
  if (m.type == matte || m.type == gltfpbr || m.type == glossy) {
    roughness = jclamp(roughness, min_roughness, 1.0f);
  } else if (m.type == volumetric) {
    roughness = 0.0f;
  } else if (roughness < min_roughness) {
    roughness = 0.0f;
  }
  p.roughness = roughness;
  return p;
}

// src/scene.jl:901-914
inline V3 eval_environment(const Scene& scene, const Environment& env, V3 direction) {
  V3 wl = transform_direction(inverse(env.frame, false), direction);
  V2 texcoord{jt_atan2f(wl.z, wl.x) / (2.0f * pif), jt_acosf(jclamp(wl.y, -1.0f, 1.0f)) / pif};
  if (texcoord.x < 0.0f) texcoord.x = texcoord.x + 1.0f;
  return env.emission * xyz(eval_texture(scene, env.emission_tex, texcoord, false));
}
// src/scene.jl:893-899
inline V3 eval_environment(const Scene& scene, V3 direction) {
  V3 emission{0, 0, 0};
  for (const Environment& env : scene.environments) emission = emission + eval_environment(scene, env, direction);
  return emission;
}

// src/scene.jl:916-920
inline bool is_delta(const MaterialPoint& m) {
  return (m.type == reflective && m.roughness == 0.0f) || (m.type == refractive && m.roughness == 0.0f) ||
         (m.type == transparent && m.roughness == 0.0f) || (m.type == volumetric);
}
// src/scene.jl:922-928
inline bool is_volumetric(const Scene& scene, const Instance& inst) {
  int32_t t = scene.materials[inst.material - 1].type;
  return t == refractive || t == volumetric || t == subsurface;
}

}  // namespace orc
