// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// Restatement of /root/reference/src/trace.jl: make_trace_lights :117-187, make_trace_state
// :189-213, trace_samples :215-274, trace_path :276-469, trace_naive :471-573, eval_emission
// :575-580, trace_sample :584-649, sample_camera :651-674, bsdf / delta dispatch :692-966,
// sample_lights :968-1008, sample_lights_pdf :1010-1084, volume scattering :1086-1115.
//
// The reference's unseeded `rand(Float32)` is replaced by the shared counter-based stream of
// jt_rng.h, consumed in the reference's draw order.
#pragma once
#include "../julia-raytracer_b200/csrc/jt_rng.h"
#include "orc_eval.h"
#include "orc_shading.h"

namespace orc {

struct Params {  // src/cli.jl:90-108 (fields read by the hot path)
  int camera = 1, resolution = 1280, samples = 512, bounces = 8, sampler = 1, clamp = 10;
  bool nocaustics = false, envhidden = false, tentfilter = false;
  int batch = 1;
  uint64_t seed = 0;
  int accumulate = 0;  // 0 = running-mean lerp (Q13), 1 = sums
};

struct State {  // src/trace.jl:87-96
  int64_t width = 0, height = 0, samples = 0;
  std::vector<V4> image;
  std::vector<V3> albedo, normal;
  std::vector<int64_t> hits;
};

struct Rng {
  uint64_t key;
  uint32_t draw;
  float next() { return jt_rng_float(key, draw++); }  // rand1f
  V2 next2() { float a = next(); float b = next(); return V2{a, b}; }  // rand2f: left to right
};

// ---- lights ------------------------------------------------------------------------------------
inline void make_trace_lights(Scene& scene) {  // :117-187
  scene.lights.clear();
  for (size_t h = 0; h < scene.instances.size(); h++) {
    const Instance& inst = scene.instances[h];
    const Material& m = scene.materials[inst.material - 1];
    if (m.emission == V3{0, 0, 0}) continue;
    const Shape& s = scene.shapes[inst.shape - 1];
    if (s.ntri() == 0 && s.nquad() == 0) continue;
    Light l{(int64_t)h + 1, invalid_id, {}};
    if (s.ntri() != 0) {
      l.cdf.resize(s.ntri());
      for (int64_t i = 0; i < s.ntri(); i++) {
        const int64_t* t = &s.triangles[3 * i];
        l.cdf[i] = triangle_area(s.positions[t[0] - 1], s.positions[t[1] - 1], s.positions[t[2] - 1]);
        if (i != 0) l.cdf[i] += l.cdf[i - 1];
      }
    }
    if (s.nquad() != 0) {
      l.cdf.resize(s.nquad());
      for (int64_t i = 0; i < s.nquad(); i++) {
        const int64_t* q = &s.quads[4 * i];
        l.cdf[i] = quad_area(s.positions[q[0] - 1], s.positions[q[1] - 1], s.positions[q[2] - 1],
                             s.positions[q[3] - 1]);
        if (i != 0) l.cdf[i] += l.cdf[i - 1];
      }
    }
    scene.lights.push_back(std::move(l));
  }
  for (size_t h = 0; h < scene.environments.size(); h++) {
    const Environment& env = scene.environments[h];
    if (env.emission == V3{0, 0, 0}) continue;
    Light l{invalid_id, (int64_t)h + 1, {}};
    if (env.emission_tex != invalid_id) {
      const Texture& tex = scene.textures[env.emission_tex - 1];
      l.cdf.resize(tex.width * tex.height);
      for (int64_t idx = 0; idx < (int64_t)l.cdf.size(); idx++) {
        int64_t i = idx % tex.width, j = idx / tex.width;
        float th = (((float)j + 0.5f) * pif) / (float)tex.height;
        V4 value = lookup_texture(tex, i, j, false);
        l.cdf[idx] = maximum(value) * jt_sinf(th);  // Q8: max over RGBA (alpha = 1)
        if (idx != 0) l.cdf[idx] += l.cdf[idx - 1];
      }
    }
    scene.lights.push_back(std::move(l));
  }
}

// ---- dispatch :692-966 ---------------------------------------------------------------------------
inline V3 eval_emission(const MaterialPoint& m, V3 normal, V3 outgoing) {  // :575-580
  return dot(normal, outgoing) >= 0.0f ? m.emission : V3{0, 0, 0};
}
inline V3 eval_bsdfcos(const MaterialPoint& m, V3 n, V3 o, V3 i) {
  if (m.roughness == 0.0f) return V3{0, 0, 0};
  switch (m.type) {
    case matte: return eval_matte(m.color, n, o, i);
    case glossy: return eval_glossy(m.color, m.ior, m.roughness, n, o, i);
    case reflective: return eval_reflective(m.color, m.roughness, n, o, i);
    case transparent: return eval_transparent(m.color, m.ior, m.roughness, n, o, i);
    case refractive: return eval_refractive(m.color, m.ior, m.roughness, n, o, i);
    case subsurface: return eval_refractive(m.color, m.ior, m.roughness, n, o, i);
    default: return V3{0, 0, 0};  // gltfpbr throws in the reference
  }
}
inline V3 eval_delta(const MaterialPoint& m, V3 n, V3 o, V3 i) {
  if (m.roughness != 0.0f) return V3{0, 0, 0};
  switch (m.type) {
    case reflective: return eval_reflective_delta(m.color, n, o, i);
    case transparent: return eval_transparent_delta(m.color, m.ior, n, o, i);
    case refractive: return eval_refractive_delta(m.color, m.ior, n, o, i);
    case volumetric: return eval_passthrough(m.color, n, o, i);
    default: return V3{0, 0, 0};
  }
}
inline V3 sample_bsdfcos(const MaterialPoint& m, V3 n, V3 o, float rnl, V2 rn) {
  if (m.roughness == 0.0f) return V3{0, 0, 0};
  switch (m.type) {
    case matte: return sample_matte(m.color, n, o, rn);
    case glossy: return sample_glossy(m.color, m.ior, m.roughness, n, o, rnl, rn);
    case reflective: return sample_reflective(m.color, m.roughness, n, o, rn);
    case transparent: return sample_transparent(m.color, m.ior, m.roughness, n, o, rnl, rn);
    case refractive: return sample_refractive(m.color, m.ior, m.roughness, n, o, rnl, rn);
    case subsurface: return sample_refractive(m.color, m.ior, m.roughness, n, o, rnl, rn);
    default: return V3{0, 0, 0};
  }
}
inline V3 sample_delta(const MaterialPoint& m, V3 n, V3 o, float rnl) {
  if (m.roughness != 0.0f) return V3{0, 0, 0};
  switch (m.type) {
    case reflective: return sample_reflective_delta(m.color, n, o);
    case transparent: return sample_transparent_delta(m.color, m.ior, n, o, rnl);
    case refractive: return sample_refractive_delta(m.color, m.ior, n, o, rnl);
    case volumetric: return sample_passthrough(m.color, n, o);
    default: return V3{0, 0, 0};
  }
}
inline float sample_bsdfcos_pdf(const MaterialPoint& m, V3 n, V3 o, V3 i) {
  if (m.roughness == 0.0f) return 0.0f;
  switch (m.type) {
    case matte: return sample_matte_pdf(m.color, n, o, i);
    case glossy: return sample_glossy_pdf(m.color, m.ior, m.roughness, n, o, i);
    case reflective: return sample_reflective_pdf(m.color, m.roughness, n, o, i);
    case transparent: return sample_transparent_pdf(m.color, m.ior, m.roughness, n, o, i);
    case refractive: return sample_refractive_pdf(m.color, m.ior, m.roughness, n, o, i);
    case subsurface: return sample_refractive_pdf(m.color, m.ior, m.roughness, n, o, i);
    default: return 0.0f;
  }
}
inline float sample_delta_pdf(const MaterialPoint& m, V3 n, V3 o, V3 i) {
  if (m.roughness != 0.0f) return 0.0f;
  switch (m.type) {
    case reflective: return sample_reflective_delta_pdf(m.color, n, o, i);
    case transparent: return sample_transparent_delta_pdf(m.color, m.ior, n, o, i);
    case refractive: return sample_refractive_delta_pdf(m.color, m.ior, n, o, i);
    case volumetric: return sample_passthrough_pdf(m.color, n, o, i);
    default: return 0.0f;
  }
}

// ---- volume scattering :1086-1115 --------------------------------------------------------------
inline V3 eval_scattering(const MaterialPoint& m, V3 outgoing, V3 incoming) {
  if (m.density == V3{0, 0, 0}) return V3{0, 0, 0};
  return (m.scattering * m.density) * eval_phasefunction(m.scanisotropy, outgoing, incoming);
}
inline V3 sample_scattering(const MaterialPoint& m, V3 outgoing, float /*rnl*/, V2 rn) {
  if (m.density == V3{0, 0, 0}) return V3{0, 0, 0};
  return sample_phasefunction(m.scanisotropy, outgoing, rn);
}
inline float sample_scattering_pdf(const MaterialPoint& m, V3 outgoing, V3 incoming) {
  if (m.density == V3{0, 0, 0}) return 0.0f;
  return sample_phasefunction_pdf(m.scanisotropy, outgoing, incoming);
}

// ---- lights :968-1084 ----------------------------------------------------------------------------
inline V3 sample_lights(const Scene& scene, V3 position, float rl, float rel, V2 ruv) {
  int64_t light_id = sample_uniform((int64_t)scene.lights.size(), rl);
  const Light& light = scene.lights[light_id - 1];
  if (light.instance != invalid_id) {
    const Instance& inst = scene.instances[light.instance - 1];
    const Shape& s = scene.shapes[inst.shape - 1];
    int64_t element = sample_discrete(light.cdf.data(), (int64_t)light.cdf.size(), rel);
    V2 uv = s.ntri() != 0 ? sample_triangle(ruv) : ruv;
    V3 lposition = eval_position(scene, inst, element, uv);
    return normalize(lposition - position);
  } else if (light.environment != invalid_id) {
    const Environment& env = scene.environments[light.environment - 1];
    if (env.emission_tex != invalid_id) {
      const Texture& tex = scene.textures[env.emission_tex - 1];
      int64_t idx = sample_discrete(light.cdf.data(), (int64_t)light.cdf.size(), rel);
      // Q7: 1-based idx % width, and Int/Int -> Float64 for the row
      float u = ((float)(idx % tex.width) + 0.5f) / (float)tex.width;
      float v = (float)((((double)idx / (double)tex.width) + (double)0.5f) / (double)tex.height);
      float up = (u * 2.0f) * pif, vp = v * pif;
      return transform_direction(env.frame, V3{jt_cosf(up) * jt_sinf(vp), jt_cosf(vp), jt_sinf(up) * jt_sinf(vp)});
    }
    return V3{0, 0, 0};  // sample_sphere is undefined in the reference
  }
  return V3{0, 0, 0};
}

inline float sample_lights_pdf(const Scene& scene, V3 position, V3 direction, Counters* cnt) {
  float pdf = 0.0f;
  for (const Light& light : scene.lights) {
    if (light.instance != invalid_id) {
      const Instance& inst = scene.instances[light.instance - 1];
      float lpdf = 0.0f;
      V3 next_position = position;
      for (int bounce = 0; bounce <= 99; bounce++) {
        SceneIsec isec = intersect_instance_bvh(scene, light.instance, make_ray(next_position, direction), false, cnt);
        if (!isec.hit) break;
        V3 lposition = eval_position(scene, inst, isec.element, isec.uv);
        V3 lnormal = eval_element_normal(scene, inst, isec.element);
        float area = light.cdf.back();
        lpdf += distance_squared(lposition, position) / (fabsf(dot(lnormal, direction)) * area);
        next_position = lposition + direction * 0.001f;
      }
      pdf += lpdf;
    } else if (light.environment != invalid_id) {
      const Environment& env = scene.environments[light.environment - 1];
      if (env.emission_tex != invalid_id) {
        const Texture& tex = scene.textures[env.emission_tex - 1];
        V3 wl = transform_direction(inverse(env.frame, false), direction);
        V2 texcoord{jt_atan2f(wl.z, wl.x) / (2.0f * pif), jt_acosf(jclamp(wl.y, -1.0f, 1.0f)) / pif};
        if (texcoord.x < 0.0f) texcoord.x = texcoord.x + 1.0f;
        int64_t i = jclampi((int64_t)(texcoord.x * (float)tex.width), 0, tex.width - 1);
        int64_t j = jclampi((int64_t)(texcoord.y * (float)tex.height), 0, tex.height - 1);
        float prob = sample_discrete_pdf(light.cdf.data(), j * tex.width + i + 1) / light.cdf.back();
        float angle = (((2.0f * pif) / (float)tex.width) * (pif / (float)tex.height)) *
                      jt_sinf((pif * ((float)j + 0.5f)) / (float)tex.height);
        pdf += prob / angle;
      } else {
        pdf += 1.0f / (4.0f * pif);
      }
    }
  }
  pdf *= sample_uniform_pdf((int64_t)scene.lights.size());
  return pdf;
}

// ---- trace_path :276-469 ---------------------------------------------------------------------------
struct TraceResult { V3 radiance; bool hit; V3 albedo, normal; };

inline TraceResult trace_path(const Scene& scene, Ray ray, const Params& params, Rng& rng, Counters* cnt) {
  V3 radiance{0, 0, 0}, weight{1, 1, 1};
  int cur_volume = 0;
  MaterialPoint volume_stack[1];  // Q14: depth never exceeds 1
  float max_roughness = 0.0f;
  bool hit = false;
  V3 hit_albedo{0, 0, 0}, hit_normal{0, 0, 0};
  int opbounce = 0;
  int bounce = -1;
  while (bounce < params.bounces) {
    bounce += 1;
    SceneIsec isec = intersect_scene_bvh(scene, ray, false, cnt);
    if (!isec.hit) {
      if (bounce > 0 || !params.envhidden) radiance = radiance + weight * eval_environment(scene, ray.d);
      break;
    }
    bool in_volume = false;
    if (cur_volume != 0) {
      const MaterialPoint& vsdf = volume_stack[cur_volume - 1];
      float r1 = rng.next();
      float r2 = rng.next();
      float distance = sample_transmittance(vsdf.density, isec.distance, r1, r2);
      weight = (weight * eval_transmittance(vsdf.density, distance)) /
               sample_transmittance_pdf(vsdf.density, distance, isec.distance);
      in_volume = distance < isec.distance;
      isec.distance = distance;
    }
    if (!in_volume) {
      V3 outgoing = -ray.d;
      const Instance& inst = scene.instances[isec.instance - 1];
      V3 position = eval_shading_position(scene, inst, isec.element, isec.uv, outgoing);
      V3 normal = eval_shading_normal(scene, inst, isec.element, isec.uv, outgoing);
      MaterialPoint material = eval_material(scene, inst, isec.element, isec.uv);
      if (params.nocaustics) {
        max_roughness = jmax(material.roughness, max_roughness);
        material.roughness = max_roughness;
      }
      if (material.opacity < 1.0f && rng.next() >= material.opacity) {
        if (opbounce > 128) break;
        opbounce += 1;
        ray = make_ray(position + ray.d * 0.01f, ray.d);
        bounce -= 1;
        continue;
      }
      if (bounce == 0) {
        hit = true;
        hit_albedo = material.color;
        hit_normal = normal;
      }
      radiance = radiance + weight * eval_emission(material, normal, outgoing);
      V3 incoming{0, 0, 0};
      if (!is_delta(material)) {
        if (rng.next() < 0.5f) {
          float rnl = rng.next();
          V2 rn = rng.next2();
          incoming = sample_bsdfcos(material, normal, outgoing, rnl, rn);
        } else {
          float rl = rng.next();
          float rel = rng.next();
          V2 ruv = rng.next2();
          incoming = sample_lights(scene, position, rl, rel, ruv);
        }
        if (incoming == V3{0, 0, 0}) break;
        weight = (weight * eval_bsdfcos(material, normal, outgoing, incoming)) /
                 (0.5f * sample_bsdfcos_pdf(material, normal, outgoing, incoming) +
                  0.5f * sample_lights_pdf(scene, position, incoming, cnt));
      } else {
        incoming = sample_delta(material, normal, outgoing, rng.next());
        weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
                 sample_delta_pdf(material, normal, outgoing, incoming);
      }
      if (is_volumetric(scene, inst) && dot(normal, outgoing) * dot(normal, incoming) < 0.0f) {
        if (cur_volume == 0) {
          material = eval_material(scene, inst, isec.element, isec.uv);
          cur_volume += 1;
          volume_stack[cur_volume - 1] = material;
        } else {
          cur_volume -= 1;
        }
      }
      ray = make_ray(position, incoming);
    } else {
      V3 outgoing = -ray.d;
      V3 position = ray.o + ray.d * isec.distance;
      const MaterialPoint& vsdf = volume_stack[cur_volume - 1];
      V3 incoming{0, 0, 0};
      if (rng.next() < 0.5f) {
        float rnl = rng.next();
        V2 rn = rng.next2();
        incoming = sample_scattering(vsdf, outgoing, rnl, rn);
      } else {
        float rl = rng.next();
        float rel = rng.next();
        V2 ruv = rng.next2();
        incoming = sample_lights(scene, position, rl, rel, ruv);
      }
      if (incoming == V3{0, 0, 0}) break;
      weight = (weight * eval_scattering(vsdf, outgoing, incoming)) /
               (0.5f * sample_scattering_pdf(vsdf, outgoing, incoming) +
                0.5f * sample_lights_pdf(scene, position, incoming, cnt));
      ray = make_ray(position, incoming);
    }
    if (weight == V3{0, 0, 0} || !all_finite(weight)) break;
    if (bounce > 3) {
      float rr_prob = jmin(0.99f, maximum(weight));
      if (rng.next() >= rr_prob) break;
      weight = weight * (1.0f / rr_prob);
    }
  }
  return TraceResult{radiance, hit, hit_albedo, hit_normal};
}

// ---- trace_naive :471-573 ---------------------------------------------------------------------------
inline TraceResult trace_naive(const Scene& scene, Ray ray, const Params& params, Rng& rng, Counters* cnt) {
  V3 radiance{0, 0, 0}, weight{1, 1, 1};
  bool hit = false;
  V3 hit_albedo{0, 0, 0}, hit_normal{0, 0, 0};
  int opbounce = 0;
  int bounce = -1;
  while (bounce < params.bounces) {
    bounce += 1;
    SceneIsec isec = intersect_scene_bvh(scene, ray, false, cnt);
    if (!isec.hit) {
      if (bounce > 0 || !params.envhidden) radiance = radiance + weight * eval_environment(scene, ray.d);
      break;
    }
    V3 outgoing = -ray.d;
    const Instance& inst = scene.instances[isec.instance - 1];
    V3 position = eval_shading_position(scene, inst, isec.element, isec.uv, outgoing);
    V3 normal = eval_shading_normal(scene, inst, isec.element, isec.uv, outgoing);
    MaterialPoint material = eval_material(scene, inst, isec.element, isec.uv);
    if (material.opacity < 1.0f && rng.next() >= material.opacity) {
      if (opbounce > 128) break;
      opbounce += 1;
      ray = make_ray(position + ray.d * 0.01f, ray.d);
      bounce -= 1;
      continue;
    }
    if (bounce == 0) {
      hit = true;
      hit_albedo = material.color;
      hit_normal = normal;
    }
    radiance = radiance + weight * eval_emission(material, normal, outgoing);
    V3 incoming{0, 0, 0};
    if (material.roughness != 0.0f) {
      float rnl = rng.next();
      V2 rn = rng.next2();
      incoming = sample_bsdfcos(material, normal, outgoing, rnl, rn);
      if (incoming == V3{0, 0, 0}) break;
      weight = (weight * eval_bsdfcos(material, normal, outgoing, incoming)) /
               sample_bsdfcos_pdf(material, normal, outgoing, incoming);
    } else {
      incoming = sample_delta(material, normal, outgoing, rng.next());
      if (incoming == V3{0, 0, 0}) break;
      weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
               sample_delta_pdf(material, normal, outgoing, incoming);
    }
    if (weight == V3{0, 0, 0} || !all_finite(weight)) break;
    if (bounce > 3) {
      float rr_prob = jmin(0.99f, maximum(weight));
      if (rng.next() >= rr_prob) break;
      weight = weight * (1.0f / rr_prob);
    }
    ray = make_ray(position, incoming);
  }
  return TraceResult{radiance, hit, hit_albedo, hit_normal};
}

// ---- sample_camera :651-674 ----------------------------------------------------------------------------
inline Ray sample_camera(const Camera& camera, int64_t i, int64_t j, int64_t w, int64_t h, V2 puv, V2 luv,
                         bool tent) {
  if (!tent) {
    V2 uv{((float)i + puv.x) / (float)w, ((float)j + puv.y) / (float)h};
    return eval_camera(camera, uv, sample_disk(luv));
  }
  const float width = 2.0f, offset = 0.5f;
  float fx = puv.x < 0.5f ? sqrtf(2.0f * puv.x) - 1.0f : 1.0f - sqrtf(2.0f - 2.0f * puv.x);
  float fy = puv.y < 0.5f ? sqrtf(2.0f * puv.y) - 1.0f : 1.0f - sqrtf(2.0f - 2.0f * puv.y);
  V2 fuv{width * fx + offset, width * fy + offset};
  V2 uv{((float)i + fuv.x) / (float)w, ((float)j + fuv.y) / (float)h};
  return eval_camera(camera, uv, sample_disk(luv));
}

// ---- trace_sample :584-649 ---------------------------------------------------------------------------------
inline void trace_sample(State& state, const Scene& scene, int64_t i, int64_t j, int64_t sample,
                         const Params& params, Counters* cnt) {
  const Camera& camera = scene.cameras[params.camera - 1];
  int64_t idx = state.width * j + i;  // 0-based here
  Rng rng{jt_rng_key(params.seed, (uint32_t)idx, (uint32_t)sample), 0};
  V2 puv = rng.next2();
  V2 luv = rng.next2();
  Ray ray = sample_camera(camera, i, j, state.width, state.height, puv, luv, params.tentfilter);
  if (cnt) cnt->camera_paths++;
  TraceResult r;
  if (params.sampler == 1) r = trace_path(scene, ray, params, rng, cnt);
  else r = trace_naive(scene, ray, params, rng, cnt);
  V3 radiance = r.radiance;
  if (!all_finite(radiance)) radiance = V3{0, 0, 0};
  if (maximum(radiance) > (float)params.clamp) radiance = radiance * ((float)params.clamp / maximum(radiance));
  bool has_env = !params.envhidden && !scene.environments.empty();
  if (params.accumulate == 0) {
    float weight = 1.0f / (float)(sample + 1);
    if (r.hit) {
      state.image[idx] = lerp(state.image[idx], V4{radiance.x, radiance.y, radiance.z, 1.0f}, weight);
      state.albedo[idx] = lerp(state.albedo[idx], r.albedo, weight);
      state.normal[idx] = lerp(state.normal[idx], r.normal, weight);
      state.hits[idx] += 1;
    } else if (has_env) {
      state.image[idx] = lerp(state.image[idx], V4{radiance.x, radiance.y, radiance.z, 1.0f}, weight);
      state.albedo[idx] = lerp(state.albedo[idx], V3{1, 1, 1}, weight);
      state.normal[idx] = lerp(state.normal[idx], -ray.d, weight);
      state.hits[idx] += 1;
    } else {
      state.image[idx] = lerp(state.image[idx], V4{0, 0, 0, 0}, weight);
      state.albedo[idx] = lerp(state.albedo[idx], V3{0, 0, 0}, weight);
      state.normal[idx] = lerp(state.normal[idx], -ray.d, weight);
    }
  } else {
    // sum mode (what multi-GPU sharding reduces): same three cases, added
    if (r.hit) {
      state.image[idx] = state.image[idx] + V4{radiance.x, radiance.y, radiance.z, 1.0f};
      state.albedo[idx] = state.albedo[idx] + r.albedo;
      state.normal[idx] = state.normal[idx] + r.normal;
      state.hits[idx] += 1;
    } else if (has_env) {
      state.image[idx] = state.image[idx] + V4{radiance.x, radiance.y, radiance.z, 1.0f};
      state.albedo[idx] = state.albedo[idx] + V3{1, 1, 1};
      state.normal[idx] = state.normal[idx] + (-ray.d);
      state.hits[idx] += 1;
    } else {
      state.normal[idx] = state.normal[idx] + (-ray.d);
    }
  }
}

}  // namespace orc
