// jt_dev_trace.cuh -- the integrators: trace_path / trace_naive / trace_sample and light sampling.
//
// Device implementation of src/trace.jl:276-649 (trace_path, trace_naive, eval_emission,
// trace_sample) and :968-1084 (sample_lights, sample_lights_pdf), plus src/sampling.jl:29-58.
// Random numbers come from the per-(pixel, sample) counter stream of jt_rng.h, consumed in the
// reference's draw order.
#pragma once
#include "jt_dev_shade.cuh"
#include "jt_dev_traverse.cuh"

struct DevParams {
  int camera;  // 0-based
  int width, height;
  int bounces, sampler, clamp;
  int nocaustics, envhidden, tentfilter;
  int accumulate;
  unsigned long long seed;
};

struct DevState {   // TraceState, src/trace.jl:87-96 (device-resident accumulators)
  float4* image;    // RGBA
  float4* albedo;   // xyz (+pad)
  float4* normal;   // xyz (+pad)
  int* hits;
};

struct PathCounters {
  unsigned int scene_rays, light_rays;
};

struct Rng {
  unsigned long long key;
  unsigned int draw;
  JT_DEV float next() { return jt_rng_float(key, draw++); }
  JT_DEV f2 next2() {
    float a = next();
    float b = next();
    return f2{a, b};
  }
};

// ---- sampling.jl --------------------------------------------------------------------------------
JT_DEV int sample_uniform(int size, float r) { return jl_clampi((int)(r * (float)size) + 1, 1, size); }  // 1-based
JT_DEV float sample_uniform_pdf(int size) { return (float)(1.0 / (double)size); }
JT_DEV int upper_bound(const float* cdf, int n, float limit) {  // src/sampling.jl:42-56, 1-based
  int idx = 0, l = 1, r = n;
  while (l <= r) {
    int m = (l + r) / 2;
    if (__ldg(cdf + m - 1) > limit) {
      idx = m;
      r = m - 1;
    } else {
      l = m + 1;
    }
  }
  return idx;
}
JT_DEV int sample_discrete(const float* cdf, int n, float r) {  // :33-37, 1-based
  float last = __ldg(cdf + n - 1);
  r = jl_clamp(r * last, 0.0f, last - 0.00001f);
  return jl_clampi(upper_bound(cdf, n, r), 1, n);
}
// sample_discrete of a light's CDF: the same bisection, started from the bracket the light's guide table gives
// for the search key (jt_stage.cpp build_cdf_guide proves the answer lies inside it), so the returned index is
// the reference's; 17-21 dependent loads become 2 + a few.
JT_DEV int sample_discrete_light(const JtDevScene& S, const JtLightRec& L, float r) {
  const float* cdf = S.light_cdf + L.cdf_off;
  const int n = L.cdf_len;
  if (L.guide_len == 0) return sample_discrete(cdf, n, r);
  float last = __ldg(cdf + n - 1);
  float limit = jl_clamp(r * last, 0.0f, last - 0.00001f);
  int b = (int)(limit * L.guide_scale);
  b = b < 0 ? 0 : (b > L.guide_len - 1 ? L.guide_len - 1 : b);
  const int32_t* G = S.light_guide + L.guide_off;
  int lo = __ldg(G + b) + 1, hi = __ldg(G + b + 1) + 1;
  if (hi > n) hi = n;
  int idx = 0;
  while (lo <= hi) {
    int m = (lo + hi) / 2;
    if (__ldg(cdf + m - 1) > limit) {
      idx = m;
      hi = m - 1;
    } else {
      lo = m + 1;
    }
  }
  return jl_clampi(idx, 1, n);
}
JT_DEV float sample_discrete_pdf(const float* cdf, int idx) {  // :39-40, idx 1-based
  return idx == 1 ? __ldg(cdf) : __ldg(cdf + idx - 1) - __ldg(cdf + idx - 2);
}

// ---- sample_lights, src/trace.jl:968-1008 ----------------------------------------------------------
JT_DEV f3 sample_lights(const JtDevScene& S, f3 position, float rl, float rel, f2 ruv) {
  int light_id = sample_uniform(S.num_lights, rl);
  const JtLightRec L = S.lights[light_id - 1];
  const float* cdf = S.light_cdf + L.cdf_off;
  if (L.instance >= 0) {
    const JtInstanceRec& I = S.instances[L.instance];
    int element = sample_discrete_light(S, L, rel) - 1;
    ElemRef E = elem_ref(S, I, element);
    f2 uv = ruv;
    if (E.sh->kind == 1) {  // sample_triangle, src/sampling.jl:58
      float sq = sqrtf(ruv.x);
      uv = f2{1.0f - sq, ruv.y * sq};
    }
    f3 lposition = eval_position(S, I, E, uv.x, uv.y);
    return normalize3(lposition - position);
  } else if (L.environment >= 0) {
    const JtEnvRec& En = S.environments[L.environment];
    if (En.emission_tex >= 0) {
      const JtTextureRec T = S.textures[En.emission_tex];
      int idx = sample_discrete_light(S, L, rel);  // 1-based, used as is (Q7)
      float u = ((float)(idx % T.width) + 0.5f) / (float)T.width;
      float v = (float)((((double)idx / (double)T.width) + (double)0.5f) / (double)T.height);
      float up = (u * 2.0f) * JT_PIF, vp = v * JT_PIF;
      return xform_direction(En.frame, f3{jt_cosf(up) * jt_sinf(vp), jt_cosf(vp), jt_sinf(up) * jt_sinf(vp)});
    }
    return f3{0.0f, 0.0f, 0.0f};  // sample_sphere is undefined in the reference
  }
  return f3{0.0f, 0.0f, 0.0f};
}

// Conservative ray / padded-box overlap for t >= 0 (never rejects a ray the instance-space walk could hit).
JT_DEV bool ray_reaches_box(f3 o, f3 d, float4 lo, float4 hi) {
  float ix = guarded_rcp(d.x), iy = guarded_rcp(d.y), iz = guarded_rcp(d.z);
  float ax = (lo.x - o.x) * ix, bx = (hi.x - o.x) * ix;
  float ay = (lo.y - o.y) * iy, by = (hi.y - o.y) * iy;
  float az = (lo.z - o.z) * iz, bz = (hi.z - o.z) * iz;
  float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
  float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
  return t0 <= t1 * 1.00001f + 1e-30f;
}

// ---- sample_lights_pdf, src/trace.jl:1010-1084 --------------------------------------------------------
// WALK = false is the variant the wide-mode shade kernel runs inline: it evaluates the sum as long as no area
// light can be reached by the ray (the box early-out below makes those terms exact zeros) and reports
// *need_walk = true (result unused) as soon as one can -- the slot then goes to the probe kernel, which runs
// the full WALK = true version. Same expression either way.
template <int MODE, bool WALK>
JT_DEV float sample_lights_pdf_impl(const JtDevScene& S, f3 position, f3 direction, PathCounters& cnt, bool* need_walk) {
  float pdf = 0.0f;
  unsigned probes = 0u;
  for (int li = 0; li < S.num_lights; li++) {
    const JtLightRec L = S.lights[li];
    const float* cdf = S.light_cdf + L.cdf_off;
    if (L.instance >= 0) {
      const JtInstanceRec& I = S.instances[L.instance];
      float lpdf = 0.0f;
      f3 next_position = position;
      float area = WALK ? __ldg(cdf + L.cdf_len - 1) : 0.0f;
      for (int bounce = 0; bounce < 100; bounce++) {
        probes++;
        // wide mode: skip the BLAS walk when the ray cannot reach the light's padded world box
        if (MODE == MODE_WIDE &&
            !ray_reaches_box(next_position, direction, __ldg(S.inst_bounds + 2 * L.instance), __ldg(S.inst_bounds + 2 * L.instance + 1)))
          break;
        if (!WALK) {
          *need_walk = true;
          return 0.0f;
        }
        DHit h = intersect_instance<MODE>(S, L.instance, DRay{next_position, direction, JT_RAY_EPS, INFINITY});
        if (h.inst < 0) break;
        ElemRef E = elem_ref(S, I, h.elem);
        f3 lposition = eval_position(S, I, E, h.u, h.v);
        f3 lnormal = eval_element_normal(S, I, E);
        f3 dp = lposition - position;
        lpdf += dot3(dp, dp) / (fabsf(dot3(lnormal, direction)) * area);
        next_position = lposition + direction * 0.001f;
      }
      pdf += lpdf;
    } else if (L.environment >= 0) {
      const JtEnvRec& En = S.environments[L.environment];
      if (En.emission_tex >= 0) {
        const JtTextureRec T = S.textures[En.emission_tex];
        f3 wl = normalize3(xform_vector_transposed(En.frame, direction));
        f2 tc = f2{jt_atan2f(wl.z, wl.x) / (2.0f * JT_PIF), jt_acosf(jl_clamp(wl.y, -1.0f, 1.0f)) / JT_PIF};
        if (tc.x < 0.0f) tc.x = tc.x + 1.0f;
        int i = jl_clampi((int)(tc.x * (float)T.width), 0, T.width - 1);
        int j = jl_clampi((int)(tc.y * (float)T.height), 0, T.height - 1);
        float prob = sample_discrete_pdf(cdf, j * T.width + i + 1) / __ldg(cdf + L.cdf_len - 1);
        float angle = (((2.0f * JT_PIF) / (float)T.width) * (JT_PIF / (float)T.height)) *
                      jt_sinf((JT_PIF * ((float)j + 0.5f)) / (float)T.height);
        pdf += prob / angle;
      } else {
        pdf += 1.0f / (4.0f * JT_PIF);
      }
    }
  }
  pdf *= sample_uniform_pdf(S.num_lights);
  cnt.light_rays += probes;
  return pdf;
}
template <int MODE>
JT_DEV float sample_lights_pdf(const JtDevScene& S, f3 position, f3 direction, PathCounters& cnt) {
  bool unused = false;
  return sample_lights_pdf_impl<MODE, true>(S, position, direction, cnt, &unused);
}

struct TraceOut {
  f3 radiance;
  bool hit;
  f3 albedo, normal;
};

// ---- trace_path, src/trace.jl:276-469 ----------------------------------------------------------------
template <int MODE>
JT_DEV TraceOut trace_path(const JtDevScene& S, DRay ray, const DevParams& P, Rng& rng, PathCounters& cnt) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f};
  f3 radiance = zero, weight = f3{1.0f, 1.0f, 1.0f};
  bool in_medium = false;  // cur_volume != 0 (Q14: the stack never holds more than one entry)
  VolPoint medium;
  medium.density = zero; medium.scattering = zero; medium.scanisotropy = 0.0f;
  float max_roughness = 0.0f;
  TraceOut out;
  out.hit = false; out.albedo = zero; out.normal = zero;
  int opbounce = 0;
  int bounce = -1;
  while (bounce < P.bounces) {
    bounce += 1;
    cnt.scene_rays++;
    DHit isec = intersect_scene<MODE>(S, ray);
    if (isec.inst < 0) {
      if (bounce > 0 || !P.envhidden) radiance = radiance + weight * eval_environment(S, ray.d);
      break;
    }
    bool in_volume = false;
    float distance = isec.t;
    if (in_medium) {
      float r1 = rng.next();
      float r2 = rng.next();
      float dist = sample_transmittance(medium.density, isec.t, r1, r2);
      weight = (weight * eval_transmittance(medium.density, dist)) / sample_transmittance_pdf(medium.density, dist, isec.t);
      in_volume = dist < isec.t;
      distance = dist;
    }
    f3 outgoing = -ray.d;
    if (!in_volume) {
      const JtInstanceRec& I = S.instances[isec.inst];
      const JtMaterialRec& M = S.materials[I.material];
      ElemRef E = elem_ref(S, I, isec.elem);
      f3 position = eval_position(S, I, E, isec.u, isec.v);
      f3 normal = eval_shading_normal(S, I, E, M, isec.u, isec.v, outgoing);
      MatPoint material = eval_material(S, E, M, isec.u, isec.v);
      if (P.nocaustics) {
        max_roughness = jl_max(material.roughness, max_roughness);
        material.roughness = max_roughness;
      }
      if (material.opacity < 1.0f && rng.next() >= material.opacity) {
        if (opbounce > 128) break;
        opbounce += 1;
        ray = DRay{position + ray.d * 0.01f, ray.d, JT_RAY_EPS, INFINITY};
        bounce -= 1;
        continue;
      }
      if (bounce == 0) {
        out.hit = true;
        out.albedo = material.color;
        out.normal = normal;
      }
      if (dot3(normal, outgoing) >= 0.0f) radiance = radiance + weight * material.emission;  // eval_emission
      else radiance = radiance + weight * zero;
      f3 incoming;
      if (!is_delta(material)) {
        if (rng.next() < 0.5f) {
          float rnl = rng.next();
          f2 rn = rng.next2();
          incoming = sample_bsdfcos(material, normal, outgoing, rnl, rn);
        } else {
          float rl = rng.next();
          float rel = rng.next();
          f2 ruv = rng.next2();
          incoming = sample_lights(S, position, rl, rel, ruv);
        }
        if (is_zero3(incoming)) break;
        float pb = sample_bsdfcos_pdf(material, normal, outgoing, incoming);
        float pl = sample_lights_pdf<MODE>(S, position, incoming, cnt);
        weight = (weight * eval_bsdfcos(material, normal, outgoing, incoming)) / (0.5f * pb + 0.5f * pl);
      } else {
        incoming = sample_delta(material, normal, outgoing, rng.next());
        weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
                 sample_delta_pdf(material, normal, outgoing, incoming);
      }
      if (is_volumetric_type(M.type) && dot3(normal, outgoing) * dot3(normal, incoming) < 0.0f) {
        if (!in_medium) {
          // the reference re-evaluates the material here (src/trace.jl:410-415); the volume code only
          // ever reads density / scattering / scanisotropy, which the nocaustics ratchet never touches
          medium.density = material.density;
          medium.scattering = material.scattering;
          medium.scanisotropy = material.scanisotropy;
          in_medium = true;
        } else {
          in_medium = false;
        }
      }
      ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
    } else {
      f3 position = ray.o + ray.d * distance;
      f3 incoming;
      if (rng.next() < 0.5f) {
        float rnl = rng.next();
        (void)rnl;
        f2 rn = rng.next2();
        incoming = sample_scattering(medium, outgoing, rn);
      } else {
        float rl = rng.next();
        float rel = rng.next();
        f2 ruv = rng.next2();
        incoming = sample_lights(S, position, rl, rel, ruv);
      }
      if (is_zero3(incoming)) break;
      float ps = sample_scattering_pdf(medium, outgoing, incoming);
      float pl = sample_lights_pdf<MODE>(S, position, incoming, cnt);
      weight = (weight * eval_scattering(medium, outgoing, incoming)) / (0.5f * ps + 0.5f * pl);
      ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
    }
    if (is_zero3(weight) || !finite3(weight)) break;
    if (bounce > 3) {
      float rr_prob = jl_min(0.99f, max3(weight));
      if (rng.next() >= rr_prob) break;
      weight = weight * (1.0f / rr_prob);
    }
  }
  out.radiance = radiance;
  return out;
}

// ---- trace_naive, src/trace.jl:471-573 -----------------------------------------------------------------
template <int MODE>
JT_DEV TraceOut trace_naive(const JtDevScene& S, DRay ray, const DevParams& P, Rng& rng, PathCounters& cnt) {
  const f3 zero = f3{0.0f, 0.0f, 0.0f};
  f3 radiance = zero, weight = f3{1.0f, 1.0f, 1.0f};
  TraceOut out;
  out.hit = false; out.albedo = zero; out.normal = zero;
  int opbounce = 0;
  int bounce = -1;
  while (bounce < P.bounces) {
    bounce += 1;
    cnt.scene_rays++;
    DHit isec = intersect_scene<MODE>(S, ray);
    if (isec.inst < 0) {
      if (bounce > 0 || !P.envhidden) radiance = radiance + weight * eval_environment(S, ray.d);
      break;
    }
    f3 outgoing = -ray.d;
    const JtInstanceRec& I = S.instances[isec.inst];
    const JtMaterialRec& M = S.materials[I.material];
    ElemRef E = elem_ref(S, I, isec.elem);
    f3 position = eval_position(S, I, E, isec.u, isec.v);
    f3 normal = eval_shading_normal(S, I, E, M, isec.u, isec.v, outgoing);
    MatPoint material = eval_material(S, E, M, isec.u, isec.v);
    if (material.opacity < 1.0f && rng.next() >= material.opacity) {
      if (opbounce > 128) break;
      opbounce += 1;
      ray = DRay{position + ray.d * 0.01f, ray.d, JT_RAY_EPS, INFINITY};
      bounce -= 1;
      continue;
    }
    if (bounce == 0) {
      out.hit = true;
      out.albedo = material.color;
      out.normal = normal;
    }
    if (dot3(normal, outgoing) >= 0.0f) radiance = radiance + weight * material.emission;
    else radiance = radiance + weight * zero;
    f3 incoming;
    if (material.roughness != 0.0f) {
      float rnl = rng.next();
      f2 rn = rng.next2();
      incoming = sample_bsdfcos(material, normal, outgoing, rnl, rn);
      if (is_zero3(incoming)) break;
      weight = (weight * eval_bsdfcos(material, normal, outgoing, incoming)) /
               sample_bsdfcos_pdf(material, normal, outgoing, incoming);
    } else {
      incoming = sample_delta(material, normal, outgoing, rng.next());
      if (is_zero3(incoming)) break;
      weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
               sample_delta_pdf(material, normal, outgoing, incoming);
    }
    if (is_zero3(weight) || !finite3(weight)) break;
    if (bounce > 3) {
      float rr_prob = jl_min(0.99f, max3(weight));
      if (rng.next() >= rr_prob) break;
      weight = weight * (1.0f / rr_prob);
    }
    ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
  }
  out.radiance = radiance;
  return out;
}

// ---- accumulation: tail of trace_sample, src/trace.jl:625-648 -------------------------------------------
JT_DEV float4 lerp4(float4 a, float4 b, float u) {
  float w = 1.0f - u;
  return make_float4(a.x * w + b.x * u, a.y * w + b.y * u, a.z * w + b.z * u, a.w * w + b.w * u);
}
// `old_*`: the pixel's accumulators as loaded by the caller (k_wf_regen issues these loads together with the slot's
// other loads, before it knows whether it is the pixel's turn).
JT_DEV void accumulate_loaded(const DevState& st, const DevParams& P, bool has_env, int idx, int sample, const TraceOut& r,
                              f3 ray_d, float4 old_img, float4 old_alb, float4 old_nrm, int old_hits) {
  f3 radiance = r.radiance;
  if (!finite3(radiance)) radiance = f3{0.0f, 0.0f, 0.0f};
  float mx = max3(radiance);
  if (mx > (float)P.clamp) radiance = radiance * ((float)P.clamp / mx);
  bool env = !P.envhidden && has_env;
  float4 img, alb, nrm;
  bool count;
  if (r.hit) {
    img = make_float4(radiance.x, radiance.y, radiance.z, 1.0f);
    alb = make_float4(r.albedo.x, r.albedo.y, r.albedo.z, 0.0f);
    nrm = make_float4(r.normal.x, r.normal.y, r.normal.z, 0.0f);
    count = true;
  } else if (env) {
    img = make_float4(radiance.x, radiance.y, radiance.z, 1.0f);
    alb = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    nrm = make_float4(-ray_d.x, -ray_d.y, -ray_d.z, 0.0f);
    count = true;
  } else {
    img = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    alb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    nrm = make_float4(-ray_d.x, -ray_d.y, -ray_d.z, 0.0f);
    count = false;
  }
  if (P.accumulate == 0) {  // running mean, Q13
    float w = 1.0f / (float)(sample + 1);
    st.image[idx] = lerp4(old_img, img, w);
    st.albedo[idx] = lerp4(old_alb, alb, w);
    st.normal[idx] = lerp4(old_nrm, nrm, w);
  } else {  // plain sums: what gets reduced across GPUs
    float4 a = old_img, b = old_alb, c = old_nrm;
    st.image[idx] = make_float4(a.x + img.x, a.y + img.y, a.z + img.z, a.w + img.w);
    st.albedo[idx] = make_float4(b.x + alb.x, b.y + alb.y, b.z + alb.z, 0.0f);
    st.normal[idx] = make_float4(c.x + nrm.x, c.y + nrm.y, c.z + nrm.z, 0.0f);
  }
  if (count) st.hits[idx] = old_hits + 1;
}
JT_DEV void accumulate_sample(const DevState& st, const DevParams& P, bool has_env, int idx, int sample,
                              const TraceOut& r, f3 ray_d) {
  accumulate_loaded(st, P, has_env, idx, sample, r, ray_d, st.image[idx], st.albedo[idx], st.normal[idx], st.hits[idx]);
}
