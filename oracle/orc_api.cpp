// ORACLE -- TEST INFRASTRUCTURE ONLY. C entry points (ctypes) of the CPU restatement.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library; the product (libjtrace_b200.so and the Python host package) never does.
//
// Parity status: UNPINNED against the reference executable (see orc_math.h).
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "orc_trace.h"

using namespace orc;

namespace {

Frame to_frame(const jt_frame& f) {
  return Frame{V3{f.x[0], f.x[1], f.x[2]}, V3{f.y[0], f.y[1], f.y[2]}, V3{f.z[0], f.z[1], f.z[2]},
               V3{f.o[0], f.o[1], f.o[2]}};
}
V3 to_v3(const float* p) { return V3{p[0], p[1], p[2]}; }

BvhTree to_tree(const jt_bvh_desc& d) {
  BvhTree t;
  t.nodes.resize(d.num_nodes);
  for (int64_t i = 0; i < d.num_nodes; i++) {
    const jt_bvh_node& n = d.nodes[i];
    t.nodes[i] = BvhNode{Bbox{to_v3(n.bbox_min), to_v3(n.bbox_max)}, n.start, n.num, n.axis, n.internal != 0};
  }
  t.primitives.assign(d.primitives, d.primitives + d.num_primitives);
  return t;
}

struct Oracle {
  Scene scene;
  State state;
  Params params;
  Counters counters;
};

thread_local std::string g_err;

}  // namespace

extern "C" {

#define ORC_API __attribute__((visibility("default")))

ORC_API const char* orc_last_error() { return g_err.c_str(); }

// use_desc_bvh / use_desc_lights: 0 = the oracle builds its own (make_scene_bvh /
// make_trace_lights restatements), 1 = take the ones in the description.
ORC_API void* orc_create(const jt_scene_desc* d, int use_desc_bvh, int use_desc_lights, int high_quality) {
  Oracle* o = new Oracle();
  Scene& s = o->scene;
  for (int64_t i = 0; i < d->num_cameras; i++) {
    const jt_camera& c = d->cameras[i];
    s.cameras.push_back(Camera{to_frame(c.frame), c.orthographic != 0, c.lens, c.film, c.aspect, c.focus, c.aperture});
  }
  for (int64_t i = 0; i < d->num_instances; i++)
    s.instances.push_back(Instance{to_frame(d->instances[i].frame), d->instances[i].shape, d->instances[i].material});
  for (int64_t i = 0; i < d->num_environments; i++)
    s.environments.push_back(Environment{to_frame(d->environments[i].frame), to_v3(d->environments[i].emission),
                                         d->environments[i].emission_tex});
  for (int64_t i = 0; i < d->num_materials; i++) {
    const jt_material& m = d->materials[i];
    s.materials.push_back(Material{m.type, to_v3(m.emission), to_v3(m.color), m.roughness, m.metallic, m.ior,
                                   to_v3(m.scattering), m.scanisotropy, m.trdepth, m.opacity, m.emission_tex,
                                   m.color_tex, m.roughness_tex, m.scattering_tex, m.normal_tex});
  }
  for (int64_t i = 0; i < d->num_textures; i++) {
    const jt_texture_desc& t = d->textures[i];
    Texture x;
    x.width = t.width;
    x.height = t.height;
    x.linear = t.linear != 0;
    int64_t n = t.width * t.height;
    if (t.pixelsf) {
      x.pixelsf.resize(n);
      memcpy(x.pixelsf.data(), t.pixelsf, sizeof(V4) * n);
    } else if (t.pixelsb) {
      x.pixelsb.assign(t.pixelsb, t.pixelsb + 4 * n);
    }
    s.textures.push_back(std::move(x));
  }
  for (int64_t i = 0; i < d->num_shapes; i++) {
    const jt_shape_desc& h = d->shapes[i];
    Shape x;
    x.positions.resize(h.num_positions);
    if (h.num_positions) memcpy(x.positions.data(), h.positions, sizeof(V3) * h.num_positions);
    x.normals.resize(h.num_normals);
    if (h.num_normals) memcpy(x.normals.data(), h.normals, sizeof(V3) * h.num_normals);
    x.texcoords.resize(h.num_texcoords);
    if (h.num_texcoords) memcpy(x.texcoords.data(), h.texcoords, sizeof(V2) * h.num_texcoords);
    x.colors.resize(h.num_colors);
    if (h.num_colors) memcpy(x.colors.data(), h.colors, sizeof(V4) * h.num_colors);
    x.triangles.assign(h.triangles, h.triangles + 3 * h.num_triangles);
    x.quads.assign(h.quads, h.quads + 4 * h.num_quads);
    if (use_desc_bvh) x.bvh = to_tree(h.bvh);
    s.shapes.push_back(std::move(x));
  }
  if (use_desc_bvh) s.bvh = to_tree(d->bvh);
  else make_scene_bvh(s, high_quality != 0);
  if (use_desc_lights) {
    for (int64_t i = 0; i < d->num_lights; i++) {
      Light l{d->lights[i].instance, d->lights[i].environment, {}};
      l.cdf.assign(d->lights[i].elements_cdf, d->lights[i].elements_cdf + d->lights[i].num_elements);
      s.lights.push_back(std::move(l));
    }
  } else {
    make_trace_lights(s);
  }
  return o;
}

ORC_API void orc_destroy(void* h) { delete (Oracle*)h; }

// ---- BVH / lights introspection (to compare the product's host builders with the oracle's) ----
ORC_API int64_t orc_bvh_num_nodes(void* h, int64_t shape /*0 = TLAS, else 1-based shape*/) {
  Oracle* o = (Oracle*)h;
  const BvhTree& t = shape == 0 ? o->scene.bvh : o->scene.shapes[shape - 1].bvh;
  return (int64_t)t.nodes.size();
}
ORC_API int64_t orc_bvh_num_primitives(void* h, int64_t shape) {
  Oracle* o = (Oracle*)h;
  const BvhTree& t = shape == 0 ? o->scene.bvh : o->scene.shapes[shape - 1].bvh;
  return (int64_t)t.primitives.size();
}
ORC_API void orc_bvh_get(void* h, int64_t shape, jt_bvh_node* nodes, int64_t* prims) {
  Oracle* o = (Oracle*)h;
  const BvhTree& t = shape == 0 ? o->scene.bvh : o->scene.shapes[shape - 1].bvh;
  for (size_t i = 0; i < t.nodes.size(); i++) {
    const BvhNode& n = t.nodes[i];
    jt_bvh_node x;
    memset(&x, 0, sizeof(x));
    x.bbox_min[0] = n.bbox.mn.x; x.bbox_min[1] = n.bbox.mn.y; x.bbox_min[2] = n.bbox.mn.z;
    x.bbox_max[0] = n.bbox.mx.x; x.bbox_max[1] = n.bbox.mx.y; x.bbox_max[2] = n.bbox.mx.z;
    x.start = n.start; x.num = n.num; x.axis = n.axis; x.internal = n.internal ? 1 : 0;
    nodes[i] = x;
  }
  memcpy(prims, t.primitives.data(), sizeof(int64_t) * t.primitives.size());
}
ORC_API int64_t orc_num_lights(void* h) { return (int64_t)((Oracle*)h)->scene.lights.size(); }
ORC_API int64_t orc_light_info(void* h, int64_t i, int64_t* instance, int64_t* environment) {
  const Light& l = ((Oracle*)h)->scene.lights[i];
  *instance = l.instance;
  *environment = l.environment;
  return (int64_t)l.cdf.size();
}
ORC_API void orc_light_cdf(void* h, int64_t i, float* out) {
  const Light& l = ((Oracle*)h)->scene.lights[i];
  memcpy(out, l.cdf.data(), sizeof(float) * l.cdf.size());
}

// ---- "identical rays" hooks ----------------------------------------------------------------------
// counters_out (optional): 11 x uint64 {scene_rays, light_rays, camera_paths, tlas_nodes, blas_nodes,
// instance_visits, tri_tests, quad_tests, probe_blas_nodes, probe_tri_tests, probe_quad_tests}
static void export_counters(const Counters& c, uint64_t* out) {
  out[0] = c.scene_rays; out[1] = c.light_rays; out[2] = c.camera_paths; out[3] = c.tlas_nodes;
  out[4] = c.blas_nodes; out[5] = c.instance_visits; out[6] = c.tri_tests; out[7] = c.quad_tests;
  out[8] = c.probe_blas_nodes; out[9] = c.probe_tri_tests; out[10] = c.probe_quad_tests;
}

ORC_API void orc_intersect(void* h, const jt_ray* rays, int64_t n, jt_hit* out, uint64_t* counters_out, int threads) {
  Oracle* o = (Oracle*)h;
  Counters total;
  if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
  {
    Counters local;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
      Ray r{to_v3(rays[i].o), to_v3(rays[i].d), rays[i].tmin, rays[i].tmax};
      SceneIsec s = intersect_scene_bvh(o->scene, r, false, &local);
      jt_hit x;
      memset(&x, 0, sizeof(x));
      x.instance = s.instance; x.element = s.element; x.uv[0] = s.uv.x; x.uv[1] = s.uv.y;
      x.distance = s.distance; x.hit = s.hit ? 1 : 0;
      out[i] = x;
    }
#pragma omp critical
    total.add(local);
  }
  if (counters_out) export_counters(total, counters_out);
}

ORC_API void orc_intersect_instance(void* h, const jt_ray* rays, const int64_t* instances, int64_t n, jt_hit* out) {
  Oracle* o = (Oracle*)h;
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; i++) {
    Ray r{to_v3(rays[i].o), to_v3(rays[i].d), rays[i].tmin, rays[i].tmax};
    SceneIsec s = intersect_instance_bvh(o->scene, instances[i], r, false, nullptr);
    jt_hit x;
    memset(&x, 0, sizeof(x));
    x.instance = s.instance; x.element = s.element; x.uv[0] = s.uv.x; x.uv[1] = s.uv.y;
    x.distance = s.distance; x.hit = s.hit ? 1 : 0;
    out[i] = x;
  }
}

ORC_API void orc_sample_camera(void* h, int camera, int tent, int32_t width, int32_t height, const int32_t* ij,
                               const float* puv_luv, int64_t n, jt_ray* out) {
  Oracle* o = (Oracle*)h;
  for (int64_t k = 0; k < n; k++) {
    Ray r = sample_camera(o->scene.cameras[camera - 1], ij[2 * k], ij[2 * k + 1], width, height,
                          V2{puv_luv[4 * k], puv_luv[4 * k + 1]}, V2{puv_luv[4 * k + 2], puv_luv[4 * k + 3]},
                          tent != 0);
    out[k] = jt_ray{{r.o.x, r.o.y, r.o.z}, {r.d.x, r.d.y, r.d.z}, r.tmin, r.tmax};
  }
}

// ---- state + trace_samples -------------------------------------------------------------------------
static Params to_params(const jt_params* p) {
  Params q;
  q.camera = p->camera; q.resolution = p->resolution; q.samples = p->samples; q.bounces = p->bounces;
  q.sampler = p->sampler; q.clamp = p->clamp; q.nocaustics = p->nocaustics != 0;
  q.envhidden = p->envhidden != 0; q.tentfilter = p->tentfilter != 0; q.batch = p->batch;
  q.seed = p->seed; q.accumulate = p->accumulate;
  return q;
}

// make_trace_state, src/trace.jl:189-213
ORC_API void orc_make_state(void* h, const jt_params* p, int32_t* width, int32_t* height) {
  Oracle* o = (Oracle*)h;
  o->params = to_params(p);
  const Camera& cam = o->scene.cameras[p->camera - 1];
  int64_t w, hgt;
  if (cam.aspect >= 1.0f) {
    w = p->resolution;
    hgt = (int64_t)nearbyintf((float)p->resolution / cam.aspect);  // round half to even
  } else {
    hgt = p->resolution;
    w = (int64_t)nearbyintf((float)p->resolution * cam.aspect);
  }
  State& st = o->state;
  st.width = w; st.height = hgt; st.samples = 0;
  st.image.assign(w * hgt, V4{0, 0, 0, 0});
  st.albedo.assign(w * hgt, V3{0, 0, 0});
  st.normal.assign(w * hgt, V3{0, 0, 0});
  st.hits.assign(w * hgt, 0);
  *width = (int32_t)w;
  *height = (int32_t)hgt;
}

// trace_samples for an explicit global sample range (src/trace.jl:215-274 drives [samples, target))
ORC_API void orc_trace_range(void* h, const jt_params* p, int32_t begin, int32_t end, int threads) {
  Oracle* o = (Oracle*)h;
  Params q = to_params(p);
  State& st = o->state;
  if (threads <= 0) threads = omp_get_max_threads();
  Counters total;
#pragma omp parallel num_threads(threads)
  {
    Counters local;
#pragma omp for schedule(dynamic, 1) collapse(1)
    for (int64_t j = 0; j < st.height; j++)
      for (int64_t i = 0; i < st.width; i++)
        for (int64_t s = begin; s < end; s++) trace_sample(st, o->scene, i, j, s, q, &local);
#pragma omp critical
    total.add(local);
  }
  o->counters.add(total);
  st.samples += end - begin;
}

ORC_API void orc_trace_samples(void* h, const jt_params* p, int threads) {
  Oracle* o = (Oracle*)h;
  if (o->state.samples >= p->samples) return;
  int64_t target = std::min<int64_t>(o->state.samples + p->batch, p->samples);
  orc_trace_range(h, p, (int32_t)o->state.samples, (int32_t)target, threads);
}

// One pixel, one sample, into caller-provided scratch (debug hook for divergence hunting)
ORC_API void orc_trace_pixel(void* h, const jt_params* p, int32_t i, int32_t j, int32_t sample, float* radiance_hit) {
  Oracle* o = (Oracle*)h;
  Params q = to_params(p);
  const Camera& camera = o->scene.cameras[q.camera - 1];
  int64_t idx = o->state.width * j + i;
  Rng rng{jt_rng_key(q.seed, (uint32_t)idx, (uint32_t)sample), 0};
  V2 puv = rng.next2();
  V2 luv = rng.next2();
  Ray ray = sample_camera(camera, i, j, o->state.width, o->state.height, puv, luv, q.tentfilter);
  TraceResult r = q.sampler == 1 ? trace_path(o->scene, ray, q, rng, nullptr) : trace_naive(o->scene, ray, q, rng, nullptr);
  radiance_hit[0] = r.radiance.x; radiance_hit[1] = r.radiance.y; radiance_hit[2] = r.radiance.z;
  radiance_hit[3] = r.hit ? 1.0f : 0.0f;
  radiance_hit[4] = (float)rng.draw;
}

// The closest-hit queries of one (pixel, sample) path in call order: rays (jt_ray layout) and, per ray, the probed
// instance (1-based) or -1 for a scene query. Returns the number of queries (at most `max` are written).
ORC_API int64_t orc_trace_pixel_rays(void* h, const jt_params* p, int32_t i, int32_t j, int32_t sample, jt_ray* rays,
                                     int64_t* instances, int64_t max) {
  Oracle* o = (Oracle*)h;
  Params q = to_params(p);
  const Camera& camera = o->scene.cameras[q.camera - 1];
  int64_t idx = o->state.width * j + i;
  Rng rng{jt_rng_key(q.seed, (uint32_t)idx, (uint32_t)sample), 0};
  V2 puv = rng.next2();
  V2 luv = rng.next2();
  Ray ray = sample_camera(camera, i, j, o->state.width, o->state.height, puv, luv, q.tentfilter);
  std::vector<Counters::LoggedRay> log;
  Counters cnt;
  cnt.log = &log;
  if (q.sampler == 1) trace_path(o->scene, ray, q, rng, &cnt);
  else trace_naive(o->scene, ray, q, rng, &cnt);
  for (int64_t k = 0; k < (int64_t)log.size() && k < max; k++) {
    const Ray& r = log[(size_t)k].ray;
    rays[k].o[0] = r.o.x; rays[k].o[1] = r.o.y; rays[k].o[2] = r.o.z;
    rays[k].d[0] = r.d.x; rays[k].d[1] = r.d.y; rays[k].d[2] = r.d.z;
    rays[k].tmin = r.tmin; rays[k].tmax = r.tmax;
    instances[k] = log[(size_t)k].instance;
  }
  return (int64_t)log.size();
}

ORC_API void orc_get_state(void* h, float* image, float* albedo, float* normal, int64_t* hits, int32_t* samples) {
  Oracle* o = (Oracle*)h;
  const State& st = o->state;
  int64_t n = st.width * st.height;
  float inv = (o->params.accumulate == 1 && st.samples > 0) ? 1.0f / (float)st.samples : 1.0f;
  bool sums = o->params.accumulate == 1;
  if (image) for (int64_t i = 0; i < n; i++) {
    V4 v = sums ? st.image[i] * inv : st.image[i];
    image[4 * i] = v.x; image[4 * i + 1] = v.y; image[4 * i + 2] = v.z; image[4 * i + 3] = v.w;
  }
  if (albedo) for (int64_t i = 0; i < n; i++) {
    V3 v = sums ? st.albedo[i] * inv : st.albedo[i];
    albedo[3 * i] = v.x; albedo[3 * i + 1] = v.y; albedo[3 * i + 2] = v.z;
  }
  if (normal) for (int64_t i = 0; i < n; i++) {
    V3 v = sums ? st.normal[i] * inv : st.normal[i];
    normal[3 * i] = v.x; normal[3 * i + 1] = v.y; normal[3 * i + 2] = v.z;
  }
  if (hits) memcpy(hits, st.hits.data(), sizeof(int64_t) * n);
  if (samples) *samples = (int32_t)st.samples;
}

ORC_API void orc_get_counters(void* h, uint64_t* out, int reset) {
  Oracle* o = (Oracle*)h;
  export_counters(o->counters, out);
  if (reset) o->counters = Counters();
}

// ---- scalar probes for unit tests (known-answer / furnace tests of the lobes) -----------------------
// kind: 0 eval_bsdfcos, 1 sample_bsdfcos_pdf, 2 eval_delta, 3 sample_delta_pdf
ORC_API void orc_bsdf_eval(const float* mat /*type,color3,roughness,ior*/, const float* n, const float* o_,
                           const float* i_, int kind, float* out) {
  MaterialPoint m;
  memset(&m, 0, sizeof(m));
  m.type = (int32_t)mat[0]; m.color = V3{mat[1], mat[2], mat[3]}; m.roughness = mat[4]; m.ior = mat[5];
  m.opacity = 1.0f;
  V3 nn = to_v3(n), oo = to_v3(o_), ii = to_v3(i_);
  if (kind == 0) { V3 r = eval_bsdfcos(m, nn, oo, ii); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
  else if (kind == 1) { out[0] = sample_bsdfcos_pdf(m, nn, oo, ii); }
  else if (kind == 2) { V3 r = eval_delta(m, nn, oo, ii); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
  else { out[0] = sample_delta_pdf(m, nn, oo, ii); }
}
ORC_API void orc_bsdf_sample(const float* mat, const float* n, const float* o_, float rnl, float r1, float r2,
                             int delta, float* out) {
  MaterialPoint m;
  memset(&m, 0, sizeof(m));
  m.type = (int32_t)mat[0]; m.color = V3{mat[1], mat[2], mat[3]}; m.roughness = mat[4]; m.ior = mat[5];
  V3 r = delta ? sample_delta(m, to_v3(n), to_v3(o_), rnl) : sample_bsdfcos(m, to_v3(n), to_v3(o_), rnl, V2{r1, r2});
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API float orc_rng_float(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t draw) {
  return jt_rng_float(jt_rng_key(seed, pixel, sample), draw);
}
ORC_API float orc_fmath(int fn, float x, float y) {
  switch (fn) {
    case 0: return jt_sinf(x);
    case 1: return jt_cosf(x);
    case 2: return jt_atanf(x);
    case 3: return jt_atan2f(x, y);
    case 4: return jt_acosf(x);
    case 5: return jt_expf(x);
    case 6: return jt_logf(x);
    case 7: return srgb_to_rgb(x);
  }
  return 0.0f;
}
ORC_API void orc_fmath_array(int fn, const float* x, const float* y, int64_t n, float* out) {
  for (int64_t i = 0; i < n; i++) out[i] = orc_fmath(fn, x[i], y ? y[i] : 0.0f);
}

}  // extern "C"
