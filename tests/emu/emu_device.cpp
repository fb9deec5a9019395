// TEST INFRASTRUCTURE -- CPU single-stepping of the device code (see emu_shims.h).
#include "emu_shims.h"
#define JT_DEV static inline
#include "../../julia-raytracer_b200/csrc/jt_dev_trace.cuh"
#include "../../julia-raytracer_b200/csrc/jt_dev_persist.cuh"
#include "../../julia-raytracer_b200/csrc/jt_dev_wavefront.cuh"

#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

struct Emu {
  JtStagedScene staged;
  JtDevScene dev;
};

extern "C" {
#define EMU_API __attribute__((visibility("default")))

EMU_API const char* emu_last_error() { return jt_last_error(); }

EMU_API void* emu_create(const jt_scene_desc* d) {
  Emu* e = new Emu();
  if (jt_stage_scene(d, &e->staged) != 0) {
    delete e;
    return nullptr;
  }
  JtStagedScene& S = e->staged;
  JtStagedPointers P;
  P.ref_nodes = S.ref_nodes.data(); P.ref_prims = S.ref_prims.data(); P.shapes = S.shape_recs.data();
  P.positions = S.positions.data(); P.normals = S.normals.data(); P.texcoords = S.texcoords.data();
  P.colors = S.colors.data(); P.elements = S.elements.data(); P.instances = S.inst_recs.data();
  P.materials = S.mats.data(); P.textures = S.texs.data(); P.texels_f = S.texels_f.data();
  P.texels_b = S.texels_b.data(); P.srgb_lut = S.lut.data(); P.environments = S.envs.data();
  P.lights = S.lights.data(); P.light_cdf = S.cdf.data(); P.light_guide = S.cdf_guide.data(); P.cameras = S.cams.data();
  P.wnodes = (const float4*)S.wide.nodes.data(); P.wtris = (const float4*)S.wide.tris.data();
  P.tri_rank = S.tri_rank.data(); P.inst_rank = S.inst_rank.data(); P.inst_bounds = S.inst_bounds.data();
  jt_fill_dev_scene(S, P, &e->dev);
  return e;
}
EMU_API void emu_destroy(void* h) { delete (Emu*)h; }

EMU_API void emu_stats(void* h, int64_t* out) {
  Emu* e = (Emu*)h;
  out[0] = (int64_t)e->staged.wide.nodes.size();
  out[1] = (int64_t)e->staged.wide.tris.size();
  out[2] = e->staged.wide.inlined_instances;
  out[3] = e->staged.wide.instanced_instances;
  out[6] = e->staged.wide.flattened_instances;
  out[4] = e->staged.depth;
  out[5] = e->staged.blas_depth;
  out[7] = e->staged.wide_from_cache ? 1 : 0;
}

static void put_hit(jt_hit* o, const DHit& h) {
  memset(o, 0, sizeof(*o));
  if (h.inst >= 0) {
    o->instance = h.inst + 1; o->element = h.elem + 1; o->uv[0] = h.u; o->uv[1] = h.v; o->distance = h.t; o->hit = 1;
  } else {
    o->instance = -1; o->element = -1;
  }
}

// the persistent-warp traversal state machine (jt_dev_persist.cuh), one single-lane "warp" per ray
static DHit persist_one(const JtDevScene& S, const DRay& r) {
  uint2 stack_local[JT_WIDE_STACK - JT_SMEM_STACK];
  TravStack stack;
  stack.local = stack_local;
#if JT_SMEM_STACK > 0 && !defined(JT_EMU_COUNT)
  __shared__ uint2 stack_shared[JT_SMEM_STACK * JT_PERSIST_BLOCK];
  stack.shared = stack_shared + threadIdx.x;
#endif
  PersistLane L;
  persist_init(L, S, r.o, r.d, r.tmin, r.tmax, S.wide_root, -1);
  bool live = S.wide_root >= 0;
  persist_traverse(S, L, stack, live, false);
  return DHit{L.best.t, L.best.u, L.best.v, L.best.inst, L.best.elem};
}
static unsigned long long g_wide_counts[4] = {0, 0, 0, 0};
EMU_API void emu_wide_counts(unsigned long long* out, int reset) {
  for (int k = 0; k < 4; k++) { out[k] = g_wide_counts[k]; if (reset) g_wide_counts[k] = 0; }
}
EMU_API void emu_intersect(void* h, const jt_ray* rays, int64_t n, int traversal, jt_hit* out) {
  Emu* e = (Emu*)h;
#pragma omp parallel
  {
    jt_emu_counts = jt_emu_counts_t{0, 0, 0, 0};
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
      DRay r{f3{rays[i].o[0], rays[i].o[1], rays[i].o[2]}, f3{rays[i].d[0], rays[i].d[1], rays[i].d[2]}, rays[i].tmin, rays[i].tmax};
      put_hit(out + i, traversal == 1 ? intersect_scene<MODE_REF>(e->dev, r)
                                      : (traversal == 3 ? persist_one(e->dev, r) : intersect_scene<MODE_WIDE>(e->dev, r)));
    }
#pragma omp critical
    { g_wide_counts[0] += jt_emu_counts.wide_nodes; g_wide_counts[1] += jt_emu_counts.wide_prims; g_wide_counts[2] += jt_emu_counts.wide_instances; g_wide_counts[3] += jt_emu_counts.wide_xforms; }
  }
}
EMU_API void emu_intersect_instance(void* h, const jt_ray* rays, const int64_t* inst, int64_t n, int traversal, jt_hit* out) {
  Emu* e = (Emu*)h;
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < n; i++) {
    DRay r{f3{rays[i].o[0], rays[i].o[1], rays[i].o[2]}, f3{rays[i].d[0], rays[i].d[1], rays[i].d[2]}, rays[i].tmin, rays[i].tmax};
    put_hit(out + i, traversal == 1 ? intersect_instance<MODE_REF>(e->dev, (int)inst[i] - 1, r)
                                    : intersect_instance<MODE_WIDE>(e->dev, (int)inst[i] - 1, r));
  }
}

// sample_discrete through a guide table (sample_discrete_light) next to the plain bisection, for arbitrary CDFs
EMU_API int emu_sample_discrete(const float* cdf, int64_t n, const float* r, int64_t m, int32_t* plain, int32_t* guided) {
  JtLightRec L;
  memset(&L, 0, sizeof(L));
  std::vector<int32_t> guide;
  L.cdf_off = 0;
  L.cdf_len = (int32_t)n;
  jt_build_cdf_guide(cdf, n, &L, &guide);
  JtDevScene S;
  memset(&S, 0, sizeof(S));
  S.light_cdf = cdf;
  S.light_guide = guide.data();
  for (int64_t i = 0; i < m; i++) {
    plain[i] = sample_discrete(cdf, (int)n, r[i]);
    guided[i] = sample_discrete_light(S, L, r[i]);
  }
  return L.guide_len;
}

// trace_sample for samples [begin, end) of every pixel into host accumulators (float4 image/albedo/normal, int hits)
EMU_API void emu_trace_range(void* h, const jt_params* p, int width, int height, int begin, int end, float* image,
                             float* albedo, float* normal, int* hits, uint64_t* counters) {
  Emu* e = (Emu*)h;
  DevParams P;
  P.camera = p->camera - 1; P.width = width; P.height = height; P.bounces = p->bounces; P.sampler = p->sampler;
  P.clamp = p->clamp; P.nocaustics = p->nocaustics; P.envhidden = p->envhidden; P.tentfilter = p->tentfilter;
  P.accumulate = p->accumulate; P.seed = p->seed;
  DevState st{(float4*)image, (float4*)albedo, (float4*)normal, hits};
  bool has_env = e->dev.num_environments != 0;
  uint64_t scene_rays = 0, light_rays = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : scene_rays, light_rays)
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++) {
      int idx = width * j + i;
      for (int s = begin; s < end; s++) {
        PathCounters cnt{0u, 0u};
        Rng rng{jt_rng_key(P.seed, (uint32_t)idx, (uint32_t)s), 0u};
        f2 puv = rng.next2();
        f2 luv = rng.next2();
        DRay ray = sample_camera(e->dev.cameras[P.camera], i, j, width, height, puv, luv, P.tentfilter != 0);
        TraceOut r = p->traversal == 1
                         ? (P.sampler == 1 ? trace_path<MODE_REF>(e->dev, ray, P, rng, cnt) : trace_naive<MODE_REF>(e->dev, ray, P, rng, cnt))
                         : (P.sampler == 1 ? trace_path<MODE_WIDE>(e->dev, ray, P, rng, cnt) : trace_naive<MODE_WIDE>(e->dev, ray, P, rng, cnt));
        accumulate_sample(st, P, has_env, idx, s, r, ray.d);
        scene_rays += cnt.scene_rays;
        light_rays += cnt.light_rays;
      }
    }
  if (counters) { counters[0] += scene_rays; counters[1] += light_rays; }
}

// > 0: the wavefront's closest hits come from the persistent extend kernel (one emulated single-lane warp walking the whole
// queue), and every ray is SUSPENDED after that many traversal iterations, re-queued and resumed by the next launch:
// the park / resume path of jt_dev_persist.cuh, which on the GPU only runs in the tail of a launch.
EMU_API void emu_set_suspend_every(int n) { jt_emu_suspend_every = n; }
static unsigned long long g_resumed = 0;
EMU_API unsigned long long emu_resumed_rays(int reset) { unsigned long long r = g_resumed; if (reset) g_resumed = 0; return r; }

EMU_API unsigned long long emu_stolen_samples(int reset) { unsigned long long r = jt_emu_steals; if (reset) jt_emu_steals = 0; return r; }

// The wavefront integrator stepped sequentially (one emulated thread = one single-lane warp).
EMU_API int emu_trace_wavefront(void* h, const jt_params* p, int width, int height, int begin, int end, float* image,
                                float* albedo, float* normal, int* hits, uint64_t* counters) {
  Emu* e = (Emu*)h;
  DevParams P;
  P.camera = p->camera - 1; P.width = width; P.height = height; P.bounces = p->bounces; P.sampler = p->sampler;
  P.clamp = p->clamp; P.nocaustics = p->nocaustics; P.envhidden = p->envhidden; P.tentfilter = p->tentfilter;
  P.accumulate = p->accumulate; P.seed = p->seed;
  DevState st{(float4*)image, (float4*)albedo, (float4*)normal, hits};
  int n = width * height;
  std::vector<float4> ga((size_t)4 * n, float4{0, 0, 0, 0}), gb((size_t)4 * n, float4{0, 0, 0, 0}),
      gc((size_t)2 * n, float4{0, 0, 0, 0}), gd((size_t)2 * n, float4{0, 0, 0, 0});
  std::vector<int> q0(n), q1(n), qs((size_t)n * WF_NKEY), qp(n), counts(WF_C_TOTAL, 0);
  WfBuffers B;
  B.bind(ga.data(), gb.data(), gc.data(), gd.data());
  B.q_ext[0] = q0.data(); B.q_ext[1] = q1.data(); B.q_shade = qs.data(); B.q_probe = qp.data();
  std::vector<unsigned char> regen((size_t)n + 32, 0);
  B.regen = regen.data();
  std::vector<uint2> parked((size_t)n * JT_SUSPEND_STACK);
  B.parked = parked.data();
  std::vector<int> next_sample(n), commit(n);
  B.next_sample = next_sample.data(); B.commit = commit.data();
  std::vector<float4> held((size_t)3 * n, float4{0, 0, 0, 0});
  B.held = held.data();
  B.counts = counts.data(); B.n = n; B.pixel_base = 0;
  unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto run = [&](int threads, auto&& kernel) {
    for (int t = 0; t < threads; t++) { emu_blockIdx.x = (unsigned)t; kernel(); }
  };
  run(n, [&] { k_wf_generate(e->dev, B, P, begin, end, cnt); });
  int cur = 0, iterations = 0;
  while (counts[WF_C_DONE] < n) {  // the host loop of jt_api.cu: until every slot is idle for good
    int next = cur ^ 1;
    int ne = counts[WF_C_EXT(cur)];
    if (p->traversal == 1) run(ne, [&] { k_wf_extend<MODE_REF>(e->dev, B, cur, cnt); });
    else if (jt_emu_suspend_every > 0) {
      for (int q = 0; q < ne; q++) g_resumed += (__float_as_int(B.ray1[B.q_ext[cur][q]].z) & WF_RAY_SUSPENDED) != 0;
      run(1, [&] { k_wf_extend_persist(e->dev, B, cur, cnt); });
    } else run(ne, [&] { k_wf_extend<MODE_WIDE>(e->dev, B, cur, cnt); });
    int ns = 0;
    for (int k = 0; k < WF_NKEY; k++) ns += counts[WF_C_SHADEK(k)];
    if (P.sampler == 1) {
      if (p->traversal == 1) run(ns, [&] { k_wf_shade<1, MODE_REF>(e->dev, B, st, P, next, end, cnt); });
      else run(ns, [&] { k_wf_shade<1, MODE_WIDE>(e->dev, B, st, P, next, end, cnt); });
      int np = counts[WF_C_PROBE];
      cnt[5] += (unsigned long long)np;  // slots that still needed the probe kernel
      cnt[6] += (unsigned long long)ns;
      if (p->traversal == 1) run(np, [&] { k_wf_probe<MODE_REF>(e->dev, B, st, P, next, end, cnt); });
      else run(np, [&] { k_wf_probe<MODE_WIDE>(e->dev, B, st, P, next, end, cnt); });
    } else {
      run(ns, [&] { k_wf_shade<2, MODE_REF>(e->dev, B, st, P, next, end, cnt); });
    }
    // k_wf_regen, sequentially: flagged slots in slot order -> commit, claim, queue; counters recycled
    for (int sl = 0; sl < n; sl++)
      if (regen[sl]) {
        unsigned char flag_out = 0;
        int what = wf_regen_slot(e->dev, B, st, P, sl, cur, end, iterations, regen[sl], &flag_out);
        regen[sl] = flag_out;
        if (what == WF_REGEN_QUEUED || what == WF_REGEN_STOLEN) B.q_ext[next][B.counts[WF_C_EXT(next)]++] = sl;
        if (what == WF_REGEN_DONE) counts[WF_C_DONE]++;
      }
    counts[WF_C_EXT(cur)] = 0; counts[WF_C_PROBE] = 0; counts[WF_C_FETCH] = 0;
    for (int k = 0; k < WF_NKEY; k++) counts[WF_C_SHADEK(k)] = 0;
    cur = next;
    iterations++;
  }
  if (counters) { counters[0] += cnt[1]; counters[1] += cnt[2]; counters[2] += cnt[0]; }
  if (getenv("JT_EMU_VERBOSE")) fprintf(stderr, "emu wavefront: %d iterations, shade slots %llu, probe-kernel slots %llu, light probes %llu\n", iterations, cnt[6], cnt[5], cnt[2]);
  return iterations;
}
}
