// jt_api.cu -- C ABI of libjtrace_b200.so (include/jtrace_b200.h): scene upload, device-resident
// TraceState, the render-loop launches and the parity hooks. Kernels live at the bottom.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "jt_dev_output.cuh"
#include "jt_dev_trace.cuh"
#include "jt_dev_wavefront.cuh"
#include "jt_internal.h"

// =================================================================================================
// error channel
// =================================================================================================
static thread_local std::string g_error;

int jt_set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

#define JT_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return jt_set_error(JT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                          __LINE__);                                                                    \
  } while (0)

extern "C" const char* jt_last_error(void) { return g_error.c_str(); }
extern "C" const char* jt_version(void) { return "jtrace_b200 0.1 (sm_100a)"; }
extern "C" int jt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// =================================================================================================
// scene
// =================================================================================================
#define JT_MAX_PIPES 4
// The host keeps JT_WF_LOOKAHEAD batches (of 4 wavefront iterations per pipeline) enqueued beyond the one whose queue
// counters it is waiting for: one batch (~2.7 ms of GPU work at 1280x720) was not enough slack for a host thread that
// gets descheduled for a few milliseconds; two cost one more batch of empty launches at the end of a range.
#ifndef JT_WF_LOOKAHEAD
#define JT_WF_LOOKAHEAD 2
#endif
#define JT_WF_RING (JT_WF_LOOKAHEAD + 1)
#define JT_EXT_EVENTS (JT_MAX_PIPES * 2 * 4 * JT_WF_RING)
struct jt_scene {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  bool timing_open = false;
  std::vector<void*> allocs;
  JtDevScene dev;
  int num_cameras = 0, num_instances = 0;
  unsigned long long* d_counters = nullptr;  // [0] camera paths [1] scene rays [2] light rays [3] stolen samples [4] resumed rays [7] jt_intersect fetch
  uint64_t launches = 0;
  unsigned persist_blocks = 0, intersect_blocks = 0;
  // device time of the dominant kernel (extend), measured with CUDA events around every launch
  cudaEvent_t ext_ev[JT_EXT_EVENTS] = {};  // [pipeline][2 * 4 * JT_WF_RING]: start/stop pairs around the extend launches of the batches in flight
  double extend_ms = 0.0;
  uint64_t extend_launches = 0;
  std::vector<jt_state*> states;  // for flushing lazily batched sample ranges
  jt_scene_stats stats;
  int64_t device_bytes = 0;
};


struct jt_state {
  jt_scene* scene = nullptr;
  int width = 0, height = 0, samples = 0;
  int accumulate = 0;
  DevState dev;
  // wavefront integrator (allocated on first use)
  bool wf_ready = false;
  int npipe = 1;              // the image is split over npipe independent wavefront pipelines (own stream each)
  WfBuffers wf[JT_MAX_PIPES];
  cudaStream_t pipe_stream[JT_MAX_PIPES] = {};
  cudaEvent_t pipe_done[JT_MAX_PIPES] = {};
  std::vector<void*> wf_allocs;
  int* h_counts = nullptr;  // pinned mirror of the pipelines' counters (WF_C_TOTAL ints each), double-buffered
  cudaEvent_t poll_ev[JT_WF_RING][JT_MAX_PIPES] = {};
  // download staging (allocated on first download)
  void* d_pack = nullptr;
  void* h_pack = nullptr;
  uint64_t wf_iterations = 0;
  // lazily batched work: trace_samples is called samples/batch times (batch defaults to 1, src/cli.jl:78-81);
  // contiguous requests are merged and launched in chunks of JT_LAZY_SPP samples or at the next sync point
  bool has_pending = false;
  int pending_begin = 0, pending_end = 0;
  jt_params pending_params;
};

#define JT_LAZY_SPP 512 /* each flushed chunk ends in a drain phase (shrinking queues): 8 / 32 / 128 / 512 / 2048 spp per chunk: 305 / 322 / 333 (v3), 383 / 391 / 393 (v4) Msamples/s */
#ifndef JT_DEFAULT_PIPES
#define JT_DEFAULT_PIPES 2 /* measured on B200: 338 -> 357 Msamples/s (classroom), profiles/r01/tuning_variants.txt */
#endif
static int flush_state(jt_state* st);
static int flush_scene(jt_scene* sc) {
  for (jt_state* st : sc->states) {
    int rc = flush_state(st);
    if (rc) return rc;
  }
  return JT_OK;
}

template <class T, class A>
static int upload(jt_scene* sc, const std::vector<T, A>& v, const T** out) {
  *out = nullptr;
  size_t bytes = sizeof(T) * std::max<size_t>(v.size(), 1);
  void* p = nullptr;
  JT_CUDA(cudaMalloc(&p, bytes));
  sc->allocs.push_back(p);
  sc->device_bytes += (int64_t)bytes;
  if (!v.empty()) JT_CUDA(cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return JT_OK;
}

static void state_release_device(jt_state* st);

extern "C" void jt_scene_destroy(jt_scene* sc) {
  if (!sc) return;
  cudaSetDevice(sc->device);
  if (sc->stream) cudaStreamSynchronize(sc->stream);
  // states outlive their scene as orphans: their device memory goes with the scene, every later entry point on them
  // returns JT_ERR_INVALID, and jt_state_destroy only frees the host struct
  for (jt_state* st : sc->states) {
    state_release_device(st);
    st->scene = nullptr;
  }
  sc->states.clear();
  for (void* p : sc->allocs) cudaFree(p);
  for (int k = 0; k < 64; k++)
    if (sc->ext_ev[k]) cudaEventDestroy(sc->ext_ev[k]);
  if (sc->ev_start) cudaEventDestroy(sc->ev_start);
  if (sc->ev_stop) cudaEventDestroy(sc->ev_stop);
  if (sc->stream) cudaStreamDestroy(sc->stream);
  delete sc;
}

static int scene_upload_impl(const JtStagedScene& staged, int device, jt_scene* sc) {
  int ndev = jt_device_count();
  if (ndev <= 0) return jt_set_error(JT_ERR_NO_DEVICE, "no CUDA device visible: libjtrace_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return jt_set_error(JT_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  sc->device = device;
  JT_CUDA(cudaSetDevice(device));
  JT_CUDA(cudaStreamCreateWithFlags(&sc->stream, cudaStreamNonBlocking));
  JT_CUDA(cudaEventCreate(&sc->ev_start));
  JT_CUDA(cudaEventCreate(&sc->ev_stop));
  for (int k = 0; k < JT_EXT_EVENTS; k++) JT_CUDA(cudaEventCreate(&sc->ext_ev[k]));
  memset(&sc->dev, 0, sizeof(sc->dev));
  memset(&sc->stats, 0, sizeof(sc->stats));

  int rc = JT_OK;
  const auto& wide = staged.wide;

  // ---- upload -----------------------------------------------------------------------------------------------
  JtDevScene& D = sc->dev;
  JtStagedPointers P;
  const JtWideNode* wn = nullptr;
  const JtWideTri* wt = nullptr;
  if ((rc = upload(sc, staged.ref_nodes, &P.ref_nodes)) || (rc = upload(sc, staged.ref_prims, &P.ref_prims)) ||
      (rc = upload(sc, staged.shape_recs, &P.shapes)) || (rc = upload(sc, staged.positions, &P.positions)) ||
      (rc = upload(sc, staged.normals, &P.normals)) || (rc = upload(sc, staged.texcoords, &P.texcoords)) ||
      (rc = upload(sc, staged.colors, &P.colors)) || (rc = upload(sc, staged.elements, &P.elements)) ||
      (rc = upload(sc, staged.inst_recs, &P.instances)) || (rc = upload(sc, staged.mats, &P.materials)) ||
      (rc = upload(sc, staged.texs, &P.textures)) || (rc = upload(sc, staged.texels_f, &P.texels_f)) ||
      (rc = upload(sc, staged.texels_b, &P.texels_b)) || (rc = upload(sc, staged.lut, &P.srgb_lut)) ||
      (rc = upload(sc, staged.envs, &P.environments)) || (rc = upload(sc, staged.lights, &P.lights)) ||
      (rc = upload(sc, staged.cdf, &P.light_cdf)) || (rc = upload(sc, staged.cams, &P.cameras)) ||
      (rc = upload(sc, wide.nodes, &wn)) || (rc = upload(sc, wide.tris, &wt)) ||
      (rc = upload(sc, staged.tri_rank, &P.tri_rank)) || (rc = upload(sc, staged.inst_rank, &P.inst_rank)) ||
      (rc = upload(sc, staged.inst_bounds, &P.inst_bounds)) || (rc = upload(sc, staged.cdf_guide, &P.light_guide)))
    return rc;
  P.wnodes = (const float4*)wn;
  P.wtris = (const float4*)wt;
  jt_fill_dev_scene(staged, P, &D);
  sc->num_cameras = (int)staged.num_cameras;
  sc->num_instances = (int)staged.num_instances;
  void* cnt = nullptr;
  JT_CUDA(cudaMalloc(&cnt, 8 * sizeof(unsigned long long)));
  sc->allocs.push_back(cnt);
  JT_CUDA(cudaMemset(cnt, 0, 8 * sizeof(unsigned long long)));
  sc->d_counters = (unsigned long long*)cnt;

  sc->stats.wide_nodes = (int64_t)wide.nodes.size();
  sc->stats.wide_node_bytes = (int64_t)wide.nodes.size() * 80;
  sc->stats.prim_records = (int64_t)wide.tris.size();
  sc->stats.prim_record_bytes = (int64_t)wide.tris.size() * 48;
  sc->stats.inlined_instances = wide.inlined_instances;
  sc->stats.instanced_instances = wide.instanced_instances;
  sc->stats.texture_bytes = (int64_t)(staged.texels_f.size() * 16 + staged.texels_b.size() * 4);
  sc->stats.total_device_bytes = sc->device_bytes;
  sc->stats.wide_depth_top = staged.depth;
  sc->stats.wide_depth_blas = staged.blas_depth;
  sc->stats.opened_instances = wide.flattened_instances;
  sc->stats.wide_bvh_from_cache = staged.wide_from_cache ? 1 : 0;
  return JT_OK;
}

// Upload an already staged scene (host vectors, wide BVH built) to `device`. jt_group stages once and uploads N times.
int jt_scene_create_staged(const JtStagedScene& staged, int device, jt_scene** out) {
  *out = nullptr;
  jt_scene* sc = new (std::nothrow) jt_scene();
  if (!sc) return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  int rc;
  try {
    rc = scene_upload_impl(staged, device, sc);
  } catch (const std::exception& e) {
    rc = jt_set_error(JT_ERR_INTERNAL, "jt_scene_create: %s", e.what());
  } catch (...) {
    rc = jt_set_error(JT_ERR_INTERNAL, "jt_scene_create: unknown exception");
  }
  if (rc != JT_OK) {
    std::string keep = g_error;
    jt_scene_destroy(sc);
    g_error = keep;
    return rc;
  }
  *out = sc;
  return JT_OK;
}

int jt_stage_scene_checked(const jt_scene_desc* desc, JtStagedScene* staged) {
  try {
    return jt_stage_scene(desc, staged);
  } catch (const std::exception& e) {
    return jt_set_error(JT_ERR_INTERNAL, "jt_scene_create: %s", e.what());
  } catch (...) {
    return jt_set_error(JT_ERR_INTERNAL, "jt_scene_create: unknown exception");
  }
}

extern "C" int jt_scene_create(const jt_scene_desc* desc, int device, jt_scene** out) {
  if (!desc || !out) return jt_set_error(JT_ERR_INVALID, "jt_scene_create: null argument");
  *out = nullptr;
  int ndev = jt_device_count();  // before the (expensive) staging
  if (ndev <= 0) return jt_set_error(JT_ERR_NO_DEVICE, "no CUDA device visible: libjtrace_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return jt_set_error(JT_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
  JtStagedScene staged;
  int rc = jt_stage_scene_checked(desc, &staged);
  if (rc != JT_OK) return rc;
  return jt_scene_create_staged(staged, device, out);
}

extern "C" int jt_scene_get_stats(jt_scene* sc, jt_scene_stats* out) {
  if (!sc || !out) return jt_set_error(JT_ERR_INVALID, "jt_scene_get_stats: null argument");
  *out = sc->stats;
  return JT_OK;
}

extern "C" int jt_scene_counters(jt_scene* sc, jt_counters* out, int reset) {
  if (!sc || !out) return jt_set_error(JT_ERR_INVALID, "jt_scene_counters: null argument");
  JT_CUDA(cudaSetDevice(sc->device));
  int frc = flush_scene(sc);
  if (frc) return frc;
  JT_CUDA(cudaStreamSynchronize(sc->stream));
  unsigned long long h[8];
  JT_CUDA(cudaMemcpy(h, sc->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->camera_paths = h[0];
  out->scene_rays = h[1];
  out->light_rays = h[2];
  out->kernel_launches = sc->launches;
  out->extend_kernel_us = (uint64_t)(sc->extend_ms * 1000.0);  // microseconds in the extend (closest-hit) kernel
  out->extend_launches = sc->extend_launches;
  out->stolen_samples = h[3];
  out->resumed_rays = h[4];
  if (reset) {
    JT_CUDA(cudaMemset(sc->d_counters, 0, sizeof(h)));
    sc->launches = 0;
    sc->extend_ms = 0.0;
    sc->extend_launches = 0;
  }
  return JT_OK;
}

// =================================================================================================
// kernels
// =================================================================================================
struct HitOut {  // jt_hit
  long long instance, element;
  float u, v, distance;
  unsigned int hit;
};
static_assert(sizeof(HitOut) == 32, "jt_hit layout");

JT_DEV void store_hit(HitOut* out, const DHit& h) {
  HitOut o;
  if (h.inst >= 0) {
    o.instance = h.inst + 1; o.element = h.elem + 1; o.u = h.u; o.v = h.v; o.distance = h.t; o.hit = 1u;
  } else {
    o.instance = -1; o.element = -1; o.u = 0.0f; o.v = 0.0f; o.distance = 0.0f; o.hit = 0u;
  }
  *out = o;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_intersect(JtDevScene S, const jt_ray* __restrict__ rays, long long n,
                                                   HitOut* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jt_ray r = rays[i];
  DRay ray{f3{r.o[0], r.o[1], r.o[2]}, f3{r.d[0], r.d[1], r.d[2]}, r.tmin, r.tmax};
  store_hit(out + i, intersect_scene<MODE>(S, ray));
}

// Wide-BVH closest hit with persistent warps + dynamic fetch (jt_dev_persist.cuh).
__global__ void __launch_bounds__(JT_PERSIST_BLOCK) k_intersect_persist(JtDevScene S, const jt_ray* __restrict__ rays,
                                                                        long long n, HitOut* __restrict__ out,
                                                                        int* fetch_counter) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int count = (int)n;
  uint2 stack_local[JT_WIDE_STACK - JT_SMEM_STACK];
  TravStack stack;
  stack.local = stack_local;
#if JT_SMEM_STACK > 0 && !defined(JT_EMU_COUNT)
  __shared__ uint2 stack_shared[JT_SMEM_STACK * JT_PERSIST_BLOCK];
  stack.shared = stack_shared + threadIdx.x;
#endif
  PersistLane L;
  bool live = false, more = true;
  int s = -1;
  for (;;) {
    __syncwarp();
    if (s >= 0 && !live) {
      store_hit(out + s, DHit{L.best.t, L.best.u, L.best.v, L.best.inst, L.best.elem});
      s = -1;
    }
    if (more) {
      bool want = !live;
      int idx = persist_fetch(fetch_counter, want, count);
      if (idx >= 0) {
        s = idx;
        jt_ray r = rays[idx];
        persist_init(L, S, f3{r.o[0], r.o[1], r.o[2]}, f3{r.d[0], r.d[1], r.d[2]}, r.tmin, r.tmax, S.wide_root, -1);
        live = S.wide_root >= 0;
      }
      if (__ballot_sync(FULL, want && idx < 0)) more = false;
    }
    if (__ballot_sync(FULL, live || s >= 0) == 0u) break;
    persist_traverse(S, L, stack, live, more);
  }
}

template <int MODE>
__global__ void __launch_bounds__(128) k_intersect_instance(JtDevScene S, const jt_ray* __restrict__ rays,
                                                            const long long* __restrict__ instances, long long n,
                                                            HitOut* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  jt_ray r = rays[i];
  DRay ray{f3{r.o[0], r.o[1], r.o[2]}, f3{r.d[0], r.d[1], r.d[2]}, r.tmin, r.tmax};
  int inst = (int)instances[i] - 1;
  DHit h{0.0f, 0.0f, 0.0f, -1, -1};
  if (inst >= 0 && inst < S.num_instances) h = intersect_instance<MODE>(S, inst, ray);
  store_hit(out + i, h);
}

__global__ void k_sample_camera(JtDevScene S, int camera, int tent, int width, int height,
                                const int* __restrict__ ij, const float* __restrict__ r, long long n, jt_ray* out) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  DRay ray = sample_camera(S.cameras[camera], ij[2 * k], ij[2 * k + 1], width, height, f2{r[4 * k], r[4 * k + 1]},
                           f2{r[4 * k + 2], r[4 * k + 3]}, tent != 0);
  jt_ray o;
  o.o[0] = ray.o.x; o.o[1] = ray.o.y; o.o[2] = ray.o.z;
  o.d[0] = ray.d.x; o.d[1] = ray.d.y; o.d[2] = ray.d.z;
  o.tmin = ray.tmin; o.tmax = ray.tmax;
  out[k] = o;
}

// One thread per pixel, samples [begin, end) in order: the whole of trace_sample (src/trace.jl:584-649).
// This "megakernel" form is the bit-exact parity integrator; warps cover 16x2 pixel tiles.
template <int MODE>
__global__ void __launch_bounds__(128) k_trace_mega(JtDevScene S, DevState st, DevParams P, int begin, int end,
                                                    unsigned long long* counters) {
  int i = blockIdx.x * 16 + (threadIdx.x & 15);
  int j = blockIdx.y * 8 + (threadIdx.x >> 4);
  PathCounters cnt{0u, 0u};
  unsigned int paths = 0;
  if (i < P.width && j < P.height) {
    int idx = P.width * j + i;
    const JtCameraRec& C = S.cameras[P.camera];
    bool has_env = S.num_environments != 0;
    for (int s = begin; s < end; s++) {
      Rng rng{jt_rng_key(P.seed, (uint32_t)idx, (uint32_t)s), 0u};
      f2 puv = rng.next2();
      f2 luv = rng.next2();
      DRay ray = sample_camera(C, i, j, P.width, P.height, puv, luv, P.tentfilter != 0);
      TraceOut r = P.sampler == 1 ? trace_path<MODE>(S, ray, P, rng, cnt) : trace_naive<MODE>(S, ray, P, rng, cnt);
      accumulate_sample(st, P, has_env, idx, s, r, ray.d);
      paths++;
    }
  }
  // one atomic per warp per counter
  unsigned int a = __reduce_add_sync(0xFFFFFFFFu, paths);
  unsigned int b = __reduce_add_sync(0xFFFFFFFFu, cnt.scene_rays);
  unsigned int c = __reduce_add_sync(0xFFFFFFFFu, cnt.light_rays);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counters + 0, (unsigned long long)a);
    atomicAdd(counters + 1, (unsigned long long)b);
    atomicAdd(counters + 2, (unsigned long long)c);
  }
}

__global__ void k_finalize(const float4* __restrict__ image, const float4* __restrict__ albedo,
                           const float4* __restrict__ normal, const int* __restrict__ hits, long long n, float scale,
                           float* out_image, float* out_albedo, float* out_normal, long long* out_hits) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = image[i], b = albedo[i], c = normal[i];
  if (out_image) {
    out_image[4 * i] = a.x * scale; out_image[4 * i + 1] = a.y * scale; out_image[4 * i + 2] = a.z * scale;
    out_image[4 * i + 3] = a.w * scale;
  }
  if (out_albedo) {
    out_albedo[3 * i] = b.x * scale; out_albedo[3 * i + 1] = b.y * scale; out_albedo[3 * i + 2] = b.z * scale;
  }
  if (out_normal) {
    out_normal[3 * i] = c.x * scale; out_normal[3 * i + 1] = c.y * scale; out_normal[3 * i + 2] = c.z * scale;
  }
  if (out_hits) out_hits[i] = hits[i];
}

// N3: the save path on the GPU (per-pixel conversion in jt_dev_output.cuh).
__global__ void k_srgb8(const float4* __restrict__ image, long long n, float scale, uchar4* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = image[i];
  out[i] = jt_srgb8_pixel(make_float4(p.x * scale, p.y * scale, p.z * scale, p.w * scale));
}

// =================================================================================================
// state
// =================================================================================================
static int check_params(jt_scene* sc, const jt_params* p) {
  if (!p) return jt_set_error(JT_ERR_INVALID, "null params");
  if (p->camera < 1 || p->camera > sc->num_cameras) return jt_set_error(JT_ERR_INVALID, "camera %d out of range (1..%d)", p->camera, sc->num_cameras);
  if (p->sampler != 1 && p->sampler != 2) return jt_set_error(JT_ERR_INVALID, "sampler must be 1 (path) or 2 (naive)");
  if (p->resolution < 1 || p->resolution > 32768) return jt_set_error(JT_ERR_INVALID, "bad resolution %d", p->resolution);
  if (p->traversal != 0 && p->traversal != 1) return jt_set_error(JT_ERR_INVALID, "traversal must be 0 (wide) or 1 (reference)");
  if (p->accumulate != 0 && p->accumulate != 1) return jt_set_error(JT_ERR_INVALID, "accumulate must be 0 or 1");
  if (p->integrator != 0 && p->integrator != 1) return jt_set_error(JT_ERR_INVALID, "integrator must be 0 (wavefront) or 1 (megakernel)");
  // the wavefront integrator packs bounce + 1 into 8 bits of the per-slot control word (jt_dev_wavefront.cuh) and always
  // traces the camera segment; outside 0..254 it would silently diverge from the reference loop `for bounce in 0:bounces`
  if (p->integrator == 0 && (p->bounces < 0 || p->bounces > 254))
    return jt_set_error(JT_ERR_INVALID, "bounces %d outside 0..254 (wavefront integrator; integrator = 1 accepts any value)", p->bounces);
  if (p->sampler == 1 && sc->dev.num_lights == 0)
    return jt_set_error(JT_ERR_UNSUPPORTED, "path sampler on a scene without lights: sample_lights indexes an empty array in the reference");
  return JT_OK;
}

extern "C" int jt_state_create(jt_scene* sc, const jt_params* p, jt_state** out) {
  if (!sc || !out) return jt_set_error(JT_ERR_INVALID, "jt_state_create: null argument");
  *out = nullptr;
  int rc = check_params(sc, p);
  if (rc) return rc;
  JT_CUDA(cudaSetDevice(sc->device));
  JtCameraRec cam;
  JT_CUDA(cudaMemcpy(&cam, sc->dev.cameras + (p->camera - 1), sizeof(cam), cudaMemcpyDeviceToHost));
  int w, h;  // make_trace_state, src/trace.jl:189-197 (Float32 division, round half to even)
  if (cam.aspect >= 1.0f) {
    w = p->resolution;
    h = (int)nearbyintf((float)p->resolution / cam.aspect);
  } else {
    h = p->resolution;
    w = (int)nearbyintf((float)p->resolution * cam.aspect);
  }
  if (w < 1 || h < 1) return jt_set_error(JT_ERR_INVALID, "degenerate image size %dx%d", w, h);
  jt_state* st = new (std::nothrow) jt_state();
  if (!st) return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  st->scene = sc;
  st->width = w;
  st->height = h;
  st->accumulate = p->accumulate;
  size_t n = (size_t)w * h;
  void *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
  if (cudaMalloc(&a, n * 16) != cudaSuccess || cudaMalloc(&b, n * 16) != cudaSuccess ||
      cudaMalloc(&c, n * 16) != cudaSuccess || cudaMalloc(&d, n * 4) != cudaSuccess) {
    cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(d);
    delete st;
    return jt_set_error(JT_ERR_CUDA, "cudaMalloc of the trace state failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  st->dev.image = (float4*)a; st->dev.albedo = (float4*)b; st->dev.normal = (float4*)c; st->dev.hits = (int*)d;
  sc->states.push_back(st);
  *out = st;
  return jt_state_reset(st);
}

#define JT_LIVE_STATE(st, who)                                                                    \
  do {                                                                                           \
    if (!(st)) return jt_set_error(JT_ERR_INVALID, who ": null argument");                       \
    if (!(st)->scene) return jt_set_error(JT_ERR_INVALID, who ": the state's scene was destroyed"); \
  } while (0)

extern "C" int jt_state_reset(jt_state* st) {
  JT_LIVE_STATE(st, "jt_state_reset");
  jt_scene* sc = st->scene;
  st->has_pending = false;
  JT_CUDA(cudaSetDevice(sc->device));
  size_t n = (size_t)st->width * st->height;
  JT_CUDA(cudaMemsetAsync(st->dev.image, 0, n * 16, sc->stream));
  JT_CUDA(cudaMemsetAsync(st->dev.albedo, 0, n * 16, sc->stream));
  JT_CUDA(cudaMemsetAsync(st->dev.normal, 0, n * 16, sc->stream));
  JT_CUDA(cudaMemsetAsync(st->dev.hits, 0, n * 4, sc->stream));
  st->samples = 0;
  return JT_OK;
}

// Frees everything the state owns on the device (the caller has made the scene's device current and idle).
static void state_release_device(jt_state* st) {
  st->has_pending = false;
  cudaFree(st->dev.image); cudaFree(st->dev.albedo); cudaFree(st->dev.normal); cudaFree(st->dev.hits);
  st->dev.image = st->dev.albedo = st->dev.normal = nullptr;
  st->dev.hits = nullptr;
  for (void* p : st->wf_allocs) cudaFree(p);
  st->wf_allocs.clear();
  st->wf_ready = false;
  for (int k = 1; k < JT_MAX_PIPES; k++)
    if (st->pipe_stream[k]) cudaStreamDestroy(st->pipe_stream[k]);
  for (int k = 0; k < JT_MAX_PIPES; k++) {
    st->pipe_stream[k] = nullptr;
    if (st->pipe_done[k]) cudaEventDestroy(st->pipe_done[k]);
    st->pipe_done[k] = nullptr;
    for (int b = 0; b < 2; b++) {
      if (st->poll_ev[b][k]) cudaEventDestroy(st->poll_ev[b][k]);
      st->poll_ev[b][k] = nullptr;
    }
  }
  if (st->h_counts) cudaFreeHost(st->h_counts);
  if (st->d_pack) cudaFree(st->d_pack);
  if (st->h_pack) cudaFreeHost(st->h_pack);
  st->h_counts = nullptr;
  st->d_pack = st->h_pack = nullptr;
}

extern "C" void jt_state_destroy(jt_state* st) {
  if (!st) return;
  if (st->scene) {  // an orphan (scene destroyed first) has nothing left on the device
    auto& v = st->scene->states;
    for (size_t i = 0; i < v.size(); i++)
      if (v[i] == st) {
        v.erase(v.begin() + (long)i);
        break;
      }
    cudaSetDevice(st->scene->device);
    cudaStreamSynchronize(st->scene->stream);
    state_release_device(st);
  }
  delete st;
}

extern "C" int jt_state_size(jt_state* st, int32_t* width, int32_t* height, int32_t* samples) {
  if (!st) return jt_set_error(JT_ERR_INVALID, "jt_state_size: null argument");
  if (width) *width = st->width;
  if (height) *height = st->height;
  if (samples) *samples = st->samples;
  return JT_OK;
}

extern "C" int jt_state_set_samples(jt_state* st, int32_t samples) {
  if (!st || samples < 0) return jt_set_error(JT_ERR_INVALID, "jt_state_set_samples: bad argument");
  st->samples = samples;
  return JT_OK;
}

extern "C" int jt_state_device_buffers(jt_state* st, void** image, void** albedo, void** normal, void** hits,
                                       int64_t* count) {
  JT_LIVE_STATE(st, "jt_state_device_buffers");
  JT_CUDA(cudaSetDevice(st->scene->device));
  int frc = flush_state(st);
  if (frc) return frc;
  // the caller reduces these buffers on ITS stream (NCCL / torch): everything the library enqueued must have landed.
  // (The pipelines' streams are joined into scene->stream at the end of every flushed range.)
  JT_CUDA(cudaStreamSynchronize(st->scene->stream));
  if (image) *image = st->dev.image;
  if (albedo) *albedo = st->dev.albedo;
  if (normal) *normal = st->dev.normal;
  if (hits) *hits = st->dev.hits;
  if (count) *count = (int64_t)st->width * st->height;
  return JT_OK;
}

extern "C" int jt_state_download(jt_state* st, float* image, float* albedo, float* normal, int64_t* hits) {
  JT_LIVE_STATE(st, "jt_state_download");
  jt_scene* sc = st->scene;
  JT_CUDA(cudaSetDevice(sc->device));
  int frc = flush_state(st);
  if (frc) return frc;
  long long n = (long long)st->width * st->height;
  // persistent staging: one packed device buffer + one pinned host buffer (48 B per pixel)
  if (!st->d_pack) JT_CUDA(cudaMalloc(&st->d_pack, (size_t)n * 48));
  if (!st->h_pack) JT_CUDA(cudaHostAlloc(&st->h_pack, (size_t)n * 48, cudaHostAllocDefault));
  char* dp = (char*)st->d_pack;
  float* di = (float*)dp;
  float* da = (float*)(dp + n * 16);
  float* dn = (float*)(dp + n * 28);
  long long* dh = (long long*)(dp + n * 40);
  float scale = (st->accumulate == 1 && st->samples > 0) ? 1.0f / (float)st->samples : 1.0f;
  k_finalize<<<(unsigned)((n + 255) / 256), 256, 0, sc->stream>>>(st->dev.image, st->dev.albedo, st->dev.normal,
                                                                   st->dev.hits, n, scale, image ? di : nullptr,
                                                                   albedo ? da : nullptr, normal ? dn : nullptr,
                                                                   hits ? dh : nullptr);
  sc->launches++;
  JT_CUDA(cudaGetLastError());
  char* hp = (char*)st->h_pack;
  if (image) JT_CUDA(cudaMemcpyAsync(hp, di, (size_t)n * 16, cudaMemcpyDeviceToHost, sc->stream));
  if (albedo) JT_CUDA(cudaMemcpyAsync(hp + n * 16, da, (size_t)n * 12, cudaMemcpyDeviceToHost, sc->stream));
  if (normal) JT_CUDA(cudaMemcpyAsync(hp + n * 28, dn, (size_t)n * 12, cudaMemcpyDeviceToHost, sc->stream));
  if (hits) JT_CUDA(cudaMemcpyAsync(hp + n * 40, dh, (size_t)n * 8, cudaMemcpyDeviceToHost, sc->stream));
  JT_CUDA(cudaStreamSynchronize(sc->stream));
  if (image) memcpy(image, hp, (size_t)n * 16);
  if (albedo) memcpy(albedo, hp + n * 16, (size_t)n * 12);
  if (normal) memcpy(normal, hp + n * 28, (size_t)n * 12);
  if (hits) memcpy(hits, hp + n * 40, (size_t)n * 8);
  return JT_OK;
}

extern "C" int jt_state_download_srgb8(jt_state* st, uint8_t* rgba8) {
  if (!rgba8) return jt_set_error(JT_ERR_INVALID, "jt_state_download_srgb8: null argument");
  JT_LIVE_STATE(st, "jt_state_download_srgb8");
  jt_scene* sc = st->scene;
  JT_CUDA(cudaSetDevice(sc->device));
  int frc = flush_state(st);
  if (frc) return frc;
  long long n = (long long)st->width * st->height;
  if (!st->d_pack) JT_CUDA(cudaMalloc(&st->d_pack, (size_t)n * 48));
  if (!st->h_pack) JT_CUDA(cudaHostAlloc(&st->h_pack, (size_t)n * 48, cudaHostAllocDefault));
  float scale = (st->accumulate == 1 && st->samples > 0) ? 1.0f / (float)st->samples : 1.0f;
  k_srgb8<<<(unsigned)((n + 255) / 256), 256, 0, sc->stream>>>(st->dev.image, n, scale, (uchar4*)st->d_pack);
  sc->launches++;
  JT_CUDA(cudaGetLastError());
  JT_CUDA(cudaMemcpyAsync(st->h_pack, st->d_pack, (size_t)n * 4, cudaMemcpyDeviceToHost, sc->stream));
  JT_CUDA(cudaStreamSynchronize(sc->stream));
  memcpy(rgba8, st->h_pack, (size_t)n * 4);
  return JT_OK;
}

// =================================================================================================
// the hot path
// =================================================================================================
__global__ void k_wf_extend_persist(JtDevScene, WfBuffers, int, unsigned long long*);
// Persistent kernels: one resident wave of blocks (SM count x occupancy), each warp loops over the queue.
static unsigned persist_grid(jt_scene* sc) {
  if (sc->persist_blocks == 0) {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_extend_persist, JT_PERSIST_BLOCK, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device);
    sc->persist_blocks = (unsigned)std::max(1, per_sm) * (unsigned)std::max(1, sms);
  }
  return sc->persist_blocks;
}

static DevParams dev_params(const jt_params* p, const jt_state* st) {
  DevParams P;
  P.camera = p->camera - 1;
  P.width = st->width; P.height = st->height;
  P.bounces = p->bounces; P.sampler = p->sampler; P.clamp = p->clamp;
  P.nocaustics = p->nocaustics; P.envhidden = p->envhidden; P.tentfilter = p->tentfilter;
  P.accumulate = st->accumulate;
  P.seed = p->seed;
  return P;
}

static unsigned persist_grid(jt_scene* sc);

// Experiment knob (JT_CARVEOUT = per cent of the unified L1 / shared memory given to shared memory): one preference for
// every wavefront kernel, so that kernels of the two pipelines never ask an SM for different configurations.
static void wf_set_carveout() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = getenv("JT_CARVEOUT");
  if (!e) return;
  const int pct = atoi(e);
  cudaFuncSetAttribute(k_wf_extend_persist, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_wf_regen, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_wf_generate, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_wf_shade<1, MODE_WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_wf_shade<2, MODE_WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(k_wf_probe<MODE_WIDE>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

static int wf_prepare(jt_scene* sc, jt_state* st) {
  if (st->wf_ready) return JT_OK;
  wf_set_carveout();
  const size_t total = (size_t)st->width * st->height;
  int npipe = getenv("JT_PIPELINES") ? atoi(getenv("JT_PIPELINES")) : JT_DEFAULT_PIPES;
  npipe = std::max(1, std::min(JT_MAX_PIPES, npipe));
  if (total < 65536) npipe = 1;
  st->npipe = npipe;
  auto alloc = [&](void** out, size_t bytes) -> int {
    JT_CUDA(cudaMalloc(out, bytes));
    st->wf_allocs.push_back(*out);
    return JT_OK;
  };
  size_t base = 0;
  for (int k = 0; k < npipe; k++) {
    // contiguous pixel ranges, warp-aligned
    size_t n = (k == npipe - 1) ? total - base : ((total / (size_t)npipe + 31) & ~(size_t)31);
    WfBuffers& B = st->wf[k];
    memset(&B, 0, sizeof(B));
    B.n = (int)n;
    B.pixel_base = (int)base;
    base += n;
    int rc;
    float4 *ga = nullptr, *gb = nullptr, *gc = nullptr, *gd = nullptr;
    if ((rc = alloc((void**)&ga, n * 64)) || (rc = alloc((void**)&gb, n * 64)) || (rc = alloc((void**)&gc, n * 32)) ||
        (rc = alloc((void**)&gd, n * 32)) ||
        (rc = alloc((void**)&B.parked, n * JT_SUSPEND_STACK * sizeof(uint2))) ||
        (rc = alloc((void**)&B.held, (JT_HELD_RESULT ? n * 3 : 1) * sizeof(float4))) ||
        (rc = alloc((void**)&B.q_ext[0], n * 4)) || (rc = alloc((void**)&B.q_ext[1], n * 4)) ||
        (rc = alloc((void**)&B.next_sample, n * 4)) || (rc = alloc((void**)&B.commit, n * 4)) ||
        (rc = alloc((void**)&B.regen, ((n + 15) & ~(size_t)15) + 16)) ||
        (rc = alloc((void**)&B.q_shade, n * 4 * WF_NKEY)) || (rc = alloc((void**)&B.q_probe, n * 4)) ||
        (rc = alloc((void**)&B.counts, WF_C_TOTAL * 4)))
      return rc;
    B.bind(ga, gb, gc, gd);
    JT_CUDA(cudaMemset(B.regen, 0, ((n + 15) & ~(size_t)15) + 16));
    st->pipe_stream[k] = sc->stream;
    if (k > 0) JT_CUDA(cudaStreamCreateWithFlags(&st->pipe_stream[k], cudaStreamNonBlocking));
    JT_CUDA(cudaEventCreateWithFlags(&st->pipe_done[k], cudaEventDisableTiming));
    for (int b = 0; b < JT_WF_RING; b++) JT_CUDA(cudaEventCreateWithFlags(&st->poll_ev[b][k], cudaEventDisableTiming));
  }
  JT_CUDA(cudaHostAlloc((void**)&st->h_counts, JT_WF_RING * JT_MAX_PIPES * WF_C_TOTAL * 4, cudaHostAllocDefault));
  st->wf_ready = true;
  return JT_OK;
}

// Host-driven wavefront loop: generate, then (extend, shade, probe, advance) until the extend queue is
// empty. Queue sizes stay on the device; the host only polls the next-queue length every few iterations.
// The image is split over npipe pipelines that run on their own streams with a share of the persistent grid,
// so one pipeline's latency-bound shade / probe kernels and kernel tails overlap the other's issue-bound extend.
template <int MODE>
static int launch_wavefront(jt_scene* sc, jt_state* st, const DevParams& P, int begin, int end) {
  int rc = wf_prepare(sc, st);
  if (rc) return rc;
  const int npipe = st->npipe;
  constexpr int poll_every = 4;
  constexpr int ev_per_pipe = 2 * poll_every * JT_WF_RING;
  static_assert(poll_every == 4 && ev_per_pipe * JT_MAX_PIPES == JT_EXT_EVENTS, "ext_ev holds the start/stop events of every batch in flight");
  int cur[JT_MAX_PIPES], remaining[JT_MAX_PIPES];
  bool active[JT_MAX_PIPES];
  // the other pipelines start after everything already enqueued on the main stream
  for (int k = 1; k < npipe; k++) {
    JT_CUDA(cudaEventRecord(st->pipe_done[0], sc->stream));
    JT_CUDA(cudaStreamWaitEvent(st->pipe_stream[k], st->pipe_done[0], 0));
  }
  for (int k = 0; k < npipe; k++) {
    WfBuffers& B = st->wf[k];
    k_wf_generate<<<(B.n + 255) / 256, 256, 0, st->pipe_stream[k]>>>(sc->dev, B, P, begin, end, sc->d_counters);
    sc->launches++;
    cur[k] = 0;
    remaining[k] = B.n;
    active[k] = true;
  }
  FILE* iter_log = getenv("JT_ITER_LOG") ? fopen(getenv("JT_ITER_LOG"), "a") : nullptr;
  const auto iter_t0 = std::chrono::steady_clock::now();
  const int pdiv = getenv("JT_PGRID_DIV") ? atoi(getenv("JT_PGRID_DIV")) : 1;
  const unsigned pgrid = std::max(1u, persist_grid(sc) / (unsigned)std::max(1, pdiv));
  // One batch = poll_every iterations of every active pipeline, followed by an async copy of the queue counters.
  // The host enqueues batch j + 1 BEFORE it waits for the counters of batch j, so the GPU never drains while the
  // host looks at queue lengths; the price is one batch of empty launches after a pipeline has run dry.
  auto enqueue_batch = [&](int batch) -> int {
    for (int sub = 0; sub < poll_every; sub++) {
      const int it = batch * poll_every + sub;
      for (int k = 0; k < npipe; k++) {
        if (!active[k]) continue;
        WfBuffers& B = st->wf[k];
        cudaStream_t q = st->pipe_stream[k];
        const int next = cur[k] ^ 1;
        // grids sized from the last polled number of slots that are not idle for good (an upper bound of every queue:
        // a slot appears at most once per iteration, and idle slots never come back within a range)
        unsigned ge = (unsigned)((remaining[k] + 127) / 128);
        unsigned gs = (unsigned)((remaining[k] + JT_SHADE_BLOCK - 1) / JT_SHADE_BLOCK + WF_NKEY);  // every key's queue is padded to a whole block
        unsigned gpr = (unsigned)((remaining[k] + JT_PROBE_BLOCK - 1) / JT_PROBE_BLOCK);
        const int evi = 2 * (it % (poll_every * JT_WF_RING));
        JT_CUDA(cudaEventRecord(sc->ext_ev[ev_per_pipe * k + evi], q));
        if (MODE == MODE_WIDE) {
          unsigned gp = std::min<unsigned>((unsigned)((remaining[k] + JT_PERSIST_BLOCK - 1) / JT_PERSIST_BLOCK), pgrid);
          k_wf_extend_persist<<<gp, JT_PERSIST_BLOCK, 0, q>>>(sc->dev, B, cur[k], sc->d_counters);
        } else {
          k_wf_extend<MODE><<<ge, 128, 0, q>>>(sc->dev, B, cur[k], sc->d_counters);
        }
        JT_CUDA(cudaEventRecord(sc->ext_ev[ev_per_pipe * k + evi + 1], q));
        if (P.sampler == 1) {
          k_wf_shade<1, MODE><<<gs, JT_SHADE_BLOCK, 0, q>>>(sc->dev, B, st->dev, P, next, end, sc->d_counters);
          k_wf_probe<MODE><<<gpr, JT_PROBE_BLOCK, 0, q>>>(sc->dev, B, st->dev, P, next, end, sc->d_counters);
          sc->launches += 4;
        } else {
          k_wf_shade<2, MODE><<<gs, JT_SHADE_BLOCK, 0, q>>>(sc->dev, B, st->dev, P, next, end, sc->d_counters);
          sc->launches += 3;
        }
        k_wf_regen<<<(unsigned)((B.n + WF_REGEN_BLOCK * WF_REGEN_PER_THREAD - 1) / (WF_REGEN_BLOCK * WF_REGEN_PER_THREAD)), WF_REGEN_BLOCK, 0, q>>>(sc->dev, B, st->dev, P, cur[k], end, it, sc->d_counters);
        cur[k] = next;
      }
      st->wf_iterations++;
    }
    int* hc = st->h_counts + (size_t)(batch % JT_WF_RING) * JT_MAX_PIPES * WF_C_TOTAL;
    for (int k = 0; k < npipe; k++) {
      if (!active[k]) continue;
      JT_CUDA(cudaMemcpyAsync(hc + k * WF_C_TOTAL, st->wf[k].counts, WF_C_TOTAL * 4, cudaMemcpyDeviceToHost, st->pipe_stream[k]));
      JT_CUDA(cudaEventRecord(st->poll_ev[batch % JT_WF_RING][k], st->pipe_stream[k]));
    }
    return JT_OK;
  };
  // returns through *any whether some pipeline still has queued rays after `batch`
  auto collect_batch = [&](int batch, bool polled[JT_MAX_PIPES], bool* any) -> int {
    const int* hc = st->h_counts + (size_t)(batch % JT_WF_RING) * JT_MAX_PIPES * WF_C_TOTAL;
    *any = false;
    for (int k = 0; k < npipe; k++) {
      if (!polled[k]) continue;
      JT_CUDA(cudaEventSynchronize(st->poll_ev[batch % JT_WF_RING][k]));
      // per-launch duration of every extend launch of every pipeline (the roofline of bench.py divides the bytes of
      // ALL scene rays by this sum; launches of different pipelines overlap in wall time, each is timed on its stream)
      float batch_ext_ms = 0.0f;
      for (int sub = 0; sub < poll_every; sub++) {
        const int evi = ev_per_pipe * k + 2 * ((batch * poll_every + sub) % (poll_every * JT_WF_RING));
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sc->ext_ev[evi], sc->ext_ev[evi + 1]) == cudaSuccess) sc->extend_ms += ms;
        batch_ext_ms += ms;
        sc->extend_launches++;
      }
      remaining[k] = std::min(remaining[k], st->wf[k].n - hc[k * WF_C_TOTAL + WF_C_DONE]);
      if (iter_log)  // JT_ITER_LOG=<file>: batch, pipeline, queue length after the batch, extend ms of the batch, host ms
        fprintf(iter_log, "%d %d %d %.4f %.3f\n", batch, k, hc[k * WF_C_TOTAL + WF_C_EXT(cur[k])], batch_ext_ms,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - iter_t0).count());
      if (remaining[k] == 0) active[k] = false;
      *any = *any || active[k];
    }
    return JT_OK;
  };
  {
    bool polled[JT_WF_RING][JT_MAX_PIPES];
    int rc2;
    for (int b = 0; b < JT_WF_RING; b++)
      for (int k = 0; k < JT_MAX_PIPES; k++) polled[b][k] = false;
    // prologue: JT_WF_LOOKAHEAD batches in flight before the first wait
    for (int batch = 0; batch < JT_WF_LOOKAHEAD; batch++) {
      for (int k = 0; k < npipe; k++) polled[batch % JT_WF_RING][k] = active[k];
      if ((rc2 = enqueue_batch(batch))) return rc2;
    }
    for (int batch = JT_WF_LOOKAHEAD;; batch++) {
      for (int k = 0; k < JT_MAX_PIPES; k++) polled[batch % JT_WF_RING][k] = k < npipe && active[k];
      if ((rc2 = enqueue_batch(batch))) return rc2;
      bool any = false;
      if ((rc2 = collect_batch(batch - JT_WF_LOOKAHEAD, polled[(batch - JT_WF_LOOKAHEAD) % JT_WF_RING], &any))) return rc2;
      if (!any) {
        // the batches already enqueued ran on empty queues; drain their polls so the events / staging can be reused
        for (int b = batch - JT_WF_LOOKAHEAD + 1; b <= batch; b++)
          if ((rc2 = collect_batch(b, polled[b % JT_WF_RING], &any))) return rc2;
        break;
      }
    }
  }
  if (iter_log) {
    fprintf(iter_log, "end\n");
    fclose(iter_log);
  }
  // later work on the main stream (download, the next range) waits for every pipeline
  for (int k = 1; k < npipe; k++) {
    JT_CUDA(cudaEventRecord(st->pipe_done[k], st->pipe_stream[k]));
    JT_CUDA(cudaStreamWaitEvent(sc->stream, st->pipe_done[k], 0));
  }
  JT_CUDA(cudaGetLastError());
  return JT_OK;
}

static int launch_range(jt_scene* sc, jt_state* st, const jt_params* p, int begin, int end) {
  if (end <= begin) return JT_OK;
  DevParams P = dev_params(p, st);
  if (!sc->timing_open) {
    JT_CUDA(cudaEventRecord(sc->ev_start, sc->stream));
    sc->timing_open = true;
  }
  if (p->integrator == 1) {
    dim3 grid((unsigned)((st->width + 15) / 16), (unsigned)((st->height + 7) / 8));
    if (p->traversal == 1) k_trace_mega<MODE_REF><<<grid, 128, 0, sc->stream>>>(sc->dev, st->dev, P, begin, end, sc->d_counters);
    else k_trace_mega<MODE_WIDE><<<grid, 128, 0, sc->stream>>>(sc->dev, st->dev, P, begin, end, sc->d_counters);
    sc->launches++;
    JT_CUDA(cudaGetLastError());
  } else {
    int rc = p->traversal == 1 ? launch_wavefront<MODE_REF>(sc, st, P, begin, end) : launch_wavefront<MODE_WIDE>(sc, st, P, begin, end);
    if (rc) return rc;
  }
  JT_CUDA(cudaEventRecord(sc->ev_stop, sc->stream));
  return JT_OK;
}

static int flush_state(jt_state* st) {
  if (!st->has_pending) return JT_OK;
  st->has_pending = false;
  jt_scene* sc = st->scene;
  JT_CUDA(cudaSetDevice(sc->device));
  return launch_range(sc, st, &st->pending_params, st->pending_begin, st->pending_end);
}

static bool same_render_params(const jt_params& a, const jt_params& b) {
  return a.camera == b.camera && a.resolution == b.resolution && a.bounces == b.bounces && a.sampler == b.sampler &&
         a.clamp == b.clamp && a.nocaustics == b.nocaustics && a.envhidden == b.envhidden &&
         a.tentfilter == b.tentfilter && a.traversal == b.traversal && a.seed == b.seed &&
         a.accumulate == b.accumulate && a.integrator == b.integrator;
}

extern "C" int jt_trace_sample_range(jt_scene* sc, jt_state* st, const jt_params* p, int32_t begin, int32_t end) {
  if (!sc || !st || st->scene != sc) return jt_set_error(JT_ERR_INVALID, "jt_trace_sample_range: bad scene/state");
  int rc = check_params(sc, p);
  if (rc) return rc;
  if (begin < 0 || end < begin) return jt_set_error(JT_ERR_INVALID, "bad sample range [%d, %d)", begin, end);
  if (end >= (1 << 23)) return jt_set_error(JT_ERR_INVALID, "sample indices are limited to 2^23 - 1 (got %d)", end);
  if (end == begin) return JT_OK;
  if (st->has_pending && (begin != st->pending_end || !same_render_params(*p, st->pending_params))) {
    if ((rc = flush_state(st))) return rc;
  }
  if (st->has_pending) {
    st->pending_end = end;
  } else {
    st->has_pending = true;
    st->pending_begin = begin;
    st->pending_end = end;
    st->pending_params = *p;
  }
  st->samples += end - begin;
  if (st->pending_end - st->pending_begin >= JT_LAZY_SPP) return flush_state(st);
  return JT_OK;
}

extern "C" int jt_trace_samples(jt_scene* sc, jt_state* st, const jt_params* p) {
  if (!sc || !st || st->scene != sc) return jt_set_error(JT_ERR_INVALID, "jt_trace_samples: bad scene/state");
  int rc = check_params(sc, p);
  if (rc) return rc;
  if (st->samples >= p->samples) return JT_OK;  // src/trace.jl:225-227
  int target = std::min(st->samples + std::max(p->batch, 1), p->samples);
  return jt_trace_sample_range(sc, st, p, st->samples, target);
}

extern "C" int jt_synchronize(jt_scene* sc) {
  if (!sc) return jt_set_error(JT_ERR_INVALID, "jt_synchronize: null argument");
  JT_CUDA(cudaSetDevice(sc->device));
  int frc = flush_scene(sc);
  if (frc) return frc;
  JT_CUDA(cudaStreamSynchronize(sc->stream));
  return JT_OK;
}

extern "C" int jt_elapsed_ms(jt_scene* sc, float* ms) {
  if (!sc || !ms) return jt_set_error(JT_ERR_INVALID, "jt_elapsed_ms: null argument");
  JT_CUDA(cudaSetDevice(sc->device));
  *ms = 0.0f;
  int frc = flush_scene(sc);
  if (frc) return frc;
  if (!sc->timing_open) return JT_OK;
  JT_CUDA(cudaEventSynchronize(sc->ev_stop));
  JT_CUDA(cudaEventElapsedTime(ms, sc->ev_start, sc->ev_stop));
  sc->timing_open = false;
  return JT_OK;
}

// =================================================================================================
// parity hooks
// =================================================================================================
extern "C" int jt_intersect_device(jt_scene* sc, const void* d_rays, int64_t n, int traversal, void* d_hits) {
  if (!sc || (n > 0 && (!d_rays || !d_hits)) || n < 0) return jt_set_error(JT_ERR_INVALID, "jt_intersect_device: bad argument");
  if (traversal < 0 || traversal > 2) return jt_set_error(JT_ERR_INVALID, "traversal must be 0, 1 or 2");
  if (n == 0) return JT_OK;
  JT_CUDA(cudaSetDevice(sc->device));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (!sc->timing_open) {
    JT_CUDA(cudaEventRecord(sc->ev_start, sc->stream));
    sc->timing_open = true;
  }
  if (traversal == 1) {
    k_intersect<MODE_REF><<<blocks, 128, 0, sc->stream>>>(sc->dev, (const jt_ray*)d_rays, n, (HitOut*)d_hits);
  } else if (traversal == 2) {  // plain one-thread-per-ray wide walk (kept for A/B measurements)
    k_intersect<MODE_WIDE><<<blocks, 128, 0, sc->stream>>>(sc->dev, (const jt_ray*)d_rays, n, (HitOut*)d_hits);
  } else {
    if (n > 0x7FFFFFFF) return jt_set_error(JT_ERR_INVALID, "jt_intersect_device: more than 2^31-1 rays per call");
    int* fetch = (int*)(sc->d_counters + 7);
    JT_CUDA(cudaMemsetAsync(fetch, 0, 4, sc->stream));
    if (sc->intersect_blocks == 0) {  // one resident wave of THIS kernel (its register count differs from the extend kernel's)
      int per_sm = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_intersect_persist, JT_PERSIST_BLOCK, 0);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device);
      sc->intersect_blocks = (unsigned)std::max(1, per_sm) * (unsigned)std::max(1, sms);
    }
    unsigned pblocks = std::min<unsigned>((unsigned)((n + JT_PERSIST_BLOCK - 1) / JT_PERSIST_BLOCK), sc->intersect_blocks);
    k_intersect_persist<<<pblocks, JT_PERSIST_BLOCK, 0, sc->stream>>>(sc->dev, (const jt_ray*)d_rays, n, (HitOut*)d_hits, fetch);
  }
  sc->launches++;
  JT_CUDA(cudaGetLastError());
  JT_CUDA(cudaEventRecord(sc->ev_stop, sc->stream));
  return JT_OK;
}

extern "C" int jt_intersect(jt_scene* sc, const jt_ray* rays, int64_t n, int traversal, jt_hit* out) {
  if (!sc || (n > 0 && (!rays || !out)) || n < 0) return jt_set_error(JT_ERR_INVALID, "jt_intersect: bad argument");
  if (n == 0) return JT_OK;
  JT_CUDA(cudaSetDevice(sc->device));
  void *dr = nullptr, *dh = nullptr;
  JT_CUDA(cudaMalloc(&dr, n * sizeof(jt_ray)));
  cudaError_t e = cudaMalloc(&dh, n * sizeof(jt_hit));
  if (e != cudaSuccess) {
    cudaFree(dr);
    return jt_set_error(JT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
  }
  int rc = JT_OK;
  if ((e = cudaMemcpyAsync(dr, rays, n * sizeof(jt_ray), cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  if (!rc) rc = jt_intersect_device(sc, dr, n, traversal, dh);
  if (!rc && (e = cudaMemcpyAsync(out, dh, n * sizeof(jt_hit), cudaMemcpyDeviceToHost, sc->stream)) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(sc->stream)) != cudaSuccess && !rc)
    rc = jt_set_error(JT_ERR_CUDA, "jt_intersect: %s", cudaGetErrorString(e));
  cudaFree(dr);
  cudaFree(dh);
  return rc;
}

extern "C" int jt_intersect_instance(jt_scene* sc, const jt_ray* rays, const int64_t* instances, int64_t n,
                                     int traversal, jt_hit* out) {
  if (!sc || (n > 0 && (!rays || !out || !instances)) || n < 0) return jt_set_error(JT_ERR_INVALID, "jt_intersect_instance: bad argument");
  if (traversal != 0 && traversal != 1) return jt_set_error(JT_ERR_INVALID, "traversal must be 0 or 1");
  if (n == 0) return JT_OK;
  JT_CUDA(cudaSetDevice(sc->device));
  void *dr = nullptr, *dh = nullptr, *di = nullptr;
  cudaError_t e;
  int rc = JT_OK;
  if ((e = cudaMalloc(&dr, n * sizeof(jt_ray))) != cudaSuccess || (e = cudaMalloc(&dh, n * sizeof(jt_hit))) != cudaSuccess ||
      (e = cudaMalloc(&di, n * 8)) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
  if (!rc && ((e = cudaMemcpyAsync(dr, rays, n * sizeof(jt_ray), cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess ||
              (e = cudaMemcpyAsync(di, instances, n * 8, cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess))
    rc = jt_set_error(JT_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  if (!rc) {
    unsigned blocks = (unsigned)((n + 127) / 128);
    if (traversal == 1) k_intersect_instance<MODE_REF><<<blocks, 128, 0, sc->stream>>>(sc->dev, (const jt_ray*)dr, (const long long*)di, n, (HitOut*)dh);
    else k_intersect_instance<MODE_WIDE><<<blocks, 128, 0, sc->stream>>>(sc->dev, (const jt_ray*)dr, (const long long*)di, n, (HitOut*)dh);
    sc->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) rc = jt_set_error(JT_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
  }
  if (!rc && (e = cudaMemcpyAsync(out, dh, n * sizeof(jt_hit), cudaMemcpyDeviceToHost, sc->stream)) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(sc->stream)) != cudaSuccess && !rc) rc = jt_set_error(JT_ERR_CUDA, "jt_intersect_instance: %s", cudaGetErrorString(e));
  cudaFree(dr); cudaFree(dh); cudaFree(di);
  return rc;
}

extern "C" int jt_sample_camera(jt_scene* sc, const jt_params* p, int32_t width, int32_t height, const int32_t* ij,
                                const float* puv_luv, int64_t n, jt_ray* out) {
  if (!sc || !p || (n > 0 && (!ij || !puv_luv || !out)) || n < 0) return jt_set_error(JT_ERR_INVALID, "jt_sample_camera: bad argument");
  if (p->camera < 1 || p->camera > sc->num_cameras) return jt_set_error(JT_ERR_INVALID, "camera out of range");
  if (n == 0) return JT_OK;
  JT_CUDA(cudaSetDevice(sc->device));
  void *dij = nullptr, *dr = nullptr, *dout = nullptr;
  cudaError_t e;
  int rc = JT_OK;
  if ((e = cudaMalloc(&dij, n * 8)) != cudaSuccess || (e = cudaMalloc(&dr, n * 16)) != cudaSuccess ||
      (e = cudaMalloc(&dout, n * sizeof(jt_ray))) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
  if (!rc && ((e = cudaMemcpyAsync(dij, ij, n * 8, cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess ||
              (e = cudaMemcpyAsync(dr, puv_luv, n * 16, cudaMemcpyHostToDevice, sc->stream)) != cudaSuccess))
    rc = jt_set_error(JT_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  if (!rc) {
    k_sample_camera<<<(unsigned)((n + 127) / 128), 128, 0, sc->stream>>>(sc->dev, p->camera - 1, p->tentfilter, width, height,
                                                                         (const int*)dij, (const float*)dr, n, (jt_ray*)dout);
    sc->launches++;
    if ((e = cudaGetLastError()) != cudaSuccess) rc = jt_set_error(JT_ERR_CUDA, "launch failed: %s", cudaGetErrorString(e));
  }
  if (!rc && (e = cudaMemcpyAsync(out, dout, n * sizeof(jt_ray), cudaMemcpyDeviceToHost, sc->stream)) != cudaSuccess)
    rc = jt_set_error(JT_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
  if ((e = cudaStreamSynchronize(sc->stream)) != cudaSuccess && !rc) rc = jt_set_error(JT_ERR_CUDA, "jt_sample_camera: %s", cudaGetErrorString(e));
  cudaFree(dij); cudaFree(dr); cudaFree(dout);
  return rc;
}
