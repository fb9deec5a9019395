// jt_group.cu -- the in-library multi-GPU path (include/jtrace_b200.h "multi-GPU group").
//
// The reference's caller (src/jtrace.jl:83-94) is one thread calling trace_samples in a loop, so the sharding cannot
// live in the host program: a jt_group owns one (device, scene) member per requested device, one worker thread per
// member, and splits every requested range of GLOBAL sample indices into contiguous per-member sub-ranges. Members
// accumulate sums (jt_params.accumulate = 1). The merge is ONE kernel on member 0, k_group_finalize: it reads every
// member's image / albedo / normal / hits buffers through peer-mapped pointers (NVLink P2P loads through NVSwitch;
// 52 B per pixel per remote member, 47.9 MB for 1280x720 on 8 GPUs), adds them in member order (deterministic for a
// given group), applies 1/samples and writes the reference's host layouts (or the 8-bit sRGB image) into the packed
// staging buffer that is copied to the host -- reduce, divide and pack fused, no intermediate buffer, no NCCL ring.
// Members on devices without peer access are copied into a staging buffer on member 0's device first.
//
// Everything below goes through the single-device entry points of jt_api.cu; only the staged-scene upload is internal.
#include <cuda_runtime.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "jt_dev_output.cuh"
#include "jt_internal.h"

int jt_scene_create_staged(const JtStagedScene& staged, int device, jt_scene** out);
int jt_stage_scene_checked(const jt_scene_desc* desc, JtStagedScene* staged);

#define JT_GROUP_MAX 16
#define JT_GROUP_LAZY_SPP 512 /* per member: the single-device chunk size (jt_api.cu JT_LAZY_SPP) */

namespace {

struct Member {
  int index = 0, device = 0;
  jt_scene* scene = nullptr;
  bool peer = false;    // member 0 reads this member's buffers through peer access
  bool staged = false;  // ... through a copy into member 0's device memory
  std::thread th;
  std::mutex m;
  std::condition_variable cv_job, cv_idle;
  std::deque<std::function<int()>> jobs;
  bool busy = false, quit = false;
  int err = 0;
  std::string errmsg;

  void run() {
    for (;;) {
      std::function<int()> job;
      {
        std::unique_lock<std::mutex> lk(m);
        cv_job.wait(lk, [&] { return quit || !jobs.empty(); });
        if (jobs.empty()) return;  // quit
        job = std::move(jobs.front());
        jobs.pop_front();
        busy = true;
      }
      int rc = err ? err : job();  // after a failure the remaining jobs are skipped
      {
        std::lock_guard<std::mutex> lk(m);
        if (rc && !err) {
          err = rc;
          errmsg = jt_last_error();  // thread-local in the worker: carry it to the caller's thread
        }
        busy = false;
        if (jobs.empty()) cv_idle.notify_all();
      }
    }
  }
  void post(std::function<int()> job) {
    {
      std::lock_guard<std::mutex> lk(m);
      jobs.push_back(std::move(job));
    }
    cv_job.notify_one();
  }
  void wait_idle() {
    std::unique_lock<std::mutex> lk(m);
    cv_idle.wait(lk, [&] { return jobs.empty() && !busy; });
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
    }
    cv_job.notify_one();
    if (th.joinable()) th.join();
  }
};

struct GroupPtrs {  // by-value kernel argument: the members' accumulators as seen from member 0's device
  const float4* image[JT_GROUP_MAX];
  const float4* albedo[JT_GROUP_MAX];
  const float4* normal[JT_GROUP_MAX];
  const int* hits[JT_GROUP_MAX];
  int n;
};

}  // namespace

struct jt_group {
  std::vector<Member*> members;
  std::vector<jt_group_state*> states;
  cudaStream_t stream = nullptr;  // on member 0's device: the merge kernel and the download copies
  jt_group_stats stats;
};

struct jt_group_state {
  jt_group* group = nullptr;
  std::vector<jt_state*> st;  // one sum-mode state per member
  int width = 0, height = 0, samples = 0;
  bool has_pending = false;
  int pending_begin = 0, pending_end = 0;
  jt_params pending_params;
  void* d_pack = nullptr;   // member 0's device: packed host layouts (48 B per pixel)
  void* h_pack = nullptr;   // pinned
  std::vector<void*> d_stage;  // per member: 52 B per pixel on member 0's device for members without peer access
};

#define JTG_CUDA(call)                                                                                    \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return jt_set_error(JT_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                          __LINE__);                                                                      \
  } while (0)

// ---- the fused reduce + finalize kernel ---------------------------------------------------------------------------
// One thread per pixel: for every member (in member order) one 128-bit load from each of its three float4 buffers and
// one 32-bit load of its hit count -- remote members over NVLink --, then the mean and the packed stores. HBM/NVLink
// bound: 52 B read per member + 48 B (or 4 B) written per pixel.
__global__ void __launch_bounds__(256) k_group_finalize(GroupPtrs P, long long n, float scale, float* out_image,
                                                         float* out_albedo, float* out_normal, long long* out_hits,
                                                         uchar4* out_srgb8) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool aov = out_albedo || out_normal || out_hits;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, c = a;
  long long h = 0;
  for (int g = 0; g < P.n; g++) {
    float4 x = P.image[g][i];
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    if (aov) {
      float4 y = P.albedo[g][i], z = P.normal[g][i];
      b.x += y.x; b.y += y.y; b.z += y.z;
      c.x += z.x; c.y += z.y; c.z += z.z;
      h += P.hits[g][i];
    }
  }
  a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
  if (out_image) {
    out_image[4 * i] = a.x; out_image[4 * i + 1] = a.y; out_image[4 * i + 2] = a.z; out_image[4 * i + 3] = a.w;
  }
  if (out_albedo) {
    out_albedo[3 * i] = b.x * scale; out_albedo[3 * i + 1] = b.y * scale; out_albedo[3 * i + 2] = b.z * scale;
  }
  if (out_normal) {
    out_normal[3 * i] = c.x * scale; out_normal[3 * i + 1] = c.y * scale; out_normal[3 * i + 2] = c.z * scale;
  }
  if (out_hits) out_hits[i] = h;
  if (out_srgb8) out_srgb8[i] = jt_srgb8_pixel(a);
}

// ---- group ----------------------------------------------------------------------------------------------------------
static int first_error(jt_group* g) {
  for (Member* m : g->members) {
    std::lock_guard<std::mutex> lk(m->m);
    if (m->err) {
      int code = m->err;
      return jt_set_error(code, "group member %d (device %d): %s", m->index, m->device, m->errmsg.c_str());
    }
  }
  return JT_OK;
}

static int wait_all(jt_group* g) {
  for (Member* m : g->members) m->wait_idle();
  return first_error(g);
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" void jt_group_destroy(jt_group* g) {
  if (!g) return;
  for (Member* m : g->members) m->wait_idle();
  // states outlive the group as orphans (same rule as jt_state / jt_scene): release what they own on the devices now
  for (jt_group_state* s : g->states) {
    for (jt_state* st : s->st) jt_state_destroy(st);
    s->st.clear();
    if (!g->members.empty()) cudaSetDevice(g->members[0]->device);
    if (s->d_pack) cudaFree(s->d_pack);
    if (s->h_pack) cudaFreeHost(s->h_pack);
    for (void* p : s->d_stage)
      if (p) cudaFree(p);
    s->d_pack = s->h_pack = nullptr;
    s->d_stage.clear();
    s->group = nullptr;
  }
  for (Member* m : g->members) {
    m->stop();
    if (m->scene) jt_scene_destroy(m->scene);
  }
  if (g->stream && !g->members.empty()) {
    cudaSetDevice(g->members[0]->device);
    cudaStreamDestroy(g->stream);
  }
  for (Member* m : g->members) delete m;
  delete g;
}

extern "C" int jt_group_create(const jt_scene_desc* desc, const int* devices, int n, jt_group** out) {
  if (!desc || !out || !devices) return jt_set_error(JT_ERR_INVALID, "jt_group_create: null argument");
  *out = nullptr;
  if (n < 1 || n > JT_GROUP_MAX) return jt_set_error(JT_ERR_INVALID, "jt_group_create: %d members (1..%d)", n, JT_GROUP_MAX);
  int ndev = jt_device_count();
  if (ndev <= 0) return jt_set_error(JT_ERR_NO_DEVICE, "no CUDA device visible: libjtrace_b200 has no CPU fallback");
  for (int k = 0; k < n; k++)
    if (devices[k] < 0 || devices[k] >= ndev)
      return jt_set_error(JT_ERR_INVALID, "jt_group_create: device %d out of range (0..%d)", devices[k], ndev - 1);
  // stage once (flatten + wide-BVH build on the host) ...
  double t0 = now_s();
  JtStagedScene* staged = new (std::nothrow) JtStagedScene();
  if (!staged) return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  int rc = jt_stage_scene_checked(desc, staged);
  if (rc) {
    delete staged;
    return rc;
  }
  double t1 = now_s();
  jt_group* g = new (std::nothrow) jt_group();
  if (!g) {
    delete staged;
    return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  }
  memset(&g->stats, 0, sizeof(g->stats));
  g->stats.members = n;
  g->stats.stage_seconds = t1 - t0;
  // ... and upload N times in parallel, each member on its own worker thread
  for (int k = 0; k < n; k++) {
    Member* m = new Member();
    m->index = k;
    m->device = devices[k];
    g->members.push_back(m);
    m->th = std::thread([m] { m->run(); });
    m->post([m, staged] { return jt_scene_create_staged(*staged, m->device, &m->scene); });
  }
  rc = wait_all(g);
  delete staged;
  g->stats.upload_seconds = now_s() - t1;
  if (rc) {
    std::string keep = jt_last_error();
    jt_group_destroy(g);
    return jt_set_error(rc, "%s", keep.c_str());
  }
  // peer access from member 0's device to every other member's device
  const int dev0 = g->members[0]->device;
  std::vector<int> seen;
  cudaError_t e = cudaSetDevice(dev0);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    jt_group_destroy(g);
    return jt_set_error(JT_ERR_CUDA, "jt_group_create: %s", cudaGetErrorString(e));
  }
  for (Member* m : g->members) {
    bool dup = false;
    for (int d : seen) dup = dup || d == m->device;
    if (!dup) seen.push_back(m->device);
    if (m->device == dev0) continue;  // same device: plain pointers
    int can = 0;
    cudaDeviceCanAccessPeer(&can, dev0, m->device);
    if (can) {
      e = cudaDeviceEnablePeerAccess(m->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        e = cudaSuccess;
      }
      can = e == cudaSuccess;
      if (!can) cudaGetLastError();
    }
    m->peer = can != 0;
    m->staged = !m->peer;
    if (m->peer) g->stats.peer_members++;
    else g->stats.staged_members++;
  }
  g->stats.distinct_devices = (int32_t)seen.size();
  *out = g;
  return JT_OK;
}

extern "C" int jt_group_get_stats(jt_group* g, jt_group_stats* out) {
  if (!g || !out) return jt_set_error(JT_ERR_INVALID, "jt_group_get_stats: null argument");
  *out = g->stats;
  return JT_OK;
}

extern "C" int jt_group_scene(jt_group* g, int member, jt_scene** out) {
  if (!g || !out) return jt_set_error(JT_ERR_INVALID, "jt_group_scene: null argument");
  if (member < 0 || member >= (int)g->members.size()) return jt_set_error(JT_ERR_INVALID, "jt_group_scene: member %d out of range", member);
  *out = g->members[(size_t)member]->scene;
  return JT_OK;
}

static int group_flush(jt_group_state* s);

extern "C" int jt_group_synchronize(jt_group* g) {
  if (!g) return jt_set_error(JT_ERR_INVALID, "jt_group_synchronize: null argument");
  for (jt_group_state* s : g->states) {
    int rc = group_flush(s);
    if (rc) return rc;
  }
  return wait_all(g);
}

extern "C" int jt_group_counters(jt_group* g, jt_counters* out, int reset) {
  if (!g || !out) return jt_set_error(JT_ERR_INVALID, "jt_group_counters: null argument");
  int rc = jt_group_synchronize(g);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  for (Member* m : g->members) {
    jt_counters c;
    if ((rc = jt_scene_counters(m->scene, &c, reset))) return rc;
    out->camera_paths += c.camera_paths;
    out->scene_rays += c.scene_rays;
    out->light_rays += c.light_rays;
    out->kernel_launches += c.kernel_launches;
    out->extend_launches += c.extend_launches;
    out->stolen_samples += c.stolen_samples;
    out->resumed_rays += c.resumed_rays;
    if (c.extend_kernel_us > out->extend_kernel_us) out->extend_kernel_us = c.extend_kernel_us;
  }
  return JT_OK;
}

// ---- group state ------------------------------------------------------------------------------------------------------
#define JTG_LIVE(s, who)                                                                            \
  do {                                                                                              \
    if (!(s)) return jt_set_error(JT_ERR_INVALID, who ": null argument");                           \
    if (!(s)->group) return jt_set_error(JT_ERR_INVALID, who ": the state's group was destroyed");  \
  } while (0)

extern "C" void jt_group_state_destroy(jt_group_state* s) {
  if (!s) return;
  if (s->group) {
    jt_group* g = s->group;
    for (Member* m : g->members) m->wait_idle();
    for (size_t i = 0; i < g->states.size(); i++)
      if (g->states[i] == s) {
        g->states.erase(g->states.begin() + (long)i);
        break;
      }
    for (jt_state* st : s->st) jt_state_destroy(st);
    cudaSetDevice(g->members[0]->device);
    if (s->d_pack) cudaFree(s->d_pack);
    if (s->h_pack) cudaFreeHost(s->h_pack);
    for (void* p : s->d_stage)
      if (p) cudaFree(p);
  }
  delete s;
}

extern "C" int jt_group_state_create(jt_group* g, const jt_params* p, jt_group_state** out) {
  if (!g || !p || !out) return jt_set_error(JT_ERR_INVALID, "jt_group_state_create: null argument");
  *out = nullptr;
  int rc = wait_all(g);
  if (rc) return rc;
  jt_group_state* s = new (std::nothrow) jt_group_state();
  if (!s) return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  s->group = g;
  jt_params q = *p;
  q.accumulate = 1;  // members hold sums; the running mean of Q13 differs from sum / N by rounding only
  for (Member* m : g->members) {
    jt_state* st = nullptr;
    rc = jt_state_create(m->scene, &q, &st);
    if (rc) {
      std::string keep = jt_last_error();
      for (jt_state* t : s->st) jt_state_destroy(t);
      delete s;
      return jt_set_error(rc, "%s", keep.c_str());
    }
    s->st.push_back(st);
  }
  int32_t w = 0, h = 0;
  jt_state_size(s->st[0], &w, &h, nullptr);
  s->width = w;
  s->height = h;
  s->d_stage.assign(g->members.size(), nullptr);
  g->states.push_back(s);
  *out = s;
  return JT_OK;
}

extern "C" int jt_group_state_size(jt_group_state* s, int32_t* width, int32_t* height, int32_t* samples) {
  if (!s) return jt_set_error(JT_ERR_INVALID, "jt_group_state_size: null argument");
  if (width) *width = s->width;
  if (height) *height = s->height;
  if (samples) *samples = s->samples;
  return JT_OK;
}

extern "C" int jt_group_state_reset(jt_group_state* s) {
  JTG_LIVE(s, "jt_group_state_reset");
  s->has_pending = false;
  int rc = wait_all(s->group);
  if (rc) return rc;
  for (jt_state* st : s->st)
    if ((rc = jt_state_reset(st))) return rc;
  s->samples = 0;
  return JT_OK;
}

// Split the pending range into contiguous per-member sub-ranges and hand them to the workers. Each job renders its
// share and waits for the device, so "every worker idle" == "every sample accumulated".
static int group_flush(jt_group_state* s) {
  if (!s->has_pending) return JT_OK;
  s->has_pending = false;
  jt_group* g = s->group;
  const int n = (int)g->members.size();
  const long long b = s->pending_begin, c = (long long)s->pending_end - s->pending_begin;
  jt_params q = s->pending_params;
  q.accumulate = 1;
  for (int k = 0; k < n; k++) {
    const int32_t lo = (int32_t)(b + c * k / n), hi = (int32_t)(b + c * (k + 1) / n);
    if (hi <= lo) continue;
    Member* m = g->members[(size_t)k];
    jt_state* st = s->st[(size_t)k];
    m->post([m, st, q, lo, hi] {
      int rc = jt_trace_sample_range(m->scene, st, &q, lo, hi);
      if (rc) return rc;
      return jt_synchronize(m->scene);
    });
  }
  return JT_OK;
}

static bool same_group_params(const jt_params& a, const jt_params& b) {
  return a.camera == b.camera && a.resolution == b.resolution && a.bounces == b.bounces && a.sampler == b.sampler &&
         a.clamp == b.clamp && a.nocaustics == b.nocaustics && a.envhidden == b.envhidden &&
         a.tentfilter == b.tentfilter && a.traversal == b.traversal && a.seed == b.seed && a.integrator == b.integrator;
}

extern "C" int jt_group_trace_sample_range(jt_group* g, jt_group_state* s, const jt_params* p, int32_t begin, int32_t end) {
  if (!g || !s || s->group != g || !p) return jt_set_error(JT_ERR_INVALID, "jt_group_trace_sample_range: bad group/state");
  if (begin < 0 || end < begin) return jt_set_error(JT_ERR_INVALID, "bad sample range [%d, %d)", begin, end);
  int rc = first_error(g);
  if (rc) return rc;
  if (end == begin) return JT_OK;
  if (s->has_pending && (begin != s->pending_end || !same_group_params(*p, s->pending_params))) {
    if ((rc = group_flush(s))) return rc;
  }
  if (s->has_pending) {
    s->pending_end = end;
  } else {
    s->has_pending = true;
    s->pending_begin = begin;
    s->pending_end = end;
    s->pending_params = *p;
  }
  s->samples += end - begin;
  if (s->pending_end - s->pending_begin >= JT_GROUP_LAZY_SPP * (int)g->members.size()) return group_flush(s);
  return JT_OK;
}

extern "C" int jt_group_trace_samples(jt_group* g, jt_group_state* s, const jt_params* p) {
  if (!g || !s || s->group != g || !p) return jt_set_error(JT_ERR_INVALID, "jt_group_trace_samples: bad group/state");
  if (s->samples >= p->samples) return JT_OK;  // src/trace.jl:225-227
  int batch = p->batch > 1 ? p->batch : 1;
  int target = s->samples + batch < p->samples ? s->samples + batch : p->samples;
  return jt_group_trace_sample_range(g, s, p, s->samples, target);
}

// Fused merge on member 0. Exactly one of (image / albedo / normal / hits) or rgba8 is requested.
static int group_download(jt_group_state* s, float* image, float* albedo, float* normal, int64_t* hits, uint8_t* rgba8) {
  jt_group* g = s->group;
  int rc = group_flush(s);
  if (rc) return rc;
  if ((rc = wait_all(g))) return rc;
  const int n = (int)g->members.size();
  const long long npix = (long long)s->width * s->height;
  const int dev0 = g->members[0]->device;
  JTG_CUDA(cudaSetDevice(dev0));
  GroupPtrs P;
  memset(&P, 0, sizeof(P));
  P.n = n;
  const bool aov = albedo || normal || hits;
  int64_t remote = 0;
  for (int k = 0; k < n; k++) {
    void *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr;
    if ((rc = jt_state_device_buffers(s->st[(size_t)k], &a, &b, &c, &d, nullptr))) return rc;  // flushed + idle
    JTG_CUDA(cudaSetDevice(dev0));
    Member* m = g->members[(size_t)k];
    if (m->staged) {  // no peer mapping: bring the member's buffers over with peer copies (through the host if needed)
      if (!s->d_stage[(size_t)k]) JTG_CUDA(cudaMalloc(&s->d_stage[(size_t)k], (size_t)npix * 52));
      char* base = (char*)s->d_stage[(size_t)k];
      JTG_CUDA(cudaMemcpyPeerAsync(base, dev0, a, m->device, (size_t)npix * 16, g->stream));
      if (aov) {
        JTG_CUDA(cudaMemcpyPeerAsync(base + npix * 16, dev0, b, m->device, (size_t)npix * 16, g->stream));
        JTG_CUDA(cudaMemcpyPeerAsync(base + npix * 32, dev0, c, m->device, (size_t)npix * 16, g->stream));
        JTG_CUDA(cudaMemcpyPeerAsync(base + npix * 48, dev0, d, m->device, (size_t)npix * 4, g->stream));
      }
      a = base; b = base + npix * 16; c = base + npix * 32; d = base + npix * 48;
    }
    if (m->device != dev0) remote += npix * (aov ? 52 : 16);
    P.image[k] = (const float4*)a; P.albedo[k] = (const float4*)b; P.normal[k] = (const float4*)c; P.hits[k] = (const int*)d;
  }
  if (!s->d_pack) JTG_CUDA(cudaMalloc(&s->d_pack, (size_t)npix * 48));
  if (!s->h_pack) JTG_CUDA(cudaHostAlloc(&s->h_pack, (size_t)npix * 48, cudaHostAllocDefault));
  char* dp = (char*)s->d_pack;
  float* di = (float*)dp;
  float* da = (float*)(dp + npix * 16);
  float* dn = (float*)(dp + npix * 28);
  long long* dh = (long long*)(dp + npix * 40);
  const float scale = s->samples > 0 ? 1.0f / (float)s->samples : 1.0f;
  k_group_finalize<<<(unsigned)((npix + 255) / 256), 256, 0, g->stream>>>(
      P, npix, scale, image ? di : nullptr, albedo ? da : nullptr, normal ? dn : nullptr, hits ? dh : nullptr,
      rgba8 ? (uchar4*)dp : nullptr);
  JTG_CUDA(cudaGetLastError());
  char* hp = (char*)s->h_pack;
  if (rgba8) JTG_CUDA(cudaMemcpyAsync(hp, dp, (size_t)npix * 4, cudaMemcpyDeviceToHost, g->stream));
  if (image) JTG_CUDA(cudaMemcpyAsync(hp, di, (size_t)npix * 16, cudaMemcpyDeviceToHost, g->stream));
  if (albedo) JTG_CUDA(cudaMemcpyAsync(hp + npix * 16, da, (size_t)npix * 12, cudaMemcpyDeviceToHost, g->stream));
  if (normal) JTG_CUDA(cudaMemcpyAsync(hp + npix * 28, dn, (size_t)npix * 12, cudaMemcpyDeviceToHost, g->stream));
  if (hits) JTG_CUDA(cudaMemcpyAsync(hp + npix * 40, dh, (size_t)npix * 8, cudaMemcpyDeviceToHost, g->stream));
  JTG_CUDA(cudaStreamSynchronize(g->stream));
  if (rgba8) memcpy(rgba8, hp, (size_t)npix * 4);
  if (image) memcpy(image, hp, (size_t)npix * 16);
  if (albedo) memcpy(albedo, hp + npix * 16, (size_t)npix * 12);
  if (normal) memcpy(normal, hp + npix * 28, (size_t)npix * 12);
  if (hits) memcpy(hits, hp + npix * 40, (size_t)npix * 8);
  g->stats.reduce_bytes_remote += remote;
  g->stats.downloads++;
  return JT_OK;
}

extern "C" int jt_group_state_download(jt_group_state* s, float* image, float* albedo, float* normal, int64_t* hits) {
  JTG_LIVE(s, "jt_group_state_download");
  return group_download(s, image, albedo, normal, hits, nullptr);
}

extern "C" int jt_group_state_download_srgb8(jt_group_state* s, uint8_t* rgba8) {
  JTG_LIVE(s, "jt_group_state_download_srgb8");
  if (!rgba8) return jt_set_error(JT_ERR_INVALID, "jt_group_state_download_srgb8: null argument");
  return group_download(s, nullptr, nullptr, nullptr, nullptr, rgba8);
}
