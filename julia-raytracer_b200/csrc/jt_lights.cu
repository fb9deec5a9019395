// jt_lights.cu -- make_trace_lights (src/trace.jl:117-187) on the GPU.
//
// The per-element weights (shape-local triangle / quad areas; max(texel) * sin(theta) per environment texel) are computed
// by one thread per element in the reference's Float32 operation order (the TU is compiled -fmad=false). The CDF itself
// is a SEQUENTIAL Float32 prefix sum in the reference (`cdf[i] = cdf[i-1] + w[i]`, :172-181) and a tree scan would round
// differently, so it stays sequential: one warp per light streams the weights with coalesced loads and every lane replays
// the same chain of 32 additions through shuffles -- the exact order of the host loop, bit for bit -- keeping the value
// of its own element. 131 072 texels take ~0.2 ms.
//
// JT_LIGHTS_ENV_LUMINANCE fixes quirk Q8 behind a flag: the reference weights an environment texel by the maximum over
// RGBA *including alpha = 1*, which makes the CDF ignore the image whenever the texels are <= 1 (always, with the
// reference loader's clamped HDR values); with the flag the weight is the maximum over RGB. Sampling and pdf both read
// the same CDF, so the estimator stays unbiased; results are no longer the reference's.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "jt_internal.h"

namespace {

__device__ __forceinline__ float area3(const float* p0, const float* p1, const float* p2) {  // triangle_area, src/geometry.jl:260-262
  float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
  float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
  float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  float d = (cx * cx + cy * cy) + cz * cz;
  return sqrtf(d) / 2.0f;
}

// elements: 1-based vertex ids, `width` (3 or 4) per element
__global__ void k_area_weights(const float* __restrict__ pos, const long long* __restrict__ elems, int width, long long n,
                               float* __restrict__ w) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long* e = elems + (long long)width * i;
  if (width == 3) {
    w[i] = area3(pos + 3 * (e[0] - 1), pos + 3 * (e[1] - 1), pos + 3 * (e[2] - 1));
  } else {  // quad_area, src/geometry.jl:264-267
    w[i] = area3(pos + 3 * (e[0] - 1), pos + 3 * (e[1] - 1), pos + 3 * (e[3] - 1)) +
           area3(pos + 3 * (e[2] - 1), pos + 3 * (e[3] - 1), pos + 3 * (e[1] - 1));
  }
}

__global__ void k_env_weights(const float4* __restrict__ texf, const uchar4* __restrict__ texb, int width, long long n,
                              const float* __restrict__ sin_row, int rgb_only, float* __restrict__ w) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v;
  if (texf) {
    v = texf[i];
  } else {
    uchar4 b = texb[i];
    v = make_float4((float)b.x / 255.0f, (float)b.y / 255.0f, (float)b.z / 255.0f, (float)b.w / 255.0f);
  }
  float value = rgb_only ? fmaxf(fmaxf(v.x, v.y), v.z) : fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
  w[i] = value * sin_row[i / width];
}

// One warp per light: cdf[i] = (i ? cdf[i - 1] : 0) + w[i], in exactly that order.
struct ScanJob {
  const float* w;
  float* cdf;
  long long n;
};
__global__ void k_sequential_cdf(const ScanJob* __restrict__ jobs, int njobs) {
  const int job = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (job >= njobs) return;
  const ScanJob J = jobs[job];
  const int lane = threadIdx.x & 31;
  float acc = 0.0f;
  for (long long base = 0; base < J.n; base += 32) {
    const long long i = base + lane;
    const float mine = i < J.n ? J.w[i] : 0.0f;
    float keep = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; k++) {
      const float wk = __shfl_sync(0xFFFFFFFFu, mine, k);
      // the host loop starts with cdf[0] = w[0] (no 0 + w[0]: identical bits anyway, but keep the reference's form)
      acc = (base == 0 && k == 0) ? wk : acc + wk;
      if (k == lane) keep = acc;
    }
    if (i < J.n) J.cdf[i] = keep;
  }
}

}  // namespace

struct jt_lights {
  std::vector<jt_light_desc> descs;
  std::vector<std::vector<float>> cdfs;
};

#define JT_L_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      for (void* p_ : allocs) cudaFree(p_);                                                          \
      return jt_set_error(JT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));              \
    }                                                                                                \
  } while (0)

extern "C" int jt_lights_create(const jt_scene_desc* d, int device, int flags, jt_lights** out) {
  if (!d || !out) return jt_set_error(JT_ERR_INVALID, "jt_lights_create: null argument");
  *out = nullptr;
  int ndev = jt_device_count();
  if (ndev <= 0) return jt_set_error(JT_ERR_NO_DEVICE, "no CUDA device visible: libjtrace_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return jt_set_error(JT_ERR_INVALID, "device %d out of range", device);
  std::vector<void*> allocs;
  JT_L_CUDA(cudaSetDevice(device));
  auto dev_copy = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 4);
    if (e != cudaSuccess) return e;
    allocs.push_back(*dst);
    return bytes ? cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
  };
  struct Pending {
    int64_t instance, environment;
    float* d_w;
    float* d_cdf;
    long long n;
  };
  std::vector<Pending> pend;
  // area lights: emissive material, non-empty shape (src/trace.jl:119-150), in instance order
  for (int64_t h = 0; h < d->num_instances; h++) {
    const jt_instance& I = d->instances[h];
    if (I.material < 1 || I.material > d->num_materials || I.shape < 1 || I.shape > d->num_shapes) {
      for (void* p : allocs) cudaFree(p);
      return jt_set_error(JT_ERR_INVALID, "instance %lld: shape / material id out of range", (long long)h);
    }
    const jt_material& M = d->materials[I.material - 1];
    if (M.emission[0] == 0.0f && M.emission[1] == 0.0f && M.emission[2] == 0.0f) continue;
    const jt_shape_desc& S = d->shapes[I.shape - 1];
    if (S.num_triangles == 0 && S.num_quads == 0) continue;
    // a second `if` in the reference: quads overwrite triangles
    const bool quads = S.num_quads > 0;
    const long long n = quads ? S.num_quads : S.num_triangles;
    const int width = quads ? 4 : 3;
    const int64_t* elems = quads ? S.quads : S.triangles;
    for (long long k = 0; k < n * width; k++)
      if (elems[k] < 1 || elems[k] > S.num_positions) {
        for (void* p : allocs) cudaFree(p);
        return jt_set_error(JT_ERR_INVALID, "shape %lld: vertex id out of range", (long long)(I.shape - 1));
      }
    void *dp = nullptr, *de = nullptr, *dw = nullptr, *dc = nullptr;
    JT_L_CUDA(dev_copy(S.positions, (size_t)S.num_positions * 12, &dp));
    JT_L_CUDA(dev_copy(elems, (size_t)n * width * 8, &de));
    JT_L_CUDA(cudaMalloc(&dw, (size_t)n * 4));
    allocs.push_back(dw);
    JT_L_CUDA(cudaMalloc(&dc, (size_t)n * 4));
    allocs.push_back(dc);
    k_area_weights<<<(unsigned)((n + 255) / 256), 256>>>((const float*)dp, (const long long*)de, width, n, (float*)dw);
    pend.push_back(Pending{h + 1, -1, (float*)dw, (float*)dc, n});
  }
  // environment lights (:152-186): any non-zero emission; the CDF needs a texture
  for (int64_t h = 0; h < d->num_environments; h++) {
    const jt_environment& E = d->environments[h];
    if (E.emission[0] == 0.0f && E.emission[1] == 0.0f && E.emission[2] == 0.0f) continue;
    if (E.emission_tex == -1) {
      pend.push_back(Pending{-1, h + 1, nullptr, nullptr, 0});
      continue;
    }
    if (E.emission_tex < 1 || E.emission_tex > d->num_textures) {
      for (void* p : allocs) cudaFree(p);
      return jt_set_error(JT_ERR_INVALID, "environment %lld: texture id out of range", (long long)h);
    }
    const jt_texture_desc& T = d->textures[E.emission_tex - 1];
    const long long n = (long long)T.width * T.height;
    std::vector<float> sin_row((size_t)T.height);
    const float pi = (float)M_PI;
    for (int64_t j = 0; j < T.height; j++) {  // sin(Float32) evaluated on the host: 1 value per row, the reference's rounding
      float th = (((float)j + 0.5f) * pi) / (float)T.height;
      sin_row[(size_t)j] = (float)sin((double)th);
    }
    void *dt = nullptr, *ds = nullptr, *dw = nullptr, *dc = nullptr;
    JT_L_CUDA(dev_copy(T.pixelsf ? (const void*)T.pixelsf : (const void*)T.pixelsb, (size_t)n * (T.pixelsf ? 16 : 4), &dt));
    JT_L_CUDA(dev_copy(sin_row.data(), sin_row.size() * 4, &ds));
    JT_L_CUDA(cudaMalloc(&dw, (size_t)(n ? n : 1) * 4));
    allocs.push_back(dw);
    JT_L_CUDA(cudaMalloc(&dc, (size_t)(n ? n : 1) * 4));
    allocs.push_back(dc);
    if (n > 0)
      k_env_weights<<<(unsigned)((n + 255) / 256), 256>>>(T.pixelsf ? (const float4*)dt : nullptr, T.pixelsf ? nullptr : (const uchar4*)dt,
                                                          (int)T.width, n, (const float*)ds, (flags & JT_LIGHTS_ENV_LUMINANCE) ? 1 : 0,
                                                          (float*)dw);
    pend.push_back(Pending{-1, h + 1, (float*)dw, (float*)dc, n});
  }
  std::vector<ScanJob> jobs;
  for (const Pending& p : pend)
    if (p.n > 0) jobs.push_back(ScanJob{p.d_w, p.d_cdf, p.n});
  if (!jobs.empty()) {
    void* dj = nullptr;
    JT_L_CUDA(dev_copy(jobs.data(), jobs.size() * sizeof(ScanJob), &dj));
    k_sequential_cdf<<<(unsigned)((jobs.size() + 3) / 4), 128>>>((const ScanJob*)dj, (int)jobs.size());
  }
  JT_L_CUDA(cudaGetLastError());
  JT_L_CUDA(cudaDeviceSynchronize());
  jt_lights* L = new (std::nothrow) jt_lights();
  if (!L) {
    for (void* p : allocs) cudaFree(p);
    return jt_set_error(JT_ERR_INTERNAL, "out of host memory");
  }
  L->cdfs.resize(pend.size());
  L->descs.resize(pend.size());
  for (size_t i = 0; i < pend.size(); i++) {
    L->cdfs[i].resize((size_t)pend[i].n);
    if (pend[i].n > 0) {
      cudaError_t e = cudaMemcpy(L->cdfs[i].data(), pend[i].d_cdf, (size_t)pend[i].n * 4, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        delete L;
        for (void* p : allocs) cudaFree(p);
        return jt_set_error(JT_ERR_CUDA, "D2H copy of a light CDF failed: %s", cudaGetErrorString(e));
      }
    }
    L->descs[i].instance = pend[i].instance;
    L->descs[i].environment = pend[i].environment;
    L->descs[i].elements_cdf = pend[i].n ? L->cdfs[i].data() : nullptr;
    L->descs[i].num_elements = pend[i].n;
  }
  for (void* p : allocs) cudaFree(p);
  *out = L;
  return JT_OK;
}

extern "C" int jt_lights_desc(jt_lights* L, const jt_light_desc** descs, int64_t* count) {
  if (!L || !descs || !count) return jt_set_error(JT_ERR_INVALID, "jt_lights_desc: null argument");
  *descs = L->descs.empty() ? nullptr : L->descs.data();
  *count = (int64_t)L->descs.size();
  return JT_OK;
}

extern "C" void jt_lights_destroy(jt_lights* L) { delete L; }
