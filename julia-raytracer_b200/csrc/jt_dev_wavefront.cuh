// jt_dev_wavefront.cuh -- the wavefront form of trace_samples (src/trace.jl:215-649).
//
// The megakernel (k_trace_mega) executes one whole trace_sample per thread; ncu shows 5.65 of 32
// lanes active per warp instruction on classroom (profiles/r01). Here the same arithmetic, the same
// RNG streams and the same per-pixel sample order are reorganised into stages that run as separate
// kernels over compacted queues of path slots:
//
//   generate : slot = pixel; every slot starts its first sample (camera ray)             -> q_extend
//   extend   : closest hit for every queued ray (intersect_scene_bvh)                     -> q_shade[key]
//              key = material type of the hit instance, or MISS: material-sorted shading queues
//   shade    : miss -> environment; volume distance sampling; eval position / normal / material;
//              opacity pass-through; emission; BSDF-or-light direction sampling; delta lobes and the
//              naive sampler finish their bounce here; so do (wide mode) the MIS bounces whose ray
//              cannot reach any area light, for which sample_lights_pdf needs no BLAS walk
//                                                                       -> q_probe | q_extend | regen flag
//   probe    : sample_lights_pdf (the chained intersect_instance_bvh probes), MIS weight,
//              Russian roulette                                                   -> q_extend | regen flag
//   regen    : slots whose sample ended this iteration (flagged by shade / probe) are visited in slot (= pixel)
//              order: the sample is accumulated (running mean or sum), the next sample of the same pixel is
//              started in place and the slot joins the next extend queue; consumed queue counters are recycled
// A slot holds one sample of ITS pixel at a time, so samples of one pixel are still accumulated in
// order (Q13) with no atomics on the image. Queue appends are warp-aggregated: one atomicAdd per warp per queue,
// lane offsets from ballot / match_any masks; the per-slot state is interleaved (WfBuffers).
#pragma once
#include "jt_dev_persist.cuh"
#include "jt_dev_trace.cuh"

#define WF_NKEY 9     /* material types 0..7 + miss */
#define WF_KEY_MISS 8
// counter slots (ints). Every counter sits on its own 128-byte line: they are all hit by one atomic per warp per
// kernel, and counters sharing a line would serialise in one L2 slice.
#define WF_CS 32 /* ints between counters */
#define WF_C_EXT(cur) ((cur) * WF_CS)            /* the two extend queues */
#define WF_C_PROBE (2 * WF_CS)
#define WF_C_FETCH (3 * WF_CS)                   /* next unfetched index of the extend queue (persistent extend kernel) */
#define WF_C_SHADEK(key) ((4 + (key)) * WF_CS)   /* WF_NKEY shading queues */
#define WF_C_TOTAL ((4 + WF_NKEY) * WF_CS)

// ctl.z flags
#define WF_F_MEDIUM 1u  /* cur_volume != 0 */
#define WF_F_HIT 2u     /* first-hit outputs valid */
#define WF_F_VOLSCAT 4u /* this bounce is an in-volume scattering event (probe stage uses eval_scattering data) */

// Per-slot path state, interleaved so that what one stage touches for a slot shares 32-byte sectors: slots are
// visited in queue order (scattered), so every separately allocated 16-byte field costs its own sector.
//   group A (64 B): ray0, ray1 | hit0, hit1   -- extend reads the first sector and writes the second
//   group B (64 B): wgt, rad   | bsdf, ctl
//   group C (32 B): alb, nrm (first-hit outputs)      group D (32 B): med0, med1 (volumes only)
template <class T, int STRIDE>
struct WfField {
  T* p;
  JT_DEV T& operator[](int s) const { return p[(size_t)STRIDE * (size_t)s]; }
};
struct WfBuffers {
  WfField<float4, 4> ray0;  // o.xyz, d.x
  WfField<float4, 4> ray1;  // d.y, d.z, -, -
  WfField<float4, 4> hit0;  // bits(inst), bits(elem), u, v
  WfField<float4, 4> hit1;  // t_hit, -, -, -
  WfField<float4, 4> wgt;   // weight.xyz, -
  WfField<float4, 4> rad;   // radiance.xyz, -
  WfField<float4, 4> bsdf;  // f.xyz, pdf_bsdf   (shade -> probe)
  WfField<uint4, 4> ctl;    // x = sample, y = draw counter, z = bounce | opbounce << 8 | flags << 16, w = max_roughness bits
  WfField<float4, 2> alb;   // first-hit albedo
  WfField<float4, 2> nrm;   // first-hit normal
  WfField<float4, 2> med0;  // density.xyz, scanisotropy (only when the scene has volumetric materials)
  WfField<float4, 2> med1;  // scattering.xyz
  int* q_ext[2];
  unsigned char* regen;  // per slot: 1 = a new camera ray was started in place this iteration (k_wf_regen queues it)
  int* q_shade;  // WF_NKEY segments of n
  int* q_probe;
  int* counts;   // WF_C_TOTAL ints
  int n;         // slots of this pipeline
  int pixel_base;  // slot s renders pixel pixel_base + s (the image may be split over several pipelines)
  // point the fields at the four interleaved allocations (16 * {4, 4, 2, 2} * n bytes)
  void bind(float4* a, float4* b, float4* c, float4* d) {
    ray0.p = a; ray1.p = a + 1; hit0.p = a + 2; hit1.p = a + 3;
    wgt.p = b; rad.p = b + 1; bsdf.p = b + 2; ctl.p = (uint4*)(b + 3);
    alb.p = c; nrm.p = c + 1;
    med0.p = d; med1.p = d + 1;
  }
};

JT_DEV unsigned lane_id() { return threadIdx.x & 31u; }

// Warp-aggregated append of `slot` to queue q (counter *cnt) for the lanes with pred set. Must be
// reached by all 32 lanes of the warp.
JT_DEV void wf_append(int* q, int* cnt, bool pred, int slot) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if (m == 0u) return;
  int leader = __ffs((int)m) - 1;
  int base = 0;
  if ((int)lane_id() == leader) base = atomicAdd(cnt, __popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (pred) q[base + __popc(m & ((1u << lane_id()) - 1u))] = slot;
}

// Same, into one of WF_NKEY queues selected by key (key < 0: no append). All 32 lanes must call.
JT_DEV void wf_append_keyed(int* q_shade, int* counts, int n, int key, int slot) {
  unsigned valid = __ballot_sync(0xFFFFFFFFu, key >= 0);
  if (valid == 0u) return;
  unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
  if (key >= 0) {
    int leader = __ffs((int)peers) - 1;
    int base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(counts + WF_C_SHADEK(key), __popc(peers));
    base = __shfl_sync(peers, base, leader);
    q_shade[(size_t)key * n + base + __popc(peers & ((1u << lane_id()) - 1u))] = slot;
  }
}

struct WfPath {  // registers of one slot during a stage
  unsigned sample, draw;
  int bounce, opbounce;
  unsigned flags;
  float max_roughness;
};
JT_DEV WfPath wf_load_ctl(const WfBuffers& B, int s) {
  uint4 c = B.ctl[s];
  WfPath p;
  p.sample = c.x;
  p.draw = c.y;
  p.bounce = (int)(c.z & 0xFFu) - 1;
  p.opbounce = (int)((c.z >> 8) & 0xFFu);
  p.flags = c.z >> 16;
  p.max_roughness = __uint_as_float(c.w);
  return p;
}
JT_DEV void wf_store_ctl(const WfBuffers& B, int s, const WfPath& p) {
  // 8 bits each for bounce + 1 and opbounce (check_params bounds bounces to 0..254 for this integrator; opbounce <= 129)
  B.ctl[s] = make_uint4(p.sample, p.draw,
                        ((unsigned)(p.bounce + 1) & 0xFFu) | (((unsigned)p.opbounce & 0xFFu) << 8) | (p.flags << 16),
                        __float_as_uint(p.max_roughness));
}

// Start sample `sample` of pixel `s`: RNG draws 0..3, camera ray, unit weight (src/trace.jl:597-608, :286-296).
JT_DEV void wf_start_sample(const JtDevScene& S, const WfBuffers& B, const DevParams& P, int s, unsigned sample) {
  const int pix = B.pixel_base + s;
  Rng rng{jt_rng_key(P.seed, (uint32_t)pix, sample), 0u};
  f2 puv = rng.next2();
  f2 luv = rng.next2();
  int i = pix % P.width, j = pix / P.width;
  DRay ray = sample_camera(S.cameras[P.camera], i, j, P.width, P.height, puv, luv, P.tentfilter != 0);
  B.ray0[s] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
  B.ray1[s] = make_float4(ray.d.y, ray.d.z, 0.0f, 0.0f);
  B.wgt[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
  B.rad[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  WfPath p;
  p.sample = sample; p.draw = rng.draw; p.bounce = -1; p.opbounce = 0; p.flags = 0u; p.max_roughness = 0.0f;
  wf_store_ctl(B, s, p);
}

// End of trace_sample for slot s (src/trace.jl:625-648); returns true if another sample was started.
JT_DEV bool wf_finish_sample(const JtDevScene& S, const WfBuffers& B, const DevState& st, const DevParams& P, int s,
                             const WfPath& p, f3 radiance, int sample_end, unsigned long long* paths_done) {
  TraceOut r;
  r.radiance = radiance;
  r.hit = (p.flags & WF_F_HIT) != 0u;
  f3 cam_d = f3{0.0f, 0.0f, 0.0f};
  if (r.hit) {
    float4 a = B.alb[s], n = B.nrm[s];
    r.albedo = f3{a.x, a.y, a.z};
    r.normal = f3{n.x, n.y, n.z};
  } else {
    r.albedo = f3{0.0f, 0.0f, 0.0f};
    r.normal = f3{0.0f, 0.0f, 0.0f};
    // the camera ray direction (normal AOV of a miss) is a pure function of the RNG stream: recompute
    const int pix = B.pixel_base + s;
    Rng rng{jt_rng_key(P.seed, (uint32_t)pix, p.sample), 0u};
    f2 puv = rng.next2();
    f2 luv = rng.next2();
    cam_d = sample_camera(S.cameras[P.camera], pix % P.width, pix / P.width, P.width, P.height, puv, luv, P.tentfilter != 0).d;
  }
  accumulate_sample(st, P, S.num_environments != 0, B.pixel_base + s, (int)p.sample, r, cam_d);
  (void)paths_done;
  if ((int)p.sample + 1 < sample_end) {
    wf_start_sample(S, B, P, s, p.sample + 1u);
    return true;
  }
  return false;
}

// The end of a sample is DEFERRED to k_wf_regen (JT_DEFER_FINISH, default): shade / probe only park the final
// radiance and the control word and flag the slot. The accumulate (3 read-modify-writes of image buffers), the
// recomputation of the camera direction for the normal AOV of a miss and the next sample's camera ray are the most
// divergent tail of the shading kernels (4.9-9.3 of 32 lanes in profiles/r01/hot_lines_shade_v4.txt) and scatter
// over the image in queue order; in k_wf_regen the same work runs compacted, in pixel order, in full warps.
#ifndef JT_DEFER_FINISH
#define JT_DEFER_FINISH 1
#endif
#define WF_REGEN_CONTINUE 1 /* sample ended, another one follows for this pixel */
#define WF_REGEN_LAST 2     /* sample ended, it was the last one of the range */
JT_DEV void wf_end_sample(const JtDevScene& S, const WfBuffers& B, const DevState& st, const DevParams& P, int s,
                          const WfPath& p, f3 radiance, int sample_end, unsigned long long* counters) {
#if JT_DEFER_FINISH
  B.rad[s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
  wf_store_ctl(B, s, p);
  B.regen[s] = ((int)p.sample + 1 < sample_end) ? WF_REGEN_CONTINUE : WF_REGEN_LAST;
#else
  if (wf_finish_sample(S, B, st, P, s, p, radiance, sample_end, counters)) B.regen[s] = WF_REGEN_CONTINUE;
#endif
}
// k_wf_regen's per-slot work under JT_DEFER_FINISH: accumulate the parked sample, start the next one.
JT_DEV void wf_regen_slot(const JtDevScene& S, const WfBuffers& B, const DevState& st, const DevParams& P, int s,
                          int sample_end) {
  WfPath p = wf_load_ctl(B, s);
  float4 r = B.rad[s];
  wf_finish_sample(S, B, st, P, s, p, f3{r.x, r.y, r.z}, sample_end, nullptr);
}

// ---- generate ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(JtDevScene S, WfBuffers B, DevParams P, int sample_begin,
                                                     int sample_end, unsigned long long* counters) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.n) return;
  wf_start_sample(S, B, P, s, (unsigned)sample_begin);
  B.q_ext[0][s] = s;
  if (s == 0) {
    for (int k = 0; k < WF_C_TOTAL; k++) B.counts[k] = 0;
    B.counts[WF_C_EXT(0)] = B.n;
    atomicAdd(counters, (unsigned long long)B.n * (unsigned long long)(sample_end - sample_begin));
  }
}

// ---- extend ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k_wf_extend(JtDevScene S, WfBuffers B, int cur, unsigned long long* counters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int count = B.counts[WF_C_EXT(cur)];
  int key = -1, s = -1;
  if (t < count) {
    s = B.q_ext[cur][t];
    float4 r0 = B.ray0[s], r1 = B.ray1[s];
    DRay ray{f3{r0.x, r0.y, r0.z}, f3{r0.w, r1.x, r1.y}, JT_RAY_EPS, INFINITY};
    DHit h = intersect_scene<MODE>(S, ray);
    if (h.inst >= 0) {
      B.hit0[s] = make_float4(__int_as_float(h.inst), __int_as_float(h.elem), h.u, h.v);
      B.hit1[s] = make_float4(h.t, 0.0f, 0.0f, 0.0f);
      key = S.materials[S.instances[h.inst].material].type;
    } else {
      B.hit0[s] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.0f, 0.0f);
      key = WF_KEY_MISS;
    }
  }
  wf_append_keyed(B.q_shade, B.counts, B.n, key, s);
  unsigned m = __ballot_sync(0xFFFFFFFFu, key >= 0);
  if (lane_id() == 0u && m) atomicAdd(counters + 1, (unsigned long long)__popc(m));
}

// Persistent-warp extend over the wide BVH (jt_dev_persist.cuh): lanes refill from the queue as they finish.
#ifndef JT_EXTEND_MINBLOCKS
#define JT_EXTEND_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(JT_PERSIST_BLOCK, JT_EXTEND_MINBLOCKS) k_wf_extend_persist(JtDevScene S, WfBuffers B, int cur,
                                                                        unsigned long long* counters) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int count = B.counts[WF_C_EXT(cur)];
  const int* queue = B.q_ext[cur];
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
  unsigned nrays = 0u;
  for (;;) {
    __syncwarp();
    int key = -1;
    if (s >= 0 && !live) {  // retire a finished ray: hit record + material-sorted shading queue
      if (L.best.inst >= 0) {
        B.hit0[s] = make_float4(__int_as_float(L.best.inst), __int_as_float(L.best.elem), L.best.u, L.best.v);
        B.hit1[s] = make_float4(L.best.t, 0.0f, 0.0f, 0.0f);
        key = S.materials[S.instances[L.best.inst].material].type;
      } else {
        B.hit0[s] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.0f, 0.0f);
        key = WF_KEY_MISS;
      }
      nrays++;
    }
    wf_append_keyed(B.q_shade, B.counts, B.n, key, s);
    if (!live) s = -1;
    if (more) {
      bool want = !live;
      int idx = persist_fetch(B.counts + WF_C_FETCH, want, count);
      if (idx >= 0) {
        s = queue[idx];
        float4 r0 = B.ray0[s], r1 = B.ray1[s];
        persist_init(L, S, f3{r0.x, r0.y, r0.z}, f3{r0.w, r1.x, r1.y}, JT_RAY_EPS, INFINITY, S.wide_root, -1);
        live = S.wide_root >= 0;
        if (!live) L.best.inst = -1;
      }
      if (__ballot_sync(FULL, want && idx < 0)) more = false;
    }
    unsigned pending = __ballot_sync(FULL, live || s >= 0);
    if (pending == 0u) break;
    persist_traverse(S, L, stack, live, more);
  }
  unsigned total = __reduce_add_sync(FULL, nrays);
  if (lane_id() == 0u && total) atomicAdd(counters + 1, (unsigned long long)total);
}

// Bounce bookkeeping shared by shade (delta / naive / volume-free finishes) and probe:
// zero / non-finite weight check and Russian roulette (src/trace.jl:455-465).
// Returns true if the path continues (weight updated in place).
JT_DEV bool wf_roulette(f3& weight, WfPath& p, uint64_t key) {
  if (is_zero3(weight) || !finite3(weight)) return false;
  if (p.bounce > 3) {
    float rr_prob = jl_min(0.99f, max3(weight));
    float r = jt_rng_float(key, p.draw++);
    if (r >= rr_prob) return false;
    weight = weight * (1.0f / rr_prob);
  }
  return true;
}

// ---- shade ---------------------------------------------------------------------------------------------
// One thread per queued slot; queues are laid out key after key, each padded to a warp multiple so a
// warp only ever sees one material type.
#ifndef JT_SHADE_BLOCK
#define JT_SHADE_BLOCK 128
#endif
#ifndef JT_PROBE_BLOCK
#define JT_PROBE_BLOCK 128
#endif
#ifndef JT_SHADE_MINBLOCKS
#define JT_SHADE_MINBLOCKS 5 /* tuned on B200: profiles/r01/tuning_variants.txt */
#endif
#ifndef JT_PROBE_MINBLOCKS
#define JT_PROBE_MINBLOCKS 6
#endif
// MODE_WIDE only: when the sampled direction cannot reach any area light's padded box, sample_lights_pdf needs no
// BLAS walk (every area term is an exact zero), so the MIS weight and the Russian roulette of src/trace.jl:386-397,
// :455-465 are finished right here and the slot skips the probe kernel's state round trip.
template <int MODE>
JT_DEV bool wf_inline_mis(const JtDevScene& S, f3 position, f3 incoming, f3 f, float pdf_bsdf, f3& weight,
                          PathCounters& cnt) {
  if (MODE != MODE_WIDE) return false;
  bool need_walk = false;
  float pl = sample_lights_pdf_impl<MODE, false>(S, position, incoming, cnt, &need_walk);
  if (need_walk) return false;
  weight = (weight * f) / (0.5f * pdf_bsdf + 0.5f * pl);
  return true;
}

template <int SAMPLER, int MODE>
__global__ void __launch_bounds__(JT_SHADE_BLOCK, JT_SHADE_MINBLOCKS) k_wf_shade(JtDevScene S, WfBuffers B, DevState st, DevParams P, int next,
                                                  int sample_end, unsigned long long* counters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int key = -1, s = -1;
  PathCounters cnt{0u, 0u};
  {
    int off = 0;
#pragma unroll
    for (int k = 0; k < WF_NKEY; k++) {
      int c = B.counts[WF_C_SHADEK(k)];
      int padded = (c + 31) & ~31;
      if (key < 0 && t >= off && t < off + padded) {
        if (t - off < c) {
          key = k;
          s = B.q_shade[(size_t)k * B.n + (t - off)];
        } else {
          key = -2;  // padding lane
        }
      }
      off += padded;
    }
  }
  bool to_extend = false, to_probe = false;
  if (s >= 0) {
    const f3 zero = f3{0.0f, 0.0f, 0.0f};
    WfPath p = wf_load_ctl(B, s);
    uint64_t rkey = jt_rng_key(P.seed, (uint32_t)(B.pixel_base + s), p.sample);
    Rng rng{rkey, p.draw};
    float4 r0 = B.ray0[s], r1 = B.ray1[s], w4 = B.wgt[s], rad4 = B.rad[s];
    DRay ray{f3{r0.x, r0.y, r0.z}, f3{r0.w, r1.x, r1.y}, JT_RAY_EPS, INFINITY};
    f3 weight = f3{w4.x, w4.y, w4.z}, radiance = f3{rad4.x, rad4.y, rad4.z};
    bool alive = true;
    p.bounce += 1;  // top of the while loop, src/trace.jl:295-297
    if (key == WF_KEY_MISS) {
      if (p.bounce > 0 || !P.envhidden) radiance = radiance + weight * eval_environment(S, ray.d);
      alive = false;
    } else {
      float4 h0 = B.hit0[s];
      int inst = __float_as_int(h0.x), elem = __float_as_int(h0.y);
      float hu = h0.z, hv = h0.w, ht = B.hit1[s].x;
      bool in_volume = false;
      float distance = ht;
      VolPoint medium;
      medium.density = zero; medium.scattering = zero; medium.scanisotropy = 0.0f;
      if (SAMPLER == 1 && (p.flags & WF_F_MEDIUM)) {
        float4 m0 = B.med0[s], m1 = B.med1[s];
        medium.density = f3{m0.x, m0.y, m0.z};
        medium.scanisotropy = m0.w;
        medium.scattering = f3{m1.x, m1.y, m1.z};
        float q1 = rng.next();
        float q2 = rng.next();
        float dist = sample_transmittance(medium.density, ht, q1, q2);
        weight = (weight * eval_transmittance(medium.density, dist)) / sample_transmittance_pdf(medium.density, dist, ht);
        in_volume = dist < ht;
        distance = dist;
      }
      f3 outgoing = -ray.d;
      if (!in_volume) {
        const JtInstanceRec& I = S.instances[inst];
        const JtMaterialRec& M = S.materials[I.material];
        ElemRef E = elem_ref(S, I, elem);
        f3 position = eval_position(S, I, E, hu, hv);
        f3 normal = eval_shading_normal(S, I, E, M, hu, hv, outgoing);
        MatPoint material = eval_material(S, E, M, hu, hv);
        if (SAMPLER == 1 && P.nocaustics) {
          p.max_roughness = jl_max(material.roughness, p.max_roughness);
          material.roughness = p.max_roughness;
        }
        bool passthrough = false;
        if (material.opacity < 1.0f && rng.next() >= material.opacity) {
          if (p.opbounce > 128) {
            alive = false;
          } else {
            p.opbounce += 1;
            ray = DRay{position + ray.d * 0.01f, ray.d, JT_RAY_EPS, INFINITY};
            p.bounce -= 1;  // the `continue` re-enters the loop head, which adds it back
            to_extend = true;
          }
          passthrough = true;
        }
        if (!passthrough) {
          if (p.bounce == 0) {
            p.flags |= WF_F_HIT;
            B.alb[s] = make_float4(material.color.x, material.color.y, material.color.z, 0.0f);
            B.nrm[s] = make_float4(normal.x, normal.y, normal.z, 0.0f);
          }
          if (dot3(normal, outgoing) >= 0.0f) radiance = radiance + weight * material.emission;
          else radiance = radiance + weight * zero;
          f3 incoming;
          if (SAMPLER == 1) {
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
              if (is_zero3(incoming)) {
                alive = false;
              } else {
                f3 f = eval_bsdfcos(material, normal, outgoing, incoming);
                float pb = sample_bsdfcos_pdf(material, normal, outgoing, incoming);
                if (!wf_inline_mis<MODE>(S, position, incoming, f, pb, weight, cnt)) {
                  B.bsdf[s] = make_float4(f.x, f.y, f.z, pb);
                  to_probe = true;
                }
              }
            } else {
              incoming = sample_delta(material, normal, outgoing, rng.next());
              weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
                       sample_delta_pdf(material, normal, outgoing, incoming);
            }
            if (alive && is_volumetric_type(M.type) && dot3(normal, outgoing) * dot3(normal, incoming) < 0.0f) {
              if (!(p.flags & WF_F_MEDIUM)) {
                B.med0[s] = make_float4(material.density.x, material.density.y, material.density.z, material.scanisotropy);
                B.med1[s] = make_float4(material.scattering.x, material.scattering.y, material.scattering.z, 0.0f);
                p.flags |= WF_F_MEDIUM;
              } else {
                p.flags &= ~WF_F_MEDIUM;
              }
            }
            if (alive) {
              ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
              if (!to_probe) {  // delta lobe, or MIS weight already applied: finish the bounce here
                p.draw = rng.draw;
                alive = wf_roulette(weight, p, rkey);
                rng.draw = p.draw;
                to_extend = alive;
              }
            }
          } else {  // trace_naive, src/trace.jl:538-571
            if (material.roughness != 0.0f) {
              float rnl = rng.next();
              f2 rn = rng.next2();
              incoming = sample_bsdfcos(material, normal, outgoing, rnl, rn);
              if (is_zero3(incoming)) alive = false;
              else weight = (weight * eval_bsdfcos(material, normal, outgoing, incoming)) /
                            sample_bsdfcos_pdf(material, normal, outgoing, incoming);
            } else {
              incoming = sample_delta(material, normal, outgoing, rng.next());
              if (is_zero3(incoming)) alive = false;
              else weight = (weight * eval_delta(material, normal, outgoing, incoming)) /
                            sample_delta_pdf(material, normal, outgoing, incoming);
            }
            if (alive) {
              p.draw = rng.draw;
              alive = wf_roulette(weight, p, rkey);
              rng.draw = p.draw;
              ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
              to_extend = alive;
            }
          }
        }
      } else {  // scattering event inside the medium, src/trace.jl:423-453
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
        if (is_zero3(incoming)) {
          alive = false;
        } else {
          f3 f = eval_scattering(medium, outgoing, incoming);
          float ps = sample_scattering_pdf(medium, outgoing, incoming);
          ray = DRay{position, incoming, JT_RAY_EPS, INFINITY};
          if (wf_inline_mis<MODE>(S, position, incoming, f, ps, weight, cnt)) {
            p.draw = rng.draw;
            alive = wf_roulette(weight, p, rkey);
            rng.draw = p.draw;
            to_extend = alive;
          } else {
            B.bsdf[s] = make_float4(f.x, f.y, f.z, ps);
            to_probe = true;
          }
        }
      }
    }
    // the while condition (bounce < bounces) is checked when the next iteration would start
    if (to_extend && !(p.bounce < P.bounces)) {
      to_extend = false;
      alive = false;
    }
    p.draw = rng.draw;
    if (alive) {
      B.ray0[s] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
      B.ray1[s] = make_float4(ray.d.y, ray.d.z, 0.0f, 0.0f);
      B.wgt[s] = make_float4(weight.x, weight.y, weight.z, 0.0f);
      B.rad[s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
      wf_store_ctl(B, s, p);
    } else {
      to_probe = false;
      wf_end_sample(S, B, st, P, s, p, radiance, sample_end, counters);  // accumulated + regenerated in pixel order by k_wf_regen
    }
  }
  wf_append(B.q_probe, B.counts + WF_C_PROBE, to_probe, s);
  wf_append(B.q_ext[next], B.counts + WF_C_EXT(next), to_extend, s);
  if (SAMPLER == 1 && MODE == MODE_WIDE) {
    unsigned lr = __reduce_add_sync(0xFFFFFFFFu, cnt.light_rays);
    if (lane_id() == 0u && lr) atomicAdd(counters + 2, (unsigned long long)lr);
  }
}

// ---- probe ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(JT_PROBE_BLOCK, JT_PROBE_MINBLOCKS) k_wf_probe(JtDevScene S, WfBuffers B, DevState st, DevParams P, int next,
                                                  int sample_end, unsigned long long* counters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int count = B.counts[WF_C_PROBE];
  int s = -1;
  bool to_extend = false;
  PathCounters cnt{0u, 0u};
  if (t < count) {
    s = B.q_probe[t];
    WfPath p = wf_load_ctl(B, s);
    float4 r0 = B.ray0[s], r1 = B.ray1[s], w4 = B.wgt[s], fb = B.bsdf[s];
    f3 position = f3{r0.x, r0.y, r0.z}, incoming = f3{r0.w, r1.x, r1.y};
    float pl = sample_lights_pdf<MODE>(S, position, incoming, cnt);
    f3 weight = (f3{w4.x, w4.y, w4.z} * f3{fb.x, fb.y, fb.z}) / (0.5f * fb.w + 0.5f * pl);
    uint64_t rkey = jt_rng_key(P.seed, (uint32_t)(B.pixel_base + s), p.sample);
    bool alive = wf_roulette(weight, p, rkey);
    if (alive && !(p.bounce < P.bounces)) alive = false;
    if (alive) {
      B.wgt[s] = make_float4(weight.x, weight.y, weight.z, 0.0f);
      wf_store_ctl(B, s, p);
      to_extend = true;
    } else {
#if JT_DEFER_FINISH
      // radiance and the control word (sample index, first-hit flag) were parked by the shade kernel
      B.regen[s] = ((int)p.sample + 1 < sample_end) ? WF_REGEN_CONTINUE : WF_REGEN_LAST;
#else
      float4 rad4 = B.rad[s];
      if (wf_finish_sample(S, B, st, P, s, p, f3{rad4.x, rad4.y, rad4.z}, sample_end, counters)) B.regen[s] = WF_REGEN_CONTINUE;
#endif
    }
  }
  wf_append(B.q_ext[next], B.counts + WF_C_EXT(next), to_extend, s);
  unsigned lr = __reduce_add_sync(0xFFFFFFFFu, cnt.light_rays);
  if (lane_id() == 0u && lr) atomicAdd(counters + 2, (unsigned long long)lr);
}

// ---- regen + advance -------------------------------------------------------------------------------------------
// Closes an iteration: (1) the slots whose sample ended this iteration (flagged by shade / probe) are compacted IN SLOT
// (= PIXEL) ORDER, block by block; their sample is accumulated, the next sample of the pixel is started in place and the
// slot is appended to the next extend queue, so that camera rays of neighbouring pixels sit in neighbouring lanes of the
// extend kernel -- queues built by atomics scatter them among the bounce rays, and the same mix traverses 11 % slower
// (tools/exp_coherence.py: 2 534 vs 2 827 Mrays/s on classroom); (2) the consumed queues' counters are recycled.
#define WF_REGEN_BLOCK 256
#ifndef WF_REGEN_PER_THREAD
#define WF_REGEN_PER_THREAD 16 /* slots per thread: one 128-bit (16) or one 32-bit (4) load of flags */
#endif
#ifndef JT_EMU_COUNT
__global__ void __launch_bounds__(WF_REGEN_BLOCK) k_wf_regen(JtDevScene S, WfBuffers B, DevState st, DevParams P, int cur,
                                                             int sample_end) {
  __shared__ int warp_sums[2][WF_REGEN_BLOCK / 32];
  __shared__ int block_base;
#if JT_DEFER_FINISH
  __shared__ int ended[WF_REGEN_BLOCK * WF_REGEN_PER_THREAD];  // slots of this block whose sample ended, in slot order
#endif
  const int next = cur ^ 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    B.counts[WF_C_EXT(cur)] = 0;
    B.counts[WF_C_PROBE] = 0;
    B.counts[WF_C_FETCH] = 0;
    for (int k = 0; k < WF_NKEY; k++) B.counts[WF_C_SHADEK(k)] = 0;
  }
  const int first = (blockIdx.x * WF_REGEN_BLOCK + threadIdx.x) * WF_REGEN_PER_THREAD;
  unsigned flags = 0u;  // bit i: slot first + i continues with a new camera ray
  unsigned last = 0u;   // bit i: slot first + i ended its LAST sample (accumulate only)
  static_assert(WF_REGEN_PER_THREAD == 16 || WF_REGEN_PER_THREAD == 4, "flag bytes are read with one 128- or 32-bit load");
  if (first + WF_REGEN_PER_THREAD <= B.n) {
#if WF_REGEN_PER_THREAD == 16
    uint4 v = *reinterpret_cast<const uint4*>(B.regen + first);
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#else
    unsigned w[1] = {*reinterpret_cast<const unsigned*>(B.regen + first)};
#endif
#pragma unroll
    for (int k = 0; k < WF_REGEN_PER_THREAD / 4; k++)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        flags |= ((w[k] >> (8 * j)) & 1u) << (4 * k + j);
        last |= ((w[k] >> (8 * j + 1)) & 1u) << (4 * k + j);
      }
    if (flags | last) {
#if WF_REGEN_PER_THREAD == 16
      *reinterpret_cast<uint4*>(B.regen + first) = make_uint4(0u, 0u, 0u, 0u);
#else
      *reinterpret_cast<unsigned*>(B.regen + first) = 0u;
#endif
    }
  } else {
    for (int i = 0; i < WF_REGEN_PER_THREAD && first + i < B.n; i++) {
      unsigned char f = B.regen[first + i];
      if (f) {
        if (f == WF_REGEN_CONTINUE) flags |= 1u << i;
        else last |= 1u << i;
        B.regen[first + i] = 0;
      }
    }
  }
  const int mine = __popc(flags);
  const int mine_all = __popc(flags | last);
  // block-wide exclusive scans of `mine` (queue positions) and `mine_all` (positions in the ended list)
  int incl = mine, incl_all = mine_all;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    int u = __shfl_up_sync(0xFFFFFFFFu, incl_all, d);
    if ((int)lane_id() >= d) {
      incl += t;
      incl_all += u;
    }
  }
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 31u) {
    warp_sums[0][warp] = incl;
    warp_sums[1][warp] = incl_all;
  }
  __syncthreads();
  int warp_off = 0, total = 0, warp_off_all = 0, total_all = 0;
#pragma unroll
  for (int k = 0; k < WF_REGEN_BLOCK / 32; k++) {
    int v = warp_sums[0][k], u = warp_sums[1][k];
    if (k < warp) {
      warp_off += v;
      warp_off_all += u;
    }
    total += v;
    total_all += u;
  }
  if (threadIdx.x == 0) block_base = total ? atomicAdd(B.counts + WF_C_EXT(next), total) : 0;
  __syncthreads();
  int at = block_base + warp_off + incl - mine;
#if JT_DEFER_FINISH
  int at_all = warp_off_all + incl_all - mine_all;
  unsigned both = flags | last;
  while (both) {
    int i = __ffs((int)both) - 1;
    both &= both - 1u;
    ended[at_all++] = first + i;
    if (flags & (1u << i)) B.q_ext[next][at++] = first + i;
  }
  __syncthreads();
  // consecutive threads take consecutive ended slots = neighbouring pixels: coalesced accumulator updates, full warps
  for (int j = threadIdx.x; j < total_all; j += WF_REGEN_BLOCK) wf_regen_slot(S, B, st, P, ended[j], sample_end);
#else
  (void)total_all; (void)warp_off_all; (void)last;
  while (flags) {
    int i = __ffs((int)flags) - 1;
    flags &= flags - 1u;
    B.q_ext[next][at++] = first + i;
  }
#endif
}
#endif
