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
//   regen    : slots whose sample ended this iteration (flagged by shade / probe) are visited in slot order: the sample
//              is COMMITTED (accumulated: running mean or sum) when it is its pixel's turn, the slot claims the next
//              unstarted sample -- of its own pixel, or, when that pixel has none left, of a pixel another slot is
//              still working on (sample-level work stealing) -- and joins the next extend queue; consumed queue
//              counters are recycled
// A slot holds one sample of one pixel at a time. Samples of a pixel may be traced by several slots at once, but they
// are accumulated strictly in sample order (Q13: the running mean is order-sensitive) through the per-pixel `commit`
// turn counter, with no atomics on the image; RNG streams are keyed by (pixel, sample), so which slot traces a sample
// changes nothing. Queue appends are warp-aggregated: one atomicAdd per warp per queue, lane offsets from ballot /
// match_any masks; the per-slot state is interleaved (WfBuffers).
#pragma once
#include "jt_dev_persist.cuh"
#include "jt_dev_trace.cuh"

// Shading queues: one per (material type, predicted MIS coin) pair + one for misses. The path sampler splits every
// non-delta bounce in half -- BSDF sampling or light sampling, decided by one RNG draw (src/trace.jl:366-377) -- and the
// two halves share no code until eval_bsdfcos. The draw's index is known when the ray is created (camera: draw 4;
// bounce: the creator's counter, + 1 if the probe kernel will still play Russian roulette, + 2 inside a medium), so the
// creator evaluates it and leaves the bit in the ray record; the extend kernel files the hit under (type, coin).
// A wrong prediction (a stochastic-opacity draw comes first) only costs coherence: shading redraws the coin itself.
#ifndef JT_COIN_QUEUES
#define JT_COIN_QUEUES 1
#endif
#define WF_TYPE_MISS 8 /* what the shading body calls the miss key */
#if JT_COIN_QUEUES
#define WF_NKEY 17
#define WF_KEY_MISS 16
#define WF_QUEUE_KEY(type, coin) (((type) << 1) | (coin))
#define WF_KEY_TYPE(key) ((key) >> 1)
#else
#define WF_NKEY 9 /* material types 0..7 + miss */
#define WF_KEY_MISS 8
#define WF_QUEUE_KEY(type, coin) (type)
#define WF_KEY_TYPE(key) (key)
#endif
// ray1.z carries two bits (as raw integer bits): 1 = suspended traversal to resume, 2 = predicted coin
#define WF_RAY_SUSPENDED 1
#define WF_RAY_COIN 2
// counter slots (ints). Every counter sits on its own 128-byte line: they are all hit by one atomic per warp per
// kernel, and counters sharing a line would serialise in one L2 slice.
#define WF_CS 32 /* ints between counters */
#define WF_C_EXT(cur) ((cur) * WF_CS)            /* the two extend queues */
#define WF_C_PROBE (2 * WF_CS)
#define WF_C_FETCH (3 * WF_CS)                   /* next unfetched index of the extend queue (persistent extend kernel) */
#define WF_C_SHADEK(key) ((4 + (key)) * WF_CS)   /* WF_NKEY shading queues */
#define WF_C_LASTN ((4 + WF_NKEY) * WF_CS)       /* length of the extend queue consumed this iteration (stable copy) */
#define WF_C_ACTIVE ((5 + WF_NKEY) * WF_CS)      /* pixels that still have unstarted samples */
#define WF_C_DONE ((6 + WF_NKEY) * WF_CS)        /* slots that are idle for good: the range is finished when == n */
#define WF_C_TOTAL ((7 + WF_NKEY) * WF_CS)

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
  WfField<float4, 4> ray1;  // d.y, d.z, bits(WF_RAY_SUSPENDED | WF_RAY_COIN), bits(pixel)
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
  uint2* parked;  // JT_SUSPEND_STACK entries per slot: traversal stack of a ray suspended in the extend kernel's tail
  float4* held;   // 3 per slot: a finished sample set aside until its pixel's turn (wf_regen_slot)
  int* q_ext[2];
  int* next_sample;  // per pixel of this pipeline (index = pixel - pixel_base): next sample index to hand out
  int* commit;       // per pixel: next sample index to accumulate (samples are accumulated in order)
  unsigned char* regen;  // per slot: 1 = k_wf_regen must visit it (sample ended / waiting for its turn / looking for work)
  int* q_shade;  // WF_NKEY segments of n
  int* q_probe;
  int* counts;   // WF_C_TOTAL ints
  int n;         // slots of this pipeline
  int pixel_base;  // this pipeline renders pixels pixel_base .. pixel_base + n - 1; slot s starts on pixel pixel_base + s
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

// Two appends at once (shade: probe queue + next extend queue): both atomics are issued before either result is used, so
// the warp waits for one L2 round trip instead of two. Must be reached by all 32 lanes.
JT_DEV void wf_append2(int* qa, int* ca, bool pa, int* qb, int* cb, bool pb, int slot) {
  const unsigned ma = __ballot_sync(0xFFFFFFFFu, pa), mb = __ballot_sync(0xFFFFFFFFu, pb);
  if ((ma | mb) == 0u) return;
  const int la = ma ? __ffs((int)ma) - 1 : 0, lb = mb ? __ffs((int)mb) - 1 : 0;
  int base_a = 0, base_b = 0;
  if (ma && (int)lane_id() == la) base_a = atomicAdd(ca, __popc(ma));
  if (mb && (int)lane_id() == lb) base_b = atomicAdd(cb, __popc(mb));
  base_a = __shfl_sync(0xFFFFFFFFu, base_a, la);
  base_b = __shfl_sync(0xFFFFFFFFu, base_b, lb);
  const unsigned below = (1u << lane_id()) - 1u;
  if (pa) qa[base_a + __popc(ma & below)] = slot;
  if (pb) qb[base_b + __popc(mb & below)] = slot;
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

// ctl.z flag (bits 16..): set while a slot whose result is already committed looks for new work
#define WF_F_COMMITTED 8u

// Start sample `sample` of pixel `pix` in slot s: RNG draws 0..3, camera ray, unit weight (src/trace.jl:597-608, :286-296).
// The pixel (absolute index) whose sample slot s is tracing travels in the spare word of the ray record.
JT_DEV int wf_slot_pixel(const WfBuffers& B, int s) { return __float_as_int((&B.ray1[s].x)[3]); }

JT_DEV void wf_start_sample(const JtDevScene& S, const WfBuffers& B, const DevParams& P, int s, int pix, unsigned sample) {
  Rng rng{jt_rng_key(P.seed, (uint32_t)pix, sample), 0u};
  f2 puv = rng.next2();
  f2 luv = rng.next2();
  int i = pix % P.width, j = pix / P.width;
  DRay ray = sample_camera(S.cameras[P.camera], i, j, P.width, P.height, puv, luv, P.tentfilter != 0);
  const int coin = (JT_COIN_QUEUES && P.sampler == 1 && jt_rng_float(rng.key, rng.draw) < 0.5f) ? WF_RAY_COIN : 0;
  B.ray0[s] = make_float4(ray.o.x, ray.o.y, ray.o.z, ray.d.x);
  B.ray1[s] = make_float4(ray.d.y, ray.d.z, __int_as_float(coin), __int_as_float(pix));
  B.wgt[s] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
  B.rad[s] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  WfPath p;
  p.sample = sample; p.draw = rng.draw; p.bounce = -1; p.opbounce = 0; p.flags = 0u; p.max_roughness = 0.0f;
  wf_store_ctl(B, s, p);
}

// The end of a sample is DEFERRED to k_wf_regen: shade / probe only park the final radiance and the control word and
// flag the slot. The accumulate (3 read-modify-writes of image buffers), the recomputation of the camera direction for
// the normal AOV of a miss and the next sample's camera ray are the most divergent tail of the shading kernels (4.9-9.3
// of 32 lanes in profiles/r01/hot_lines_shade_v4.txt) and scatter over the image in queue order; in k_wf_regen the same
// work runs compacted, in slot order, in full warps.
JT_DEV void wf_end_sample(const WfBuffers& B, int s, const WfPath& p, f3 radiance) {
  B.rad[s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
  wf_store_ctl(B, s, p);
  B.regen[s] = 1;
}

// The per-pixel turn counter holds (next sample to accumulate) << 8 | (epoch of the k_wf_regen launch that wrote it).
// A sample is committed only when it is the pixel's turn AND the turn was handed over by an EARLIER launch: everything
// that launch stored (the accumulators) is then visible without fences or cache-bypassing loads, and no thread ever
// waits on another thread of its own launch. Epochs are iteration numbers mod 256; the rare alias (a turn handed over
// exactly 256 iterations ago) costs the slot one extra iteration.
#define WF_COMMIT_WORD(sample, epoch) (((sample) << 8) | ((epoch) & 255))

// Sample-level work stealing (JT_WORK_STEALING, default on). With one slot per pixel the slowest pixel of a chunk sets
// the number of wavefront iterations and the queues run dry long before that: on features1 35 % of a 512-spp step was
// spent at < 10 % queue fill, 6-14 % on classroom (JT_ITER_LOG, profiles/r02). A slot whose pixel has no unstarted
// sample left now claims the next sample of a pixel that is still being worked on (picked through a random entry of the
// extend queue just consumed), at most WF_STEAL_WINDOW samples ahead of that pixel's commit counter.
#ifndef JT_WORK_STEALING
#define JT_WORK_STEALING 1
#endif
#ifdef JT_EMU_COUNT
static unsigned long long jt_emu_steals = 0;  // host emulation: samples traced by a slot that started on another pixel
#endif
#ifndef WF_STEAL_TRIES
#define WF_STEAL_TRIES 4
#endif
#ifndef WF_STEAL_WINDOW
#define WF_STEAL_WINDOW 8
#endif
#define WF_REGEN_QUEUED 1  /* a new sample was started: the slot joins the next extend queue */
#define WF_REGEN_RETRY 2   /* waiting for the pixel's turn, or found nothing to steal yet: visit again next iteration */
#define WF_REGEN_DONE 3    /* nothing left to start anywhere: idle for good */
#define WF_REGEN_STOLEN 4  /* like QUEUED, with a sample of another slot's pixel */

JT_DEV bool wf_claim(const JtDevScene& S, const WfBuffers& B, const DevParams& P, int s, int pix, int sample_end) {
  const int lp = pix - B.pixel_base;
  const int k = atomicAdd(B.next_sample + lp, 1);
  if (k >= sample_end) return false;
  if (k == sample_end - 1) atomicAdd(B.counts + WF_C_ACTIVE, -1);  // the pixel's last sample has been handed out
  wf_start_sample(S, B, P, s, pix, (unsigned)k);
  return true;
}

// k_wf_regen's per-slot work: commit the parked sample when it is the pixel's turn (src/trace.jl:625-648), then find the
// slot its next sample. `cur` is the extend queue consumed this iteration (victim lookup).
// One result of a finished sample: committed into the accumulators when it is its pixel's turn.
struct WfResult {
  int pix, sample;
  bool hit;
  f3 radiance, albedo, normal;
};
JT_DEV bool wf_try_commit(const JtDevScene& S, const WfBuffers& B, const DevState& st, const DevParams& P, const WfResult& R,
                          int epoch) {
  const int lp = R.pix - B.pixel_base;
  // every load the commit needs is issued before the turn is examined: one memory round trip instead of two
  const int turn = B.commit[lp];
  const float4 old_img = st.image[R.pix], old_alb = st.albedo[R.pix], old_nrm = st.normal[R.pix];
  const int old_hits = st.hits[R.pix];
  // an earlier sample of the pixel is still in flight, or was committed by this very launch
  if ((turn >> 8) != R.sample || (turn & 255) == (epoch & 255)) return false;
  TraceOut r;
  r.radiance = R.radiance;
  r.hit = R.hit;
  f3 cam_d = f3{0.0f, 0.0f, 0.0f};
  if (r.hit) {
    r.albedo = R.albedo;
    r.normal = R.normal;
  } else {
    r.albedo = f3{0.0f, 0.0f, 0.0f};
    r.normal = f3{0.0f, 0.0f, 0.0f};
    // the camera ray direction (normal AOV of a miss) is a pure function of the RNG stream: recompute
    Rng rng{jt_rng_key(P.seed, (uint32_t)R.pix, (uint32_t)R.sample), 0u};
    f2 puv = rng.next2();
    f2 luv = rng.next2();
    cam_d = sample_camera(S.cameras[P.camera], R.pix % P.width, R.pix / P.width, P.width, P.height, puv, luv, P.tentfilter != 0).d;
  }
  accumulate_loaded(st, P, S.num_environments != 0, R.pix, R.sample, r, cam_d, old_img, old_alb, old_nrm, old_hits);
  B.commit[lp] = WF_COMMIT_WORD(R.sample + 1, epoch);
  return true;
}

// A slot can hold a SECOND finished sample (WfBuffers::held: 3 x float4 per slot). With several slots on one pixel a
// sample often finishes before its predecessor; instead of idling until its turn the slot sets the result aside and
// starts another sample (JT_ITER_LOG on features1: the extend queue ran at 0.5-0.75 fill because of such waits). The
// slot keeps visiting k_wf_regen (flag WF_FLAG_HELD) until the held result is committed; it only stalls when the next
// sample ends while the held one is still waiting.
//   held[0] = {radiance.xyz, bits(pixel)}  held[1] = {albedo.xyz, bits(sample)}  held[2] = {normal.xyz, bits(1 valid | 2 hit)}
#ifndef JT_HELD_RESULT
#define JT_HELD_RESULT 0 /* measured: features1 +1 %, classroom -1 % (profiles/r02/tuning_variants.txt item 22) */
#endif
#define WF_FLAG_ENDED 1 /* regen flag byte: the slot's current sample ended (set by shade / probe) */
#define WF_FLAG_HELD 2  /* the slot only has a held result to commit; its current path is in flight */

// k_wf_regen's per-slot work: commit what can be committed (src/trace.jl:625-648), then find the slot its next sample.
// `cur` is the extend queue consumed this iteration (victim lookup). *flag_out = the slot's regen flag afterwards.
JT_DEV int wf_regen_slot(const JtDevScene& S, const WfBuffers& B, const DevState& st, const DevParams& P, int s, int cur,
                         int sample_end, int epoch, int flag_in, unsigned char* flag_out) {
  bool held = false;
#if JT_HELD_RESULT
  {
    const float4 h2 = B.held[3 * (size_t)s + 2];
    const int hbits = __float_as_int(h2.w);
    if (hbits & 1) {
      const float4 h0 = B.held[3 * (size_t)s], h1 = B.held[3 * (size_t)s + 1];
      WfResult R{__float_as_int(h0.w), __float_as_int(h1.w), (hbits & 2) != 0, f3{h0.x, h0.y, h0.z}, f3{h1.x, h1.y, h1.z},
                 f3{h2.x, h2.y, h2.z}};
      if (wf_try_commit(S, B, st, P, R, epoch)) B.held[3 * (size_t)s + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      else held = true;
    }
  }
#endif
  if (!(flag_in & WF_FLAG_ENDED)) {  // the current path is still in flight: nothing else to do
    *flag_out = held ? WF_FLAG_HELD : 0;
    return WF_REGEN_RETRY;
  }
  WfPath p = wf_load_ctl(B, s);
  const int pix = wf_slot_pixel(B, s);
  if (!(p.flags & WF_F_COMMITTED)) {
    const float4 rad = B.rad[s], a = B.alb[s], n = B.nrm[s];
    WfResult R{pix, (int)p.sample, (p.flags & WF_F_HIT) != 0u, f3{rad.x, rad.y, rad.z}, f3{a.x, a.y, a.z}, f3{n.x, n.y, n.z}};
    if (!wf_try_commit(S, B, st, P, R, epoch)) {
#if JT_HELD_RESULT
      if (!held) {  // set it aside; the slot is free for another sample
        B.held[3 * (size_t)s] = make_float4(rad.x, rad.y, rad.z, __int_as_float(pix));
        B.held[3 * (size_t)s + 1] = make_float4(a.x, a.y, a.z, __int_as_float((int)p.sample));
        B.held[3 * (size_t)s + 2] = make_float4(n.x, n.y, n.z, __int_as_float(1 | (R.hit ? 2 : 0)));
        held = true;
      } else
#endif
      {
        *flag_out = WF_FLAG_ENDED;  // both results wait for their turns: the slot stalls
        return WF_REGEN_RETRY;
      }
    }
  }
  *flag_out = held ? WF_FLAG_HELD : 0;
  if (wf_claim(S, B, P, s, pix, sample_end)) return WF_REGEN_QUEUED;
#if JT_WORK_STEALING
  if (B.counts[WF_C_ACTIVE] > 0) {
    const int lastn = B.counts[WF_C_LASTN];
    uint32_t h = (uint32_t)s * 2654435761u ^ (p.sample * 40503u + p.draw);
    for (int t = 0; t < WF_STEAL_TRIES && lastn > 0; t++) {
      h = h * 1664525u + 1013904223u;
      const int victim = wf_slot_pixel(B, B.q_ext[cur][(h >> 8) % (uint32_t)lastn]);
      const int lv = victim - B.pixel_base;
      const int nx = B.next_sample[lv];
      if (nx < sample_end && nx - (B.commit[lv] >> 8) < WF_STEAL_WINDOW && wf_claim(S, B, P, s, victim, sample_end)) {
#ifdef JT_EMU_COUNT
        jt_emu_steals++;
#endif
        return WF_REGEN_STOLEN;
      }
    }
  }
#endif
  // nothing to start: the slot stays flagged as "ended, result taken care of" while a held result or unstarted samples
  // somewhere keep it busy, and is idle for good otherwise
  if (!(p.flags & WF_F_COMMITTED)) {
    p.flags |= WF_F_COMMITTED;
    wf_store_ctl(B, s, p);
  }
  if (held || (JT_WORK_STEALING && B.counts[WF_C_ACTIVE] > 0)) {
    *flag_out = WF_FLAG_ENDED;
    return WF_REGEN_RETRY;
  }
  *flag_out = 0;
  return WF_REGEN_DONE;
}

// ---- generate ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_generate(JtDevScene S, WfBuffers B, DevParams P, int sample_begin,
                                                     int sample_end, unsigned long long* counters) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B.n) return;
  wf_start_sample(S, B, P, s, B.pixel_base + s, (unsigned)sample_begin);
  B.next_sample[s] = sample_begin + 1;
  B.commit[s] = WF_COMMIT_WORD(sample_begin, 255);  // the first k_wf_regen launch has epoch 0
  B.regen[s] = 0;
  if (JT_HELD_RESULT) B.held[3 * (size_t)s + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  B.q_ext[0][s] = s;
  if (s == 0) {
    for (int k = 0; k < WF_C_TOTAL; k++) B.counts[k] = 0;
    B.counts[WF_C_EXT(0)] = B.n;
    B.counts[WF_C_ACTIVE] = sample_begin + 1 < sample_end ? B.n : 0;
    atomicAdd(counters, (unsigned long long)B.n * (unsigned long long)(sample_end - sample_begin));
  }
}

// ---- extend ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k_wf_extend(JtDevScene S, WfBuffers B, int cur, unsigned long long* counters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int count = B.counts[WF_C_EXT(cur)];
  if (t == 0) B.counts[WF_C_LASTN] = count;
  int key = -1, s = -1;
  if (t < count) {
    s = B.q_ext[cur][t];
    float4 r0 = B.ray0[s], r1 = B.ray1[s];
    DRay ray{f3{r0.x, r0.y, r0.z}, f3{r0.w, r1.x, r1.y}, JT_RAY_EPS, INFINITY};
    DHit h = intersect_scene<MODE>(S, ray);
    if (h.inst >= 0) {
      B.hit0[s] = make_float4(__int_as_float(h.inst), __int_as_float(h.elem), h.u, h.v);
      B.hit1[s] = make_float4(h.t, 0.0f, 0.0f, 0.0f);
      key = WF_QUEUE_KEY(S.materials[S.instances[h.inst].material].type, (__float_as_int(r1.z) & WF_RAY_COIN) ? 1 : 0);
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
#ifndef JT_SUSPEND_MIN_QUEUE
#define JT_SUSPEND_MIN_QUEUE 8192 /* shorter queues: the launch is all tail and cheap; parking would only add iterations */
#endif
__global__ void __launch_bounds__(JT_PERSIST_BLOCK, JT_EXTEND_MINBLOCKS) k_wf_extend_persist(JtDevScene S, WfBuffers B, int cur,
                                                                        unsigned long long* counters) {
  const unsigned FULL = 0xFFFFFFFFu;
  const int count = B.counts[WF_C_EXT(cur)];
  if (blockIdx.x == 0 && threadIdx.x == 0) B.counts[WF_C_LASTN] = count;
  const int* queue = B.q_ext[cur];
  uint2 stack_local[JT_WIDE_STACK - JT_SMEM_STACK];
  TravStack stack;
  stack.local = stack_local;
#if JT_SMEM_STACK > 0 && !defined(JT_EMU_COUNT)
  __shared__ uint2 stack_shared[JT_SMEM_STACK * JT_PERSIST_BLOCK];
  stack.shared = stack_shared + threadIdx.x;
#endif
  PersistLane L;
  bool live = false, more = true, parked = false;
  bool allow_suspend = B.parked != nullptr && count >= JT_SUSPEND_MIN_QUEUE;
  int s = -1, coin = 0;  // coin: predicted MIS coin of the lane's ray (queue key), from the ray record
  unsigned nrays = 0u, nresumed = 0u;
  for (;;) {
    __syncwarp();
    int key = -1;
    if (s >= 0 && !live && !parked) {  // retire a finished ray: hit record + material-sorted shading queue
      if (L.best.inst >= 0) {
        B.hit0[s] = make_float4(__int_as_float(L.best.inst), __int_as_float(L.best.elem), L.best.u, L.best.v);
        B.hit1[s] = make_float4(L.best.t, 0.0f, 0.0f, 0.0f);
        key = WF_QUEUE_KEY(S.materials[S.instances[L.best.inst].material].type, coin);
      } else {
        B.hit0[s] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.0f, 0.0f);
        key = WF_KEY_MISS;
      }
      nrays++;
    }
    wf_append_keyed(B.q_shade, B.counts, B.n, key, s);
    // suspended rays skip this iteration's shading: straight to the next extend queue
    wf_append(B.q_ext[cur ^ 1], B.counts + WF_C_EXT(cur ^ 1), parked, s);
    parked = false;
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
        const int rbits = __float_as_int(r1.z);
        coin = (rbits & WF_RAY_COIN) ? 1 : 0;
        if (rbits & WF_RAY_SUSPENDED) {  // a ray suspended by the previous launch: pick its traversal up where it stopped
          persist_resume(L, S, stack, B.hit0[s], B.hit1[s], B.parked + (size_t)s * JT_SUSPEND_STACK);
          (&B.ray1[s].x)[2] = __int_as_float(rbits & ~WF_RAY_SUSPENDED);
          nresumed++;
        }
      }
      if (__ballot_sync(FULL, want && idx < 0)) more = false;
    }
    unsigned pending = __ballot_sync(FULL, live || s >= 0);
    if (pending == 0u) break;
    if (persist_traverse(S, L, stack, live, more, allow_suspend)) {
      if (live && persist_can_park(L)) {
        persist_park(L, stack, &B.hit0[s], &B.hit1[s], B.parked + (size_t)s * JT_SUSPEND_STACK);
        (&B.ray1[s].x)[2] = __int_as_float(WF_RAY_SUSPENDED | (coin ? WF_RAY_COIN : 0));
        live = false;
        parked = true;
      }
#ifndef JT_EMU_COUNT
      allow_suspend = false;  // lanes too deep to park run to the end
#endif
    }
  }
  unsigned total = __reduce_add_sync(FULL, nrays);
  if (lane_id() == 0u && total) atomicAdd(counters + 1, (unsigned long long)total);
  total = __reduce_add_sync(FULL, nresumed);
  if (lane_id() == 0u && total) atomicAdd(counters + 4, (unsigned long long)total);
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
// One thread per queued slot; queues are laid out key after key, each padded to a block multiple so a
// block only ever sees one material type.
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

// Per-material specialisation (KEY = the shading queue's material type, or -1 = decided at run time). Queues are padded
// to BLOCK multiples, so a block only ever sees one key and jumps to the body compiled for it: the lobe switches of
// jt_dev_shade.cuh fold to one case, the dead lobes leave the instruction stream (the generic kernel is 20.6 k SASS
// instructions and 17 % of its stall samples were instruction-cache misses, profiles/r02/hot_lines_shade_r02f.txt).
#ifndef JT_SHADE_SPECIALISE
#define JT_SHADE_SPECIALISE 0 /* measured: 1-1.5 % slower (96 instead of 80 registers; profiles/r02/tuning_variants.txt) */
#endif
template <int SAMPLER, int MODE, int KEY>
JT_DEV void wf_shade_slot(const JtDevScene& S, const WfBuffers& B, const DevParams& P, int s, int key_rt, bool& to_extend,
                          bool& to_probe, PathCounters& cnt) {
  const int key = KEY >= 0 ? KEY : key_rt;
  if (s >= 0) {
    const f3 zero = f3{0.0f, 0.0f, 0.0f};
    WfPath p = wf_load_ctl(B, s);
    float4 r0 = B.ray0[s], r1 = B.ray1[s], w4 = B.wgt[s], rad4 = B.rad[s];
    const int pix = __float_as_int(r1.w);
    uint64_t rkey = jt_rng_key(P.seed, (uint32_t)pix, p.sample);
    Rng rng{rkey, p.draw};
    DRay ray{f3{r0.x, r0.y, r0.z}, f3{r0.w, r1.x, r1.y}, JT_RAY_EPS, INFINITY};
    f3 weight = f3{w4.x, w4.y, w4.z}, radiance = f3{rad4.x, rad4.y, rad4.z};
    bool alive = true;
    p.bounce += 1;  // top of the while loop, src/trace.jl:295-297
    if (key == WF_TYPE_MISS) {
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
        const JtMaterialRec& Mg = S.materials[I.material];
        JtMaterialRec Mk;  // specialised bodies: a copy whose type is a compile-time constant (the queue key IS the
        if (KEY >= 0) {    // material type), so the lobe switches fold to one case; scalarised, dead fields never load
          Mk = Mg;
          Mk.type = KEY;
        }
        const JtMaterialRec& M = KEY >= 0 ? Mk : Mg;
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
      int coin = 0;
      if (JT_COIN_QUEUES && SAMPLER == 1) {
        // index of the next bounce's coin: this counter, + 1 if the probe kernel still plays Russian roulette for this
        // bounce (p.bounce > 3, src/trace.jl:459), + 2 if the next bounce starts with the medium's distance sampling
        const unsigned at = p.draw + ((to_probe && p.bounce > 3) ? 1u : 0u) + ((p.flags & WF_F_MEDIUM) ? 2u : 0u);
        coin = jt_rng_float(rkey, at) < 0.5f ? WF_RAY_COIN : 0;
      }
      B.ray1[s] = make_float4(ray.d.y, ray.d.z, __int_as_float(coin), __int_as_float(pix));
      B.wgt[s] = make_float4(weight.x, weight.y, weight.z, 0.0f);
      B.rad[s] = make_float4(radiance.x, radiance.y, radiance.z, 0.0f);
      wf_store_ctl(B, s, p);
    } else {
      to_probe = false;
      wf_end_sample(B, s, p, radiance);  // committed + regenerated by k_wf_regen
    }
  }
}

template <int SAMPLER, int MODE>
__global__ void __launch_bounds__(JT_SHADE_BLOCK, JT_SHADE_MINBLOCKS) k_wf_shade(JtDevScene S, WfBuffers B, DevState st, DevParams P, int next,
                                                  int sample_end, unsigned long long* counters) {
  // block -> (key, first entry of the block inside that key's queue); every key's segment is a whole number of blocks
  int key = -1, s = -1;
  PathCounters cnt{0u, 0u};
  {
    int first_block = 0;
#pragma unroll
    for (int k = 0; k < WF_NKEY; k++) {
      const int c = B.counts[WF_C_SHADEK(k)];
      const int blocks = (c + JT_SHADE_BLOCK - 1) / JT_SHADE_BLOCK;
      if (key < 0 && (int)blockIdx.x < first_block + blocks) {
        key = k == WF_KEY_MISS ? WF_TYPE_MISS : WF_KEY_TYPE(k);  // the body only needs the material type
        const int at = ((int)blockIdx.x - first_block) * JT_SHADE_BLOCK + (int)threadIdx.x;
        if (at < c) s = B.q_shade[(size_t)k * B.n + at];
      }
      first_block += blocks;
    }
  }
  bool to_extend = false, to_probe = false;
#if JT_SHADE_SPECIALISE
  if (SAMPLER == 1 && MODE == MODE_WIDE) {  // the benchmarked combination; the others share the generic body
    switch (key) {
      case 0: wf_shade_slot<SAMPLER, MODE, 0>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 1: wf_shade_slot<SAMPLER, MODE, 1>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 2: wf_shade_slot<SAMPLER, MODE, 2>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 3: wf_shade_slot<SAMPLER, MODE, 3>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 4: wf_shade_slot<SAMPLER, MODE, 4>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 5: wf_shade_slot<SAMPLER, MODE, 5>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case 6: wf_shade_slot<SAMPLER, MODE, 6>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      case WF_TYPE_MISS: wf_shade_slot<SAMPLER, MODE, WF_TYPE_MISS>(S, B, P, s, key, to_extend, to_probe, cnt); break;
      default: break;  // blocks beyond the queues; key 7 (gltfpbr) is rejected at staging
    }
  } else
#endif
  {
    wf_shade_slot<SAMPLER, MODE, -1>(S, B, P, s, key, to_extend, to_probe, cnt);
  }
  wf_append2(B.q_probe, B.counts + WF_C_PROBE, to_probe, B.q_ext[next], B.counts + WF_C_EXT(next), to_extend, s);
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
    uint64_t rkey = jt_rng_key(P.seed, (uint32_t)__float_as_int(r1.w), p.sample);
    bool alive = wf_roulette(weight, p, rkey);
    if (alive && !(p.bounce < P.bounces)) alive = false;
    if (alive) {
      B.wgt[s] = make_float4(weight.x, weight.y, weight.z, 0.0f);
      wf_store_ctl(B, s, p);
      to_extend = true;
    } else {
      B.regen[s] = 1;  // radiance and the control word (sample index, first-hit flag) were parked by the shade kernel
    }
  }
  wf_append(B.q_ext[next], B.counts + WF_C_EXT(next), to_extend, s);
  unsigned lr = __reduce_add_sync(0xFFFFFFFFu, cnt.light_rays);
  if (lane_id() == 0u && lr) atomicAdd(counters + 2, (unsigned long long)lr);
}

// ---- regen + advance -------------------------------------------------------------------------------------------
// Closes an iteration: (1) the flagged slots (sample ended / waiting for their pixel's turn / looking for work) are
// compacted IN SLOT ORDER, block by block, and visited in groups of 32 consecutive entries per warp: commit, claim the next
// sample (wf_regen_slot), append the slots that started a camera ray to the next extend queue in that order -- so the
// camera rays of neighbouring pixels sit in neighbouring lanes of the extend kernel (queues built by per-thread atomics
// scatter them among the bounce rays, and the same mix traverses 11 % slower, tools/exp_coherence.py: 2 534 vs 2 827
// Mrays/s on classroom); (2) the consumed queues' counters are recycled.
#define WF_REGEN_BLOCK 256
#ifndef WF_REGEN_PER_THREAD
#define WF_REGEN_PER_THREAD 4 /* slots per thread: one 128-bit (16) or one 32-bit (4) load of flags; 4 = 4x the blocks */
#endif
#ifndef JT_EMU_COUNT
__global__ void __launch_bounds__(WF_REGEN_BLOCK) k_wf_regen(JtDevScene S, WfBuffers B, DevState st, DevParams P, int cur,
                                                             int sample_end, int epoch, unsigned long long* counters) {
  __shared__ int warp_sums[WF_REGEN_BLOCK / 32];
  __shared__ int flagged[WF_REGEN_BLOCK * WF_REGEN_PER_THREAD];  // flagged slots of this block, in slot order
  const int next = cur ^ 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    B.counts[WF_C_EXT(cur)] = 0;
    B.counts[WF_C_PROBE] = 0;
    B.counts[WF_C_FETCH] = 0;
    for (int k = 0; k < WF_NKEY; k++) B.counts[WF_C_SHADEK(k)] = 0;
  }
  const int first = (blockIdx.x * WF_REGEN_BLOCK + threadIdx.x) * WF_REGEN_PER_THREAD;
  unsigned flags = 0u;  // bit i: slot first + i is flagged
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
      for (int j = 0; j < 4; j++) flags |= (((w[k] >> (8 * j)) & 3u) != 0u ? 1u : 0u) << (4 * k + j);
  } else {
    for (int i = 0; i < WF_REGEN_PER_THREAD && first + i < B.n; i++)
      if (B.regen[first + i]) flags |= 1u << i;
  }
  const int mine = __popc(flags);
  int incl = mine;  // block-wide exclusive scan: positions in the flagged list
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if ((int)lane_id() >= d) incl += t;
  }
  const int warp = threadIdx.x >> 5;
  if (lane_id() == 31u) warp_sums[warp] = incl;
  __syncthreads();
  int warp_off = 0, total = 0;
#pragma unroll
  for (int k = 0; k < WF_REGEN_BLOCK / 32; k++) {
    int v = warp_sums[k];
    if (k < warp) warp_off += v;
    total += v;
  }
  int at = warp_off + incl - mine;
  while (flags) {
    int i = __ffs((int)flags) - 1;
    flags &= flags - 1u;
    flagged[at++] = first + i;
  }
  __syncthreads();
  // consecutive lanes take consecutive flagged slots (= neighbouring pixels, mostly): coalesced accumulator updates
  int done = 0, stolen = 0;
  for (int base = warp * 32; base < total; base += WF_REGEN_BLOCK) {
    const int j = base + (int)lane_id();
    int slot = -1, what = 0;
    if (j < total) {
      slot = flagged[j];
      const int flag_in = B.regen[slot];
      unsigned char flag_out = 0;
      what = wf_regen_slot(S, B, st, P, slot, cur, sample_end, epoch, flag_in, &flag_out);
      if (flag_out != (unsigned char)flag_in) B.regen[slot] = flag_out;
      done += what == WF_REGEN_DONE;
      stolen += what == WF_REGEN_STOLEN;
    }
    wf_append(B.q_ext[next], B.counts + WF_C_EXT(next), what == WF_REGEN_QUEUED || what == WF_REGEN_STOLEN, slot);
  }
  done = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)done);
  stolen = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)stolen);
  if (lane_id() == 0u && done) atomicAdd(B.counts + WF_C_DONE, done);
  if (lane_id() == 0u && stolen) atomicAdd(counters + 3, (unsigned long long)stolen);
}
#endif
