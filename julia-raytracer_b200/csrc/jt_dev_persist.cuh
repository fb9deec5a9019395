// jt_dev_persist.cuh -- persistent-warp closest-hit kernel over the wide BVH.
//
// ncu on the first wavefront extend kernel (profiles/r01/ncu_full_wavefront_v1_*): 6.3 of 32 lanes
// active per warp instruction, issue-bound. Two causes, two remedies (Aila & Laine 2009; Ylitie,
// Karras & Laine 2017):
//   * rays of one warp need very different numbers of steps -> persistent warps that REFILL idle
//     lanes from the ray queue whenever fewer than JT_FETCH_THRESHOLD lanes are still traversing;
//   * node steps and triangle tests interleave differently per lane -> every loop iteration is one
//     node step for all lanes that have one, and triangle tests run only when at least 1/4 of the
//     live lanes hold triangles (others postpone theirs onto the traversal stack).
// The loop is warp-uniform (all 32 lanes iterate together, idle lanes predicated off) so every
// ballot is full-mask. Results are identical to wide_walk(): the closest hit and its tie-break do
// not depend on the order in which candidates are tested.
#pragma once
#include "jt_dev_traverse.cuh"

#ifndef JT_FETCH_THRESHOLD
#define JT_FETCH_THRESHOLD 20
#endif
#ifndef JT_PERSIST_BLOCK
#define JT_PERSIST_BLOCK 384 /* x 2 blocks per SM; 128 x 6: -5 %, 256 x 3: -1 %, 768 x 1: -1 % (tuning_variants.txt) */
#endif

// Traversal stack: entries 0 .. JT_SMEM_STACK-1 live in shared memory (one column per thread, so a warp's accesses to
// one level are conflict-free 8-byte words), deeper ones in local memory. JT_SMEM_STACK = 0: all local.
#ifndef JT_SMEM_STACK
#define JT_SMEM_STACK 0
#endif
struct TravStack {
  uint2* local;
#if JT_SMEM_STACK > 0 && !defined(JT_EMU_COUNT)
  uint2* shared;  // this thread's column: element i at shared[i * JT_PERSIST_BLOCK]
  JT_DEV void put(int i, uint2 v) const {
    if (i < JT_SMEM_STACK) shared[i * JT_PERSIST_BLOCK] = v;
    else local[i - JT_SMEM_STACK] = v;
  }
  JT_DEV uint2 get(int i) const { return i < JT_SMEM_STACK ? shared[i * JT_PERSIST_BLOCK] : local[i - JT_SMEM_STACK]; }
#else
  JT_DEV void put(int i, uint2 v) const { local[i] = v; }
  JT_DEV uint2 get(int i) const { return local[i]; }
#endif
};

// Postponed work is popped several iterations after it was pushed: warm the L1 with its first record meanwhile.
#ifndef JT_PREFETCH_POSTPONED
#define JT_PREFETCH_POSTPONED 0
#endif
JT_DEV void prefetch_l1(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

struct PersistLane {
  WideRay R;            // current space (world or instance)
  uint2 ngroup, tgroup;
  int sp, blas_sp, cur_inst;
  WideBest best;
  float best_t, tmin;
  uint32_t world_oct;
  f3 wo, wd;            // world ray (kept to leave a BLAS)
};

JT_DEV void persist_init(PersistLane& L, const JtDevScene& S, f3 o, f3 d, float tmin, float tmax, int root,
                         int single_inst) {
  wide_ray_setup(L.R, o, d);
  L.wo = o;
  L.wd = d;
  L.world_oct = L.R.rank_oct;
  L.ngroup = make_uint2((uint32_t)root, 0x80000000u);
  L.tgroup = make_uint2(0u, 0u);
  L.sp = 0;
  L.blas_sp = -1;
  L.cur_inst = single_inst;
  L.best = WideBest{0.0f, 0.0f, 0.0f, -1, -1, -1};
  L.best_t = tmax;
  L.tmin = tmin;
}

// Launch tail (tools/sim_persist_tail.py: 27 % of a full launch's warp iterations run after the queue is empty, at ~10
// of 32 lanes; 66 % for a quarter-full queue): once nothing is left to fetch, a warp that is down to fewer than
// JT_SUSPEND_BELOW live lanes SUSPENDS its stragglers instead of dragging them to the end at a few lanes per
// instruction. The traversal state (closest hit so far, the pending node / triangle groups) is parked in the slot, the
// slot goes straight to the next iteration's extend queue, and the next launch resumes it inside a full warp. A slot
// holds one ray of one pixel at a time, so delaying it by an iteration changes no result. Every ray runs at least
// JT_SUSPEND_MIN_ITERS loop iterations per launch (progress).
#ifndef JT_SUSPEND_BELOW
#define JT_SUSPEND_BELOW 8 /* 0 = never suspend; 8 / 12 / 16 / 24 measured in profiles/r02/tuning_variants.txt */
#endif
#ifndef JT_SUSPEND_MIN_ITERS
#define JT_SUSPEND_MIN_ITERS 6
#endif
#define JT_SUSPEND_STACK 16 /* parked stack entries per slot (8 B each); deeper lanes simply run on */
#ifdef JT_EMU_COUNT
// host emulation (single-lane warps that always find `more`): tests force a suspension every N iterations instead
static int jt_emu_suspend_every = 0;
JT_DEV bool persist_wants_suspend(bool more, int lanes, int iters) {
  (void)more; (void)lanes;
  return jt_emu_suspend_every > 0 && iters >= jt_emu_suspend_every;
}
#else
JT_DEV bool persist_wants_suspend(bool more, int lanes, int iters) {
  return JT_SUSPEND_BELOW > 0 && !more && lanes < JT_SUSPEND_BELOW && iters >= JT_SUSPEND_MIN_ITERS;
}
#endif

// One warp-uniform traversal episode: runs until every lane has finished its ray, or (when `more` rays are waiting)
// fewer than JT_FETCH_THRESHOLD lanes remain live, or (allow_suspend) the warp should park its stragglers: returns true
// in that last case. `live` is updated per lane.
JT_DEV bool persist_traverse(const JtDevScene& S, PersistLane& L, const TravStack& stack, bool& live, bool more,
                             bool allow_suspend = false) {
  const unsigned FULL = 0xFFFFFFFFu;
  int iters = 0;
  for (;; iters++) {
    unsigned am = __ballot_sync(FULL, live);
    if (am == 0u) break;
    if (more && __popc(am) < JT_FETCH_THRESHOLD) break;
    if (allow_suspend && persist_wants_suspend(more, __popc(am), iters)) return true;
    // ---- A: one node step -------------------------------------------------------------------------
    if (live) {
      if (L.ngroup.y > 0x00FFFFFFu) {
        uint32_t hits = L.ngroup.y;
        uint32_t bit = 31u - (uint32_t)__clz(hits);
        hits &= ~(1u << bit);
        L.ngroup.y = hits;
        if (hits > 0x00FFFFFFu) stack.put(L.sp++, L.ngroup);
        uint32_t slot = (bit - 24u) ^ L.R.oct;
        uint32_t rel = __popc(hits & 0xFFu & ~(0xFFFFFFFFu << slot));
        wide_node_hits(S.wnodes, L.ngroup.x + rel, L.R, L.tmin, L.best_t, &L.ngroup, &L.tgroup);
      } else {
        L.tgroup = L.ngroup;  // the popped entry was a (postponed) triangle group
        L.ngroup = make_uint2(0u, 0u);
      }
    }
    // ---- B: triangle tests, only while enough lanes have some ----------------------------------------
#ifndef JT_TRI_QUORUM_DIV
#define JT_TRI_QUORUM_DIV 4
#endif
    const int quorum = __popc(am) / JT_TRI_QUORUM_DIV;
    const int threshold = quorum > 1 ? quorum : 1;
    for (;;) {
      bool has = live && L.tgroup.y != 0u;
      unsigned tm = __ballot_sync(FULL, has);
      if (tm == 0u || __popc(tm) < threshold) break;
      if (has) {
        uint32_t bit = (uint32_t)__ffs((int)L.tgroup.y) - 1u;
        L.tgroup.y &= ~(1u << bit);
        uint32_t wtri = L.tgroup.x + bit;
        const float4* tp = S.wtris + 3 * (size_t)wtri;
        float4 r0 = __ldg(tp), r1 = __ldg(tp + 1), r2 = __ldg(tp + 2);
        uint32_t flags = __float_as_uint(r2.w);
        if (flags & 0x100u) {  // instance record: park world-level work, enter the BLAS
          int inst = __float_as_int(r1.w);
          if (L.tgroup.y != 0u) stack.put(L.sp++, L.tgroup);
          if (L.ngroup.y > 0x00FFFFFFu) stack.put(L.sp++, L.ngroup);
          L.blas_sp = L.sp;
          L.cur_inst = inst;
          const JtInstanceRec& I = S.instances[inst];
          wide_ray_setup(L.R, xform_point(I.inv, L.wo), xform_vector(I.inv, L.wd));
          const int entry = __float_as_int(r0.w);  // braided sub-tree of the BLAS, or -1: the whole shape
          L.ngroup = make_uint2((uint32_t)(entry >= 0 ? entry : S.shapes[I.shape].wide_root), 0x80000000u);
          L.tgroup = make_uint2(0u, 0u);
        } else {
          float t, u, v;
          f3 to = L.R.o, td = L.R.d;
          uint32_t local_oct = L.R.rank_oct;
          if (flags & 0x200u) flat_ray(S, __float_as_int(r1.w), L.wo, L.wd, to, td, local_oct);
          if (tri_test(to, td, L.tmin, L.best_t, f3{r0.x, r0.y, r0.z}, f3{r1.x, r1.y, r1.z},
                       f3{r2.x, r2.y, r2.z}, &t, &u, &v)) {
            int inst = L.cur_inst >= 0 ? L.cur_inst : __float_as_int(r1.w);
            if (flags & 1u) {
              u = 1.0f - u;
              v = 1.0f - v;
            }
            wide_accept(S, L.best, t, u, v, inst, __float_as_int(r0.w), (int)wtri, L.world_oct, local_oct);
            L.best_t = L.best.t;
          }
        }
      }
    }
    // ---- C: postpone leftovers, pop the next group, leave the BLAS, or finish ---------------------------
    if (live) {
      if (L.tgroup.y != 0u) {
        stack.put(L.sp++, L.tgroup);
#if JT_PREFETCH_POSTPONED
        {
          const float4* tp = S.wtris + 3 * (size_t)(L.tgroup.x + (uint32_t)__ffs((int)L.tgroup.y) - 1u);
          prefetch_l1(tp);
          prefetch_l1(tp + 2);
        }
#endif
        L.tgroup = make_uint2(0u, 0u);
      }
      if (L.ngroup.y <= 0x00FFFFFFu) {
        if (L.sp == L.blas_sp) {
          L.blas_sp = -1;
          L.cur_inst = -1;
          wide_ray_setup(L.R, L.wo, L.wd);
        }
        if (L.sp == 0) live = false;
        else L.ngroup = stack.get(--L.sp);
      }
    }
  }
  return false;
}

// Park / resume a live lane at the top of the traversal loop (tgroup is empty there; ngroup is the current entry).
//   hit0 = {inst, elem, u, v}   hit1 = {t, wtri, entries | (blas_sp + 1) << 8, cur_inst}   entries = sp + 1 <= JT_SUSPEND_STACK
JT_DEV bool persist_can_park(const PersistLane& L) { return L.sp + 1 <= JT_SUSPEND_STACK; }
JT_DEV void persist_park(const PersistLane& L, const TravStack& stack, float4* hit0, float4* hit1, uint2* parked) {
  *hit0 = make_float4(__int_as_float(L.best.inst), __int_as_float(L.best.elem), L.best.u, L.best.v);
  *hit1 = make_float4(L.best.t, __int_as_float(L.best.wtri), __int_as_float((L.sp + 1) | ((L.blas_sp + 1) << 8)),
                      __int_as_float(L.cur_inst));
  for (int i = 0; i < L.sp; i++) parked[i] = stack.get(i);
  parked[L.sp] = L.ngroup;
}
// after persist_init(L, ...) with the slot's world ray
JT_DEV void persist_resume(PersistLane& L, const JtDevScene& S, const TravStack& stack, float4 hit0, float4 hit1,
                           const uint2* parked) {
  L.best.inst = __float_as_int(hit0.x);
  L.best.elem = __float_as_int(hit0.y);
  L.best.u = hit0.z;
  L.best.v = hit0.w;
  L.best.t = hit1.x;
  L.best.wtri = __float_as_int(hit1.y);
  if (L.best.inst >= 0) L.best_t = L.best.t;
  const int packed = __float_as_int(hit1.z);
  const int entries = packed & 0xFF;
  L.blas_sp = (packed >> 8) - 1;
  L.cur_inst = __float_as_int(hit1.w);
  if (L.cur_inst >= 0) {  // the same arithmetic as at BLAS entry: identical instance-space ray
    const JtInstanceRec& I = S.instances[L.cur_inst];
    wide_ray_setup(L.R, xform_point(I.inv, L.wo), xform_vector(I.inv, L.wd));
  }
  L.sp = entries - 1;
  for (int i = 0; i < L.sp; i++) stack.put(i, parked[i]);
  L.ngroup = parked[L.sp];
  L.tgroup = make_uint2(0u, 0u);
}

// Warp-aggregated fetch of the next queue indices for the lanes that `want` one. All 32 lanes call.
JT_DEV int persist_fetch(int* counter, bool want, int count) {
  const unsigned FULL = 0xFFFFFFFFu;
  unsigned m = __ballot_sync(FULL, want);
  if (m == 0u) return -1;
  unsigned lane = threadIdx.x & 31u;
  int leader = __ffs((int)m) - 1;
  int base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(FULL, base, leader);
  int idx = base + __popc(m & ((1u << lane) - 1u));
  return (want && idx < count) ? idx : -1;
}
