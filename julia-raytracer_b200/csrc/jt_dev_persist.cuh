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

// One warp-uniform traversal episode: runs until every lane has finished its ray or (when `more` rays
// are waiting) fewer than JT_FETCH_THRESHOLD lanes remain live. `live` is updated per lane.
JT_DEV void persist_traverse(const JtDevScene& S, PersistLane& L, const TravStack& stack, bool& live, bool more) {
  const unsigned FULL = 0xFFFFFFFFu;
  for (;;) {
    unsigned am = __ballot_sync(FULL, live);
    if (am == 0u) break;
    if (more && __popc(am) < JT_FETCH_THRESHOLD) break;
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
