// jt_wide_bvh.cpp -- collapses the host-built binary BVHs into the 8-wide quantised BVH the fast
// traversal kernel walks (compressed-wide-BVH layout in the spirit of Ylitie, Karras & Laine 2017).
//
// Node = 80 B = 5 x 128-bit loads:
//   p[3]        f32  origin of the quantisation grid (node box minimum)
//   e[3]        u8   per-axis biased exponents: grid step = 2^(e-127)
//   imask       u8   bit s set <=> slot s holds an internal child
//   child_base  u32  index of the first internal child (internal children are contiguous, slot order)
//   prim_base   u32  index of the first triangle record of this node's leaf children
//   meta[8]     u8   empty: 0 | internal: 0b001_11sss (s = slot) | leaf: unary count << 5 | offset
//   qlo[3][8], qhi[3][8]  u8 child boxes on the grid (floor / ceil: always conservative)
// Slots are assigned so that slot s (bit0 = +x, bit1 = +y, bit2 = +z) holds the child lying in
// that direction from the node centre; visiting slots by decreasing (s XOR ray_octant) is then a
// near-to-far order without any distance sort.
//
// Triangle record = 48 B = 3 x 128-bit loads: {p1, element} {p2-p1, instance} {p3-p1, flags}.
// Quads become one or two records (src/geometry.jl:238-258: (p1,p2,p4) and (p3,p4,p2)).
//
// Exactness contract: the leaf test is the reference's Moeller-Trumbore with the reference's
// operation order, so t/u/v are bit-identical; the tree only decides WHICH primitives are tested.
// The closest hit is order independent; exact-t ties are resolved with the per-octant visit
// ranks computed here from the reference tree (SURVEY.md §8a "tie-break contract").
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <functional>
#include <limits>

#include "jt_internal.h"

namespace {

struct Box3 {
  float lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::numeric_limits<float>::infinity();
      hi[k] = -std::numeric_limits<float>::infinity();
    }
  }
  void add(const Box3& b) {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], b.lo[k]);
      hi[k] = std::max(hi[k], b.hi[k]);
    }
  }
  void add(const float* p) {
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], p[k]);
      hi[k] = std::max(hi[k], p[k]);
    }
  }
  float area() const {
    float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    if (!(x >= 0) || !(y >= 0) || !(z >= 0)) return 0.0f;
    return 2.0f * (x * y + y * z + z * x);
  }
  bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
};

struct Rec {  // one leaf record + its reference visit ranks
  Box3 box;
  JtWideTri tri;
  uint32_t rank[8];
};

struct BNode {  // intermediate binary tree. No default initialisers: the node array of a flattened scene is 1.4 GB and
  Box3 box;     // is sized up front; every creation site writes all five fields
  int left, right;   // children (internal), -1 / -1 for a leaf
  int first, count;  // records (leaf, count <= 3)
  bool leaf() const { return left < 0; }
};

struct Ref {  // what the SAH build streams and partitions in place: box + record id (28 B, contiguous)
  float lo[3], hi[3];
  int rec;
};

enum { NODE_CHUNK = 256, TASK_MIN = 8192 };

struct Tree {
  JtBigVec<Rec> recs;  // (JtBigVec: resize() does not zero -- these are GB-sized for a flattened scene)
  // filled by the (multi-threaded) SAH build: storage is sized up front, slots are handed out atomically
  JtBigVec<BNode> nodes;
  JtBigVec<int> leaf_store;  // record ids of the leaves, contiguous per leaf
  std::atomic<int> next_node{0}, next_leaf{0};
  void reserve_for(size_t nrecs) {
    // 2n - 1 nodes at most, plus the tails of the per-task allocation chunks (NodeChunk)
    nodes.resize(2 * nrecs + 1 + NODE_CHUNK * (2 * nrecs / TASK_MIN + 64));
    leaf_store.resize(nrecs);
    next_node = 0;
    next_leaf = 0;
  }
  int alloc_nodes(int n) { return next_node.fetch_add(n); }
  int alloc_leaf(int n) { return next_leaf.fetch_add(n); }
};

// Node indices are handed out in per-task chunks: one atomic per NODE_CHUNK nodes, and the nodes one thread writes
// are neighbours in memory.
struct NodeChunk {
  int next = 0, end = 0;
  int get(Tree& t) {
    if (next == end) {
      next = t.alloc_nodes(NODE_CHUNK);
      end = next + NODE_CHUNK;
    }
    return next++;
  }
};

// Output of one emit job: a wide-node subtree in LOCAL indices (node 0 = the job's own node).
struct Sink {
  std::vector<JtWideNode> nodes;
  std::vector<JtWideTri> tris;
  std::vector<uint32_t> rank[8];
};

class Collapser {
 public:
  struct Work {
    int wide, bnode;
  };
  Collapser(const std::vector<JtHostShape>& shapes, const std::vector<JtHostInstance>& insts,
            JtWideResult* out)
      : shapes_(shapes), insts_(insts), out_(out) {}

  // ---- reference visit order (per octant) of the elements of one binary BVH --------------------
  // octant bit k set <=> d[k] < 0 (ray_dsign, src/bvh.jl:323). At an internal node with split
  // axis a: d[a] >= 0 -> child start+1 is visited first, else child start (Q1).
  static void visit_ranks(const std::vector<jt_bvh_node>& nodes, const std::vector<int64_t>& prims,
                          int64_t nprims, std::vector<uint32_t> rank[8]) {
    for (int oct = 0; oct < 8; oct++) {
      rank[oct].assign((size_t)nprims, 0u);
      if (nodes.empty()) continue;
      uint32_t counter = 0;
      std::vector<int64_t> st;
      st.push_back(0);
      while (!st.empty()) {
        const jt_bvh_node& n = nodes[(size_t)st.back()];
        st.pop_back();
        if (n.internal) {
          int64_t a = n.start - 1, b = n.start;
          bool neg = (oct >> (n.axis - 1)) & 1;
          // stack: push the one visited LAST first
          if (!neg) {
            st.push_back(a);
            st.push_back(b);
          } else {
            st.push_back(b);
            st.push_back(a);
          }
        } else {
          for (int64_t i = n.start - 1; i < n.start - 1 + n.num; i++) rank[oct][(size_t)(prims[(size_t)i] - 1)] = counter++;
        }
      }
    }
  }

  // ---- records of one element ---------------------------------------------------------------------
  void element_records(const JtHostShape& s, int64_t elem, int inst, const std::vector<uint32_t> erank[8],
                       std::vector<int>* out_recs, Tree& t) {
    const int32_t* q = &s.elems[4 * (size_t)elem];
    auto P = [&](int v) { return &s.pos[3 * (size_t)v]; };
    auto push = [&](const float* a, const float* b, const float* c, uint32_t flags) {
      Rec r;
      r.box.reset();
      r.box.add(a);
      r.box.add(b);
      r.box.add(c);
      memset(&r.tri, 0, sizeof(r.tri));
      for (int k = 0; k < 3; k++) {
        r.tri.p1[k] = a[k];
        r.tri.e1[k] = b[k] - a[k];  // edge1 = p2 - p1, src/geometry.jl:207
        r.tri.e2[k] = c[k] - a[k];  // edge2 = p3 - p1, :208
      }
      r.tri.element = (int32_t)elem;
      r.tri.instance = inst;
      r.tri.flags = flags;
      for (int o = 0; o < 8; o++) r.rank[o] = 2u * erank[o][(size_t)elem] + (flags & 1u);
      out_recs->push_back((int)t.recs.size());
      t.recs.push_back(r);
    };
    if (s.kind == 1) {
      push(P(q[0]), P(q[1]), P(q[2]), 0);
    } else {
      const float *p1 = P(q[0]), *p2 = P(q[1]), *p3 = P(q[2]), *p4 = P(q[3]);
      bool degenerate = p3[0] == p4[0] && p3[1] == p4[1] && p3[2] == p4[2];  // Q11: positions
      push(p1, p2, p4, 0);
      if (!degenerate) push(p3, p4, p2, 1);
    }
  }

  const std::vector<uint32_t>* shape_rank(int shape_id) {
    if (shape_ranks_.size() != shapes_.size()) shape_ranks_.resize(shapes_.size());
    auto& r = shape_ranks_[(size_t)shape_id];
    if (!r.built) {
      const JtHostShape& s = shapes_[(size_t)shape_id];
      visit_ranks(s.ref_nodes, s.ref_prims, s.num_elements(), r.rank);
      r.built = true;
    }
    return r.rank;
  }

  // ---- binned-SAH binary tree over records (the fast mode's own topology) -----------------------------------
  // The reference tree (centroid middle split, src/bvh.jl:185-216) only defines tie-break ranks; which boxes
  // a ray visits is free, so the wide BVH is collapsed from a surface-area-heuristic tree instead:
  // 16 bins per axis, leaves of <= 3 records (the per-slot limit of the wide node). The build streams a compact
  // array of (box, record id) that it partitions in place, and forks subtrees as OpenMP tasks (a flattened
  // ecosys is 16.8 M records).
  int sah_root(const std::vector<int>& recs, Tree& t) {
    if (recs.empty()) return -1;
    JtBigVec<Ref> refs(recs.size());
#pragma omp parallel for schedule(static) if ((int64_t)recs.size() > parallel_min_) num_threads(build_threads_)
    for (long i = 0; i < (long)recs.size(); i++) {
      const Box3& b = t.recs[(size_t)recs[i]].box;
      for (int k = 0; k < 3; k++) {
        refs[i].lo[k] = b.lo[k];
        refs[i].hi[k] = b.hi[k];
      }
      refs[i].rec = recs[i];
    }
    t.reserve_for(recs.size());
    // the exact small-node sweep is a build-time measure for huge trees (16.8 M records: 26 -> 21 s per thread); on the
    // ordinary scenes the 16-bin splits give slightly better trees (cornellbox 1.6 vs 2.4, classroom 12.1 vs 12.3 node
    // visits per ray), so they keep them
    small_sweep_ = recs.size() > ((size_t)2 << 20);
    int root = -1;
    // small trees stay on the calling thread: waking the pool costs more than they do
#pragma omp parallel if ((int64_t)recs.size() > parallel_min_) num_threads(build_threads_)
#pragma omp single nowait
    {
      NodeChunk chunk;
      root = sah_tree(refs.data(), 0, refs.size(), t, chunk);
    }
    return root;
  }

  // Nodes of a few records (most of the tree: leaves hold <= 3) skip the 3 x 16 bins, whose set-up and sweep
  // cost more than the records themselves: every split position of the centroid order of each axis is evaluated
  // exactly, with the same cost function as the binned path.
  enum { SMALL_NODE = 12 };
  int sah_small(Ref* refs, size_t a, size_t b, Tree& t, NodeChunk& chunk, const Box3& box) {
    const int n = (int)(b - a);
    auto ref_box = [](const Ref& r) {
      Box3 x;
      for (int k = 0; k < 3; k++) {
        x.lo[k] = r.lo[k];
        x.hi[k] = r.hi[k];
      }
      return x;
    };
    int order[3][SMALL_NODE];
    int best_axis = -1, best_pos = -1;
    float best_cost = std::numeric_limits<float>::infinity(), best_tri_cost = 0.0f;
    for (int ax = 0; ax < 3; ax++) {
      int* o = order[ax];
      float key[SMALL_NODE] = {};
      for (int i = 0; i < n; i++) {  // insertion sort by centroid
        float c = refs[a + (size_t)i].lo[ax] + refs[a + (size_t)i].hi[ax];
        int j = i;
        while (j > 0 && key[j - 1] > c) {
          key[j] = key[j - 1];
          o[j] = o[j - 1];
          j--;
        }
        key[j] = c;
        o[j] = i;
      }
      if (!(key[n - 1] > key[0])) continue;  // identical centroids on this axis
      float right_area[SMALL_NODE];
      Box3 acc;
      acc.reset();
      for (int i = n - 1; i > 0; i--) {
        acc.add(ref_box(refs[a + (size_t)o[i]]));
        right_area[i] = acc.area();
      }
      acc.reset();
      for (int i = 1; i < n; i++) {  // left = o[0..i), right = o[i..n)
        acc.add(ref_box(refs[a + (size_t)o[i - 1]]));
        int cl = i, cr = n - i;
        float la = acc.area();
        float cost = la * (float)((cl + leaf_max_ - 1) / leaf_max_) + right_area[i] * (float)((cr + leaf_max_ - 1) / leaf_max_);
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = ax;
          best_pos = i;
          best_tri_cost = la * (float)cl + right_area[i] * (float)cr;
        }
      }
    }
    auto make_leaf = [&]() {
      int id = chunk.get(t);
      BNode& leaf = t.nodes[(size_t)id];
      leaf.box = box;
      leaf.left = leaf.right = -1;
      leaf.first = t.alloc_leaf(n);
      leaf.count = n;
      for (int i = 0; i < n; i++) t.leaf_store[(size_t)leaf.first + (size_t)i] = refs[a + (size_t)i].rec;
      return id;
    };
    if (n <= leaf_max_) {  // leaf or split: same rule as the binned path
      if (best_axis < 0) return make_leaf();
      float leaf_cost = tri_cost_ * box.area() * (float)n;
      float split_cost = tri_cost_ * best_tri_cost + 2.0f * box.area();
      if (split_cost >= leaf_cost) return make_leaf();
    }
    size_t m;
    if (best_axis < 0) {
      m = a + (size_t)n / 2;
    } else {
      Ref tmp[SMALL_NODE];
      for (int i = 0; i < n; i++) tmp[i] = refs[a + (size_t)order[best_axis][i]];
      for (int i = 0; i < n; i++) refs[a + (size_t)i] = tmp[i];
      m = a + (size_t)best_pos;
    }
    int l = sah_tree(refs, a, m, t, chunk);
    int r = sah_tree(refs, m, b, t, chunk);
    int id = chunk.get(t);
    BNode& in = t.nodes[(size_t)id];
    in.left = l;
    in.right = r;
    in.first = in.count = 0;
    in.box = t.nodes[(size_t)l].box;
    in.box.add(t.nodes[(size_t)r].box);
    return id;
  }

  int sah_tree(Ref* refs, size_t a, size_t b, Tree& t, NodeChunk& chunk) {
    size_t n = b - a;
    if (n == 0) return -1;
    // Big nodes choose their split plane from a strided sample (~128 k references): the binned SAH estimate barely
    // moves, and the passes over the top levels -- which no task parallelism can hide -- shrink to the partition.
    // (`box` is then only an estimate too; it is used for the leaf decision of nodes with <= 3 records, never here.)
    const size_t stride = n > ((size_t)1 << 18) ? n >> 17 : 1;
    Box3 box, cbox;
    box.reset();
    cbox.reset();
    for (size_t i = a; i < b; i += stride) {
      const Ref& r = refs[i];
      for (int k = 0; k < 3; k++) {
        box.lo[k] = std::min(box.lo[k], r.lo[k]);
        box.hi[k] = std::max(box.hi[k], r.hi[k]);
        float c = 0.5f * (r.lo[k] + r.hi[k]);
        cbox.lo[k] = std::min(cbox.lo[k], c);
        cbox.hi[k] = std::max(cbox.hi[k], c);
      }
    }
    auto make_leaf = [&]() {
      int id = chunk.get(t);
      BNode& leaf = t.nodes[(size_t)id];
      leaf.box = box;
      leaf.left = leaf.right = -1;
      leaf.first = t.alloc_leaf((int)n);
      leaf.count = (int)n;
      for (size_t i = a; i < b; i++) t.leaf_store[(size_t)leaf.first + (i - a)] = refs[i].rec;
      return id;
    };
    if (n == 1) return make_leaf();
    if (small_sweep_ && n <= SMALL_NODE) return sah_small(refs, a, b, t, chunk, box);
    const int NB = 16;
    // one pass fills the bins of all three axes; boxes are kept as two 4-float vectors (lane 3 unused) so that a bin
    // update is one vector min + one vector max -- this loop is where the build of a 17 M-record scene spends its time
    typedef float v4 __attribute__((vector_size(16)));
    struct VBin {
      v4 lo, hi;
    };
    VBin vb[3][NB];
    int cnt[3][NB];
    float lo3[3], scale3[3];
    bool axis_ok[3];
    const float inf = std::numeric_limits<float>::infinity();
    for (int ax = 0; ax < 3; ax++) {
      float ext = cbox.hi[ax] - cbox.lo[ax];
      axis_ok[ax] = ext > 0.0f;
      lo3[ax] = cbox.lo[ax];
      scale3[ax] = axis_ok[ax] ? (float)NB / ext : 0.0f;
      for (int k = 0; k < NB; k++) {
        vb[ax][k].lo = v4{inf, inf, inf, inf};
        vb[ax][k].hi = v4{-inf, -inf, -inf, -inf};
        cnt[ax][k] = 0;
      }
    }
    for (size_t i = a; i < b; i += stride) {
      const Ref& r = refs[i];
      const v4 rlo = v4{r.lo[0], r.lo[1], r.lo[2], 0.0f}, rhi = v4{r.hi[0], r.hi[1], r.hi[2], 0.0f};
      for (int ax = 0; ax < 3; ax++) {
        if (!axis_ok[ax]) continue;
        int k = (int)((0.5f * (r.lo[ax] + r.hi[ax]) - lo3[ax]) * scale3[ax]);
        k = std::max(0, std::min(NB - 1, k));
        VBin& bb = vb[ax][k];
        bb.lo = rlo < bb.lo ? rlo : bb.lo;
        bb.hi = rhi > bb.hi ? rhi : bb.hi;
        cnt[ax][k]++;
      }
    }
    Box3 bins[3][NB];
    for (int ax = 0; ax < 3; ax++)
      for (int k = 0; k < NB; k++)
        for (int c = 0; c < 3; c++) {
          bins[ax][k].lo[c] = vb[ax][k].lo[c];
          bins[ax][k].hi[c] = vb[ax][k].hi[c];
        }
    int best_axis = -1, best_bin = -1;
    float best_cost = std::numeric_limits<float>::infinity(), best_tri_cost = 0.0f;
    for (int ax = 0; ax < 3; ax++) {
      if (!axis_ok[ax]) continue;
      float right_area[NB];
      int right_cnt[NB];
      Box3 acc;
      acc.reset();
      int c = 0;
      for (int k = NB - 1; k > 0; k--) {
        acc.add(bins[ax][k]);
        c += cnt[ax][k];
        right_area[k] = acc.area();
        right_cnt[k] = c;
      }
      acc.reset();
      c = 0;
      for (int k = 0; k < NB - 1; k++) {
        acc.add(bins[ax][k]);
        c += cnt[ax][k];
        if (c == 0 || right_cnt[k + 1] == 0) continue;
        // wide-node aware cost: a slot holds up to 3 records, so count ceil(n/3) leaf slots per side
        float cost = acc.area() * (float)((c + leaf_max_ - 1) / leaf_max_) +
                     right_area[k + 1] * (float)((right_cnt[k + 1] + leaf_max_ - 1) / leaf_max_);
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = ax;
          best_bin = k;
          best_tri_cost = acc.area() * (float)c + right_area[k + 1] * (float)right_cnt[k + 1];
        }
      }
    }
    if (n <= (size_t)leaf_max_) {
      // Leaf or split? On the GPU one triangle test costs about as many issue slots as a whole 8-child node
      // step (it runs at ~6 of 32 lanes, profiles/r01), so small groups are split whenever the children's
      // boxes are tighter: cost = tri_cost * sum(area_k * n_k) + slab_cost * area(box) per extra slot.
      float leaf_cost = tri_cost_ * box.area() * (float)n;
      if (best_axis < 0) return make_leaf();
      float split_cost = tri_cost_ * best_tri_cost + 2.0f * box.area();
      if (split_cost >= leaf_cost) return make_leaf();
    }
    size_t m;
    if (best_axis < 0) {
      m = a + n / 2;  // identical centroids: split the list
    } else {
      const float lo = lo3[best_axis], scale = scale3[best_axis];
      const int ax = best_axis;
      Ref* mid = std::partition(refs + a, refs + b, [&](const Ref& r) {
        int k = (int)((0.5f * (r.lo[ax] + r.hi[ax]) - lo) * scale);
        k = std::max(0, std::min(NB - 1, k));
        return k <= best_bin;
      });
      m = (size_t)(mid - refs);
      if (m == a || m == b) m = a + n / 2;
    }
    int l = -1, r = -1;
    if (n > TASK_MIN) {
#pragma omp task shared(l, t) firstprivate(refs, a, m)
      {
        NodeChunk mine;
        l = sah_tree(refs, a, m, t, mine);
      }
      r = sah_tree(refs, m, b, t, chunk);
#pragma omp taskwait
    } else {
      l = sah_tree(refs, a, m, t, chunk);
      r = sah_tree(refs, m, b, t, chunk);
    }
    int id = chunk.get(t);
    BNode& in = t.nodes[(size_t)id];
    in.left = l;
    in.right = r;
    in.first = in.count = 0;
    in.box = t.nodes[(size_t)l].box;
    in.box.add(t.nodes[(size_t)r].box);
    return id;
  }

  // all records of a shape (reference element order), optionally with a baked instance id
  void shape_records(int shape_id, int inst, std::vector<int>* recs, Tree& t) {
    const JtHostShape& s = shapes_[(size_t)shape_id];
    if (s.kind == 0 || s.ref_nodes.empty() || s.num_elements() == 0) return;
    const std::vector<uint32_t>* erank = shape_rank(shape_id);
    for (int64_t e = 0; e < s.num_elements(); e++) element_records(s, e, inst, erank, recs, t);
  }

  // ---- which binary nodes become the (up to) 8 children of a wide node ------------------------------------------------
  // Default: dynamic programming over the binary tree (after Ylitie, Karras & Laine 2017, sec. 3.1, adapted to this
  // layout where a leaf of <= 3 records is a child slot, not a node). F(n, i) = least expected cost of representing
  // the subtree of n with at most i child slots of ONE wide node:
  //   F(n, 1) = leaf(n) ? A_n * P_n * c_tri  :  A_n * c_node + D(n, 8)         (n becomes a wide node of its own)
  //   F(n, i) = min(F(n, 1), D(n, i)),   D(n, i) = min_{0<k<i} F(left, k) + F(right, i - k)   (n is dissolved)
  // with A the surface area (visit probability), c_node = one wide node step, c_tri = tri_cost_ / 8 of it per record --
  // the same currency as the SAH build above. JT_COLLAPSE=greedy restores "open the largest child until 8" (round 1).
  struct Dp {
    float f[8];
    uint8_t k[8];  // k[i - 1]: 0 = n takes one slot, else its left child gets k of the i slots
  };
  std::vector<Dp> dp_;
  bool use_dp_ = false;
  size_t dp_max_nodes_ = (size_t)6 << 20;
  void compute_dp(const Tree& t, int root) {
    use_dp_ = false;
    const char* mode = getenv("JT_COLLAPSE");
    const size_t nn = (size_t)t.next_node.load();
    // 40 B of tables per binary node: JT_COLLAPSE_DP_MAX_NODES caps the trees it is used on
    const size_t cap = getenv("JT_COLLAPSE_DP_MAX_NODES") ? (size_t)atoll(getenv("JT_COLLAPSE_DP_MAX_NODES")) : dp_max_nodes_;
    if ((mode && !strcmp(mode, "greedy")) || nn > cap) return;
    dp_.assign(nn, Dp());
    const float c_node = 8.0f, c_tri = tri_cost_;
    std::vector<std::pair<int, int>> st;  // (node, phase)
    st.push_back({root, 0});
    while (!st.empty()) {
      auto [n, phase] = st.back();
      st.pop_back();
      const BNode& b = t.nodes[(size_t)n];
      Dp& d = dp_[(size_t)n];
      if (b.leaf()) {
        const float c = b.box.area() * (float)b.count * c_tri;
        for (int i = 0; i < 8; i++) {
          d.f[i] = c;
          d.k[i] = 0;
        }
        continue;
      }
      if (phase == 0) {
        st.push_back({n, 1});
        st.push_back({b.left, 0});
        st.push_back({b.right, 0});
        continue;
      }
      const Dp& L = dp_[(size_t)b.left];
      const Dp& R = dp_[(size_t)b.right];
      float dist[9];
      uint8_t arg[9];
      for (int i = 2; i <= 8; i++) {
        float best = std::numeric_limits<float>::infinity();
        int bk = 1;
        for (int k = 1; k < i; k++) {
          const float c = L.f[k - 1] + R.f[i - k - 1];
          if (c < best) {
            best = c;
            bk = k;
          }
        }
        dist[i] = best;
        arg[i] = (uint8_t)bk;
      }
      const float own = b.box.area() * c_node + dist[8];
      d.f[0] = own;
      d.k[0] = 0;
      for (int i = 2; i <= 8; i++) {
        if (dist[i] < own) {
          d.f[i - 1] = dist[i];
          d.k[i - 1] = arg[i];
        } else {
          d.f[i - 1] = own;
          d.k[i - 1] = 0;
        }
      }
      d.k[7] = arg[8];  // a wide node rooted here always opens n itself
    }
    use_dp_ = true;
  }
  void dp_gather(const Tree& t, int n, int slots, int* kids, int* nkids) const {
    const BNode& b = t.nodes[(size_t)n];
    const int k = b.leaf() ? 0 : dp_[(size_t)n].k[slots - 1];
    if (k == 0 || slots == 1) {
      kids[(*nkids)++] = n;
      return;
    }
    dp_gather(t, b.left, k, kids, nkids);
    dp_gather(t, b.right, slots - k, kids, nkids);
  }

  // ---- emit a wide BVH from a binary tree; returns the root index in out_->nodes ---------------------------
  // The binary tree is cut into jobs: the top of the tree is expanded breadth-first by one thread until a few
  // hundred subtrees are open, those are emitted in parallel into private sinks (local indices), and everything is
  // spliced into out_ in job order -- the result does not depend on the thread count.
  // The traversal stack bounds the depth of the wide tree (jt_stage.cpp checks 2 * (top + BLAS levels) + 4 entries); the
  // dynamic programme does not, and now and then prefers long chains of sparsely filled nodes (coffee: 20 + 16 levels).
  // Such a tree is emitted again with the greedy rule.
  static constexpr int kMaxDpDepth = 13;
  int wide_depth(int root) const {
    int best = 0;
    std::vector<std::pair<int, int>> st;
    st.push_back({root, 1});
    while (!st.empty()) {
      auto [n, d] = st.back();
      st.pop_back();
      best = std::max(best, d);
      const JtWideNode& w = out_->nodes[(size_t)n];
      const int k = __builtin_popcount(w.imask);
      for (int i = 0; i < k; i++) st.push_back({(int)w.child_base + i, d + 1});
    }
    return best;
  }
  int emit(const Tree& t, int root) {
    if (root < 0) return -1;
    compute_dp(t, root);
    const size_t G0 = out_->nodes.size(), T0 = out_->tris.size();
    int r = emit_tree(t, root);
    if (use_dp_ && wide_depth(r) > kMaxDpDepth) {
      out_->nodes.resize(G0);
      out_->tris.resize(T0);
      for (int o = 0; o < 8; o++) out_->tri_rank[o].resize(T0);
      use_dp_ = false;
      r = emit_tree(t, root);
    }
    return r;
  }
  int emit_tree(const Tree& t, int root) {
    Sink top;
    top.nodes.push_back(JtWideNode());
    std::vector<Work> open;
    open.push_back(Work{0, root});
    size_t head = 0;
    while (head < open.size() && open.size() - head < 512) {
      Work w = open[head++];
      emit_node(t, top, w.wide, w.bnode, &open);
    }
    std::vector<Work> jobs(open.begin() + (long)head, open.end());
    std::vector<Sink> sinks(jobs.size());
    const bool big = (int64_t)t.recs.size() > parallel_min_;
#pragma omp parallel for schedule(dynamic, 1) if (big) num_threads(build_threads_)
    for (long j = 0; j < (long)jobs.size(); j++) {
      Sink& sk = sinks[(size_t)j];
      sk.nodes.push_back(JtWideNode());
      std::vector<Work> work;
      work.push_back(Work{0, jobs[(size_t)j].bnode});
      while (!work.empty()) {
        Work w = work.back();
        work.pop_back();
        emit_node(t, sk, w.wide, w.bnode, &work);
      }
    }
    // splice: `top` first, then the descendants of every job; a job's own node overwrites its placeholder in `top`
    const size_t G0 = out_->nodes.size(), T0 = out_->tris.size();
    size_t total_nodes = top.nodes.size(), total_tris = top.tris.size();
    for (const Sink& sk : sinks) {
      total_nodes += sk.nodes.size() - 1;
      total_tris += sk.tris.size();
    }
    out_->nodes.resize(G0 + total_nodes);
    out_->tris.resize(T0 + total_tris);
    for (int o = 0; o < 8; o++) out_->tri_rank[o].resize(T0 + total_tris);
    auto place = [&](const Sink& sk, size_t first_local, size_t node_dst, size_t child_shift, size_t tri_dst) {
      // nodes [first_local, end) of sk go to node_dst...; a local child index c maps to c + child_shift
      for (size_t i = first_local; i < sk.nodes.size(); i++) {
        JtWideNode n = sk.nodes[i];
        if (n.imask) n.child_base = (uint32_t)((size_t)n.child_base + child_shift);
        n.prim_base = (uint32_t)((size_t)n.prim_base + tri_dst);
        out_->nodes[node_dst + (i - first_local)] = n;
      }
      if (!sk.tris.empty()) memcpy(&out_->tris[tri_dst], sk.tris.data(), sk.tris.size() * sizeof(JtWideTri));
      for (int o = 0; o < 8; o++)
        if (!sk.rank[o].empty()) memcpy(&out_->tri_rank[o][tri_dst], sk.rank[o].data(), sk.rank[o].size() * 4);
    };
    place(top, 0, G0, G0, T0);
    std::vector<size_t> node_at(jobs.size()), tri_at(jobs.size());
    size_t gn = G0 + top.nodes.size(), gt = T0 + top.tris.size();
    for (size_t j = 0; j < jobs.size(); j++) {
      node_at[j] = gn;
      tri_at[j] = gt;
      gn += sinks[j].nodes.size() - 1;
      gt += sinks[j].tris.size();
    }
#pragma omp parallel for schedule(dynamic, 1) if (big) num_threads(build_threads_)
    for (long j = 0; j < (long)jobs.size(); j++) {
      const Sink& sk = sinks[(size_t)j];
      // descendants: local index c >= 1 lives at node_at + c - 1
      place(sk, 1, node_at[(size_t)j], node_at[(size_t)j] - 1, tri_at[(size_t)j]);
      JtWideNode n = sk.nodes[0];
      if (n.imask) n.child_base = (uint32_t)((size_t)n.child_base + node_at[(size_t)j] - 1);
      n.prim_base = (uint32_t)((size_t)n.prim_base + tri_at[(size_t)j]);
      out_->nodes[G0 + (size_t)jobs[(size_t)j].wide] = n;
    }
    return (int)G0;
  }

  void emit_node(const Tree& t, Sink& sk, int wide, int bnode, std::vector<Work>* work) {
    // 1. gather up to 8 children by repeatedly opening the largest internal child
    int kids[8];
    int nkids = 0;
    const BNode& top = t.nodes[(size_t)bnode];
    if (top.leaf()) {
      kids[nkids++] = bnode;
    } else if (use_dp_) {
      const int k = dp_[(size_t)bnode].k[7];
      dp_gather(t, top.left, k, kids, &nkids);
      dp_gather(t, top.right, 8 - k, kids, &nkids);
    } else {
      kids[nkids++] = top.left;
      kids[nkids++] = top.right;
      while (nkids < 8) {
        int best = -1;
        float best_area = -1.0f;
        for (int i = 0; i < nkids; i++) {
          const BNode& k = t.nodes[(size_t)kids[i]];
          if (k.leaf()) continue;
          float a = k.box.area();
          if (a > best_area) {
            best_area = a;
            best = i;
          }
        }
        if (best < 0) break;
        int open = kids[best];
        kids[best] = t.nodes[(size_t)open].left;
        kids[nkids++] = t.nodes[(size_t)open].right;
      }
    }
    // 2. node box and quantisation grid
    Box3 nb;
    nb.reset();
    for (int i = 0; i < nkids; i++) nb.add(t.nodes[(size_t)kids[i]].box);
    JtWideNode node;
    memset(&node, 0, sizeof(node));
    int ebias[3];
    float step[3];
    for (int a = 0; a < 3; a++) {
      node.p[a] = nb.lo[a];
      float ext = nb.hi[a] - nb.lo[a];
      int e = -126;
      if (ext > 0.0f && std::isfinite(ext)) {
        e = (int)std::ceil(std::log2((double)ext / 255.0));
        // make sure 255 * 2^e really covers the extent after float rounding of p + q*step
        while (std::ldexp(255.0, e) < (double)ext * (1.0 + 1e-6)) e++;
      }
      e = std::max(-126, std::min(127, e));
      ebias[a] = e + 127;
      node.e[a] = (uint8_t)ebias[a];
      step[a] = std::ldexp(1.0f, e);
    }
    // 3. slot assignment: greedy maximisation of dot(child centre - node centre, slot direction)
    float cen[3] = {0.5f * (nb.lo[0] + nb.hi[0]), 0.5f * (nb.lo[1] + nb.hi[1]), 0.5f * (nb.lo[2] + nb.hi[2])};
    int nk = nkids;
    int slot_of[8];
    bool slot_used[8] = {false, false, false, false, false, false, false, false};
    bool kid_done[8] = {false, false, false, false, false, false, false, false};
    float cost[8][8];
    for (int c = 0; c < nk; c++) {
      const Box3& b = t.nodes[(size_t)kids[c]].box;
      float d[3] = {0.5f * (b.lo[0] + b.hi[0]) - cen[0], 0.5f * (b.lo[1] + b.hi[1]) - cen[1],
                    0.5f * (b.lo[2] + b.hi[2]) - cen[2]};
      for (int s = 0; s < 8; s++)
        cost[c][s] = ((s & 1) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 4) ? d[2] : -d[2]);
    }
    for (int round = 0; round < nk; round++) {
      int bc = -1, bs = -1;
      float bv = -std::numeric_limits<float>::infinity();
      for (int c = 0; c < nk; c++) {
        if (kid_done[c]) continue;
        for (int s = 0; s < 8; s++) {
          if (slot_used[s]) continue;
          if (cost[c][s] > bv || bc < 0) {
            bv = cost[c][s];
            bc = c;
            bs = s;
          }
        }
      }
      kid_done[bc] = true;
      slot_used[bs] = true;
      slot_of[bc] = bs;
    }
    int kid_in_slot[8];
    for (int s = 0; s < 8; s++) kid_in_slot[s] = -1;
    for (int c = 0; c < nk; c++) kid_in_slot[slot_of[c]] = kids[c];
    // 4. children
    node.prim_base = (uint32_t)sk.tris.size();
    int ninternal = 0;
    for (int s = 0; s < 8; s++)
      if (kid_in_slot[s] >= 0 && !t.nodes[(size_t)kid_in_slot[s]].leaf()) ninternal++;
    node.child_base = (uint32_t)sk.nodes.size();
    sk.nodes.resize(sk.nodes.size() + (size_t)ninternal);
    int rel = 0;
    uint32_t prim_off = 0;
    for (int s = 0; s < 8; s++) {
      int k = kid_in_slot[s];
      if (k < 0) continue;
      const BNode& kn = t.nodes[(size_t)k];
      if (!kn.leaf()) {
        node.imask |= (uint8_t)(1u << s);
        node.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
        work->push_back(Work{(int)node.child_base + rel, k});
        rel++;
      } else {
        uint32_t unary = kn.count == 1 ? 1u : (kn.count == 2 ? 3u : 7u);
        node.meta[s] = (uint8_t)((unary << 5) | prim_off);
        for (int i = 0; i < kn.count; i++) {
          const Rec& r = t.recs[(size_t)t.leaf_store[(size_t)(kn.first + i)]];
          sk.tris.push_back(r.tri);
          for (int o = 0; o < 8; o++) sk.rank[o].push_back(r.rank[o]);
        }
        prim_off += (uint32_t)kn.count;
      }
      for (int a = 0; a < 3; a++) {
        double lo = ((double)kn.box.lo[a] - (double)node.p[a]) / (double)step[a];
        double hi = ((double)kn.box.hi[a] - (double)node.p[a]) / (double)step[a];
        int ql = (int)std::floor(lo), qh = (int)std::ceil(hi);
        ql = std::max(0, std::min(255, ql));
        qh = std::max(0, std::min(255, qh));
        // verify in float, as the kernel reconstructs: p + q * step
        while (ql > 0 && node.p[a] + (float)ql * step[a] > kn.box.lo[a]) ql--;
        while (qh < 255 && node.p[a] + (float)qh * step[a] < kn.box.hi[a]) qh++;
        node.qlo[a][s] = (uint8_t)ql;
        node.qhi[a][s] = (uint8_t)qh;
      }
    }
    sk.nodes[(size_t)wide] = node;
  }

  // ---- braiding of instanced BLASes into the top level ------------------------------------------------------------
  struct SubInfo {
    Box3 box;  // exact instance-space box of the sub-tree's triangles
    int count = 0;
  };
  static int slot_child(const JtWideNode& n, int s) {
    return (int)n.child_base + __builtin_popcount((unsigned)n.imask & ((1u << s) - 1u));
  }
  static int slot_count(const JtWideNode& n, int s) { return __builtin_popcount((unsigned)n.meta[s] >> 5); }
  static void tri_vertices(const JtWideTri& tr, float v[3][3]) {
    for (int k = 0; k < 3; k++) {
      v[0][k] = tr.p1[k];
      v[1][k] = tr.p1[k] + tr.e1[k];
      v[2][k] = tr.p1[k] + tr.e2[k];
    }
  }
  void subtree_info(int node, std::vector<SubInfo>& info) {
    const JtWideNode n = out_->nodes[(size_t)node];
    SubInfo r;
    r.box.reset();
    for (int s = 0; s < 8; s++) {
      if (!n.meta[s]) continue;
      if ((n.imask >> s) & 1) {
        int c = slot_child(n, s);
        subtree_info(c, info);
        r.box.add(info[(size_t)c].box);
        r.count += info[(size_t)c].count;
      } else {
        for (int i = 0; i < slot_count(n, s); i++) {
          float v[3][3];
          tri_vertices(out_->tris[(size_t)n.prim_base + (n.meta[s] & 31u) + (size_t)i], v);
          for (int k = 0; k < 3; k++) r.box.add(v[k]);
          r.count++;
        }
      }
    }
    info[(size_t)node] = r;
  }
  // world-space box of instance-space points, with a hair of slack: the exact tests run in instance space
  static void world_box(const float* frame, const float (*pts)[3], int npts, Box3* out) {
    out->reset();
    for (int i = 0; i < npts; i++) {
      float w[3];
      for (int k = 0; k < 3; k++)
        w[k] = ((frame[k] * pts[i][0] + frame[3 + k] * pts[i][1]) + frame[6 + k] * pts[i][2]) + frame[9 + k];
      out->add(w);
    }
    for (int k = 0; k < 3; k++) {
      float pad = 1e-5f * std::max(std::fabs(out->lo[k]), std::fabs(out->hi[k])) + 1e-30f;
      out->lo[k] -= pad;
      out->hi[k] += pad;
    }
  }
  static Box3 root_box(const JtHostShape& s) {  // transform_bbox's input, src/geometry.jl:70-86
    Box3 b;
    for (int k = 0; k < 3; k++) {
      b.lo[k] = s.ref_nodes[0].bbox_min[k];
      b.hi[k] = s.ref_nodes[0].bbox_max[k];
    }
    return b;
  }
  // entry record: "take the ray into instance `inst` and walk its BLAS from wide node `node`" (-1: the shape's root)
  void entry_record(int inst, const JtHostInstance& I, int node, const Box3& b, std::vector<int>* top, Tree& t) {
    float c[8][3];
    for (int i = 0; i < 8; i++)
      for (int k = 0; k < 3; k++) c[i][k] = ((i >> k) & 1) ? b.hi[k] : b.lo[k];
    Rec r;
    world_box(I.frame, c, 8, &r.box);
    memset(&r.tri, 0, sizeof(r.tri));
    r.tri.element = node;
    r.tri.instance = inst;
    r.tri.flags = 1u << 8;
    for (int o = 0; o < 8; o++) r.rank[o] = 0;
    top->push_back((int)t.recs.size());
    t.recs.push_back(r);
  }
  // world-space box of a sub-tree from its transformed triangles (tighter than the box of the 8 transformed corners
  // of its instance-space box, which is what matters for rotated plants)
  void subtree_world_box(int node, const float* frame, Box3* acc) {
    const JtWideNode n = out_->nodes[(size_t)node];
    for (int s = 0; s < 8; s++) {
      if (!n.meta[s]) continue;
      if ((n.imask >> s) & 1) {
        subtree_world_box(slot_child(n, s), frame, acc);
        continue;
      }
      for (int i = 0; i < slot_count(n, s); i++) {
        float v[3][3];
        tri_vertices(out_->tris[(size_t)n.prim_base + (n.meta[s] & 31u) + (size_t)i], v);
        Box3 b;
        world_box(frame, v, 3, &b);
        acc->add(b);
      }
    }
  }
  // number of top-level records braid() produces for the sub-tree at `node`
  int64_t braid_count(int node, const std::vector<SubInfo>& info) const {
    if (braid_max_ > 1 && info[(size_t)node].count <= braid_max_) return 1;
    const JtWideNode& n = out_->nodes[(size_t)node];
    int64_t c = 0;
    for (int s = 0; s < 8; s++) {
      if (!n.meta[s]) continue;
      c += ((n.imask >> s) & 1) ? braid_count(slot_child(n, s), info) : slot_count(n, s);
    }
    return c;
  }
  // writes the records of the sub-tree at `node` to dst[*cursor...] (dst is pre-sized: instances run in parallel)
  void braid(int node, int inst, const JtHostInstance& I, const std::vector<SubInfo>& info, Rec* dst, int64_t* cursor) {
    if (braid_max_ > 1 && info[(size_t)node].count <= braid_max_) {  // braid_max_ == 1: flatten every triangle
      Rec& r = dst[(*cursor)++];
      r.box.reset();
      subtree_world_box(node, I.frame, &r.box);
      memset(&r.tri, 0, sizeof(r.tri));
      r.tri.element = node;
      r.tri.instance = inst;
      r.tri.flags = 1u << 8;
      for (int o = 0; o < 8; o++) r.rank[o] = 0;
      return;
    }
    const JtWideNode n = out_->nodes[(size_t)node];
    for (int s = 0; s < 8; s++) {
      if (!n.meta[s]) continue;
      if ((n.imask >> s) & 1) {
        braid(slot_child(n, s), inst, I, info, dst, cursor);
        continue;
      }
      for (int i = 0; i < slot_count(n, s); i++) {  // triangles held directly by an opened node: flattened records
        size_t idx = (size_t)n.prim_base + (n.meta[s] & 31u) + (size_t)i;
        Rec& r = dst[(*cursor)++];
        r.tri = out_->tris[idx];
        r.tri.instance = inst;
        r.tri.flags |= 1u << 9;
        float v[3][3];
        tri_vertices(r.tri, v);
        world_box(I.frame, v, 3, &r.box);
        for (int o = 0; o < 8; o++) r.rank[o] = out_->tri_rank[o][idx];
      }
    }
  }

  int run(const std::vector<jt_bvh_node>& tlas_nodes, const std::vector<int64_t>& tlas_prims) {
    for (int o = 0; o < 8; o++) {
      out_->tri_rank[o].clear();
      out_->inst_rank[o].clear();
    }
    // BLAS of every non-empty shape (light probes walk single instances, so all are needed)
    const bool verbose = getenv("JT_BUILD_VERBOSE") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tb0 = now();
    out_->shape_root.assign(shapes_.size(), -1);
    for (size_t s = 0; s < shapes_.size(); s++) {
      Tree t;
      std::vector<int> recs;
      shape_records((int)s, -1, &recs, t);
      int root = sah_root(recs, t);
      out_->shape_root[s] = emit(t, root);
    }
    if (verbose) fprintf(stderr, "jt_build_wide: %zu BLAS in %.2f s\n", shapes_.size(), now() - tb0);
    visit_ranks(tlas_nodes, tlas_prims, (int64_t)insts_.size(), out_->inst_rank);
    Tree t;
    std::vector<int> top;
    std::vector<char> in_tlas(insts_.size(), 0);
    for (int64_t p : tlas_prims) in_tlas[(size_t)(p - 1)] = 1;
    // Instances with a non-identity frame are BRAIDED into the top-level tree: their BLAS is opened from the root
    // down to sub-trees of at most braid_max_ triangles; each such sub-tree becomes one entry record (instance +
    // BLAS node to start from, box = its instance-space box taken to world space), and the few triangles that sit
    // directly in the opened nodes become flattened leaf records (instance-space triangle, world-space box, ray
    // transformed at test time). Overlapping instances -- ecosys' 12.7 k plants -- are thereby separated at the
    // granularity of small sub-trees instead of whole shapes (tools/bvh_stats.py: 70.6 -> ~31 nodes per bounce
    // ray), nothing is duplicated but those few triangles, and the build stays a fraction of a second.
    // Policy (measured on B200, tools/exp_braid.sh): with a handful of instanced shapes (classroom 6, features1 6)
    // the two-level walk is as fast or faster, so they stay whole; scenes made of instances (ecosys: 12.7 k) are
    // flattened completely (braid_max_ = 1: 107 -> 214 Msamples/s; partial braids of 64 / 16 triangles give 117 /
    // 126) as long as the copies fit the record budget. JT_BRAID_MAX / JT_BRAID_MIN_INSTANCES override.
    int64_t n_instanced = 0, n_flat_records = 0;
    for (size_t inst = 0; inst < insts_.size(); inst++) {
      if (!in_tlas[inst] || insts_[inst].inlined) continue;
      const JtHostShape& s = shapes_[(size_t)insts_[inst].shape];
      n_instanced++;
      n_flat_records += s.num_elements() * (s.kind == 2 ? 2 : 1);
    }
    const bool braid_on = braid_max_ > 0 && n_instanced >= braid_min_instances_ &&
                          (braid_max_ > 1 || n_flat_records <= flatten_budget_);
    std::vector<SubInfo> info(out_->nodes.size());
    std::vector<char> info_done(shapes_.size(), 0);
    struct BraidJob {
      int inst;
      int64_t first, count;
    };
    std::vector<BraidJob> braids;
    for (size_t inst = 0; inst < insts_.size(); inst++) {
      if (!in_tlas[inst]) continue;
      const JtHostInstance& I = insts_[inst];
      const JtHostShape& s = shapes_[(size_t)I.shape];
      if (s.kind == 0 || s.ref_nodes.empty() || s.num_elements() == 0) continue;
      if (I.inlined) {
        shape_records(I.shape, (int)inst, &top, t);
        out_->inlined_instances++;
        continue;
      }
      const int root = out_->shape_root[(size_t)I.shape];
      if (braid_on && root >= 0) {
        if (!info_done[(size_t)I.shape]) {
          subtree_info(root, info);
          info_done[(size_t)I.shape] = 1;
        }
        braids.push_back(BraidJob{(int)inst, 0, braid_count(root, info)});
        out_->flattened_instances++;
      } else {
        entry_record((int)inst, I, -1, root_box(s), &top, t);
      }
      out_->instanced_instances++;
    }
    if (!braids.empty()) {
      int64_t first = (int64_t)t.recs.size();
      for (BraidJob& j : braids) {
        j.first = first;
        first += j.count;
      }
      t.recs.resize((size_t)first);
      const size_t top0 = top.size();
      top.resize(top0 + (size_t)(first - braids[0].first));
      const int64_t base = braids[0].first;
#pragma omp parallel for schedule(dynamic, 8) if (first - base > parallel_min_) num_threads(build_threads_)
      for (long j = 0; j < (long)braids.size(); j++) {
        const BraidJob& B = braids[(size_t)j];
        const JtHostInstance& I = insts_[(size_t)B.inst];
        int64_t cursor = B.first;
        braid(out_->shape_root[(size_t)I.shape], B.inst, I, info, t.recs.data(), &cursor);
        for (int64_t r = B.first; r < B.first + B.count; r++) top[top0 + (size_t)(r - base)] = (int)r;
      }
    }
    double t0 = now();
    if (verbose) fprintf(stderr, "jt_build_wide: %zu BLAS + top-level records in %.2f s\n", shapes_.size(), t0 - tb0);
    int root = sah_root(top, t);
    double t1 = now();
    out_->top_root = emit(t, root);
    if (verbose)
      fprintf(stderr, "jt_build_wide: top level %zu records: sah %.2f s, emit %.2f s, %zu wide nodes\n", top.size(), t1 - t0,
              now() - t1, out_->nodes.size());
    return JT_OK;
  }

 private:
  struct ShapeRank {
    bool built = false;
    std::vector<uint32_t> rank[8];
  };
  const std::vector<JtHostShape>& shapes_;
  const std::vector<JtHostInstance>& insts_;
  JtWideResult* out_;
  std::vector<ShapeRank> shape_ranks_;
  // Threads of the big builds: JT_BUILD_THREADS, else the OpenMP default -- except under torchrun, which exports
  // OMP_NUM_THREADS=1 to every rank: there the host cores are split evenly over the local ranks instead.
  static int default_build_threads() {
    if (const char* e = getenv("JT_BUILD_THREADS")) return std::max(1, atoi(e));
    int n = omp_get_max_threads();
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    if (lws && atoi(lws) > 0) n = std::max(n, omp_get_num_procs() / atoi(lws));
    return std::max(1, n);
  }
  int build_threads_ = default_build_threads();
  // trees below this many records are built on the calling thread (JT_BUILD_PARALLEL_MIN: lowered by the tests)
  int64_t parallel_min_ = getenv("JT_BUILD_PARALLEL_MIN") ? atoll(getenv("JT_BUILD_PARALLEL_MIN")) : 200000;
  bool small_sweep_ = false;
  int braid_max_ = getenv("JT_BRAID_MAX") ? atoi(getenv("JT_BRAID_MAX")) : 1;  // triangles per braided sub-tree; 1 = flatten; 0 = off
  int64_t braid_min_instances_ = getenv("JT_BRAID_MIN_INSTANCES") ? atoll(getenv("JT_BRAID_MIN_INSTANCES")) : 256;
  int64_t flatten_budget_ = (int64_t)48 << 20;  // records (48 B + 32 B of ranks each, plus ~0.2 nodes of 80 B)
  int leaf_max_ = getenv("JT_LEAF_MAX") ? atoi(getenv("JT_LEAF_MAX")) : 3;          // experiment knobs
  float tri_cost_ = getenv("JT_TRI_COST") ? (float)atof(getenv("JT_TRI_COST")) : 8.0f;
};

}  // namespace

int jt_build_wide(const std::vector<JtHostShape>& shapes, const std::vector<JtHostInstance>& instances,
                  const std::vector<jt_bvh_node>& tlas_nodes, const std::vector<int64_t>& tlas_prims,
                  JtWideResult* out) {
  for (const JtHostShape& s : shapes) {
    for (const jt_bvh_node& n : s.ref_nodes) {
      int64_t limit = n.internal ? (int64_t)s.ref_nodes.size() : (int64_t)s.ref_prims.size();
      int64_t last = n.internal ? n.start + 1 : n.start + n.num - 1;
      if (n.start < 1 || last > limit || (n.internal && (n.axis < 1 || n.axis > 3)))
        return jt_set_error(JT_ERR_INVALID, "shape BVH node out of range");
    }
    for (int64_t p : s.ref_prims)
      if (p < 1 || p > s.num_elements()) return jt_set_error(JT_ERR_INVALID, "shape BVH primitive id out of range");
  }
  for (const jt_bvh_node& n : tlas_nodes) {
    int64_t limit = n.internal ? (int64_t)tlas_nodes.size() : (int64_t)tlas_prims.size();
    int64_t last = n.internal ? n.start + 1 : n.start + n.num - 1;
    if (n.start < 1 || last > limit || (n.internal && (n.axis < 1 || n.axis > 3)))
      return jt_set_error(JT_ERR_INVALID, "scene BVH node out of range");
  }
  for (int64_t p : tlas_prims)
    if (p < 1 || p > (int64_t)instances.size()) return jt_set_error(JT_ERR_INVALID, "scene BVH instance id out of range");
  Collapser c(shapes, instances, out);
  return c.run(tlas_nodes, tlas_prims);
}
