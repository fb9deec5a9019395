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
#include <algorithm>
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

struct BNode {  // intermediate binary tree
  Box3 box;
  int left = -1, right = -1;  // children (internal)
  int first = 0, count = 0;   // records (leaf, count <= 3)
  bool leaf() const { return left < 0; }
};

struct Tree {
  std::vector<BNode> nodes;
  std::vector<Rec> recs;
};

// An item is either an already-built subtree or a single record.
struct Item {
  int node;  // >= 0: subtree root in Tree::nodes
  int rec;   // >= 0: record index (pending, not yet in a leaf)
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

  // ---- generic: binary tree over a list of items ------------------------------------------------------
  int tree_over(std::vector<Item>& items, size_t a, size_t b, Tree& t) {
    size_t n = b - a;
    if (n == 0) return -1;
    bool all_recs = true;
    for (size_t i = a; i < b; i++) all_recs = all_recs && items[i].rec >= 0;
    if (all_recs && n <= 3) {
      // records of one leaf must be contiguous in t.recs: they are created in order, but be safe
      BNode leaf;
      leaf.box.reset();
      leaf.first = (int)leaf_store_.size();
      leaf.count = (int)n;
      for (size_t i = a; i < b; i++) {
        leaf_store_.push_back(items[i].rec);
        leaf.box.add(t.recs[(size_t)items[i].rec].box);
      }
      t.nodes.push_back(leaf);
      return (int)t.nodes.size() - 1;
    }
    if (n == 1) return items[a].node;
    size_t m = a + n / 2;
    int l = tree_over(items, a, m, t);
    int r = tree_over(items, m, b, t);
    if (l < 0) return r;
    if (r < 0) return l;
    BNode in;
    in.left = l;
    in.right = r;
    in.box = t.nodes[(size_t)l].box;
    in.box.add(t.nodes[(size_t)r].box);
    t.nodes.push_back(in);
    return (int)t.nodes.size() - 1;
  }

  // ---- binary tree of a shape (optionally with a baked instance id) -------------------------------------
  int shape_tree(int shape_id, int inst, Tree& t) {
    const JtHostShape& s = shapes_[(size_t)shape_id];
    if (s.kind == 0 || s.ref_nodes.empty() || s.num_elements() == 0) return -1;
    const std::vector<uint32_t>* erank = shape_rank(shape_id);
    std::function<int(int64_t)> conv = [&](int64_t ni) -> int {
      const jt_bvh_node& n = s.ref_nodes[(size_t)ni];
      if (n.internal) {
        int l = conv(n.start - 1);
        int r = conv(n.start);
        if (l < 0) return r;
        if (r < 0) return l;
        BNode in;
        in.left = l;
        in.right = r;
        in.box = t.nodes[(size_t)l].box;
        in.box.add(t.nodes[(size_t)r].box);
        t.nodes.push_back(in);
        return (int)t.nodes.size() - 1;
      }
      std::vector<int> recs;
      for (int64_t i = n.start - 1; i < n.start - 1 + n.num; i++)
        element_records(s, s.ref_prims[(size_t)i] - 1, inst, erank, &recs, t);
      std::vector<Item> items;
      for (int r : recs) items.push_back(Item{-1, r});
      return tree_over(items, 0, items.size(), t);
    };
    return conv(0);
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

  // ---- binned-SAH binary tree over record indices (the fast mode's own topology) -------------------------
  // The reference tree (centroid middle split, src/bvh.jl:185-216) only defines tie-break ranks; which boxes
  // a ray visits is free, so the wide BVH is collapsed from a surface-area-heuristic tree instead:
  // 16 bins per axis, leaves of <= 3 records (the per-slot limit of the wide node).
  int sah_tree(std::vector<int>& recs, size_t a, size_t b, Tree& t) {
    size_t n = b - a;
    if (n == 0) return -1;
    Box3 box, cbox;
    box.reset();
    cbox.reset();
    for (size_t i = a; i < b; i++) {
      const Box3& rb = t.recs[(size_t)recs[i]].box;
      box.add(rb);
      float c[3] = {0.5f * (rb.lo[0] + rb.hi[0]), 0.5f * (rb.lo[1] + rb.hi[1]), 0.5f * (rb.lo[2] + rb.hi[2])};
      cbox.add(c);
    }
    auto make_leaf = [&]() {
      BNode leaf;
      leaf.box = box;
      leaf.first = (int)leaf_store_.size();
      leaf.count = (int)n;
      for (size_t i = a; i < b; i++) leaf_store_.push_back(recs[i]);
      t.nodes.push_back(leaf);
      return (int)t.nodes.size() - 1;
    };
    if (n == 1) return make_leaf();
    const int NB = 16;
    int best_axis = -1, best_bin = -1;
    float best_cost = std::numeric_limits<float>::infinity(), best_tri_cost = 0.0f;
    for (int ax = 0; ax < 3; ax++) {
      float lo = cbox.lo[ax], ext = cbox.hi[ax] - cbox.lo[ax];
      if (!(ext > 0.0f)) continue;
      Box3 bins[NB];
      int cnt[NB];
      for (int k = 0; k < NB; k++) {
        bins[k].reset();
        cnt[k] = 0;
      }
      float scale = (float)NB / ext;
      for (size_t i = a; i < b; i++) {
        const Box3& rb = t.recs[(size_t)recs[i]].box;
        int k = (int)((0.5f * (rb.lo[ax] + rb.hi[ax]) - lo) * scale);
        k = std::max(0, std::min(NB - 1, k));
        bins[k].add(rb);
        cnt[k]++;
      }
      float right_area[NB];
      int right_cnt[NB];
      Box3 acc;
      acc.reset();
      int c = 0;
      for (int k = NB - 1; k > 0; k--) {
        acc.add(bins[k]);
        c += cnt[k];
        right_area[k] = acc.area();
        right_cnt[k] = c;
      }
      acc.reset();
      c = 0;
      for (int k = 0; k < NB - 1; k++) {
        acc.add(bins[k]);
        c += cnt[k];
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
      float lo = cbox.lo[best_axis], scale = (float)NB / (cbox.hi[best_axis] - cbox.lo[best_axis]);
      auto mid = std::partition(recs.begin() + (long)a, recs.begin() + (long)b, [&](int r) {
        const Box3& rb = t.recs[(size_t)r].box;
        int k = (int)((0.5f * (rb.lo[best_axis] + rb.hi[best_axis]) - lo) * scale);
        k = std::max(0, std::min(NB - 1, k));
        return k <= best_bin;
      });
      m = (size_t)(mid - recs.begin());
      if (m == a || m == b) m = a + n / 2;
    }
    int l = sah_tree(recs, a, m, t);
    int r = sah_tree(recs, m, b, t);
    BNode in;
    in.left = l;
    in.right = r;
    in.box = t.nodes[(size_t)l].box;
    in.box.add(t.nodes[(size_t)r].box);
    t.nodes.push_back(in);
    return (int)t.nodes.size() - 1;
  }

  // all records of a shape (reference element order), optionally with a baked instance id
  void shape_records(int shape_id, int inst, std::vector<int>* recs, Tree& t) {
    const JtHostShape& s = shapes_[(size_t)shape_id];
    if (s.kind == 0 || s.ref_nodes.empty() || s.num_elements() == 0) return;
    const std::vector<uint32_t>* erank = shape_rank(shape_id);
    for (int64_t e = 0; e < s.num_elements(); e++) element_records(s, e, inst, erank, recs, t);
  }

  // ---- emit a wide BVH from a binary tree; returns the root index in out_->nodes ---------------------------
  int emit(const Tree& t, int root) {
    if (root < 0) return -1;
    int root_index = (int)out_->nodes.size();
    out_->nodes.push_back(JtWideNode());
    std::vector<Work> work;
    work.push_back(Work{root_index, root});
    while (!work.empty()) {
      Work w = work.back();
      work.pop_back();
      emit_node(t, w.wide, w.bnode, &work);
    }
    return root_index;
  }

  void emit_node(const Tree& t, int wide, int bnode, std::vector<Work>* work) {
    // 1. gather up to 8 children by repeatedly opening the largest internal child
    std::vector<int> kids;
    const BNode& top = t.nodes[(size_t)bnode];
    if (top.leaf()) {
      kids.push_back(bnode);
    } else {
      kids.push_back(top.left);
      kids.push_back(top.right);
      while (kids.size() < 8) {
        int best = -1;
        float best_area = -1.0f;
        for (size_t i = 0; i < kids.size(); i++) {
          const BNode& k = t.nodes[(size_t)kids[i]];
          if (k.leaf()) continue;
          float a = k.box.area();
          if (a > best_area) {
            best_area = a;
            best = (int)i;
          }
        }
        if (best < 0) break;
        int open = kids[(size_t)best];
        kids[(size_t)best] = t.nodes[(size_t)open].left;
        kids.push_back(t.nodes[(size_t)open].right);
      }
    }
    // 2. node box and quantisation grid
    Box3 nb;
    nb.reset();
    for (int k : kids) nb.add(t.nodes[(size_t)k].box);
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
    int nk = (int)kids.size();
    int slot_of[8];
    bool slot_used[8] = {false, false, false, false, false, false, false, false};
    bool kid_done[8] = {false, false, false, false, false, false, false, false};
    float cost[8][8];
    for (int c = 0; c < nk; c++) {
      const Box3& b = t.nodes[(size_t)kids[(size_t)c]].box;
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
    for (int c = 0; c < nk; c++) kid_in_slot[slot_of[c]] = kids[(size_t)c];
    // 4. children
    node.prim_base = (uint32_t)out_->tris.size();
    int ninternal = 0;
    for (int s = 0; s < 8; s++)
      if (kid_in_slot[s] >= 0 && !t.nodes[(size_t)kid_in_slot[s]].leaf()) ninternal++;
    node.child_base = (uint32_t)out_->nodes.size();
    out_->nodes.resize(out_->nodes.size() + (size_t)ninternal);
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
          const Rec& r = t.recs[(size_t)leaf_store_[(size_t)(kn.first + i)]];
          out_->tris.push_back(r.tri);
          for (int o = 0; o < 8; o++) out_->tri_rank[o].push_back(r.rank[o]);
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
    out_->nodes[(size_t)wide] = node;
  }

  int run(const std::vector<jt_bvh_node>& tlas_nodes, const std::vector<int64_t>& tlas_prims) {
    for (int o = 0; o < 8; o++) {
      out_->tri_rank[o].clear();
      out_->inst_rank[o].clear();
    }
    // BLAS of every non-empty shape (light probes walk single instances, so all are needed)
    out_->shape_root.assign(shapes_.size(), -1);
    for (size_t s = 0; s < shapes_.size(); s++) {
      Tree t;
      leaf_store_.clear();
      std::vector<int> recs;
      shape_records((int)s, -1, &recs, t);
      int root = sah_tree(recs, 0, recs.size(), t);
      out_->shape_root[s] = emit(t, root);
    }
    // top level: ONE surface-area tree over every triangle of the identity-frame instances (inlined: the
    // two-level structure is pure overhead for them) plus one reference record per
    // remaining instance. Order follows the reference TLAS only through the rank tables.
    visit_ranks(tlas_nodes, tlas_prims, (int64_t)insts_.size(), out_->inst_rank);
    Tree t;
    leaf_store_.clear();
    std::vector<int> top;
    std::vector<char> in_tlas(insts_.size(), 0);
    for (int64_t p : tlas_prims) in_tlas[(size_t)(p - 1)] = 1;
    for (size_t inst = 0; inst < insts_.size(); inst++) {
      if (!in_tlas[inst]) continue;
      const JtHostInstance& I = insts_[inst];
      const JtHostShape& s = shapes_[(size_t)I.shape];
      if (s.kind == 0 || s.ref_nodes.empty() || s.num_elements() == 0) continue;
      if (I.inlined) {
        shape_records(I.shape, (int)inst, &top, t);
        out_->inlined_instances++;
      } else {
        // instance-reference record; box = transform_bbox(frame, BLAS root box) (src/geometry.jl:70-86)
        const jt_bvh_node& rootn = s.ref_nodes[0];
        Rec r;
        r.box.reset();
        for (int c = 0; c < 8; c++) {
          float p[3] = {(c & 4) ? rootn.bbox_max[0] : rootn.bbox_min[0], (c & 2) ? rootn.bbox_max[1] : rootn.bbox_min[1],
                        (c & 1) ? rootn.bbox_max[2] : rootn.bbox_min[2]};
          float w[3];
          for (int k = 0; k < 3; k++)
            w[k] = ((I.frame[k] * p[0] + I.frame[3 + k] * p[1]) + I.frame[6 + k] * p[2]) + I.frame[9 + k];
          r.box.add(w);
        }
        // a hair of slack: the BLAS is tested in instance space, not against this box
        for (int k = 0; k < 3; k++) {
          float pad = 1e-5f * std::max(std::fabs(r.box.lo[k]), std::fabs(r.box.hi[k])) + 1e-30f;
          r.box.lo[k] -= pad;
          r.box.hi[k] += pad;
        }
        memset(&r.tri, 0, sizeof(r.tri));
        r.tri.element = -1;
        r.tri.instance = (int)inst;
        r.tri.flags = 1u << 8;
        for (int o = 0; o < 8; o++) r.rank[o] = 0;
        top.push_back((int)t.recs.size());
        t.recs.push_back(r);
        out_->instanced_instances++;
      }
    }
    int root = sah_tree(top, 0, top.size(), t);
    out_->top_root = emit(t, root);
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
  int leaf_max_ = getenv("JT_LEAF_MAX") ? atoi(getenv("JT_LEAF_MAX")) : 3;          // experiment knobs
  float tri_cost_ = getenv("JT_TRI_COST") ? (float)atof(getenv("JT_TRI_COST")) : 8.0f;
  std::vector<int> leaf_store_;  // record indices of leaves, contiguous per leaf
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
