// jt_host_bvh.cpp -- jt_make_bvh: the binary BVH the Julia host builds (src/bvh.jl:138-304).
//
// The GPU library consumes the host-built tree (north star: "bvh.jl builds it"); this C++
// builder exists for hosts that are not Julia (the Python mirror in this repo).
// It must produce the SAME tree as bvh.jl, node for node and slot for slot, because the tree's
// primitive order defines the reference's tie-breaks: same LIFO work order
// (right child expanded first), same Hoare partition swap sequence, same axis tie rules, and
// Julia's NaN-propagating min/max in every box merge.
//
// Works on 0-based flat arrays internally; emits the reference's 1-based 40-byte nodes.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

#include "../../include/jtrace_b200.h"
#include "jt_internal.h"

namespace {

struct Box {
  float lo[3], hi[3];
};

inline float nan_min(float a, float b) {  // Julia min(): NaN wins, -0 < +0
  if (a != a) return a;
  if (b != b) return b;
  if (a < b) return a;
  if (b < a) return b;
  return std::signbit(a) ? a : b;
}
inline float nan_max(float a, float b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a > b) return a;
  if (b > a) return b;
  return std::signbit(a) ? b : a;
}
inline Box void_box() {
  const float inf = std::numeric_limits<float>::infinity();
  return Box{{inf, inf, inf}, {-inf, -inf, -inf}};
}
inline void grow(Box& b, const float* lo, const float* hi) {
  for (int k = 0; k < 3; k++) {
    b.lo[k] = nan_min(b.lo[k], lo[k]);
    b.hi[k] = nan_max(b.hi[k], hi[k]);
  }
}

struct Builder {
  const float* boxes;  // n x 6
  int64_t n;
  std::vector<float> cent;     // n x 3, (min + max) / 2
  std::vector<int64_t> order;  // 0-based primitive ids, permuted in place

  const float* lo(int64_t p) const { return boxes + 6 * p; }
  const float* hi(int64_t p) const { return boxes + 6 * p + 3; }
  float c(int64_t slot, int axis) const { return cent[3 * order[slot] + axis]; }

  // Hoare partition over inclusive slots [first, last]; returns the last slot of the low side
  // (may be first-1). src/bvh.jl:281-304
  int64_t hoare(int axis, float pivot, int64_t first, int64_t last) {
    int64_t a = first, b = last;
    for (;;) {
      while (a <= last && c(a, axis) < pivot) ++a;
      while (b >= first && c(b, axis) >= pivot) --b;
      if (a >= b) return b;
      std::swap(order[a], order[b]);
    }
  }

  Box centroid_box(int64_t first, int64_t last) const {
    Box cb = void_box();
    for (int64_t s = first; s <= last; s++) {
      const float* p = &cent[3 * order[s]];
      grow(cb, p, p);
    }
    return cb;
  }

  static int64_t median(int64_t first, int64_t last) {
    // div(left + right + 1, 2) on 1-based inclusive bounds == this on 0-based ones
    return (first + last + 3) / 2 - 1;
  }

  // src/bvh.jl:185-216
  void middle_split(int64_t first, int64_t last, int64_t* mid, int* axis) {
    Box cb = centroid_box(first, last);
    float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
    if (ext[0] == 0.0f && ext[1] == 0.0f && ext[2] == 0.0f) {
      *mid = median(first, last);
      *axis = 0;
      return;
    }
    int ax = 0;  // three sequential >= tests: later axes win ties
    if (ext[1] >= ext[0] && ext[1] >= ext[2]) ax = 1;
    if (ext[2] >= ext[0] && ext[2] >= ext[1]) ax = 2;
    // note: if no test fires for axis 0 either (NaN extents) the default stays 0 like Julia's 1
    float pivot = (cb.lo[ax] + cb.hi[ax]) / 2.0f;
    int64_t m = hoare(ax, pivot, first, last);
    if (m < first || m > last) m = median(first, last);
    *mid = m;
    *axis = ax;
  }

  static float area(const Box& b) {  // src/bvh.jl:276-279
    float sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    return ((0.000000000001f + (2.0f * sx) * sy) + (2.0f * sx) * sz) + (2.0f * sy) * sz;
  }

  // src/bvh.jl:218-274
  void sah_split(int64_t first, int64_t last, int64_t* mid, int* axis) {
    Box cb = centroid_box(first, last);
    float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
    if (ext[0] == 0.0f && ext[1] == 0.0f && ext[2] == 0.0f) {
      *mid = median(first, last);
      *axis = 0;
      return;
    }
    const int bins = 16;
    int best_axis = 0;
    float best_pivot = 0.0f;
    float best = std::numeric_limits<float>::infinity();
    const float whole = area(cb);
    for (int ax = 0; ax < 3; ax++) {
      for (int b = 1; b < bins; b++) {
        float pivot = cb.lo[ax] + ((float)b * ext[ax]) / (float)bins;
        Box l = void_box(), r = void_box();
        int64_t nl = 0, nr = 0;
        for (int64_t s = first; s <= last; s++) {
          int64_t p = order[s];
          if (cent[3 * p + ax] < pivot) {
            grow(l, lo(p), hi(p));
            nl++;
          } else {
            grow(r, lo(p), hi(p));
            nr++;
          }
        }
        float cost = (1.0f + ((float)nl * area(l)) / whole) + ((float)nr * area(r)) / whole;
        if (cost < best) {
          best = cost;
          best_pivot = pivot;
          best_axis = ax;
        }
      }
    }
    int64_t m = hoare(best_axis, best_pivot, first, last);
    if (m == first || m == last) m = median(first, last);
    *mid = m;
    *axis = best_axis;
  }
};

}  // namespace

extern "C" int jt_make_bvh(const float* bboxes, int64_t n, int high_quality, jt_bvh_node* nodes_out,
                           int64_t* num_nodes_out, int64_t* primitives_out) {
  if ((n > 0 && !bboxes) || !nodes_out || !num_nodes_out || (n > 0 && !primitives_out) || n < 0)
    return jt_set_error(JT_ERR_INVALID, "jt_make_bvh: null argument");
  Builder B;
  B.boxes = bboxes;
  B.n = n;
  B.cent.resize(3 * (size_t)n);
  B.order.resize((size_t)n);
  for (int64_t p = 0; p < n; p++) {
    B.order[p] = p;
    for (int k = 0; k < 3; k++) B.cent[3 * p + k] = (bboxes[6 * p + k] + bboxes[6 * p + 3 + k]) / 2.0f;
  }
  struct Job {
    int64_t node, first, last;
  };
  std::vector<Job> todo;
  int64_t count = 1;
  todo.push_back(Job{0, 0, n - 1});
  const int64_t capacity = 2 * n + 1;
  while (!todo.empty()) {
    Job j = todo.back();
    todo.pop_back();
    Box bb = void_box();
    for (int64_t s = j.first; s <= j.last; s++) grow(bb, B.lo(B.order[s]), B.hi(B.order[s]));
    jt_bvh_node& out = nodes_out[j.node];
    memset(&out, 0, sizeof(out));
    for (int k = 0; k < 3; k++) {
      out.bbox_min[k] = bb.lo[k];
      out.bbox_max[k] = bb.hi[k];
    }
    int64_t size = j.last - j.first + 1;
    if (size > 4) {  // BVH_MAX_PRIMS, src/bvh.jl:32
      int64_t mid;
      int axis;
      if (high_quality) B.sah_split(j.first, j.last, &mid, &axis);
      else B.middle_split(j.first, j.last, &mid, &axis);
      if (count + 2 > capacity) return jt_set_error(JT_ERR_INTERNAL, "jt_make_bvh: node overflow");
      out.start = count + 1;  // 1-based index of the first of two consecutive children
      out.num = 2;
      out.axis = (int8_t)(axis + 1);
      out.internal = 1;
      todo.push_back(Job{count, j.first, mid});
      todo.push_back(Job{count + 1, mid + 1, j.last});
      count += 2;
    } else {
      out.start = j.first + 1;
      out.num = (int16_t)size;
      out.axis = 1;
      out.internal = 0;
    }
  }
  for (int64_t p = 0; p < n; p++) primitives_out[p] = B.order[p] + 1;
  *num_nodes_out = count;
  return JT_OK;
}
